"""Developer tool: set up a problem and run a few forward sweeps (target of ncu captures).  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = int(sys.argv[2]) if len(sys.argv) > 2 else 12
strips = int(sys.argv[3]) if len(sys.argv) > 3 else 0
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
m_hi = n if not strips else b + strips
s.setup_preconditioner(m_lo=b + 1, m_hi=m_hi)
u = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
for _ in range(3):
    s.sweep_forward(u, b + 1, m_hi - 1)
torch.cuda.synchronize()
print("status", s.sweep_status())
