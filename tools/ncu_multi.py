"""ncu target: one forward sweep of the multi-vector kernel with R right-hand sides at n (GPU box; run under ncu -k regex:hp_sweep4m)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m_hi = int(sys.argv[3]) if len(sys.argv) > 3 else n
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner(m_lo=b + 1, m_hi=m_hi)
g = torch.Generator(device="cuda").manual_seed(1)
bufs = [torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g) for _ in range(R)]
s.sweep_forward_multi_buf(bufs, 0, b + 1, min(m_hi, n - 1))
torch.cuda.synchronize()
print("status", s.sweep_status())
