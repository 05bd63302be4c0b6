import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
n, b, strips = int(sys.argv[1]), 12, int(sys.argv[2])
variant = int(sys.argv[3])
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
m_hi = b + strips
s.setup_preconditioner(m_lo=b + 1, m_hi=m_hi)
s.set_sweep_variant(variant)
print(s.layout()["P"], s.layout()["K"], flush=True)
u = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
for i in range(3):
    s.sweep_forward(u, b + 1, m_hi - 1)
    torch.cuda.synchronize()
    print("sweep", i, "ok status", s.sweep_status(), flush=True)
