"""Developer tool: where the wall time of bench.py's setup clock goes (HelmholtzSolver + setup_preconditioner after the
96^2 warm-up solve).  HP_SETUP_TRACE=1 adds the phases inside hp_setup_strips.  (GPU box)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n, b = 4096, 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
om0 = 2 * np.pi * 9.6 + 2j
c0, f0 = hp.init_layered_f1(om0, 96)
s0 = hp.HelmholtzSolver(96, b, om0, 100.0, c0)
s0.setup_preconditioner()
s0.precond_apply(torch.from_numpy(f0.ravel().astype(np.complex128)).cuda())
torch.cuda.synchronize()
s0.close()
del s0
for rep in range(2):
    t0 = time.time()
    s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
    torch.cuda.synchronize()
    t1 = time.time()
    s.setup_preconditioner()
    t2 = time.time()
    torch.cuda.synchronize()
    t3 = time.time()
    print(f"rep {rep}: solver {1e3 * (t1 - t0):.1f} ms, setup call {1e3 * (t2 - t1):.1f} ms, sync {1e3 * (t3 - t2):.1f} ms, device {s.setup_ms:.1f} ms", flush=True)
    s.close()
    del s
