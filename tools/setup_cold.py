"""Developer tool: ONE strip setup in a fresh process (what a single solve sees), optionally after keeping the GPU busy
for a while (argv[2] = milliseconds of warm-up work).  HP_SETUP_TRACE=1 prints the host-side phases.  (GPU box)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
torch.cuda.init()
if warm_ms > 0:
    a = torch.randn(8192, 8192, device="cuda")
    t0 = time.time()
    while (time.time() - t0) * 1e3 < warm_ms:
        a @ a
        torch.cuda.synchronize()
torch.cuda.synchronize()
t0 = time.time()
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
t1 = time.time()
s.setup_preconditioner()
torch.cuda.synchronize()
t2 = time.time()
print(f"warm {warm_ms:.0f} ms  scratch cap {os.environ.get('HP_SCRATCH_GB', '64')} GB:  solver {1e3 * (t1 - t0):.1f} ms, setup wall {1e3 * (t2 - t1):.1f} ms, device {s.setup_ms:.1f} ms", flush=True)
