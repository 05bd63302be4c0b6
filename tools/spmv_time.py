"""Developer probe: stencil SpMV time at n^2 (3 rotating inputs larger than L2), for the march length in HP_SPMV_ROWS. (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
g = torch.Generator(device="cuda").manual_seed(3)
xs = [torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g) for _ in range(3)]
ys = [torch.empty_like(x) for x in xs]
for i in range(6):
    s.matvec(xs[i % 3], ys[i % 3])
torch.cuda.synchronize()
reps = 60
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    s.matvec(xs[i % 3], ys[i % 3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"HP_SPMV_ROWS={os.environ.get('HP_SPMV_ROWS', 'auto')}: {1e3 * ms:.1f} us, {40 * n * n / ms / 1e6:.0f} GB/s", flush=True)
