"""Developer probe: device time of one preconditioner application for R = 1, 2, 4, 8 right-hand sides per launch (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner()
print("layout", {k: s.layout()[k] for k in ("P", "K", "CW", "QP", "NRQ", "NXG")}, "multi_max", s.multi_max, flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g) for _ in range(8)]
ref = [s.precond_apply(x) for x in xs]
for R in [int(x) for x in os.environ.get("RS", "1,2,4,8").split(",")]:
    if R > s.multi_max:
        break
    outs = [torch.empty_like(x) for x in xs[:R]]
    s.precond_apply_multi(xs[:R], outs)
    torch.cuda.synchronize()
    err = max((torch.linalg.norm(o - r) / torch.linalg.norm(r)).item() for o, r in zip(outs, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        s.precond_apply_multi(xs[:R], outs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"R={R}: {ms:8.2f} ms per application  {ms / R:8.2f} ms per right-hand side  {1e3 * ms / (2 * (n - b)):6.2f} us per strip   max rel diff vs single {err:.2e}  status {s.sweep_status()}", flush=True)
