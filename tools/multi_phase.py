"""Developer probe: per-phase cycle sums of the multi-vector sweep kernel (DBG instantiation) for R right-hand sides. (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Rs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8]
b = 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner()
L = s.layout()
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randn(n * n, dtype=torch.complex128, device="cuda", generator=g) for _ in range(8)]
names = ["pre (GL,GF,R)", "A wait x3", "B rho+bar", "C rows", "D poll", "D sum+send", "c W compute only", "w4: G wait + a + poll",
         "w4: wait V", "w4: W + tail bars", "b wait x3", "b corr+send", "c wait V", "c W barrier", "c tail", "(W chunk waits)"]
nst = n - 1 - b
for R in Rs:
    bufs = [x.clone() for x in xs[:R]]
    s.sweep_forward_multi_buf(bufs, 0, b + 1, n - 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sweep_forward_multi_buf(bufs, 0, b + 1, n - 1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    s.lib.hp_debug_phases(s.handle, 1, None)
    s.sweep_forward_multi_buf(bufs, 0, b + 1, n - 1); torch.cuda.synchronize()
    raw = np.zeros(L["G"] * (16 + 1024), dtype=np.int64)
    s.lib.hp_debug_phases(s.handle, 0, raw.ctypes.data)
    out = raw[:L["G"] * 16].reshape(L["G"], 16)
    print(f"R={R}: forward sweep {ms:.2f} ms = {1e3 * ms / nst:.2f} us/strip = {1.965e3 * 1e3 * ms / nst / 1e3:.0f} cycles/strip; phases (mean | min | max over CTAs), status {s.sweep_status()}")
    for i, nm in enumerate(names):
        if nm != "-":
            print(f"   {nm:16s} {out[:, i].mean() / nst:9.0f} {out[:, i].min() / nst:9.0f} {out[:, i].max() / nst:9.0f}")
    np.save(f"gpurun_out/multi_phase_R{R}{os.environ.get('TAG', '')}.npy", out)
    print("   sums: critical %.0f  off-path %.0f" % (out[:, :6].sum(1).mean() / nst, out[:, 8:15].sum(1).mean() / nst), flush=True)
