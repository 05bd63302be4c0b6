"""Developer probe: shared-memory event log of the single-right-hand-side cluster sweep kernel (DBG instantiation):
clock64 stamps of 32 strips starting at HP_DBG_WIN, 8 events for the critical group and 8 for the off-path group of
every CTA, plus (globaltimer, clock64) pairs at both ends of the kernel to align the SM clocks.  (GPU box)
    HP_DBG_WIN=2000 python tools/timeline4.py out.npy"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n, b = 4096, 12
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner()
L = s.layout(); G = L["G"]
u = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
nst = n - 1 - b
s.sweep_forward(u, b + 1, n - 1); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.sweep_forward(u, b + 1, n - 1); e1.record(); torch.cuda.synchronize()
print(f"plain: {1e3 * e0.elapsed_time(e1) / nst:.3f} us/strip")
s.lib.hp_debug_phases(s.handle, 1, None)
e0.record(); s.sweep_forward(u, b + 1, n - 1); e1.record(); torch.cuda.synchronize()
print(f"instrumented: {1e3 * e0.elapsed_time(e1) / nst:.3f} us/strip, status {s.sweep_status()}")
out = np.zeros(G * (16 + 1024), dtype=np.int64)
s.lib.hp_debug_phases(s.handle, 0, out.ctypes.data)
ph = out[:G * 16].reshape(G, 16)
print("polling rounds per strip (warp 0 of every CTA): mean %.2f min %.2f max %.2f; D poll cycles per strip mean %.0f -> %.0f cycles per round" % (
    ph[:, 6].mean() / nst, ph[:, 6].min() / nst, ph[:, 6].max() / nst, ph[:, 4].mean() / nst, ph[:, 4].sum() / max(1, ph[:, 6].sum())))
np.save(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline4s.npy", out)
