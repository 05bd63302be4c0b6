"""Developer probe: where does a strip step of the sweep kernel spend its cycles?  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = int(sys.argv[2]) if len(sys.argv) > 2 else 12
P = int(sys.argv[3]) if len(sys.argv) > 3 else 0
K = int(sys.argv[4]) if len(sys.argv) > 4 else 0
omega = 2 * np.pi * n / 10 + 2j
c_mat, f_mat = hp.init_layered_f1(omega, n)
s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
s.setup_preconditioner(P=P, K=K)
L = s.layout()
print({k: int(L[k]) for k in ("P", "K", "G", "QP", "CW", "NS", "NR", "PK")}, "setup ms", s.setup_ms)
u = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
for variant in (tuple(int(v) for v in sys.argv[5].split(",")) if len(sys.argv) > 5 else (3,)):
    s.set_sweep_variant(variant)
    s.sweep_forward(u, b + 1, n - 1)
    torch.cuda.synchronize()
    s.lib.hp_debug_phases(s.handle, 1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sweep_forward(u, b + 1, n - 1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    raw = np.zeros(L["G"] * (16 + 1024), dtype=np.int64)
    s.lib.hp_debug_phases(s.handle, 0, raw.ctypes.data)
    out = raw[:L["G"] * 16].reshape(L["G"], 16) if variant >= 3 else np.pad(raw[:L["G"] * 8].reshape(L["G"], 8), ((0, 0), (0, 8)))
    nst = n - 1 - b
    names = (["reduce", "B gather+bar", "B rows", "A gather+bar", "A end bar", "B tma wait", "A tma wait", "A rows", "a tma", "a gb", "b wait xs", "b corr", "c gather", "c W", "c tail", "-"]
             if variant >= 3 else ["tma_wait", "S1", "reduce_wait", "S2_poll", "S2_compute", "S3_poll", "S3"])
    print(f"variant {variant}: {ms:.2f} ms, {1e3 * ms / nst:.2f} us/strip; cycles/strip (mean over CTAs | min | max):")
    for i, nm in enumerate(names):
        print(f"   {nm:9s} {out[:, i].mean() / nst:9.0f} {out[:, i].min() / nst:9.0f} {out[:, i].max() / nst:9.0f}")
    red = np.arange(L["G"]) % L["K"] == 0
    if variant >= 3:
        print("   reducers only: C1 wait xs %.0f  C1 gpb+M %.0f  C2 wait GR %.0f" % tuple(out[red, i].mean() / nst for i in (0, 1, 2)))
    print("   total     ", out[:, :8].sum(1).mean() / nst, out[:, 8:].sum(1).mean() / nst, "status", s.lib.hp_sweep_status(s.handle))
