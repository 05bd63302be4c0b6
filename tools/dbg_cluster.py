"""Developer check of the cluster sweep kernel on the GPU box: small strips vs the oracle, then timing at 4096^2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import helmholtz_preconditioner_b200 as hp
from oracle import helmholtz_oracle as orc

def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))

stage = sys.argv[1] if len(sys.argv) > 1 else "small"
if stage == "small":
    for (n, b, P, K) in [(45, 12, 4, 2), (40, 5, 7, 1), (40, 5, 1, 3), (63, 12, 5, 3), (300, 12, 6, 4), (200, 12, 0, 0)]:
        omega = 2 * np.pi * (n / 10) + 2j
        c_mat, f_mat = orc.init_c1_f1(omega, n)
        h = 1 / (n + 1)
        Pc = orc.SweepingPreconditioner(b, 60.0, b * h, omega, h, n, c_mat)
        s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(P=P, K=K, layout="cluster")
        L = s.layout()
        print((n, b, P, K), {k: int(L[k]) for k in ("P", "K", "G", "QP", "CW", "NS", "NCB", "NRQ", "NXG", "PK", "colN")}, flush=True)
        rng = np.random.default_rng(3)
        for m in (b + 1, (n + b) // 2, n):
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            y = s.strip_apply(m, torch.from_numpy(v).cuda())
            print("   strip", m, "err", rel(y, Pc.T(m, v)), "status", s.sweep_status(), flush=True)
        f = f_mat.flatten().astype(np.complex128)
        for d in ("reference", "paper"):
            Pc.diag = d
            u = s.precond_apply(torch.from_numpy(f).cuda(), diag=d)
            print("   M f", d, "err", rel(u, Pc.apply(f)), "status", s.sweep_status(), flush=True)
        s.close()
else:
    n, b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 12
    omega = 2 * np.pi * n / 10 + 2j
    c_mat, f_mat = hp.init_layered_f1(omega, n)
    x = torch.from_numpy(f_mat.ravel().astype(np.complex128)).cuda()
    res = {}
    for layout in (("cluster",) if len(sys.argv) > 3 else ("cluster", "classic")):
        s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
        s.setup_preconditioner(layout=layout)
        L = s.layout()
        print(layout, {k: int(L[k]) for k in ("P", "K", "G", "QP", "CW", "NS", "NCB", "NRQ", "NXG", "PK", "colN")}, "setup ms", s.setup_ms,
              "GB", s.precond_bytes / 1e9, flush=True)
        u = x.clone()
        s.sweep_forward(u, b + 1, n - 1)
        torch.cuda.synchronize()
        print("   status", s.sweep_status(), flush=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u = x.clone()
        e0.record(); s.sweep_forward(u, b + 1, n - 1); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"   forward sweep {ms:.2f} ms, {1e3 * ms / (n - 1 - b):.2f} us/strip", flush=True)
        res[layout] = s.precond_apply(x).clone()
        e0.record(); s.precond_apply(x); e1.record(); torch.cuda.synchronize()
        print(f"   precond apply {e0.elapsed_time(e1):.2f} ms  status {s.sweep_status()}", flush=True)
        if layout == "cluster":
            s.lib.hp_debug_phases(s.handle, 1, None)
            u = x.clone()
            s.sweep_forward(u, b + 1, n - 1); torch.cuda.synchronize()
            raw = np.zeros(L["G"] * (16 + 1024), dtype=np.int64)
            s.lib.hp_debug_phases(s.handle, 0, raw.ctypes.data)
            out = raw[:L["G"] * 16].reshape(L["G"], 16)
            os.makedirs("gpurun_out", exist_ok=True)
            tag = os.environ.get("HP_TAG", "")
            np.save(f"gpurun_out/timeline4{tag}.npy", raw[L["G"] * 16:L["G"] * (16 + 1024)].reshape(L["G"], 64, 16))
            np.save(f"gpurun_out/phases4{tag}.npy", out)
            nst = n - 1 - b
            names = ["pre (GL,GF,R)", "A wait x3", "B rho+bar", "C rows", "D poll", "D sum+send", "-", "-",
                     "a G wait", "a gb", "b wait x3", "b corr+send", "c wait V", "c W", "c tail", "(W chunk waits)"]
            for i, nm in enumerate(names):
                print(f"   {nm:14s} {out[:, i].mean() / nst:9.0f} {out[:, i].min() / nst:9.0f} {out[:, i].max() / nst:9.0f}")
        s.close()
        del s
    if "classic" in res:
        print("cluster vs classic |M f| rel diff", rel(res["cluster"], res["classic"].cpu().numpy()))
