"""Developer check (GPU box): a wide problem (parts wider than 32 columns: one lane per column, several W chunks) on a
sub-range of strips, cluster layout against the classic layout and against a SuperLU solve of single strips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse.linalg as spla
import helmholtz_preconditioner_b200 as hp
from oracle import helmholtz_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 12
omega = 2 * np.pi * n / 10 + 2j
h = 1 / (n + 1)
c_mat, f_mat = hp.init_layered_f1(omega, n)
m_lo, m_hi = n // 2, n // 2 + 150
rng = np.random.default_rng(5)
u0 = torch.from_numpy(rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)).cuda()
res = {}
for layout in ("cluster", "classic"):
    s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat)
    s.setup_preconditioner(m_lo=m_lo, m_hi=m_hi, layout=layout)
    L = s.layout()
    print(layout, {k: int(L[k]) for k in ("P", "K", "G", "QP", "CW", "NS", "NRQ", "PK", "colN")}, "setup ms", round(s.setup_ms), flush=True)
    u = u0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.sweep_forward(u, m_lo, m_hi - 1); torch.cuda.synchronize()
    u = u0.clone()
    e0.record(); s.sweep_forward(u, m_lo, m_hi - 1); e1.record(); torch.cuda.synchronize()
    print(f"   forward {e0.elapsed_time(e1) / (m_hi - m_lo) * 1e3:.2f} us/strip", flush=True)
    s.sweep_backward(u, m_hi, m_lo, "paper")
    s.sweep_backward(u, m_hi, m_lo, "reference")
    res[layout] = u[(m_lo - 2) * n:(m_hi + 1) * n].clone()
    if layout == "cluster":
        for m in (m_lo, m_hi):
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            lu = spla.splu(orc.get_Hm(m, b, 100.0, b * h, omega, h, n, c_mat).tocsc())
            t = np.zeros(b * n, complex); t[-n:] = v
            ref = lu.solve(t)[-n:]
            y = s.strip_apply(m, torch.from_numpy(v).cuda()).cpu().numpy()
            print("   strip", m, "vs SuperLU", np.linalg.norm(y - ref) / np.linalg.norm(ref), flush=True)
    print("   status", s.sweep_status(), flush=True)
    s.close(); del s
d = (torch.linalg.norm(res["cluster"] - res["classic"]) / torch.linalg.norm(res["classic"])).item()
print("cluster vs classic after forward + 2 backward sweeps:", d)
assert d < 1e-10
