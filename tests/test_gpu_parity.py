"""Parity of the CUDA path (through the C ABI / the host mirror of the reference API) against the oracle and
against the golden vectors of the unmodified reference.  Runs on the B200 box: pytest -m gpu."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import helmholtz_oracle as orc  # noqa: E402
from tools import strip_model as sm  # noqa: E402

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))
IDS = [os.path.basename(p)[:-4] for p in CASES]


@pytest.fixture(scope="module")
def hp():
    import helmholtz_preconditioner_b200 as hp
    hp.load()
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    return hp


def _load(path):
    g = np.load(path, allow_pickle=False)
    n, b = int(g["n"]), int(g["b"])
    omega = 2 * np.pi * float(g["wave_num"]) + 1j * float(g["alpha"])
    h = 1 / (n + 1)
    return g, dict(b=b, const=float(g["const"]), eta=b * h, omega=omega, h=h, n=n)


def relerr(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return np.linalg.norm(a.ravel() - np.asarray(b).ravel()) / np.linalg.norm(np.asarray(b).ravel())


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).cuda()


@pytest.mark.parametrize("path", CASES, ids=IDS)
def test_assembly_golden(hp, path):
    """build_A_matrix: sparsity pattern bit exact, values to 1e-12 (code.py:202-219)."""
    g, p = _load(path)
    A = hp.build_A_matrix(c_mat=g["c_mat"], **p)
    indptr, indices, data = A.to_host()
    assert np.array_equal(indptr, g["A_indptr"])
    assert np.array_equal(indices, g["A_indices"])
    assert np.max(np.abs(data - g["A_data"]) / np.abs(g["A_data"])) < 1e-12
    y = A @ g["x_rand"]
    assert relerr(y, g["A_x_rand"]) < 1e-13
    y2 = A.solver.matvec(dev(g["x_rand"]))
    assert relerr(y2, g["A_x_rand"]) < 1e-13


@pytest.mark.parametrize("path", CASES, ids=IDS)
def test_strip_operator_csr_golden(hp, path):
    """get_Hm on the device (code.py:283-290): sparsity pattern bit exact, values to 1e-12, against the unmodified
    reference's first / middle / last layer; the moving-PML coefficients the factorisation is built from."""
    import scipy.sparse
    g, p = _load(path)
    for tag in ("first", "mid", "last"):
        m = int(g[f"Hm_{tag}_m"])
        H = hp.get_Hm(m, c_mat=g["c_mat"], **p)
        indptr, indices, data = H.to_host()
        assert np.array_equal(indptr, g[f"Hm_{tag}_indptr"])
        assert np.array_equal(indices, g[f"Hm_{tag}_indices"])
        assert np.max(np.abs(data - g[f"Hm_{tag}_data"]) / np.abs(g[f"Hm_{tag}_data"])) < 1e-12
    # the front block: get_Hm(b) is A[:bn, :bn]; the reference's H_F keeps its tridiagonal diagonal blocks
    b, n = p["b"], p["n"]
    A = scipy.sparse.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(n * n, n * n))[:b * n, :b * n].tocsr()
    A.sort_indices()
    ip, ix, d = hp.get_A_FF_block(c_mat=g["c_mat"], coupled=True, **p).to_host()
    assert np.array_equal(ip, A.indptr) and np.array_equal(ix, A.indices)
    assert np.max(np.abs(d - A.data) / np.abs(A.data)) < 1e-12
    HF = orc.get_A_FF_block(c_mat=g["c_mat"], **p).tocsr()
    HF.sort_indices()
    HF.eliminate_zeros()
    ip, ix, d = hp.get_A_FF_block(c_mat=g["c_mat"], **p).to_host()
    assert np.array_equal(ip, HF.indptr) and np.array_equal(ix, HF.indices)
    assert np.max(np.abs(d - HF.data) / np.abs(HF.data)) < 1e-12


@pytest.mark.parametrize("layout", ["classic", "cluster"])
@pytest.mark.parametrize("n,b,P,K", [(45, 12, 4, 2), (63, 12, 5, 3), (40, 5, 1, 3), (40, 5, 7, 1), (50, 20, 3, 4)])
def test_strip_generators_vs_model(hp, n, b, P, K, layout):
    """The packets written by the setup kernels against the numpy model of the same layout."""
    omega = 2 * np.pi * 5 + 2j
    c_mat = orc.init_c1_f1(omega, n)[0]
    s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat)
    s.setup_preconditioner(P=P, K=K, layout=layout)
    h = 1 / (n + 1)
    for m in (b + 1, (n + b) // 2, n):
        L, pk = s.strip_packets(m)
        assert (L["P"], L["K"]) == (P, K)
        mod = sm.StripModel(m, b, 60.0, b * h, omega, h, n, c_mat, P=P, K=K)
        pt = mod.part
        assert np.array_equal(L["leaf_start"], pt["leaf_start"]) and np.array_equal(L["q"], pt["q"])
        QP, CW, NS, NR = L["QP"], L["CW"], L["NS"], L["NR"]
        for l in range(P):
            q = int(pt["q"][l])
            for k in range(K):
                g_ = l * K + k
                lc0, lc1 = (q * k) // K, (q * (k + 1)) // K
                Wp = pk[g_, :CW * QP].reshape(CW, QP)
                Gp = pk[g_, CW * QP:CW * QP + 2 * b * CW].reshape(2 * b, CW)
                assert np.allclose(Wp[:lc1 - lc0, :q], mod.W[l, lc0:lc1, :q], rtol=0, atol=1e-11 * np.abs(mod.W).max())
                Gm = mod.G[l].reshape(2 * b, -1)[:, lc0:lc1]
                assert np.allclose(Gp[:, :lc1 - lc0], Gm, rtol=0, atol=1e-11 * max(np.abs(mod.G).max(), 1e-300))
        assert L["colN"] == (layout == "cluster")
        if NS:
            offN = CW * QP + 2 * b * CW
            N = np.zeros((NS, NS), complex)
            if L["colN"]:
                NRQ = L["NRQ"]      # CTA (j, k) holds N[NRQ*k : NRQ*(k+1), columns of separator j], column major
                for j in range(P - 1):
                    for k in range(K):
                        blk = pk[j * K + k, offN:offN + b * NRQ].reshape(b, NRQ)
                        r0, r1 = NRQ * k, min(NS, NRQ * (k + 1))
                        if r1 > r0:
                            N[r0:r1, j * b:(j + 1) * b] = blk[:, :r1 - r0].T
            else:
                for row in range(NS):
                    N[row] = pk[row // NR, offN + (row % NR) * NS: offN + (row % NR + 1) * NS]
            assert np.allclose(N, mod.N, rtol=0, atol=1e-11 * np.abs(mod.N).max())
    s.close()


def variants_of(s):
    """sweep kernel variants that can run on the layout the solver was set up with"""
    return (4,) if s.layout()["colN"] else (1, 2, 3)


@pytest.mark.parametrize("layout", ["classic", "auto"])
@pytest.mark.parametrize("n,b,P,K", [(63, 12, 5, 3), (40, 5, 1, 3), (40, 5, 7, 1), (200, 12, 0, 0), (130, 20, 6, 5), (300, 12, 6, 4),
                                     (45, 12, 4, 2)])
def test_strip_apply_vs_oracle(hp, n, b, P, K, layout):
    """y = T_m v against the oracle's splu solve (code.py:368-370)."""
    omega = 2 * np.pi * (n / 10) + 2j
    c_mat = orc.init_c1_f1(omega, n)[0]
    h = 1 / (n + 1)
    Pc = orc.SweepingPreconditioner(b, 60.0, b * h, omega, h, n, c_mat)
    s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(P=P, K=K, layout=layout)
    rng = np.random.default_rng(3)
    for variant in variants_of(s):
        s.set_sweep_variant(variant)
        for m in (b + 1, (n + b) // 2, n - 1, n):
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            y = s.strip_apply(m, dev(v))
            assert relerr(y, Pc.T(m, v)) < 1e-12
    assert s.sweep_status() == 0
    s.close()


GM = [p for p in CASES if "M_f" in np.load(p).files]
GM_IDS = [os.path.basename(p)[:-4] for p in GM]


@pytest.mark.parametrize("path", GM, ids=GM_IDS)
def test_preconditioner_golden(hp, path):
    """algo2_4 against the reference's own output (code.py:356-385)."""
    g, p = _load(path)
    s, _ = hp.algo2_3(c_mat=g["c_mat"], **p)
    n, b = p["n"], p["b"]
    u = hp.algo2_4(g["f_mat"].flatten(), b, n, s)
    assert relerr(u, g["M_f"]) < 1e-12
    u = hp.algo2_4(g["x_rand"], b, n, s)
    assert relerr(u, g["M_x_rand"]) < 1e-12
    m = int(g["Hm_mid_m"])
    assert relerr(s.strip_apply(m, dev(g["x_rand"][:n])), g["T_mid_v"]) < 1e-12
    # Engquist-Ying form of the diagonal solve, against the oracle
    Pp = orc.SweepingPreconditioner(c_mat=g["c_mat"], diag="paper", **p)
    assert relerr(hp.algo2_4(g["x_rand"], b, n, s, diag="paper"), Pp.apply(g["x_rand"])) < 1e-12
    s.close()


@pytest.mark.parametrize("path", GM, ids=GM_IDS)
def test_gmres_golden(hp, path):
    """run_solver: the literal reference (M ignores its argument) and the well-defined variant (M applied to
    the vector), iteration counts and residual histories against the reference's."""
    g, p = _load(path)
    args = (p["n"], p["b"], float(g["wave_num"]), p["const"], float(g["alpha"]), getattr(hp, str(g["init"])))
    r = hp.run_solver(*args, precond_input="vector", maxiter=25, verbose=False)
    assert r.niter == len(g["gmres_vector_hist"]) and r.info == int(g["gmres_vector_info"])
    assert np.allclose(r.residuals, g["gmres_vector_hist"], rtol=1e-8, atol=0)
    assert relerr(r.u, g["gmres_vector_u"]) < 1e-8
    r = hp.run_solver(*args, verbose=False)
    assert abs(r.niter - len(g["gmres_literal_hist"])) <= 2 and 1 <= r.niter <= 3
    assert r.info == int(g["gmres_literal_info"])
    assert r.residuals[-1] < 1e-12


def test_krylov_kernels(hp):
    from helmholtz_preconditioner_b200.gmres import DeviceVectors
    rng = np.random.default_rng(5)
    N, k = 100003, 7
    V = rng.standard_normal((k, N)) + 1j * rng.standard_normal((k, N))
    w = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    vec = DeviceVectors(N, torch.device("cuda:0"))
    Vd, wd = dev(V).reshape(k, N), dev(w)
    assert abs(vec.norm(wd) - np.linalg.norm(w)) < 1e-12 * np.linalg.norm(w)
    h, h1, h0 = vec.mgs(Vd, k, wd)
    wr = w.copy()
    hr = np.zeros(k, complex)
    for j in range(k):
        hr[j] = np.vdot(V[j], wr)
        wr -= hr[j] * V[j]
    assert np.allclose(h, hr, rtol=1e-12) and abs(h1 - np.linalg.norm(wr)) < 1e-11 * h1
    assert abs(h0 - np.linalg.norm(w)) < 1e-12 * h0
    assert relerr(wd, wr) < 1e-13
    # fused passes (w -= h_j v_j together with the next coefficient) against the separate dot / axpy launches: the
    # slices and the accumulation order are the same, so the results are bit-identical
    import os
    for kk in (0, 1, 2, k):
        res = []
        for unfused in (False, True):
            os.environ.pop("HP_MGS_UNFUSED", None)
            if unfused:
                os.environ["HP_MGS_UNFUSED"] = "1"
            w2 = dev(w)
            try:
                hh, a1, a0 = vec.mgs(Vd, kk, w2)
            finally:
                os.environ.pop("HP_MGS_UNFUSED", None)
            res.append((hh.copy(), a1, a0, w2.cpu().numpy()))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and res[0][2] == res[1][2]
        assert np.array_equal(res[0][3], res[1][3])
    x = dev(w)
    y = rng.standard_normal(k) + 1j * rng.standard_normal(k)
    vec.combine(Vd, y, x)
    assert relerr(x, w + y @ V) < 1e-13


@pytest.mark.parametrize("layout", ["classic", "auto"])
def test_medium_preconditioner_vs_oracle(hp, layout):
    """n = 255 (the reference's second problem size, code.py:579), automatic partition."""
    n, b, wn, const = 255, 12, 32, 62
    omega = 2 * np.pi * wn + 2j
    h = 1 / (n + 1)
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    Pc = orc.SweepingPreconditioner(b, const, b * h, omega, h, n, c_mat)
    s, _ = hp.algo2_3(b, const, b * h, omega, h, n, c_mat, layout=layout)
    f = f_mat.flatten().astype(np.complex128)
    ref = Pc.apply(f)
    Pc.diag = "paper"
    Pc_paper = Pc.apply(f)
    assert relerr(hp.algo2_4(f, b, n, s), ref) < 1e-12
    for variant in variants_of(s):
        s.set_sweep_variant(variant)
        assert relerr(hp.algo2_4(f, b, n, s), ref) < 1e-12
        assert relerr(hp.algo2_4(f, b, n, s, diag="paper"), Pc_paper) < 1e-12
    assert s.sweep_status() == 0
    s.close()


@pytest.mark.parametrize("nstrips", [1, 2, 3, 4, 5, 7])
def test_short_sweeps_cluster_vs_classic(hp, nstrips):
    """Sweeps over very few strips (ring prologues / epilogues of the pipelined kernels): cluster layout against
    the classic layout on the same strip range, forward and both backward variants."""
    n, b = 200, 12
    omega = 2 * np.pi * (n / 10) + 2j
    c_mat = orc.init_c1_f1(omega, n)[0]
    rng = np.random.default_rng(21)
    u0 = dev(rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n))
    m_lo = 60
    m_hi = m_lo + nstrips - 1
    res = {}
    for layout in ("cluster", "classic"):
        s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(m_lo=m_lo, m_hi=m_hi, layout=layout)
        u = u0.clone()
        s.sweep_forward(u, m_lo, m_hi)
        s.sweep_backward(u, m_hi, m_lo, "reference")
        s.sweep_backward(u, m_hi, m_lo, "paper")
        res[layout] = u.clone()
        assert s.sweep_status() == 0
        s.close()
    assert relerr(res["cluster"], res["classic"].cpu().numpy()) < 1e-12


def oracle_strip_T(m, v, b, const, omega, n, c_mat):
    """T_m v with one SuperLU factorisation (the oracle class factors every strip in its constructor)."""
    import scipy.sparse.linalg
    h = 1 / (n + 1)
    lu = scipy.sparse.linalg.splu(orc.get_Hm(m, b, const, b * h, omega, h, n, c_mat).tocsc())
    t = np.zeros(b * n, dtype=np.complex128)
    t[-n:] = v
    return lu.solve(t)[-n:]


SETUP_SWITCHES = {
    "shared-memory chains": ["HP_CHAIN_SMEM"],
    "unrolled pivot loop": ["HP_CHAIN_UNROLL"],
    "thread-per-chain separators / corners": ["HP_SETUP_THREAD"],
    "one inner leaf per warp in the corner kernel": ["HP_CORNER_WARP"],
    "rows of N through the row buffer": ["HP_SEP_ROWS_BUF"],
    "CTA-paced leaf kernel": ["HP_LEAF_CTA"],
    "CTA-paced leaf kernel, direct copies": ["HP_LEAF_CTA", "HP_LEAF_NOPIPE"],
    "first generation (one thread per chain)": ["HP_CHAIN_THREAD", "HP_SETUP_THREAD"],
}


@pytest.mark.parametrize("n", [300, 1000])
def test_setup_kernel_generations_agree(hp, n):
    """The b = 12 setup kernels (register-resident Schur chains, half-warp separator chains, warp-per-leaf corners,
    warp-paced leaf generators) against the earlier generations of the same recurrences, selected by the developer
    switches of csrc/hp_setup.cu, and against SuperLU (the oracle) on one strip."""
    import os
    b = 12
    omega = 2 * np.pi * (n / 10) + 2j
    h = 1 / (n + 1)
    c_mat, f_mat = orc.init_c1_f1(omega, n)
    f = dev(f_mat.flatten().astype(np.complex128))
    every = [k for sw in SETUP_SWITCHES.values() for k in sw]

    def run(switches):
        for k in every:
            os.environ.pop(k, None)
        for k in switches:
            os.environ[k] = "1"
        try:
            s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner()
            y = s.precond_apply(f).clone()
            rng = np.random.default_rng(5)
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            t = s.strip_apply(n // 2, dev(v)).clone()
            assert s.sweep_status() == 0
            s.close()
        finally:
            for k in every:
                os.environ.pop(k, None)
        return y, t, v

    y0, t0, v = run([])
    assert relerr(t0, oracle_strip_T(n // 2, v, b, 60.0, omega, n, c_mat)) < 1e-12
    for name, sw in SETUP_SWITCHES.items():
        y, t, _ = run(sw)
        assert relerr(y, y0.cpu().numpy()) < 1e-11, name
        assert relerr(t, t0.cpu().numpy()) < 1e-12, name


@pytest.mark.parametrize("n,P,K", [(2048, 8, 4), (1600, 10, 2)])
def test_wide_leaves_vs_oracle(hp, n, P, K):
    """Leaves wider than 128 columns (what an 8192^2 problem gets with at most 33 clusters): the b = 12 setup takes the
    256-thread leaf kernel there; a few strips against SuperLU, set up as a short strip range."""
    b = 12
    omega = 2 * np.pi * (n / 10) + 2j
    c_mat = orc.init_c1_f1(omega, n)[0]
    m_lo, m_hi = n // 2, n // 2 + 3
    s = hp.HelmholtzSolver(n, b, omega, 60.0, c_mat).setup_preconditioner(P=P, K=K, m_lo=m_lo, m_hi=m_hi, layout="classic")
    assert s.layout()["QP"] > 128
    rng = np.random.default_rng(11)
    for m in (m_lo, m_hi):
        v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert relerr(s.strip_apply(m, dev(v)), oracle_strip_T(m, v, b, 60.0, omega, n, c_mat)) < 1e-12
    assert s.sweep_status() == 0
    s.close()
