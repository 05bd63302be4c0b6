"""Full-size checks on the B200 (pytest -m gpu): the BASELINE.json sizes, through properties that do not need
the oracle at full size, plus oracle spot checks of single strips (a SuperLU solve of one 49k x 49k strip takes
a fraction of a second)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import helmholtz_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def hp():
    import helmholtz_preconditioner_b200 as hp
    hp.load()
    assert torch.cuda.is_available()
    return hp


def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel())


def refined_solve(H, lu, t, rounds=2):
    """SuperLU solve of H x = t followed by iterative refinement with the residual formed in extended precision
    (numpy longdouble): the reference value for the 1e-12 comparisons at n = 4096 / 6000, where a plain SuperLU solve of a
    49k-72k strip is itself only good to a few 1e-12 (condition of the strip operator near resonance)."""
    coo = H.tocoo()
    dat = coo.data.astype(np.clongdouble)
    x = lu.solve(t)
    for _ in range(rounds):
        Hx = np.zeros(H.shape[0], dtype=np.clongdouble)
        np.add.at(Hx, coo.row, dat * x.astype(np.clongdouble)[coo.col])
        r = (t.astype(np.clongdouble) - Hx).astype(np.complex128)
        x = x + lu.solve(r)
    return x


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal(n) + 1j * rng.standard_normal(n)).cuda()


def test_strips_4096_vs_oracle(hp):
    """BASELINE config 'layered 4096^2': first, middle and last strips against a SuperLU solve of the strip."""
    n, b, const = 4096, 12, 100.0
    omega = 2 * np.pi * n / 10 + 2j
    h = 1 / (n + 1)
    c_mat, _ = hp.init_layered_f1(omega, n)
    s = hp.HelmholtzSolver(n, b, omega, const, c_mat)
    import scipy.sparse.linalg as spla
    for m_lo in (b + 1, n // 2, n - 1):
        m_hi = min(n, m_lo + 1)
        refs = {}
        for m in (m_lo, m_hi):
            H = orc.get_Hm(m, b, const, b * h, omega, h, n, c_mat).tocsc()
            lu = spla.splu(H)
            t = np.zeros(b * n, complex)
            t[-n:] = rnd(n, m).cpu().numpy()
            refs[m] = refined_solve(H, lu, t)[-n:]
        for layout in ("classic", "auto"):
            s.setup_preconditioner(m_lo=m_lo, m_hi=m_hi, layout=layout)
            for m in (m_lo, m_hi):
                for variant in variants_of(s):
                    s.set_sweep_variant(variant)
                    assert rel(s.strip_apply(m, rnd(n, m)), refs[m]) < 1e-12
            s.set_sweep_variant(0)
    assert s.sweep_status() == 0
    s.close()


def variants_of(s):
    return (4,) if s.layout()["colN"] else (1, 2, 3)


@pytest.fixture(scope="module")
def solver1024c(hp):
    """the same problem as solver1024 with the classic generator layout (sweep kernel variants 1-3)"""
    n, b = 1024, 20
    omega = 2 * np.pi * n / 10 + 2j
    c_mat, f_mat = hp.init_const_f1(omega, n)
    s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner(layout="classic")
    yield s, f_mat
    s.close()


@pytest.fixture(scope="module")
def solver1024(hp):
    """BASELINE config 'constant velocity 1024^2, ~10 points per wavelength, PML width 20'."""
    n, b = 1024, 20
    omega = 2 * np.pi * n / 10 + 2j
    c_mat, f_mat = hp.init_const_f1(omega, n)
    s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner()
    yield s, f_mat
    s.close()


def test_strip_operator_properties_1024(solver1024, solver1024c):
    s, _ = solver1024c
    s4, _ = solver1024
    n = s.n
    v1, v2 = rnd(n, 1), rnd(n, 2)
    for m in (s.b + 1, 500, n):
        y1, y2 = s.strip_apply(m, v1), s.strip_apply(m, v2)
        # T_m is complex symmetric (H_m is): v2^T T v1 = v1^T T v2
        a, c = torch.sum(v2 * y1), torch.sum(v1 * y2)
        assert abs(a - c) / abs(a) < 1e-10
        # linearity
        y3 = s.strip_apply(m, (2 - 1j) * v1 + v2)
        assert rel(y3, (2 - 1j) * y1 + y2) < 1e-12
        # the block-synchronous kernels (direct, TMA-staged) run the same arithmetic; the pipelined one regroups it
        s.set_sweep_variant(1)
        yd = s.strip_apply(m, v1)
        s.set_sweep_variant(2)
        assert torch.equal(s.strip_apply(m, v1), yd)
        s.set_sweep_variant(3)
        assert rel(s.strip_apply(m, v1), yd) < 1e-13
        s.set_sweep_variant(0)
        assert rel(s4.strip_apply(m, v1), yd) < 1e-12          # cluster layout: another partition of the strip
    assert s.sweep_status() == 0 and s4.sweep_status() == 0


def test_preconditioner_properties_1024(solver1024, solver1024c):
    s, f_mat = solver1024c
    s4, _ = solver1024
    N = s.n ** 2
    x, y = rnd(N, 3), rnd(N, 4)
    Mx, My = s.precond_apply(x), s.precond_apply(y)
    assert rel(s.precond_apply((0.5 + 2j) * x - y), (0.5 + 2j) * Mx - My) < 1e-11
    # idempotent call: same input, same bits; the three kernel variants agree
    assert torch.equal(s.precond_apply(x), Mx)
    s.set_sweep_variant(1)
    Md = s.precond_apply(x)
    s.set_sweep_variant(2)
    assert torch.equal(s.precond_apply(x), Md)
    s.set_sweep_variant(3)
    assert rel(s.precond_apply(x), Md) < 1e-12
    assert rel(Mx, Md) < 1e-12
    s.set_sweep_variant(0)
    M4 = s4.precond_apply(x)
    assert rel(M4, Md) < 1e-11 and torch.equal(s4.precond_apply(x), M4)
    assert rel(s4.precond_apply(x, diag="paper"), s.precond_apply(x, diag="paper")) < 1e-11
    assert rel(s4.precond_apply((0.5 + 2j) * x - y), (0.5 + 2j) * M4 - s4.precond_apply(y)) < 1e-11
    for d in ("reference", "paper"):
        assert torch.isfinite(torch.view_as_real(s.precond_apply(x, diag=d))).all()
    assert s.sweep_status() == 0 and s4.sweep_status() == 0


def test_matvec_matches_assembled_csr_1024(solver1024):
    s, _ = solver1024
    A = s.assemble_csr()
    x = rnd(s.n ** 2, 5)
    assert rel(s.matvec(x), A @ x) < 1e-13
    indptr = A.indptr.cpu().numpy()
    assert indptr[0] == 0 and indptr[-1] == 5 * s.n ** 2 - 4 * s.n
    assert np.all(np.diff(indptr) >= 3) and np.all(np.diff(indptr) <= 5)
    idx = A.indices.cpu().numpy()
    assert np.all(np.diff(idx)[np.setdiff1d(np.arange(len(idx) - 1), indptr[1:-1] - 1)] > 0)   # sorted inside rows


@pytest.mark.parametrize("diag", ["reference", "paper"])
def test_run_solver_127_vs_oracle(hp, diag):
    """The reference's first production case, run_solver(127, 12, 16, 81, 2, init_c1_f1) (code.py:574), with
    the preconditioner applied to the Krylov vector: iteration count equal, residual history and field to 1e-8."""
    args = (127, 12, 16, 81, 2)
    r = hp.run_solver(*args, hp.init_c1_f1, precond_input="vector", diag=diag, maxiter=30, verbose=False)
    u0, hist0, niter0, info0 = orc.run_solver(*args, orc.init_c1_f1, precond_input="vector", diag=diag, maxiter=30)
    assert abs(r.niter - niter0) <= 1 and r.info == info0
    k = min(r.niter, niter0)
    assert np.allclose(r.residuals[:k], hist0[:k], rtol=1e-7)
    if r.niter == niter0:
        assert rel(r.u, u0) < 1e-8


@pytest.mark.parametrize("n,b", [(6000, 12), (3000, 20)])
def test_wide_parts_cluster_vs_classic(hp, n, b):
    """Parts wider than 32 columns (one lane per column, several W chunks per strip) and the generic (b, K)
    instantiation of the cluster kernel: a sub-range of strips against SuperLU and against the classic layout."""
    import scipy.sparse.linalg as spla
    omega = 2 * np.pi * n / 10 + 2j
    h = 1 / (n + 1)
    c_mat, _ = hp.init_layered_f1(omega, n)
    m_lo, m_hi = n // 2, n // 2 + 60
    u0 = rnd(n * n, 9)
    res = {}
    for layout in ("cluster", "classic"):
        s = hp.HelmholtzSolver(n, b, omega, 100.0, c_mat).setup_preconditioner(m_lo=m_lo, m_hi=m_hi, layout=layout)
        assert s.layout()["colN"] == (layout == "cluster")
        u = u0.clone()
        s.sweep_forward(u, m_lo, m_hi - 1)
        s.sweep_backward(u, m_hi, m_lo, "paper")
        s.sweep_backward(u, m_hi, m_lo, "reference")
        res[layout] = u[(m_lo - 2) * n:(m_hi + 1) * n].clone()
        if layout == "cluster":
            v = rnd(n, 10)
            H = orc.get_Hm(m_hi, b, 100.0, b * h, omega, h, n, c_mat).tocsc()
            lu = spla.splu(H)
            t = np.zeros(b * n, complex)
            t[-n:] = v.cpu().numpy()
            assert rel(s.strip_apply(m_hi, v), refined_solve(H, lu, t)[-n:]) < 1e-12
        assert s.sweep_status() == 0
        s.close()
    assert rel(res["cluster"], res["classic"]) < 1e-10
