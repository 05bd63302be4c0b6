"""CPU oracle for the moving-PML sweeping-preconditioner Helmholtz path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy/scipy restatement of the
reference algorithm (/root/reference/code.py) and exists to *check* the CUDA
product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``helmholtz_preconditioner_b200/`` imports it, and the product path has no CPU
fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` hold outputs of the unmodified
reference (imported in the build container by ``tests/golden/make_golden.py``,
numba enabled, matplotlib stubbed, ``tol=`` mapped to scipy>=1.14's ``rtol=``)
and ``tests/test_oracle_golden.py`` checks every function here against them.

Each function cites the reference lines it restates.  The restatement is
vectorised (numpy) where the reference uses numba scalar loops; the arithmetic
per entry is the same expression in the same order, so assembled values agree
to the last bit or two.

The third-party pieces on the path are scipy's ``splu`` (SuperLU, used as a
black-box direct solve of the strip operators) and ``scipy.sparse.linalg.gmres``
(scipy 1.18.1 in this image: left-preconditioned restarted GMRES(20), modified
Gram-Schmidt, Givens rotations via LAPACK ``lartg``, inner tolerance control of
scipy gh-8400).  ``gmres_scipy_restated`` restates that published algorithm and
is pinned against ``scipy.sparse.linalg.gmres`` itself in the tests.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

# --------------------------------------------------------------------------
# PML stretching functions                      (reference code.py:11-37)
# --------------------------------------------------------------------------


def sigma1(x, const, eta):
    """code.py:11-18 -- quadratic PML profile on both ends of x1."""
    x = np.asarray(x, dtype=np.float64)
    lo = const / eta * ((x - eta) / eta) ** 2
    hi = const / eta * ((x - 1 + eta) / eta) ** 2
    return np.where(x <= eta, lo, np.where(x >= 1 - eta, hi, 0.0))


def sigma2(x, const, eta):
    """code.py:20-25 -- quadratic PML profile on the low end of x2 only."""
    x = np.asarray(x, dtype=np.float64)
    lo = const / eta * ((x - eta) / eta) ** 2
    return np.where(x <= eta, lo, 0.0)


def s1(x, const, eta, omega):
    """code.py:27-29."""
    return (1 + 1j * sigma1(x, const, eta) / omega) ** -1


def s2(x, const, eta, omega):
    """code.py:31-33."""
    return (1 + 1j * sigma2(x, const, eta) / omega) ** -1


def s2m(x, m, b, const, eta, omega, h):
    """code.py:35-37 -- the x2 PML moved so that it ends at grid row m."""
    return (1 + 1j * sigma2(np.asarray(x) - (m - b) * h, const, eta) / omega) ** -1


# --------------------------------------------------------------------------
# Velocity fields and sources                   (reference code.py:39-66, 390-408)
# --------------------------------------------------------------------------


def init_c1_mat(r1, r2, n):
    """code.py:40-44."""
    x_i = np.linspace(0, 1, n + 2)
    xx, yy = np.meshgrid(x_i, x_i)
    return 4 / 3 * (1 - .5 * np.exp(-32 * ((xx - r1) ** 2 + (yy - r2) ** 2)))


def init_c2_mat(n):
    """code.py:47-51."""
    x_i = np.linspace(0, 1, n + 2)
    xx, yy = np.meshgrid(x_i, x_i)
    return 4 / 3 * (1 - .5 * np.exp(-32 * ((xx - .5) ** 2)))


def init_f1_mat(r1, r2, omega, n):
    """code.py:54-58."""
    x_i = np.linspace(0, 1, n + 2)
    xx, yy = np.meshgrid(x_i[1:-1], x_i[1:-1])
    return np.exp(-(4 * omega / np.pi) ** 2 * ((xx - r1) ** 2 + (yy - r2) ** 2))


def init_f2_mat(r1, r2, d1, d2, omega, n):
    """code.py:61-66."""
    x_i = np.linspace(0, 1, n + 2)
    xx, yy = np.meshgrid(x_i[1:-1], x_i[1:-1])
    return np.exp(-4 * omega * ((xx - r1) ** 2 + (yy - r2) ** 2)) \
        * np.exp(1j * omega * (xx * d1 + yy * d2))


def init_c1_f1(omega, n, cr1=.5, cr2=.5, fr1=.5, fr2=.125):
    """code.py:390-393."""
    return init_c1_mat(cr1, cr2, n), init_f1_mat(fr1, fr2, omega, n)


def init_c1_f2(omega, n, cr1=.5, cr2=.5, fr1=.125, fr2=.125, d1=1 / 2 ** .5, d2=1 / 2 ** .5):
    """code.py:395-398."""
    return init_c1_mat(cr1, cr2, n), init_f2_mat(fr1, fr2, d1, d2, omega, n)


def init_c2_f1(omega, n, r1=.5, r2=.5):
    """code.py:400-403."""
    return init_c2_mat(n), init_f1_mat(r1, r2, omega, n)


def init_c2_f2(omega, n, r1=.5, r2=.5, d1=1 / 2 ** .5, d2=1 / 2 ** .5):
    """code.py:405-408."""
    return init_c2_mat(n), init_f2_mat(r1, r2, d1, d2, omega, n)


# --------------------------------------------------------------------------
# Stencil coefficients
# --------------------------------------------------------------------------


def stencil_coeffs(rows, m_shift, b, const, eta, omega, h, n, c_mat):
    """Five stencil coefficients for grid rows ``rows`` (1-based x2 indices j).

    Restates the loop bodies of get_A_diag_block_coeffs (code.py:82-113, with
    ``m_shift=None`` -> s2) and get_Hm_coeffs (code.py:237-275, with
    ``m_shift=m`` -> s2m).  Returns c1..c5 with shape (len(rows), n); entry
    [r, i-1] belongs to grid point (i, rows[r]).  Note the reference reads the
    velocity as c_mat[i-1, j-1] from an (n+2, n+2) array (code.py:108, 270);
    that indexing is kept.
    """
    rows = np.asarray(rows, dtype=np.int64)
    i = np.arange(1, n + 1, dtype=np.float64)
    j = rows.astype(np.float64)[:, None]
    if m_shift is None:
        S2 = lambda x: s2(x, const, eta, omega)  # noqa: E731
    else:
        S2 = lambda x: s2m(x, m_shift, b, const, eta, omega, h)  # noqa: E731
    s1_lo = s1((i - .5) * h, const, eta, omega)[None, :]
    s1_hi = s1((i + .5) * h, const, eta, omega)[None, :]
    s1_c = s1(i * h, const, eta, omega)[None, :]
    s2_lo = S2((j - .5) * h)
    s2_hi = S2((j + .5) * h)
    s2_c = S2(j * h)
    c1 = 1 / h ** 2 * (s1_lo / s2_c)
    c2 = 1 / h ** 2 * (s1_hi / s2_c)
    c3 = 1 / h ** 2 * (s2_lo / s1_c)
    c4 = 1 / h ** 2 * (s2_hi / s1_c)
    ii = np.arange(0, n)[None, :]
    jj = (rows - 1)[:, None]
    cv = np.asarray(c_mat)[ii, jj]
    c5 = omega ** 2 / (s1_c * s2_c * cv ** 2) - (c1 + c2 + c3 + c4)
    return c1, c2, c3, c4, c5


def build_A_matrix(b, const, eta, omega, h, n, c_mat):
    """code.py:202-219 -- the n^2 x n^2 operator, as sorted CSR.

    Row (j-1)*n + (i-1) holds, in column order, the couplings to (i, j-1) [c3],
    (i-1, j) [c1], itself [c5], (i+1, j) [c2] and (i, j+1) [c4], with the
    out-of-grid neighbours dropped (homogeneous Dirichlet truncation).
    """
    c1, c2, c3, c4, c5 = stencil_coeffs(np.arange(1, n + 1), None, b, const, eta, omega, h, n, c_mat)
    N = n * n
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    row = (jj * n + ii).ravel()
    cols = np.stack([row - n, row - 1, row, row + 1, row + n], axis=1)
    vals = np.stack([c3.ravel(), c1.ravel(), c5.ravel(), c2.ravel(), c4.ravel()], axis=1)
    keep = np.stack([(jj > 0).ravel(), (ii > 0).ravel(), np.ones(N, bool),
                     (ii < n - 1).ravel(), (jj < n - 1).ravel()], axis=1)
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(keep.sum(axis=1), out=indptr[1:])
    A = scipy.sparse.csr_matrix((vals[keep], cols[keep].astype(np.int32), indptr), shape=(N, N))
    return A


def stencil_matvec(x, b, const, eta, omega, h, n, c_mat):
    """Matrix-free A @ x (same operator as build_A_matrix)."""
    c1, c2, c3, c4, c5 = stencil_coeffs(np.arange(1, n + 1), None, b, const, eta, omega, h, n, c_mat)
    u = np.asarray(x, dtype=np.complex128).reshape(n, n)
    y = c5 * u
    y[:, 1:] += c1[:, 1:] * u[:, :-1]
    y[:, :-1] += c2[:, :-1] * u[:, 1:]
    y[1:, :] += c3[1:, :] * u[:-1, :]
    y[:-1, :] += c4[:-1, :] * u[1:, :]
    return y.ravel()


def get_Hm(m, b, const, eta, omega, h, n, c_mat):
    """code.py:224-290 -- strip operator of grid rows m-b+1..m with moved PML."""
    rows = np.arange(m - b + 1, m + 1)
    c1, c2, c3, c4, c5 = stencil_coeffs(rows, m, b, const, eta, omega, h, n, c_mat)
    c1v = c1.ravel()[1:].copy()
    c2v = c2.ravel()[:-1].copy()
    c3v = c3.ravel()[n:]
    c4v = c4.ravel()[:-n]
    c1v[n - 1::n] = 0
    c2v[n - 1::n] = 0
    return scipy.sparse.diags(c5.ravel()) + scipy.sparse.diags(c1v, -1) + scipy.sparse.diags(c2v, 1) \
        + scipy.sparse.diags(c3v, -n) + scipy.sparse.diags(c4v, n)


def get_A_FF_block(b, const, eta, omega, h, n, c_mat, coupled=False):
    """code.py:178-183 -- the reference keeps only the b diagonal (tridiagonal)
    blocks A_11..A_bb.  ``coupled=True`` gives the full A[:bn, :bn] of the paper."""
    c1, c2, c3, c4, c5 = stencil_coeffs(np.arange(1, b + 1), None, b, const, eta, omega, h, n, c_mat)
    c1v = c1.ravel()[1:].copy()
    c2v = c2.ravel()[:-1].copy()
    c1v[n - 1::n] = 0
    c2v[n - 1::n] = 0
    HF = scipy.sparse.diags(c5.ravel()) + scipy.sparse.diags(c1v, -1) + scipy.sparse.diags(c2v, 1)
    if coupled:
        HF = HF + scipy.sparse.diags(c3.ravel()[n:], -n) + scipy.sparse.diags(c4.ravel()[:-n], n)
    return HF


class SweepingPreconditioner:
    """algo2_3 + algo2_4 (code.py:345-385).

    ``diag='reference'`` reproduces code.py:372-375 literally
    (u_m <- u_m - T_m u_m); ``diag='paper'`` is Engquist & Ying's Algorithm 2.4
    (u_m <- T_m u_m).  ``front='blockdiag'`` is the reference's H_F
    (code.py:178-183); ``front='coupled'`` is the full A_FF.
    """

    def __init__(self, b, const, eta, omega, h, n, c_mat, diag="reference", front="blockdiag"):
        assert diag in ("reference", "paper") and front in ("blockdiag", "coupled")
        self.b, self.n, self.diag, self.front = b, n, diag, front
        # algo2_3, code.py:345-353
        HF = get_A_FF_block(b, const, eta, omega, h, n, c_mat, coupled=(front == "coupled")).tocsc()
        self.lu_HF = scipy.sparse.linalg.splu(HF)
        self.lu_Hm = [scipy.sparse.linalg.splu(get_Hm(m, b, const, eta, omega, h, n, c_mat).tocsc())
                      for m in range(b + 1, n + 1)]
        _, _, c3, c4, _ = stencil_coeffs(np.arange(1, n + 1), None, b, const, eta, omega, h, n, c_mat)
        self.lo = c3  # lo[j-1] = diagonal of A_{j, j-1}   (code.py:145-154)
        self.up = c4  # up[j-1] = diagonal of A_{j, j+1}   (code.py:131-140)

    def T(self, m, v):
        """Last-row restriction of Hm^{-1} (code.py:368-370, 373-375, 378-380)."""
        b, n = self.b, self.n
        t = np.zeros(b * n, dtype=np.complex128)
        t[-n:] = v
        return self.lu_Hm[m - b - 1].solve(t)[-n:]

    def apply(self, f_vec):
        """algo2_4, code.py:356-385."""
        b, n = self.b, self.n
        u = np.array(f_vec, dtype=np.complex128).reshape(n, n).copy()
        TFuF = self.lu_HF.solve(u[:b].ravel())
        u[b] = u[b] - self.lo[b] * TFuF[-n:]                      # code.py:365
        for m in range(b + 1, n):                                  # code.py:366-370
            u[m] = u[m] - self.lo[m] * self.T(m, u[m - 1])
        for m in range(b + 1, n + 1):                              # code.py:372-375
            t = self.T(m, u[m - 1])
            u[m - 1] = (u[m - 1] - t) if self.diag == "reference" else t
        for m in range(n - 1, b, -1):                              # code.py:376-380
            u[m - 1] = u[m - 1] - self.T(m, self.up[m - 1] * u[m])
        Au = np.zeros(b * n, dtype=np.complex128)                  # code.py:381-382
        Au[-n:] = self.up[b - 1] * u[b]
        uF = TFuF - self.lu_HF.solve(Au)
        u[:b] = uF.reshape(b, n)
        return u.ravel()


# --------------------------------------------------------------------------
# scipy 1.18.1 gmres, restated
# --------------------------------------------------------------------------


def _lartg(f, g):
    """LAPACK zlartg semantics: c real, s complex, [c s; -conj(s) c] [f; g] = [r; 0]."""
    from scipy.linalg import get_lapack_funcs
    lartg = get_lapack_funcs("lartg", dtype=np.complex128)
    return lartg(f, g)


def gmres_scipy_restated(matvec, psolve, b, rtol=1e-5, atol=0.0, restart=20, maxiter=None):
    """scipy/sparse/linalg/_isolve/iterative.py::gmres (scipy 1.18.1), legacy
    callback semantics as triggered by code.py:516 (``callback=counter``).

    Returns (x, info, hist) with hist the values passed to the callback
    (preconditioned residual / ||b||, one per inner iteration).
    """
    b = np.asarray(b, dtype=np.complex128)
    n = len(b)
    x = np.zeros(n, dtype=np.complex128)
    bnrm2 = np.linalg.norm(b)
    atol = max(float(atol), float(rtol) * float(bnrm2))
    hist = []
    if bnrm2 == 0:
        return b, 0, hist
    eps = np.finfo(np.complex128).eps
    if maxiter is None:
        maxiter = n * 10
    restart = min(restart, n)
    Mb_nrm2 = np.linalg.norm(psolve(b))
    ptol_max_factor = 1.
    ptol = Mb_nrm2 * min(ptol_max_factor, atol / bnrm2)
    presid = 0.
    v = np.empty([restart + 1, n], dtype=np.complex128)
    hh = np.zeros([restart, restart + 1], dtype=np.complex128)
    givens = np.zeros([restart, 2], dtype=np.complex128)
    inner_iter = 0
    rnorm = np.inf
    for iteration in range(maxiter):
        if iteration == 0:
            r = b.copy()
            if np.linalg.norm(r) < atol:
                return x, 0, hist
        v[0, :] = psolve(r)
        tmp = np.linalg.norm(v[0, :])
        v[0, :] *= (1 / tmp)
        S = np.zeros(restart + 1, dtype=np.complex128)
        S[0] = tmp
        breakdown = False
        for col in range(restart):
            av = matvec(v[col, :])
            w = np.array(psolve(av), dtype=np.complex128).reshape(n)
            h0 = np.linalg.norm(w)
            for k in range(col + 1):
                tmp = np.vdot(v[k, :], w)
                hh[col, k] = tmp
                w -= tmp * v[k, :]
            h1 = np.linalg.norm(w)
            hh[col, col + 1] = h1
            v[col + 1, :] = w[:]
            if h1 <= eps * h0:
                hh[col, col + 1] = 0
                breakdown = True
            else:
                v[col + 1, :] *= (1 / h1)
            for k in range(col):
                c, s = givens[k, 0], givens[k, 1]
                n0, n1 = hh[col, [k, k + 1]]
                hh[col, [k, k + 1]] = [c * n0 + s * n1, -s.conj() * n0 + c * n1]
            c, s, mag = _lartg(hh[col, col], hh[col, col + 1])
            givens[col, :] = [c, s]
            hh[col, [col, col + 1]] = mag, 0
            tmp = -np.conjugate(s) * S[col]
            S[[col, col + 1]] = [c * S[col], tmp]
            presid = np.abs(tmp)
            inner_iter += 1
            hist.append(presid / bnrm2)
            if inner_iter == maxiter:
                break
            if presid <= ptol or breakdown:
                break
        if hh[col, col] == 0:
            S[col] = 0
        y = np.zeros([col + 1], dtype=np.complex128)
        y[:] = S[:col + 1]
        for k in range(col, 0, -1):
            if y[k] != 0:
                y[k] /= hh[k, k]
                tmp = y[k]
                y[:k] -= tmp * hh[k, :k]
        if y[0] != 0:
            y[0] /= hh[0, 0]
        x += y @ v[:col + 1, :]
        r = b - matvec(x)
        rnorm = np.linalg.norm(r)
        if inner_iter == maxiter:
            return x, (0 if rnorm <= atol else maxiter), hist
        if rnorm <= atol:
            break
        elif breakdown:
            break
        elif presid <= ptol:
            ptol_max_factor = max(eps, 0.25 * ptol_max_factor)
        else:
            ptol_max_factor = min(1.0, 1.5 * ptol_max_factor)
        ptol = presid * min(ptol_max_factor, atol / rnorm)
    info = 0 if (rnorm <= atol) else maxiter
    return x, info, hist


def run_solver(n, b, wave_num, const, alpha, init_func=init_c1_f1, *, c_mat=None, f_mat=None,
               diag="reference", front="blockdiag", precond_input="rhs", rtol=1e-3,
               maxiter=None, restart=20):
    """code.py:424-541 without the plotting: returns (u, hist, niter, info).

    ``precond_input='rhs'`` is the reference: its LinearOperator ignores the
    vector it is handed and always preconditions f_vec (code.py:510-511).
    ``precond_input='vector'`` applies M to the argument.
    """
    omega = 2 * np.pi * wave_num + 1j * alpha
    h = 1 / (n + 1)
    eta = b * h
    if c_mat is None or f_mat is None:
        c0, f0 = init_func(omega, n)
        c_mat = c0 if c_mat is None else c_mat
        f_mat = f0 if f_mat is None else f_mat
    f_vec = np.asarray(f_mat).flatten().astype(np.complex128)
    A = build_A_matrix(b, const, eta, omega, h, n, c_mat)
    P = SweepingPreconditioner(b, const, eta, omega, h, n, c_mat, diag=diag, front=front)
    if precond_input == "rhs":
        psolve = lambda x: P.apply(f_vec)  # noqa: E731
    else:
        psolve = P.apply
    u, info, hist = gmres_scipy_restated(lambda x: A @ x, psolve, f_vec, rtol=rtol,
                                         restart=restart, maxiter=maxiter)
    return u, hist, len(hist), info
