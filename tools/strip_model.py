"""Host model (numpy) of the GPU strip-solve data structures: leaves + separators, one level.

Development/test helper: it mirrors, array for array, what csrc/hp_setup.cu produces and what
csrc/hp_sweep.cu consumes, so that the CUDA stages can be checked one by one.  It is not part of the
product path and not the oracle (the oracle follows the reference's splu formulation).

The strip operator H_m (reference get_Hm, code.py:283-290) is block tridiagonal when the unknowns are
ordered x1-major: block row i (x1 index) holds the b unknowns of grid column i, the diagonal block D_i is
the b x b tridiagonal x2 coupling and the off-diagonal blocks L_i (to i-1), U_i (to i+1) are diagonal, with
L_{i+1} = U_i.  T_m v = (H_m^{-1} [0; v]) restricted to the last strip row (code.py:368-370).

The x1 axis is cut into P leaves separated by P-1 single separator columns:

    leaf 0 | s_0 | leaf 1 | s_1 | ... | s_{P-2} | leaf P-1

  leaf l  : Dirichlet-truncated leaf inverse G_l = (H_m restricted to the leaf)^{-1}, sampled as
            W_l [q,q]  = G_l[(c,b),(c',b)]                      (source and receiver in the last strip row)
            Gf_l[b,q]  = cpl(left sep)  * G_l[(first,k),(c,b)]  (response in the first leaf column)
            Gl_l[b,q]  = cpl(right sep) * G_l[(last,k),(c,b)]   (response in the last leaf column)
            where cpl(s)[k] = U_s[k] = L_{s+1}[k] is the x1 coupling across the cut.
  seps    : the Schur complement S of the separator unknowns (block tridiagonal, b x b blocks) is inverted
            to the dense N = S^{-1} [(P-1)b, (P-1)b].

  y = T_m v:   g_l   = [Gf_l; Gl_l] v_l                                  (leaf phase, 2b numbers per leaf)
               rho_s = e_b v_s - Gl_{l}(s) v - Gf_{l+1}(s) v              (l = leaf left of s)
               x_S   = N rho                                             (separator phase)
               y_l   = W_l v_l - Gf_l^T x_{s_{l-1}} - Gl_l^T x_{s_l},    y_s = x_s[b-1]
"""
import numpy as np

from oracle import helmholtz_oracle as orc


def partition(n, P, K):
    """Leaves/separators/parts.  Returns dict with
    leaf_start[P+1] (leaf l covers columns leaf_start[l] .. leaf_start[l]+q[l]-1, 0-based), q[P],
    sep[P-1] (separator columns), part_start[P][K+1] (absolute first column of each part)."""
    assert P >= 1 and K >= 1 and n - (P - 1) >= P
    inner = n - (P - 1)
    q = np.array([(inner * (l + 1)) // P - (inner * l) // P for l in range(P)], dtype=np.int64)
    leaf_start = np.zeros(P, dtype=np.int64)
    sep = np.zeros(max(P - 1, 0), dtype=np.int64)
    pos = 0
    for l in range(P):
        leaf_start[l] = pos
        pos += q[l]
        if l < P - 1:
            sep[l] = pos
            pos += 1
    assert pos == n
    part_start = np.zeros((P, K + 1), dtype=np.int64)
    for l in range(P):
        for k in range(K + 1):
            part_start[l, k] = leaf_start[l] + (q[l] * k) // K
    return dict(P=P, K=K, q=q, leaf_start=leaf_start, sep=sep, part_start=part_start,
                QP=int(q.max()), CW=int(max(-(-int(x) // K) for x in q)))


def strip_blocks(m, b, const, eta, omega, h, n, c_mat):
    """D (n,b,b), L (n,b), U (n,b) of the x1-major block tridiagonal form of H_m."""
    rows = np.arange(m - b + 1, m + 1)
    c1, c2, c3, c4, c5 = orc.stencil_coeffs(rows, m, b, const, eta, omega, h, n, c_mat)  # (b, n)
    D = np.zeros((n, b, b), dtype=np.complex128)
    k = np.arange(b)
    D[:, k, k] = c5.T
    D[:, k[1:], k[:-1]] = c3.T[:, 1:]
    D[:, k[:-1], k[1:]] = c4.T[:, :-1]
    return D, c1.T.copy(), c2.T.copy()


def leaf_inverse(D, L, U, i0, i1):
    """Dense inverse of the leaf operator on block rows i0..i1-1, shape (q,b,q,b)."""
    q, b = i1 - i0, D.shape[1]
    H = np.zeros((q, b, q, b), complex)
    for r in range(q):
        H[r, :, r, :] = D[i0 + r]
        if r > 0:
            H[r, np.arange(b), r - 1, np.arange(b)] = L[i0 + r]
        if r < q - 1:
            H[r, np.arange(b), r + 1, np.arange(b)] = U[i0 + r]
    return np.linalg.inv(H.reshape(q * b, q * b)).reshape(q, b, q, b)


class StripModel:
    """Generators of one strip (layer m) for a given partition, and the two-phase apply."""

    def __init__(self, m, b, const, eta, omega, h, n, c_mat, P, K=1):
        self.b, self.n, self.m = b, n, m
        self.part = pt = partition(n, P, K)
        D, L, U = strip_blocks(m, b, const, eta, omega, h, n, c_mat)
        self.D, self.L, self.U = D, L, U
        QP = pt["QP"]
        self.W = np.zeros((P, QP, QP), complex)
        self.G = np.zeros((P, 2, b, QP), complex)          # Gf (scaled), Gl (scaled)
        self.corners = np.zeros((P, 4, b, b), complex)     # pp, pt, tp, tt
        for l in range(P):
            i0, ql = int(pt["leaf_start"][l]), int(pt["q"][l])
            Gi = leaf_inverse(D, L, U, i0, i0 + ql)
            self.W[l, :ql, :ql] = Gi[:, b - 1, :, b - 1]
            cl = L[i0] if l > 0 else np.zeros(b)           # coupling to the separator on the left
            cr = U[i0 + ql - 1] if l < P - 1 else np.zeros(b)
            self.G[l, 0, :, :ql] = cl[:, None] * Gi[0, :, :, b - 1]
            self.G[l, 1, :, :ql] = cr[:, None] * Gi[ql - 1, :, :, b - 1]
            self.corners[l] = [Gi[0, :, 0, :], Gi[0, :, ql - 1, :], Gi[ql - 1, :, 0, :], Gi[ql - 1, :, ql - 1, :]]
        # separator Schur complement, block tridiagonal
        ns = P - 1
        S = np.zeros((ns, b, ns, b), complex)
        for j in range(ns):
            s = int(pt["sep"][j])
            pp_r = self.corners[j + 1][0]
            tt_l = self.corners[j][3]
            S[j, :, j, :] = D[s] - L[s][:, None] * tt_l * U[s - 1][None, :] - U[s][:, None] * pp_r * L[s + 1][None, :]
            if j + 1 < ns:
                s2 = int(pt["sep"][j + 1])
                ptm, tpm = self.corners[j + 1][1], self.corners[j + 1][2]
                S[j, :, j + 1, :] = -U[s][:, None] * ptm * U[s2 - 1][None, :]
                S[j + 1, :, j, :] = -L[s2][:, None] * tpm * L[s + 1][None, :]
        self.S = S.reshape(ns * b, ns * b)
        self.N = np.linalg.inv(self.S) if ns else np.zeros((0, 0), complex)

    def apply(self, v):
        b, pt, P = self.b, self.part, self.part["P"]
        ns = P - 1
        g = np.zeros((P, 2, b), complex)
        y = np.zeros(self.n, complex)
        for l in range(P):
            i0, ql = int(pt["leaf_start"][l]), int(pt["q"][l])
            g[l] = self.G[l, :, :, :ql] @ v[i0:i0 + ql]
        rho = np.zeros((ns, b), complex)
        for j in range(ns):
            rho[j] = -g[j, 1] - g[j + 1, 0]
            rho[j, b - 1] += v[int(pt["sep"][j])]
        xs = (self.N @ rho.ravel()).reshape(ns, b)
        for l in range(P):
            i0, ql = int(pt["leaf_start"][l]), int(pt["q"][l])
            yl = self.W[l, :ql, :ql] @ v[i0:i0 + ql]
            if l > 0:
                yl = yl - self.G[l, 0, :, :ql].T @ xs[l - 1]
            if l < P - 1:
                yl = yl - self.G[l, 1, :, :ql].T @ xs[l]
            y[i0:i0 + ql] = yl
        for j in range(ns):
            y[int(pt["sep"][j])] = xs[j, b - 1]
        return y


def half_warp_inverse_model(A):
    """Lane-level model of hp_half_inv (csrc/hp_setup.cu): lane j of a half-warp holds column j of the B x B block in
    "registers" R[k][j]; the pivot loop is rolled, after every step the rows rotate by one register so that the pivot row
    is always register 0; pivot rows are kept as 4-bit fields; the row exchanges are undone on the columns (lanes) at the
    end.  Returns (inverse, pivot word)."""
    A = np.array(A, dtype=np.complex128)
    B = A.shape[0]
    assert A.shape == (B, B) and B <= 16
    R = A.copy()                                    # R[k, j]: register k of lane j
    pivs = 0
    for p in range(B):
        # pivot search inside lane p over the live registers 0 .. B-1-p (|re| + |im|, first maximum)
        col = R[:B - p, p]
        norm = np.abs(col.real) + np.abs(col.imag)
        kr = int(np.argmax(norm))
        if norm[kr] == 0.0:
            raise ZeroDivisionError("singular block")
        pivs |= (p + kr) << (4 * p)
        if kr != 0:
            R[[0, kr], :] = R[[kr, 0], :]
        a = R[0, p]
        r = 1.0 / (a.real * a.real + a.imag * a.imag)
        d = complex(a.real * r, -a.imag * r)        # conj(a) / |a|^2
        prow = np.where(np.arange(B) == p, 1.0 + 0j, R[0, :]) * d
        f = R[1:, p].copy()                         # multipliers, broadcast from lane p
        base = R[1:, :].copy()
        base[:, p] = 0.0
        R[:B - 1, :] = base - np.outer(f, prow)     # eliminated rows move up by one register
        R[B - 1, :] = prow                          # the pivot row goes to the end
    for p in range(B - 1, -1, -1):                  # after B rotations register k holds row k again
        r = (pivs >> (4 * p)) & 15
        if r != p:
            R[:, [p, r]] = R[:, [r, p]]
    return R, pivs
