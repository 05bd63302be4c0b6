"""Host model (numpy) of the GPU strip-solve data structures.

Development/test helper: it mirrors, array for array, what csrc/hp_setup.cu
produces and what csrc/hp_sweep.cu consumes, so the CUDA stages can be checked
one by one.  It is not part of the product path and not the oracle (the oracle
follows the reference's splu formulation).

The strip operator H_m (reference get_Hm, code.py:283-290) is block tridiagonal
when the unknowns are ordered x1-major: block row i (x1 index) holds the b
unknowns of column i, the diagonal block D_i is the b x b tridiagonal x2
coupling and the off-diagonal blocks L_i (to i-1), U_i (to i+1) are diagonal.
T_m v = (H_m^{-1} [0; v]) restricted to the last strip row (code.py:368-370).

  leaves : contiguous ranges of block rows.  Per leaf the Dirichlet-truncated
           inverse G = H_leaf^{-1} is sampled: W = G[(i,b),(i',b)], the first
           and last block rows of the columns (i', b) (Gf, Gl) and the four
           b x b corner blocks.
  nodes  : a binary tree merges neighbouring segments through the 2b x 2b
           interface system of the two facing block rows.
"""
import numpy as np

from oracle import helmholtz_oracle as orc


def leaf_partition(n, qmax):
    """Leaves 0..P-1 (P a power of two); leaf l covers block rows [start[l], start[l+1])."""
    d = 0
    while -(-n // (1 << d)) > qmax:
        d += 1
    P = 1 << d
    start = np.array([(l * n) // P for l in range(P + 1)], dtype=np.int64)
    return d, P, start


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


def leaf_generators(D, L, U, i0, i1):
    """RGF on block rows i0..i1-1 (0-based, half open).  Returns dict with
    W (q,q), Gf (b,q), Gl (b,q), corners pp, pt, tp, tt (b,b)."""
    q = i1 - i0
    b = D.shape[1]
    Finv = np.zeros((q, b, b), complex)
    Binv = np.zeros((q, b, b), complex)
    F = D[i0].copy()
    Finv[0] = np.linalg.inv(F)
    for r in range(1, q):
        i = i0 + r
        F = D[i] - (L[i][:, None] * Finv[r - 1]) * U[i - 1][None, :]
        Finv[r] = np.linalg.inv(F)
    Bm = D[i1 - 1].copy()
    Binv[q - 1] = np.linalg.inv(Bm)
    for r in range(q - 2, -1, -1):
        i = i0 + r
        Bm = D[i] - (U[i][:, None] * Binv[r + 1]) * L[i + 1][None, :]
        Binv[r] = np.linalg.inv(Bm)
    W = np.zeros((q, q), complex)
    Gf = np.zeros((b, q), complex)
    Gl = np.zeros((b, q), complex)
    gcol = np.zeros((q, b), complex)
    for r in range(q):
        i = i0 + r
        # G_ii = (F_i + B_i - D_i)^{-1}; only its last column is needed
        Fi = np.linalg.inv(Finv[r])
        Bi = np.linalg.inv(Binv[r])
        x = np.linalg.solve(Fi + Bi - D[i], np.eye(b)[:, b - 1])
        W[r, r] = x[b - 1]
        gcol[r] = x
        xl = x.copy()
        for rr in range(r - 1, -1, -1):            # leftwards: x_i = -F_i^{-1} U_i x_{i+1}
            xl = -Finv[rr] @ (U[i0 + rr] * xl)
            W[rr, r] = xl[b - 1]
        Gf[:, r] = xl
        xr = x.copy()
        for rr in range(r + 1, q):                 # rightwards: x_i = -B_i^{-1} L_i x_{i-1}
            xr = -Binv[rr] @ (L[i0 + rr] * xr)
            W[rr, r] = xr[b - 1]
        Gl[:, r] = xr
    # corner blocks
    pp = Binv[0].copy()                            # G_{pp} = B_p^{-1}
    tt = Finv[q - 1].copy()                        # G_{tt} = F_t^{-1}
    X = pp.copy()
    for rr in range(1, q):
        X = -Binv[rr] @ (L[i0 + rr][:, None] * X)
    tp = X                                         # G_{t,p}
    X = tt.copy()
    for rr in range(q - 2, -1, -1):
        X = -Finv[rr] @ (U[i0 + rr][:, None] * X)
    pt = X                                         # G_{p,t}
    return dict(W=W, Gf=Gf, Gl=Gl, pp=pp, pt=pt, tp=tp, tt=tt, Finv=Finv, Binv=Binv, gcol=gcol)


def merge(c1, c2, Uq, Lq1):
    """Merge segment 1 = [p..q] and segment 2 = [q+1..t] (Uq = U_q, Lq1 = L_{q+1}).

    Interface unknowns a = x_q, c = x_{q+1}:  [[I, X], [Y, I]] [a; c] = [g1_q; g2_{q+1}],
    X = G1_tt diag(U_q), Y = G2_pp diag(L_{q+1}).  Values handed between tree levels are
    pre-scaled by the coupling they will meet: a~ = L_{q+1} a, c~ = U_q c, and a segment
    receives x~l = L_p x_{p-1}, x~r = U_t x_{t+1}.

    Node record (what hp_sweep.cu streams per node and layer):
      UP (4b x 2b), applied to in = [g_t(child1); g_p(child2)]:
          rows 0..b-1   a~        rows b..2b-1  c~          (kept as xi~)
          rows 2b..3b-1 g_p(parent) - g_p(child1)   rows 3b..4b-1 g_t(parent) - g_t(child2)
      DN (2b x 2b), applied to [x~l; x~r]:   [a~; c~] = xi~ + DN [x~l; x~r]
    """
    b = c1["pp"].shape[0]
    X = c1["tt"] * Uq[None, :]
    Y = c2["pp"] * Lq1[None, :]
    K = np.linalg.inv(np.eye(b) - X @ Y)
    Naa, Nac = K, -K @ X
    Nca, Ncc = -Y @ K, np.eye(b) + Y @ K @ X
    Ua = Lq1[:, None] * np.hstack([Naa, Nac])
    Uc = Uq[:, None] * np.hstack([Nca, Ncc])
    UP = np.vstack([Ua, Uc, -c1["pt"] @ Uc, -c2["tp"] @ Ua])
    Z = np.zeros((b, b))
    DN = -np.vstack([Ua, Uc]) @ np.block([[c1["tp"], Z], [Z, c2["pt"]]])
    pp = c1["pp"] - c1["pt"] @ (Uc[:, :b] @ c1["tp"])
    tp = -c2["tp"] @ (Ua[:, :b] @ c1["tp"])
    tt = c2["tt"] - c2["tp"] @ (Ua[:, b:] @ c2["pt"])
    pt = -c1["pt"] @ (Uc[:, b:] @ c2["pt"])
    return dict(UP=UP, DN=DN, pp=pp, pt=pt, tp=tp, tt=tt)


class StripTree:
    """All generators of one strip (one moving-PML layer m), in the packed GPU layout:
    W [P][QP][QP], G [P][2][b][QP] (Gf, Gl), nodes [P-1][12 b^2] (UP then DN),
    node (lv, t) at index lvoff[lv] + t, lvoff[1] = 0, lvoff[lv+1] = lvoff[lv] + (P >> lv)."""

    def __init__(self, m, b, const, eta, omega, h, n, c_mat, qmax=64):
        self.b, self.n = b, n
        self.d, self.P, self.start = leaf_partition(n, qmax)
        P, st = self.P, self.start
        self.QP = QP = int(max(st[1:] - st[:-1]))
        D, L, U = strip_blocks(m, b, const, eta, omega, h, n, c_mat)
        self.L, self.U = L, U
        self.leaves = [leaf_generators(D, L, U, st[l], st[l + 1]) for l in range(P)]
        self.W = np.zeros((P, QP, QP), complex)
        self.G = np.zeros((P, 2, b, QP), complex)
        for l, lf in enumerate(self.leaves):
            q = st[l + 1] - st[l]
            self.W[l, :q, :q] = lf["W"]
            self.G[l, 0, :, :q] = lf["Gf"]
            self.G[l, 1, :, :q] = lf["Gl"]
        self.lvoff = [0, 0]
        for lv in range(1, self.d + 1):
            self.lvoff.append(self.lvoff[-1] + (P >> lv))
        self.nodes = np.zeros((max(P - 1, 1), 12 * b * b), complex)
        self.corners = [self.leaves]
        prev = self.leaves
        for lv in range(1, self.d + 1):
            cur = []
            for t in range(P >> lv):
                q1 = st[((2 * t + 1) << (lv - 1))]   # first block row of the right child
                nd = merge(prev[2 * t], prev[2 * t + 1], U[q1 - 1], L[q1])
                self.nodes[self.lvoff[lv] + t, :8 * b * b] = nd["UP"].ravel()
                self.nodes[self.lvoff[lv] + t, 8 * b * b:] = nd["DN"].ravel()
                cur.append(nd)
            self.corners.append(cur)
            prev = cur

    def apply(self, v):
        """y = T_m v through leaf / up / down / leaf phases (what hp_sweep.cu does per layer)."""
        b, P, st, QP = self.b, self.P, self.start, self.QP
        vp = np.zeros((P, QP), complex)
        for l in range(P):
            vp[l, :st[l + 1] - st[l]] = v[st[l]:st[l + 1]]
        seg = [[np.concatenate([self.G[l, 0] @ vp[l], self.G[l, 1] @ vp[l]]) for l in range(P)]]  # (g_p, g_t)
        y0 = np.einsum("lrc,lc->lr", self.W, vp)
        xi = [None]
        for lv in range(1, self.d + 1):                      # upward
            cur, x = [], []
            for t in range(P >> lv):
                rec = self.nodes[self.lvoff[lv] + t]
                UP = rec[:8 * b * b].reshape(4 * b, 2 * b)
                s1_, s2_ = seg[lv - 1][2 * t], seg[lv - 1][2 * t + 1]
                o = UP @ np.concatenate([s1_[b:], s2_[:b]])
                x.append(o[:2 * b])
                cur.append(np.concatenate([s1_[:b] + o[2 * b:3 * b], s2_[b:] + o[3 * b:]]))
            seg.append(cur)
            xi.append(x)
        ext = [np.zeros(2 * b, complex)]                      # (x~l, x~r) of the root
        for lv in range(self.d, 0, -1):                       # downward
            nxt = []
            for t in range(P >> lv):
                rec = self.nodes[self.lvoff[lv] + t]
                DN = rec[8 * b * b:].reshape(2 * b, 2 * b)
                ac = xi[lv][t] + DN @ ext[t]
                nxt.append(np.concatenate([ext[t][:b], ac[b:]]))     # left child: (x~l, c~)
                nxt.append(np.concatenate([ac[:b], ext[t][b:]]))     # right child: (a~, x~r)
            ext = nxt
        y = np.zeros(self.n, complex)
        for l in range(P):
            yl = y0[l] - self.G[l, 0].T @ ext[l][:b] - self.G[l, 1].T @ ext[l][b:]
            y[st[l]:st[l + 1]] = yl[:st[l + 1] - st[l]]
        return y
