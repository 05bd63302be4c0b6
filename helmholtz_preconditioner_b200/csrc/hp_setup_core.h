// Per-task setup arithmetic of the strip solves (host/device).
//
// Replaces the reference's algo2_3 (/root/reference/code.py:345-353): instead of a sparse LU of each
// bn x bn strip operator H_m, the strip is treated as a block tridiagonal matrix over the x1 index
// (b x b blocks) and the restriction T_m = (H_m^{-1})[last row, last row] is represented by
//   * per leaf (<= QP consecutive block rows): samples of the Dirichlet-truncated leaf inverse,
//   * per tree node: the 2b x 2b interface solve that merges two neighbouring segments.
// tools/tree_prototype.py is the numpy model of exactly these arrays.
#pragma once
#include "hp_small.h"

struct HpStripCtx {
    int n, b;
    HpPml pml;
    cplx omega2;
    const cplx* s1t;      // s1 on the x1 half grid, t = 0..2n+2
    const cplx* is1t;     // 1/s1
    const double* c_mat;  // (n+2) x (n+2), row major, indexed [i-1][j-1] as the reference does
};

// Forward Schur chain of one leaf: Finv[i] = (D_i - L_i Finv[i-1] U_{i-1})^{-1}, i = i0..i1 (1-based,
// inclusive).  out points at this strip's [n][b*b] scratch.
HP_HD int hp_chain_forward(cplx* out, int i0, int i1, int m, const HpStripCtx& c) {
    const int b = c.b, bb = b * b;
    HpStripRow R;
    hp_strip_rows(R, m, b, c.pml);
    HpBlockRow B;
    cplx Uprev[HP_BMAX];
    cplx F[HP_BMAX * HP_BMAX];
    int bad = 0;
    for (int i = i0; i <= i1; ++i) {
        hp_block_row(B, R, i, m, b, c.n, c.pml, c.s1t, c.is1t, c.c_mat, c.omega2);
        hp_schur_step(F, B, B.L, i > i0 ? out + (size_t)(i - 2) * bb : out, Uprev, b, i > i0);
        bad |= hp_inv_inplace(F, b);
        cplx* dst = out + (size_t)(i - 1) * bb;
        for (int e = 0; e < bb; ++e) dst[e] = F[e];
        for (int k = 0; k < b; ++k) Uprev[k] = B.U[k];
    }
    return bad;
}

// Backward Schur chain Binv[i] = (D_i - U_i Binv[i+1] L_{i+1})^{-1}, i = i1..i0, followed by the
// ascending recurrence for the diagonal blocks of the leaf inverse
//   G_{i0,i0} = Binv[i0],   G_ii = Binv_i + Binv_i L_i G_{i-1,i-1} U_{i-1} Binv_i
// of which only the last column (source in the last strip row) is kept: gcol[i][0..b).
HP_HD int hp_chain_backward(cplx* out, cplx* gcol, int i0, int i1, int m, const HpStripCtx& c) {
    const int b = c.b, bb = b * b;
    HpStripRow R;
    hp_strip_rows(R, m, b, c.pml);
    HpBlockRow B;
    cplx Lnext[HP_BMAX];
    cplx F[HP_BMAX * HP_BMAX];
    int bad = 0;
    for (int i = i1; i >= i0; --i) {
        hp_block_row(B, R, i, m, b, c.n, c.pml, c.s1t, c.is1t, c.c_mat, c.omega2);
        hp_schur_step(F, B, B.U, i < i1 ? out + (size_t)i * bb : out, Lnext, b, i < i1);
        bad |= hp_inv_inplace(F, b);
        cplx* dst = out + (size_t)(i - 1) * bb;
        for (int e = 0; e < bb; ++e) dst[e] = F[e];
        for (int k = 0; k < b; ++k) Lnext[k] = B.L[k];
    }
    // diagonal blocks, ascending
    cplx G[HP_BMAX * HP_BMAX], T1[HP_BMAX * HP_BMAX], T2[HP_BMAX * HP_BMAX];
    cplx Uprev[HP_BMAX];
    for (int i = i0; i <= i1; ++i) {
        const cplx* Bi = out + (size_t)(i - 1) * bb;
        hp_block_row(B, R, i, m, b, c.n, c.pml, c.s1t, c.is1t, c.c_mat, c.omega2);
        if (i == i0) {
            for (int e = 0; e < bb; ++e) G[e] = Bi[e];
        } else {
            for (int r = 0; r < b; ++r)
                for (int s = 0; s < b; ++s) T1[r * b + s] = cmul(cmul(B.L[r], G[r * b + s]), Uprev[s]);
            hp_gemm(T2, Bi, T1, b, b, b, b, b, b, +1, 0);
            for (int e = 0; e < bb; ++e) G[e] = Bi[e];
            hp_gemm(G, T2, Bi, b, b, b, b, b, b, +1, 1);
        }
        for (int k = 0; k < b; ++k) gcol[(size_t)(i - 1) * b + k] = G[k * b + (b - 1)];
        for (int k = 0; k < b; ++k) Uprev[k] = B.U[k];
    }
    return bad;
}

// Corner blocks of a segment [p..t]:  pp = G_pp, pt = G_pt, tp = G_tp, tt = G_tt  (each b x b)
// Merge of segment 1 = [p..q] and segment 2 = [q+1..t]; cpl[k] = U_q[k] = L_{q+1}[k].
// Node record: UP (4b x 2b) then DN (2b x 2b), see tools/tree_prototype.py::merge.
HP_HD int hp_merge(cplx* rec, cplx* out_corners, const cplx* c1, const cplx* c2, const cplx* cpl, int b) {
    const int bb = b * b, b2 = 2 * b;
    const cplx *pp1 = c1, *pt1 = c1 + bb, *tp1 = c1 + 2 * bb, *tt1 = c1 + 3 * bb;
    const cplx *pp2 = c2, *pt2 = c2 + bb, *tp2 = c2 + 2 * bb, *tt2 = c2 + 3 * bb;
    cplx X[HP_BMAX * HP_BMAX], Y[HP_BMAX * HP_BMAX], K[HP_BMAX * HP_BMAX], KX[HP_BMAX * HP_BMAX],
        YK[HP_BMAX * HP_BMAX];
    for (int r = 0; r < b; ++r)
        for (int s = 0; s < b; ++s) {
            X[r * b + s] = cmul(tt1[r * b + s], cpl[s]);
            Y[r * b + s] = cmul(pp2[r * b + s], cpl[s]);
        }
    hp_gemm(K, X, Y, b, b, b, b, b, b, -1, 0);                     // K = -X Y
    for (int r = 0; r < b; ++r) K[r * b + r].x += 1.0;             // I - X Y
    int bad = hp_inv_inplace(K, b);
    hp_gemm(KX, K, X, b, b, b, b, b, b, +1, 0);
    hp_gemm(YK, Y, K, b, b, b, b, b, b, +1, 0);
    cplx* UP = rec;
    cplx* DN = rec + 8 * bb;
    // Ua = diag(cpl) [K, -KX]   rows 0..b-1;   Uc = diag(cpl) [-YK, I + Y KX]   rows b..2b-1
    for (int r = 0; r < b; ++r)
        for (int s = 0; s < b; ++s) {
            UP[r * b2 + s] = cmul(cpl[r], K[r * b + s]);
            UP[r * b2 + b + s] = cneg(cmul(cpl[r], KX[r * b + s]));
            UP[(b + r) * b2 + s] = cneg(cmul(cpl[r], YK[r * b + s]));
        }
    hp_gemm(X, Y, KX, b, b, b, b, b, b, +1, 0);                    // X := Y K X   (X no longer needed)
    for (int r = 0; r < b; ++r)
        for (int s = 0; s < b; ++s) {
            cplx v = X[r * b + s];
            if (r == s) v.x += 1.0;
            UP[(b + r) * b2 + b + s] = cmul(cpl[r], v);
        }
    // rows 2b..3b-1 = -pt1 * Uc ; rows 3b..4b-1 = -tp2 * Ua
    hp_gemm(UP + (size_t)2 * b * b2, pt1, UP + (size_t)b * b2, b, b, b2, b2, b, b2, -1, 0);
    hp_gemm(UP + (size_t)3 * b * b2, tp2, UP, b, b, b2, b2, b, b2, -1, 0);
    // DN[:, :b] = -[Ua; Uc][:, :b] tp1 ;  DN[:, b:] = -[Ua; Uc][:, b:] pt2
    hp_gemm(DN, UP, tp1, b2, b, b, b2, b2, b, -1, 0);
    hp_gemm(DN + b, UP + b, pt2, b2, b, b, b2, b2, b, -1, 0);
    // merged corners
    cplx *pp = out_corners, *pt = out_corners + bb, *tp = out_corners + 2 * bb, *tt = out_corners + 3 * bb;
    for (int e = 0; e < bb; ++e) { pp[e] = pp1[e]; tt[e] = tt2[e]; }
    hp_gemm(pp, pt1, DN + (size_t)b * b2, b, b, b, b, b, b2, +1, 1);          // pp1 + pt1 DN[b:2b, :b]
    hp_gemm(tp, tp2, DN, b, b, b, b, b, b2, +1, 0);                          // tp2 DN[0:b, :b]
    hp_gemm(tt, tp2, DN + b, b, b, b, b, b, b2, +1, 1);                      // tt2 + tp2 DN[0:b, b:]
    hp_gemm(pt, pt1, DN + (size_t)b * b2 + b, b, b, b, b, b, b2, +1, 0);      // pt1 DN[b:2b, b:]
    return bad;
}

// coupling between block rows q and q+1 of strip m: U_q[k] = L_{q+1}[k] = 1/h^2 s1((q+.5)h)/s2m(j_k h)
HP_HD void hp_coupling(cplx* cpl, int q, int m, const HpStripCtx& c) {
    HpStripRow R;
    hp_strip_rows(R, m, c.b, c.pml);
    double ih2 = 1.0 / (c.pml.h * c.pml.h);
    cplx s = c.s1t[2 * q + 1];
    for (int k = 0; k < c.b; ++k) cpl[k] = cscale(ih2, cmul(s, R.is2c[k]));
}
