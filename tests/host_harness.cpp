// CPU build of the host/device setup arithmetic (csrc/hp_small.h, csrc/hp_setup_core.h) so that it can
// be checked against tools/strip_model.py in the GPU-less build container.  Test infrastructure: the
// loops below stand in for the thread grids of csrc/hp_setup.cu and call the same per-thread bodies.
#include "../helmholtz_preconditioner_b200/csrc/hp_setup_core.h"
#include <vector>

extern "C" {

void hh_tables(int n, double cst, double eta, double h, double om_re, double om_im,
               cplx* s1t, cplx* is1t, cplx* s2t, cplx* is2t) {
    HpPml p{cst, eta, h, cmake(om_re, om_im)};
    for (int t = 0; t <= 2 * n + 2; ++t) hp_table_entry(t, p, s1t + t, is1t + t, s2t + t, is2t + t);
}

int hh_inv(int b, cplx* A) { return hp_inv_inplace(A, b); }

// Whole setup of one strip m for the partition (leaf_start[P], q[P], sep[P-1], all 0-based columns).
//   W [P][QP][QP], G [P][2][b][QP], N [ns*b][ns*b]
int hh_strip_setup(int n, int b, int m, int P, int QP, const int* leaf_start, const int* q, const int* sep,
                   double cst, double eta, double h, double om_re, double om_im, const double* c_mat,
                   cplx* W, cplx* G, cplx* N) {
    std::vector<cplx> s1t(2 * n + 3), is1t(2 * n + 3), s2t(2 * n + 3), is2t(2 * n + 3);
    hh_tables(n, cst, eta, h, om_re, om_im, s1t.data(), is1t.data(), s2t.data(), is2t.data());
    HpStripCtx c;
    c.n = n; c.b = b; c.pml = HpPml{cst, eta, h, cmake(om_re, om_im)};
    c.omega2 = cmul(c.pml.omega, c.pml.omega);
    c.s1t = s1t.data(); c.is1t = is1t.data(); c.c_mat = c_mat;
    const int bb = b * b, ns = P - 1;
    std::vector<cplx> Finv((size_t)n * bb), Binv((size_t)n * bb), gcol((size_t)n * b);
    int bad = 0;
    for (int l = 0; l < P; ++l) {                                           // kernel: chains
        bad |= hp_chain_forward(Finv.data(), leaf_start[l] + 1, leaf_start[l] + q[l], m, c);
        bad |= hp_chain_backward(Binv.data(), gcol.data(), leaf_start[l] + 1, leaf_start[l] + q[l], m, c);
    }
    for (int l = 0; l < P; ++l)                                             // kernel: leaf columns
        for (int r = 0; r < QP; ++r)
            hp_leaf_column(W + ((size_t)l * QP + r) * QP, G + ((size_t)l * 2 * b) * QP + r,
                           G + ((size_t)l * 2 * b + b) * QP + r, QP, Finv.data(), Binv.data(), gcol.data(),
                           leaf_start[l] + 1, q[l], QP, r, m, l > 0, l < P - 1, c);
    if (ns == 0) return bad;
    std::vector<cplx> tp((size_t)P * bb);
    for (int l = 1; l < P - 1; ++l)                                         // kernel: corners
        for (int kap = 0; kap < b; ++kap) {
            cplx col[HP_BMAX];
            hp_leaf_corner_tp(col, Binv.data(), leaf_start[l] + 1, q[l], QP, kap, m, c);
            for (int a = 0; a < b; ++a) tp[(size_t)l * bb + a * b + kap] = col[a];
        }
    std::vector<cplx> Sd((size_t)ns * bb), So((size_t)ns * bb), FX((size_t)ns * bb), FXi((size_t)ns * bb),
        PF((size_t)ns * bb), BX((size_t)ns * bb), BXi((size_t)ns * bb), PB((size_t)ns * bb), Njj((size_t)ns * bb);
    for (int j = 0; j < ns; ++j) {                                          // kernel: separator blocks
        int s = sep[j] + 1;
        hp_sep_diag(Sd.data() + (size_t)j * bb, s, m, Finv.data() + (size_t)(s - 2) * bb,
                    Binv.data() + (size_t)s * bb, c);
        if (j + 1 < ns) hp_sep_offdiag(So.data() + (size_t)j * bb, s, sep[j + 1] + 1, m, tp.data() + (size_t)(j + 1) * bb, c);
    }
    bad |= hp_sep_chain(FX.data(), FXi.data(), PF.data(), Sd.data(), So.data(), ns, +1, b);   // kernel: sep chains
    bad |= hp_sep_chain(BX.data(), BXi.data(), PB.data(), Sd.data(), So.data(), ns, -1, b);
    for (int j = 0; j < ns; ++j)
        bad |= hp_sep_diag_inverse(Njj.data() + (size_t)j * bb, FX.data() + (size_t)j * bb, BX.data() + (size_t)j * bb,
                                   Sd.data() + (size_t)j * bb, b);
    for (int j = 0; j < ns; ++j)                                            // kernel: separator rows
        for (int kap = 0; kap < b; ++kap)
            hp_sep_row(N + ((size_t)j * b + kap) * ns * b, Njj.data(), PF.data(), PB.data(), ns, j, kap, b);
    return bad;
}
}
