// CPU build of the host/device setup arithmetic (csrc/hp_small.h, csrc/hp_setup_core.h) so that it can
// be checked against tools/tree_prototype.py in the GPU-less build container.  Test infrastructure.
#include "../helmholtz_preconditioner_b200/csrc/hp_setup_core.h"
#include <vector>

extern "C" {

void hh_tables(int n, double cst, double eta, double h, double om_re, double om_im,
               cplx* s1t, cplx* is1t, cplx* s2t, cplx* is2t) {
    HpPml p{cst, eta, h, cmake(om_re, om_im)};
    for (int t = 0; t <= 2 * n + 2; ++t) hp_table_entry(t, p, s1t + t, is1t + t, s2t + t, is2t + t);
}

// Finv, Binv: [n][b*b]; gcol: [n][b]; leaf block rows i0..i1 (1-based inclusive)
int hh_leaf_chains(int n, int b, int m, int i0, int i1, double cst, double eta, double h, double om_re,
                   double om_im, const double* c_mat, cplx* Finv, cplx* Binv, cplx* gcol) {
    std::vector<cplx> s1t(2 * n + 3), is1t(2 * n + 3), s2t(2 * n + 3), is2t(2 * n + 3);
    hh_tables(n, cst, eta, h, om_re, om_im, s1t.data(), is1t.data(), s2t.data(), is2t.data());
    HpStripCtx c;
    c.n = n; c.b = b; c.pml = HpPml{cst, eta, h, cmake(om_re, om_im)};
    c.omega2 = cmul(c.pml.omega, c.pml.omega);
    c.s1t = s1t.data(); c.is1t = is1t.data(); c.c_mat = c_mat;
    int bad = hp_chain_forward(Finv, i0, i1, m, c);
    bad |= hp_chain_backward(Binv, gcol, i0, i1, m, c);
    return bad;
}

int hh_merge(int b, const cplx* c1, const cplx* c2, const cplx* cpl, cplx* rec, cplx* corners) {
    return hp_merge(rec, corners, c1, c2, cpl, b);
}

void hh_coupling(int n, int b, int m, int q, double cst, double eta, double h, double om_re, double om_im,
                 cplx* cpl) {
    std::vector<cplx> s1t(2 * n + 3), is1t(2 * n + 3), s2t(2 * n + 3), is2t(2 * n + 3);
    hh_tables(n, cst, eta, h, om_re, om_im, s1t.data(), is1t.data(), s2t.data(), is2t.data());
    HpStripCtx c;
    c.n = n; c.b = b; c.pml = HpPml{cst, eta, h, cmake(om_re, om_im)};
    c.s1t = s1t.data(); c.is1t = is1t.data(); c.c_mat = nullptr;
    hp_coupling(cpl, q, m, c);
}

int hh_inv(int b, cplx* A) { return hp_inv_inplace(A, b); }
}
