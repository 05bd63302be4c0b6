// Pipelined sweep kernel: the same strip chain as csrc/hp_sweep.cu (algo2_4, /root/reference/code.py:366-380)
// with the per-strip dependency cycle cut down by using the linearity of the recurrence.
//
// Per strip t every leaf part needs the separator solution x(t-1) of the previous strip only through
//     v(t) = vb(t) + coef * Gc(t-1)^T x(t-1)            (input of strip t on the part's columns)
//     g(t) = Gc(t) v(t) = gb(t) + M(t) x(t-1)           (interface data;  M(t) precomputed per leaf, hp_mleaf_kernel)
// where vb(t), gb(t) do not depend on x(t-1).  Two thread groups of a CTA run decoupled loops that talk only
// through the self-validating exchange ring in L2 (csrc/hp_sweep_common.cuh):
//
//   critical group (warps 0-3)                            off-path group (warps 4-11)
//   C1 leaf's first CTA: GR(t) = sum_k gb_k(t)            a  gb(t) = Gc(t) vb(t)                    -> GPb(t)
//                        + M(t) x(t-1)         -> GR(t)   b  wait x(t-1): v(t) = vb(t) + coef Gc(t-1)^T x(t-1)
//   C2 rho(t) from GR(t), VS(t);  own rows of                -> field row, V(t)
//      x(t) = N(t) rho(t)                      -> XS(t)   c  gather v_leaf(t); y0(t) = W(t) v_leaf(t); vb(t+1)
//      separator columns: field, VS(t+1)
//
// The cycle x(t-1) -> x(t) is two L2 hand-overs and two small matvecs; everything that touches the big W blocks
// hangs off it with a slack of one strip.  Packets are staged by TMA in three independent shared-memory rings
// (W+G for the off-path group, N for the critical group, M for the leaf's first CTA), each refilled by the group
// that consumes it.
#include "hp_sweep_common.cuh"

#define HP2_CRIT 128
#define HP2_OFF 256
#define HP2_THREADS (HP2_CRIT + HP2_OFF)
#define HP2_KPL 2            // interface components per lane of a warp: 2b <= 64

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define HP_STAMP(k) do { if (DBG && it >= 512 && it < 576) a.dbg[(size_t)G * 16 + ((size_t)g * 64 + (it - 512)) * 4 + (k)] = (long long)gtime(); } while (0)
__device__ __forceinline__ void bar_crit() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bar_off() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// spin until the word is valid; a runaway spin (a bug) marks the CTA dead so that all later waits fall through
__device__ __forceinline__ cplx xwait(const cplx* p, unsigned int* abort_flag, volatile unsigned int* dead) {
    cplx v;
    unsigned int spins = 0;
    while (!xtry(p, v)) {
        if (*dead) break;
        if (++spins > HP_SPIN_LIMIT) { hp_raise_abort(abort_flag); *dead = 1u; break; }
        if ((spins & 0xFFF) == 0 && *((volatile unsigned int*)abort_flag)) { *dead = 1u; break; }
    }
    return v;
}

__device__ __forceinline__ void ring_fill(unsigned char* dst, const cplx* src, unsigned int bytes, unsigned long long* bar) {
    mbar_expect_tx(bar, bytes);
    for (unsigned int o = 0; o < bytes; o += HP_BULK_CHUNK)
        bulk_g2s(dst + o, (const char*)src + o, min(HP_BULK_CHUNK, bytes - o), bar);
}

// MODE: 0 forward, 1 backward (reference diagonal), 2 backward (paper diagonal), 3 single strip apply
template <int MODE, bool DBG>
__global__ void __launch_bounds__(HP2_THREADS) hp_sweep2_kernel(HpSweepArgs a) {
    constexpr int a_mode = MODE == 0 ? 0 : (MODE == 3 ? 2 : 1);
    constexpr int a_diag = MODE == 2 ? 1 : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = a.b, b2 = 2 * a.b, n = a.n, K = a.lay.K, P = a.lay.P, G = a.lay.G, QP = a.lay.QP, CW = a.lay.CW;
    const int NS = a.lay.NS, NSP = a.lay.NSP, NR = a.lay.NR;
    const int g = blockIdx.x, l = g / K, k = g % K;
    const int tid = threadIdx.x, lane = tid & 31;
    const int q = a.leaf_q[l], ls = a.leaf_start[l];
    const int lc0 = (q * k) / K, lc1 = (q * (k + 1)) / K, ncols = lc1 - lc0, c0 = ls + lc0;
    const int row0 = g * NR, nrows = max(0, min(NR, NS - row0));
    unsigned int* abort_flag = a.bar + 1;
    const int step = a_mode == 1 ? -1 : 1;
    const int nsteps = a_mode == 2 ? 1 : (a_mode == 0 ? a.m_to - a.m_from + 1 : a.m_from - a.m_to + 1);
    const int dir = a_mode == 1 ? 1 : 0;
    const double sg = a_diag == 0 ? 1.0 : -1.0;

    // ---- shared memory carve-up (must match hp_sweep2_smem)
    const unsigned int wg_bytes = (unsigned int)(a.lay.offN * sizeof(cplx));
    const unsigned int n_bytes = (unsigned int)((size_t)NR * NSP * sizeof(cplx));
    const unsigned int m_bytes = (unsigned int)((size_t)b2 * b2 * sizeof(cplx));
    const size_t wg_st = ((size_t)wg_bytes + 127) & ~(size_t)127, n_st = ((size_t)n_bytes + 127) & ~(size_t)127,
                 m_st = ((size_t)m_bytes + 127) & ~(size_t)127;
    unsigned char* ringWG = smem_raw;
    unsigned char* ringN = ringWG + 2 * wg_st;
    unsigned char* ringM = ringN + 2 * n_st;
    cplx* Gprev = reinterpret_cast<cplx*>(ringM + 2 * m_st);     // [2b][CW]  Gc of the previous strip
    cplx* vb = Gprev + (size_t)b2 * CW;                          // [CW]
    cplx* v_leaf = vb + CW;                                      // [QP]
    cplx* y0w = v_leaf + QP;                                     // [8][CW]
    cplx* gpw = y0w + 8 * (size_t)CW;                            // [8][2b]
    cplx* xlr_o = gpw + 8 * (size_t)b2;                          // [2b]   off-path copy of (x_left, x_right)
    cplx* xlr_c = xlr_o + b2;                                    // [2b]   critical copy
    cplx* rho = xlr_c + b2;                                      // [NSP]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(rho + NSP);   // WG[2], N[2], M[2]
    volatile unsigned int* dead = reinterpret_cast<volatile unsigned int*>(mbar + 6);

    const cplx* pk_base = a.packets + (size_t)g * a.lay.PK;
    const size_t strip_stride = (size_t)G * a.lay.PK;
    const size_t m_stride = (size_t)2 * P * b2 * b2;             // complex numbers per strip in mleaf
    const cplx* m_base = a.mleaf + ((size_t)dir * P + l) * b2 * b2;
    const bool reducer = k == 0;
    const int m0 = a.m_from;

    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&mbar[i], 1);
        *dead = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (tid < HP2_CRIT) {
        // =====================================================================================================
        // critical group
        // =====================================================================================================
        const int ctid = tid, cw = ctid >> 5;
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
#define HP_TICK(i) do { if (DBG && lane == 0 && (tid == 0 || tid == HP2_CRIT)) { long long t_ = clock64(); tacc[i] += t_ - tprev; tprev = t_; } } while (0)
        if (ctid == 0) {
            for (int sidx = 0; sidx < 2 && sidx < nsteps; ++sidx) {
                int ms = m0 + sidx * step;
                if (n_bytes && nrows > 0) ring_fill(ringN + sidx * n_st, pk_base + (size_t)(ms - a.m_lo) * strip_stride + a.lay.offN, n_bytes, &mbar[2 + sidx]);
                if (reducer) ring_fill(ringM + sidx * m_st, m_base + (size_t)(ms - a.m_lo) * m_stride, m_bytes, &mbar[4 + sidx]);
            }
        }
        // rows of x_S this warp computes: rr = cw + 4*o; lane o keeps the state of row rr (separator columns)
        const int my_rr = cw + 4 * lane;
        int sep_j = -1, sep_col = -1;
        if (my_rr < nrows) {
            int row = row0 + my_rr;
            int j = row / b;
            if (row - j * b == b - 1) { sep_j = j; sep_col = a.sep[j]; }
        }
        const cplx cis1s = sep_col >= 0 ? a.is1t[2 * (sep_col + 1)] : cmake(0.0, 0.0);
        cplx usbase = cmake(0.0, 0.0);
        if (sep_col >= 0) {                                    // input value of the first strip on the separator column
            cplx v;
            if (a_mode == 2) v = a.vin[sep_col];
            else if (a_mode == 0) v = ldcg(a.u + (size_t)(m0 - 1) * n + sep_col);
            else {
                usbase = ldcg(a.u + (size_t)(m0 - 1) * n + sep_col);
                v = usbase;
                if (m0 < n) v = cfma(cscale(sg, cmul(hp_rowfac(a, m0), cis1s)), ldcg(a.u + (size_t)m0 * n + sep_col), v);
            }
            xput(a.xch + a.oVS + sep_j, v);                    // slot 0
        }
        const int rho_j0 = ctid < NS ? ctid / b : 0, rho_k0 = ctid < NS ? ctid - (ctid / b) * b : 0;

        for (int it = 0; it < nsteps; ++it) {
            const int m = m0 + it * step, mn = m + step;
            const bool more = it + 1 < nsteps;
            cplx* slot = a.xch + (size_t)(it & (HP_RING - 1)) * a.slot_stride;
            cplx* slot_prev = a.xch + (size_t)((it + 3) & (HP_RING - 1)) * a.slot_stride;     // strip it-1
            cplx* slot_next = a.xch + (size_t)((it + 1) & (HP_RING - 1)) * a.slot_stride;
            cplx* slot_arm = a.xch + (size_t)((it + 2) & (HP_RING - 1)) * a.slot_stride;      // strip it-2
            // early loads for the separator column of this lane
            cplx usep = cmake(0.0, 0.0), rfac = cmake(0.0, 0.0);
            if (sep_col >= 0) {
                rfac = hp_rowfac(a, a_mode == 1 ? mn : m);
                if (a_mode == 0) usep = ldcg(a.u + (size_t)m * n + sep_col);
                else if (a_mode == 1 && more) usep = ldcg(a.u + (size_t)(mn - 1) * n + sep_col);
            }
            if (DBG && tid == 0) tprev = clock64();
            // ---- C1: the leaf's first CTA turns the partial gb into GR(t) = sum gb + M(t) x(t-1)
            if (reducer && cw == 0) {
                // one batch of loads per lane: its x(t-1) word and the K partial gb words (one L2 round trip when ready)
                cplx gsum[HP2_KPL];
#pragma unroll 1
                for (int i = 0; i < HP2_KPL; ++i) {
                    const int t = lane + 32 * i;
                    gsum[i] = cmake(0.0, 0.0);
                    if (t < b2) {
                        const int side = t / b, kap = t - side * b, j = l - 1 + side;
                        const bool need_x = it > 0 && j >= 0 && j < P - 1;
                        const cplx* px = slot_prev + a.oXS + (size_t)j * b + kap;
                        cplx xv = cmake(0.0, 0.0), gv[16];
                        unsigned int spins = 0;
                        for (;;) {
                            bool ok = true;
                            if (need_x) ok = xtry(px, xv);
                            for (int kk = 0; kk < K; ++kk) ok = xtry(slot + a.oGP + (size_t)(g + kk) * b2 + t, gv[kk & 15]) && ok;
                            if (ok || *dead) break;
                            if (++spins > HP_SPIN_LIMIT) { hp_raise_abort(abort_flag); *dead = 1u; break; }
                            if ((spins & 0xFFF) == 0 && *((volatile unsigned int*)abort_flag)) { *dead = 1u; break; }
                        }
                        xlr_c[t] = xv;
                        for (int kk = 0; kk < K; ++kk) gsum[i] = cadd(gsum[i], gv[kk & 15]);
                    }
                }
                if (lane == 0) HP_STAMP(0);
                __syncwarp();
                // (also at it = 0, where M is not used: the stage must have landed before it is refilled below)
                mbar_wait(&mbar[4 + (it & 1)], (it >> 1) & 1);
                HP_TICK(0);
                const cplx* M = reinterpret_cast<const cplx*>(ringM + (it & 1) * m_st);
#pragma unroll 1
                for (int i = 0; i < HP2_KPL; ++i) {
                    int kp = lane + 32 * i;
                    if (kp < b2) {
                        cplx acc = gsum[i];
                        if (it > 0) {
                            cplx a1 = cmake(0.0, 0.0), a2 = cmake(0.0, 0.0), a3 = cmake(0.0, 0.0);
                            int kap = 0;
#pragma unroll 2
                            for (; kap + 3 < b2; kap += 4) {
                                acc = cfma(M[(size_t)kap * b2 + kp], xlr_c[kap], acc);
                                a1 = cfma(M[(size_t)(kap + 1) * b2 + kp], xlr_c[kap + 1], a1);
                                a2 = cfma(M[(size_t)(kap + 2) * b2 + kp], xlr_c[kap + 2], a2);
                                a3 = cfma(M[(size_t)(kap + 3) * b2 + kp], xlr_c[kap + 3], a3);
                            }
                            for (; kap < b2; ++kap) acc = cfma(M[(size_t)kap * b2 + kp], xlr_c[kap], acc);
                            acc = cadd(cadd(acc, a1), cadd(a2, a3));
                        }
                        xput(slot + a.oGR + (size_t)l * b2 + kp, acc);
                        xarm(slot_arm + a.oGR + (size_t)l * b2 + kp);
                    }
                }
                __syncwarp();
                if (lane == 0) HP_STAMP(1);
                HP_TICK(1);
                if (DBG && tid == 0) {          // in-situ latencies: one L2 round trip of an exchange word, one smem load
                    cplx tmp;
                    long long c0_ = clock64();
                    bool okk = xtry(slot + a.oGR + (size_t)l * b2, tmp);
                    long long c1_ = clock64();
                    if (okk && tmp.x == 12345.678) tacc[7] += 1;
                    tacc[5] += c1_ - c0_;
                    volatile cplx* vp = xlr_c;
                    long long c2_ = clock64();
                    double q_ = vp[0].x;
                    long long c3_ = clock64();
                    if (q_ == 12345.678) tacc[7] += 1;
                    tacc[6] += c3_ - c2_;
                    tprev = clock64();
                }
                if (lane == 0 && it + 2 < nsteps)
                    ring_fill(ringM + (it & 1) * m_st, m_base + (size_t)(m + 2 * step - a.m_lo) * m_stride, m_bytes, &mbar[4 + (it & 1)]);
            }
            // ---- C2: separator right-hand sides and own rows of x_S
            if (nrows > 0) {
                for (int e0 = ctid; e0 < NS; e0 += 4 * HP2_CRIT) {          // up to 4 entries per thread in flight
                    const cplx *pa[4], *pc[4], *pv[4];
                    cplx va[4], vc[4], vs[4];
                    bool need_v[4], have[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int e = e0 + i * HP2_CRIT;
                        have[i] = e < NS;
                        int j = rho_j0, kap = rho_k0;
                        if (e != ctid && have[i]) { j = e / b; kap = e - j * b; }
                        pa[i] = slot + a.oGR + ((size_t)j * 2 + 1) * b + kap;          // Gl of leaf j
                        pc[i] = slot + a.oGR + ((size_t)(j + 1) * 2) * b + kap;        // Gf of leaf j+1
                        pv[i] = slot + a.oVS + j;
                        need_v[i] = have[i] && kap == b - 1;
                        vs[i] = cmake(0.0, 0.0);
                    }
                    unsigned int spins = 0;
                    for (;;) {
                        bool ok = true;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (have[i]) { ok = xtry(pa[i], va[i]) && ok; ok = xtry(pc[i], vc[i]) && ok; }
                            if (need_v[i]) ok = xtry(pv[i], vs[i]) && ok;
                        }
                        if (ok || *dead) break;
                        if (++spins > HP_SPIN_LIMIT) { hp_raise_abort(abort_flag); *dead = 1u; break; }
                        if ((spins & 0xFFF) == 0 && *((volatile unsigned int*)abort_flag)) { *dead = 1u; break; }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (have[i]) rho[e0 + i * HP2_CRIT] = csub(vs[i], cadd(va[i], vc[i]));
                }
                HP_TICK(2);
                bar_crit();
                if (ctid == 0) HP_STAMP(2);
                mbar_wait(&mbar[2 + (it & 1)], (it >> 1) & 1);
                HP_TICK(3);
                const cplx* Np = reinterpret_cast<const cplx*>(ringN + (it & 1) * n_st);
                for (int o = 0; cw + 4 * o < nrows; ++o) {
                    const int rr = cw + 4 * o;
                    cplx acc = cmake(0.0, 0.0), a1 = cmake(0.0, 0.0), a2 = cmake(0.0, 0.0), a3 = cmake(0.0, 0.0);
                    const cplx* nr = Np + (size_t)rr * NSP;
                    int e = lane;
                    for (; e + 96 < NS; e += 128) {
                        acc = cfma(nr[e], rho[e], acc);
                        a1 = cfma(nr[e + 32], rho[e + 32], a1);
                        a2 = cfma(nr[e + 64], rho[e + 64], a2);
                        a3 = cfma(nr[e + 96], rho[e + 96], a3);
                    }
                    for (; e < NS; e += 32) acc = cfma(nr[e], rho[e], acc);
                    acc = hp_warp_sum2(cadd(cadd(acc, a1), cadd(a2, a3)));           // every lane holds the row sum
                    if (lane == o) {
                        xput(slot + a.oXS + row0 + rr, acc);
                        xarm(slot_arm + a.oXS + row0 + rr);
                        if (sep_col >= 0) {                                             // y_s = x_s[b-1]
                            if (a_mode == 2) a.yout[sep_col] = acc;
                            else if (a_mode == 0) {
                                cplx un = cfms(cmul(rfac, cis1s), acc, usep);
                                a.u[(size_t)m * n + sep_col] = un;
                                if (more) xput(slot_next + a.oVS + sep_j, un);
                            } else {
                                cplx un = a_diag == 0 ? csub(usbase, acc) : acc;
                                a.u[(size_t)(m - 1) * n + sep_col] = un;
                                if (more) xput(slot_next + a.oVS + sep_j, cfma(cscale(sg, cmul(rfac, cis1s)), un, usep));
                                usbase = usep;
                            }
                            xarm(slot_prev + a.oVS + sep_j);
                        }
                    }
                }
                bar_crit();
                if (ctid == 0) HP_STAMP(3);
                HP_TICK(4);
                if (ctid == 0 && it + 2 < nsteps)
                    ring_fill(ringN + (it & 1) * n_st, pk_base + (size_t)(m + 2 * step - a.m_lo) * strip_stride + a.lay.offN, n_bytes,
                              &mbar[2 + (it & 1)]);
            }
        }
        if (DBG && tid == 0)
            for (int i = 0; i < 8; ++i) a.dbg[(size_t)g * 16 + i] = tacc[i];
    } else {
        // =====================================================================================================
        // off-path group
        // =====================================================================================================
        const int ot = tid - HP2_CRIT, ow = ot >> 5;
        long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
        if (ot == 0) {
            for (int sidx = 0; sidx < 2 && sidx < nsteps; ++sidx)
                ring_fill(ringWG + sidx * wg_st, pk_base + (size_t)(m0 + sidx * step - a.m_lo) * strip_stride, wg_bytes, &mbar[sidx]);
            if (nsteps > 2) {
                const char* src = (const char*)(pk_base + (size_t)(m0 + 2 * step - a.m_lo) * strip_stride);
                unsigned int pk_bytes = (unsigned int)(a.lay.PK * sizeof(cplx));
                for (unsigned int o = 0; o < pk_bytes; o += HP_BULK_CHUNK) bulk_prefetch_l2(src + o, min(HP_BULK_CHUNK, pk_bytes - o));
            }
        }
        const bool col = ot < ncols;
        const int c = c0 + ot;
        const cplx cis1 = col ? a.is1t[2 * (c + 1)] : cmake(0.0, 0.0);
        // per-column state: vbr = vb(t), y0prev = y0(t-1), coefc = multiplier of the correction in v(t),
        // ubase = (backward) original value of the row strip t-1 overwrites, ucur = field value combined in b(t)
        cplx vbr = cmake(0.0, 0.0), y0prev = cmake(0.0, 0.0), coefc = cmake(0.0, 0.0), ubase = cmake(0.0, 0.0),
             ubase_prev = cmake(0.0, 0.0);
        if (col) {
            if (a_mode == 2) vbr = a.vin[c];
            else if (a_mode == 0) vbr = ldcg(a.u + (size_t)(m0 - 1) * n + c);
            else {
                ubase = ldcg(a.u + (size_t)(m0 - 1) * n + c);
                vbr = ubase;
                if (m0 < n) vbr = cfma(cscale(sg, cmul(hp_rowfac(a, m0), cis1)), ldcg(a.u + (size_t)m0 * n + c), vbr);
            }
            vb[ot] = vbr;
        }
        bar_off();

        for (int it = 0; it <= nsteps; ++it) {
            const int m = m0 + it * step, mn = m + step, mp = m - step;     // this, next, previous strip
            const bool live = it < nsteps, more = it + 1 < nsteps;
            cplx* slot = a.xch + (size_t)(it & (HP_RING - 1)) * a.slot_stride;
            cplx* slot_prev = a.xch + (size_t)((it + 3) & (HP_RING - 1)) * a.slot_stride;
            cplx* slot_arm = a.xch + (size_t)((it + 2) & (HP_RING - 1)) * a.slot_stride;
            const cplx* pk = reinterpret_cast<const cplx*>(ringWG + (it & 1) * wg_st);
            const cplx* Wp = pk;
            const cplx* Gp = pk + a.lay.offG;
            // early loads: the row coupling and the field value vb(t+1) is built from (written by no other thread)
            const cplx rf_it = hp_rowfac(a, a_mode == 1 ? mn : m);
            cplx unx = cmake(0.0, 0.0);
            if (col && live) {
                if (a_mode == 0) unx = ldcg(a.u + (size_t)m * n + c);
                else if (a_mode == 1 && more) unx = ldcg(a.u + (size_t)(mn - 1) * n + c);
            }
            if (DBG && ot == 0) tprev = clock64();
            // ---- a: gb(t) = Gc(t) vb(t)
            if (live) {
                mbar_wait(&mbar[it & 1], (it >> 1) & 1);
                HP_TICK(0);
                for (int kap = lane; kap < b2; kap += 32) {
                    cplx acc = cmake(0.0, 0.0);
                    const cplx* gr = Gp + (size_t)kap * CW;
                    for (int cc = ow; cc < ncols; cc += 8) acc = cfma(gr[cc], vb[cc], acc);
                    gpw[ow * b2 + kap] = acc;
                }
                bar_off();
                if (ot < b2) {
                    cplx acc = gpw[ot];
#pragma unroll
                    for (int w = 1; w < 8; ++w) acc = cadd(acc, gpw[w * b2 + ot]);
                    xput(slot + a.oGP + (size_t)g * b2 + ot, acc);
                    xarm(slot_arm + a.oGP + (size_t)g * b2 + ot);
                }
            }
            HP_TICK(1);
            // ---- b: x(t-1) arrives: finish strip t-1 on the own columns, input of strip t
            cplx v = vbr;
            if (it > 0) {
                if (ot < b2) {
                    int side = ot / b, kap = ot - side * b, j = l - 1 + side;
                    xlr_o[ot] = (j >= 0 && j < P - 1) ? xwait(slot_prev + a.oXS + (size_t)j * b + kap, abort_flag, dead) : cmake(0.0, 0.0);
                }
                bar_off();
                HP_TICK(2);
                if (col) {
                    cplx corr = cmake(0.0, 0.0), c1 = cmake(0.0, 0.0), c2 = cmake(0.0, 0.0), c3 = cmake(0.0, 0.0);
                    int kap = 0;
#pragma unroll 2
                    for (; kap + 3 < b2; kap += 4) {
                        corr = cfma(Gprev[(size_t)kap * CW + ot], xlr_o[kap], corr);
                        c1 = cfma(Gprev[(size_t)(kap + 1) * CW + ot], xlr_o[kap + 1], c1);
                        c2 = cfma(Gprev[(size_t)(kap + 2) * CW + ot], xlr_o[kap + 2], c2);
                        c3 = cfma(Gprev[(size_t)(kap + 3) * CW + ot], xlr_o[kap + 3], c3);
                    }
                    for (; kap < b2; ++kap) corr = cfma(Gprev[(size_t)kap * CW + ot], xlr_o[kap], corr);
                    corr = cadd(cadd(corr, c1), cadd(c2, c3));
                    v = cfma(coefc, corr, vbr);
                    if (a_mode == 2) a.yout[c] = csub(y0prev, corr);
                    else if (a_mode == 0) a.u[(size_t)mp * n + c] = v;                      // row m_{t-1}: final
                    else {
                        cplx un = a_diag == 0 ? cadd(csub(ubase_prev, y0prev), corr) : csub(y0prev, corr);
                        a.u[(size_t)(mp - 1) * n + c] = un;
                    }
                }
            }
            if (!live) break;
            HP_TICK(3);
            if (col) {
                v_leaf[lc0 + ot] = v;
                if (K > 1) { xput(slot + c0 + ot, v); xarm(slot_arm + c0 + ot); }
            }
            // ---- c: gather the leaf's input, leaf product y0(t) = W(t) v_leaf(t), vb(t+1)
            for (int cc = ot; cc < q; cc += HP2_OFF)
                if (cc < lc0 || cc >= lc1) v_leaf[cc] = xwait(slot + ls + cc, abort_flag, dead);
            bar_off();
            HP_TICK(4);
            // keep Gc(t) for the correction of the next strip (every read of the old copy is behind the barrier)
            for (int e = ot; e < b2 * CW; e += HP2_OFF) Gprev[e] = Gp[e];
            for (int cc = lane; cc < ncols; cc += 32) {
                cplx acc = cmake(0.0, 0.0), a1 = cmake(0.0, 0.0), a2 = cmake(0.0, 0.0), a3 = cmake(0.0, 0.0);
                const cplx* wr = Wp + (size_t)cc * QP;
                int cq = ow;
#pragma unroll 2
                for (; cq + 24 < q; cq += 32) {
                    acc = cfma(wr[cq], v_leaf[cq], acc);
                    a1 = cfma(wr[cq + 8], v_leaf[cq + 8], a1);
                    a2 = cfma(wr[cq + 16], v_leaf[cq + 16], a2);
                    a3 = cfma(wr[cq + 24], v_leaf[cq + 24], a3);
                }
                for (; cq < q; cq += 8) acc = cfma(wr[cq], v_leaf[cq], acc);
                y0w[(size_t)ow * CW + cc] = cadd(cadd(acc, a1), cadd(a2, a3));
            }
            bar_off();                                   // every thread of the group is done with the stage
            HP_TICK(5);
            if (ot == HP2_OFF - 32) {
                if (it + 2 < nsteps) ring_fill(ringWG + (it & 1) * wg_st, pk_base + (size_t)(m + 2 * step - a.m_lo) * strip_stride, wg_bytes, &mbar[it & 1]);
                if (it + 3 < nsteps) {
                    const char* src = (const char*)(pk_base + (size_t)(m + 3 * step - a.m_lo) * strip_stride);
                    unsigned int pk_bytes = (unsigned int)(a.lay.PK * sizeof(cplx));
                    for (unsigned int o = 0; o < pk_bytes; o += HP_BULK_CHUNK) bulk_prefetch_l2(src + o, min(HP_BULK_CHUNK, pk_bytes - o));
                }
            }
            if (col) {
                cplx y0 = y0w[ot];
#pragma unroll
                for (int w = 1; w < 8; ++w) y0 = cadd(y0, y0w[(size_t)w * CW + ot]);
                y0prev = y0;
                if (a_mode == 0) {
                    coefc = cmul(rf_it, cis1);                                 // A_{m+1,m}
                    vbr = cfms(coefc, y0, unx);                                // u_{m+1} - coef y0
                } else if (a_mode == 1) {
                    coefc = cmul(rf_it, cis1);                                 // A_{m-1,m}
                    vbr = a_diag == 0 ? cfma(coefc, csub(ubase, y0), unx) : cfms(coefc, y0, unx);
                    ubase_prev = ubase;
                    ubase = unx;
                }
                vb[ot] = vbr;
            }
            bar_off();                                   // vb, y0w free for the next strip
            HP_TICK(6);
        }
        if (DBG && ot == 0)
            for (int i = 0; i < 8; ++i) a.dbg[(size_t)g * 16 + 8 + i] = tacc[i];
    }
}

size_t hp_sweep2_smem(const HpLayout& L, int b) {
    size_t wg = (L.offN * sizeof(cplx) + 127) & ~(size_t)127, nn = ((size_t)L.NR * L.NSP * sizeof(cplx) + 127) & ~(size_t)127,
           mm = ((size_t)4 * b * b * sizeof(cplx) + 127) & ~(size_t)127;
    size_t small = sizeof(cplx) * ((size_t)2 * b * L.CW + L.CW + L.QP + 8 * (size_t)L.CW + 8 * (size_t)2 * b + 4 * b + L.NSP) + 6 * 8 + 16;
    return 2 * wg + 2 * nn + 2 * mm + small;
}

int hp_sweep2_launch(hp_solver* s, HpSweepArgs& a, cudaStream_t st) {
    const HpLayout& L = s->lay;
    size_t smem = hp_sweep2_smem(L, s->b);
    const int mode = a.mode == 0 ? 0 : (a.mode == 2 ? 3 : (a.diag_mode == 0 ? 1 : 2));
    const void* fns[2][4] = {{(const void*)hp_sweep2_kernel<0, false>, (const void*)hp_sweep2_kernel<1, false>,
                              (const void*)hp_sweep2_kernel<2, false>, (const void*)hp_sweep2_kernel<3, false>},
                             {(const void*)hp_sweep2_kernel<0, true>, (const void*)hp_sweep2_kernel<1, true>,
                              (const void*)hp_sweep2_kernel<2, true>, (const void*)hp_sweep2_kernel<3, true>}};
    const void* fn = fns[a.dbg ? 1 : 0][mode];
    HP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&a};
    if (s->coop) HP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(L.G), dim3(HP2_THREADS), args, smem, st));
    else HP_CUDA(cudaLaunchKernel(fn, dim3(L.G), dim3(HP2_THREADS), args, smem, st));   // contexts: see hp_context_clone
    return 0;
}
