// Fused per-subject pass, third generation (M <= 62, at most 8 * NT rows per group, NT = 3 or 5): sm_100a, FP64 DMMA.8x8x4.
//
// Measured on B200 (profiles/r02_fp64_pipe_sharing.txt): DMMA and DFMA share ONE FP64 pipe (37 TFLOP/s of DMMA, 33 of DFMA, and
// any mix of the two adds up to ~34), and the second generation of this kernel spent more shared-memory wavefronts (7.7k per
// subject and latent, 128 B per clock and SM) than FP64-pipe cycles.  This generation is laid out around those two limits:
//
//   * TRANSPOSED register layout.  Warp w owns the inducing columns j = 8w .. 8w+7 for the whole kernel; every T x M matrix
//     of a row group lives as its transpose in DMMA accumulator layout: lane (g, q) holds X^T[8w + g][8 tt + 2q + e].
//     A product whose contraction runs over the ROWS t of the group takes those registers directly as A fragments (the k
//     index of m8n8k4 may be permuted freely as long as A and B agree: DMMA e uses k = 8 tt + 2q + e), so
//         U^T = Kxz^T L^-T,   V^T = U^T L^-1      (A = own registers, B = 16-byte loads of the small T x T factor)
//         S  += U^T U           (A = own registers, B = the other warps' U^T rows; diagonal tile: own registers for both;
//                                every unordered tile pair is computed once: 4-5 tiles per warp)
//     S = U^T U rather than Kxz^T V on purpose: the products U[t][i] U[t][j] commute, so S is bitwise symmetric AND its rows
//     for inducing points with identical K0 covariates are bitwise equal however the tiles are dealt to warps.  The
//     reference's inducing points are data rows, many of them K0-duplicates; Kzz^-1 then holds +-1/(2 eps) pairs and
//     Kzz^-1 S Kzz^-1 is exact only up to how well those rows agree (measured: 2 x the error of grad_m with Kxz^T V);
//   * W = c (G - Kzz^-1) is held as A fragments in registers for the whole kernel (16 doubles): Y^T = W^T V^T costs one
//     8-byte shared load per DMMA instead of two;
//   * inducing columns M and M+1.. of the 64-column tiles are padding; two of them carry mu and r = Kxz a - mu through the
//     same products, so that  u = B^-1 r,  ng1 = sum Kxz^T B^-1 mu,  da = sum Kxz^T B^-1 r  and  A = sum r^T B^-1 r  come out
//     of V^T and of the S accumulators without any extra reduction (hence M <= 62);
//   * one CTA = 8 warps, two CTAs per SM (<= 128 registers), no warp sets: W, a and the inducing covariates are per-thread;
//   * groups that hold ONE subject filling all NT row tiles (every group of a fixed-T minibatch) run a fully unrolled,
//     predicate-free instance of the group body.
//
// Per group (whole subjects, <= 8 NT rows; plan by k_plan_groups3):
//   B0  wait for this group's cp.async data (gathered covariates, mu, block-diagonal L^-1 and L^-T, zero filled)
//   J1  Kxz^T in registers from the covariates ; un-scaled SE component values -> FC (per-thread slots) ; products with a ->
//       row sums of this warp ; warp 7 assembles r and takes mu, r as its columns 62, 63
//   J2  U^T (DMMA, lower-triangular tile range) -> registers + smem ;  J3  V^T (upper range) -> registers + smem ; u
//   B2  __syncthreads ; issue the prefetch of the next group
//   J4  S += U^T U ;  Y^T = W^T V^T
//   J5  adjoint of Kxz = 2c u a^T + 2Y against d k_c / d theta ; d_mu ; Y^T -> smem
//   B3  __syncthreads
//   J6  Q = Y V^T on the subject-diagonal upper tiles ; adjoint of B_p = -(c u u^T + Q) against d K1 / d theta and the noise
// L^-1 and L^-T rows come from k_prep3 (row-major per row, zero padded to TP).  Nothing of size T x M touches HBM.
#include <type_traits>

#include "lvae_kld.h"

namespace {

constexpr int GT = LVAE_F2_GT;     // ints per group-plan entry: row0, R, nsub, end_1 .. end_5
constexpr int CS = 4;              // covariate slots per component: SE column, up to 3 mask columns
constexpr int NWARP = 8;
constexpr int NTHR = 256;
constexpr int COL_MU = 62, COL_R = 63;
template <int V> using ic = std::integral_constant<int, V>;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// asynchronous copies; !valid writes zeros (src-size 0) and never forms an out-of-range address (falls back to `safe`)
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid, const void* safe) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    const void* src = valid ? gmem : safe;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid, const void* safe) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    const void* src = valid ? gmem : safe;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(NTHR) : "memory"); }
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NTHR) : "memory"); }
__device__ __forceinline__ void tri2(int e, int& i, int& j) {   // e-th element of a lower triangle, j <= i
    i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    j = e - i * (i + 1) / 2;
}

// shared-memory map (doubles), everything at compile-time offsets except FC (sized by the number of SE-bearing K0 components)
template <int NC0, int NC1, int NT>
struct Smem {
    static constexpr int RG = 8 * NT;
    static constexpr int LDT = RG;          // L^-1, L^-T      row stride == 8 (mod 16): conflict-free 16-byte row reads
    static constexpr int LDU = RG;          // U^T [64][LDU]   same pattern (B fragments of S); also the a-products of J1
    static constexpr int LDC = RG + 4;      // V^T, Y^T [64][LDC]  == 12 (mod 16): conflict-free 8-byte column reads
    static constexpr int NCT = NC0 + NC1;
    static constexpr int XCSZ = NCT * CS * RG;
    static constexpr int oLinv = 0;
    static constexpr int oLinvT = oLinv + RG * LDT;
    static constexpr int oUT = oLinvT + RG * LDT;
    static constexpr int oVT = oUT + 64 * LDU;
    static constexpr int oYT = oVT + 64 * LDC;
    static constexpr int oXC = oYT + 64 * LDC;                 // [2][XCSZ]
    static constexpr int oMus = oXC + 2 * XCSZ;                // [RG]
    static constexpr int oRpart = oMus + RG;                   // [NWARP][RG]
    static constexpr int oRs = oRpart + NWARP * RG;            // [RG]
    static constexpr int oUs = oRs + RG;                       // [RG]
    static constexpr int oZC = oUs + RG;                       // [NC0 * CS][64]
    static constexpr int oInts = oZC + NC0 * CS * 64;          // ints: meta[3][GT], lo[3][RG], hi[3][RG], dimtab[NCT*CS]
    static constexpr int nInts = 3 * GT + 6 * RG + NCT * CS;
    static constexpr int oFC = oInts + (nInts + 1) / 2 + (((nInts + 1) / 2) & 1);   // 16-byte aligned
    static constexpr int fcPerComp = NT * NTHR * 2;            // [NT][NTHR] double2
    __host__ __device__ static constexpr size_t doubles(int nr) { return (size_t)oFC + (size_t)nr * fcPerComp; }
};

template <int NC0, int NC1, int NT>
__global__ void __launch_bounds__(NTHR, NT <= 3 ? 2 : 1)
k_subjects_fused3(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int N_b, int TP,
                  const double* __restrict__ x, const double* __restrict__ mu, const double* __restrict__ z,
                  const double* __restrict__ ls, const double* __restrict__ os, double c, double* __restrict__ d_mu,
                  double* __restrict__ ws) {
    using S_ = Smem<NC0, NC1, NT>;
    constexpr int RG = S_::RG, LDT = S_::LDT, LDU = S_::LDU, LDC = S_::LDC, NCT = S_::NCT, XCSZ = S_::XCSZ;
    extern __shared__ __align__(16) double sm[];
    __shared__ double hil2[LVAE_MAXC], il3[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    __shared__ double hyp[NWARP][2 * LVAE_MAXC + 2];
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nh = hyp_count(sp), MM = M * M;
    const int j = 8 * wid + g;                       // this thread's inducing column (row of every transposed matrix)
    const bool cv = j < M;

    double* const Linv = sm + S_::oLinv;
    double* const LinvT = sm + S_::oLinvT;
    double* const UTs = sm + S_::oUT;
    double* const VTs = sm + S_::oVT;
    double* const YTs = sm + S_::oYT;
    double* const XC = sm + S_::oXC;
    double* const mus = sm + S_::oMus;
    double* const rpart = sm + S_::oRpart;
    double* const rs = sm + S_::oRs;
    double* const us = sm + S_::oUs;
    double* const ZC = sm + S_::oZC;
    int* const meta = reinterpret_cast<int*>(sm + S_::oInts);  // [3][GT]
    int* const lo_r = meta + 3 * GT;                            // [3][RG]  first row of the subject that owns row t
    int* const hi_r = lo_r + 3 * RG;                            // [3][RG]  one past its last row
    int* const dimtab = hi_r + 3 * RG;                          // [NCT * CS]
    double2* const FC = reinterpret_cast<double2*>(sm + S_::oFC);

    // ---- per-CTA constants -------------------------------------------------------------------------------------------
    if (tid < sp.n_ls) { const double v = ls[(size_t)tid * L + l]; hil2[tid] = 0.5 / (v * v); il3[tid] = 1.0 / (v * v * v); }
    if (tid < sp.n0 + sp.n1) osc[tid] = os[(size_t)tid * L + l];
    load_exp_table(etab);
    if (tid < NCT * CS) {
        const int cc = tid / CS, sl = tid % CS;
        int dim = -1;
        if (sl == 0) dim = sp.rbf_dim[cc];
        else if (sl - 1 < sp.n_mask[cc]) dim = sp.mask_dim[cc][sl - 1];
        dimtab[tid] = dim;
    }
    __syncthreads();
    for (int e = tid; e < NC0 * CS * 64; e += NTHR) {
        const int sl = e >> 6, jj = e & 63, dim = dimtab[sl];
        ZC[e] = (dim >= 0 && jj < M) ? z[((size_t)l * M + jj) * Q + dim] : 0.0;
    }
    // W^T as A fragments: wf[ks] = W[4 ks + q][8 wid + g]   (zero outside M x M, so the padding columns never reach Y)
    double wf[16];
    {
        const double* Wl = ws + w.W + (size_t)l * MM;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) wf[ks] = (cv && 4 * ks + q < M) ? Wl[(size_t)(4 * ks + q) * M + j] : 0.0;
    }
    const double aj = cv ? ws[w.a + (size_t)l * M + j] : 0.0;
    const double caj = c * aj;
    const int nks = (M + 3) >> 2;                    // k steps of Y^T / Q that touch real columns

    const int* gtab = reinterpret_cast<const int*>(ws + w.gtab) + (size_t)chunk * w.gstride * GT;
    const int ngroups = reinterpret_cast<const int*>(ws + w.gcount)[chunk];
    const double* Lrows = ws + w.Lrows + (size_t)l * N_b * TP;      // rows of the per-subject L^-1 (lower triangular)
    const double* Ltrows = ws + w.Ltrows + (size_t)l * N_b * TP;    // rows of the per-subject L^-T (upper triangular)

    // S tiles of this warp: (wid, (wid + d) & 7), d = 0 .. 3, and d = 4 for wid < 4 — every unordered tile pair exactly once.
    // TWO-LEVEL sum: the register accumulators are flushed into CTA-private second-level accumulators (global memory, L2
    // resident) every FLUSH groups.  One register chain over the ~110 groups of a CTA loses ~110 eps / 2 relative accuracy in
    // S, and Kzz^-1 S Kzz^-1 magnifies the last bits of S by ~1e10 (cond(Kzz) ~ 1e8): measured on a host restatement,
    // grad_m is 2.7e-6 from the exact value with chains of 112 subjects, 6.4e-7 with this two-level sum, 3e-5 with one chain of
    // 1000.
    constexpr int FLUSH = 16;
    double sacc[5][2];
#pragma unroll
    for (int d = 0; d < 5; ++d) sacc[d][0] = sacc[d][1] = 0.0;
    // (the pointer is re-formed at every use: it is needed once per FLUSH groups and must not hold two registers in between)
    auto acc2_ptr = [&]() {
        return reinterpret_cast<double2*>(ws + w.acc2 + ((size_t)blockIdx.x * L + blockIdx.y) * LVAE_F3_ACC2) + threadIdx.x;
    };
    {
        double2* acc2 = acc2_ptr();
#pragma unroll
        for (int d = 0; d < 5; ++d) acc2[d * NTHR] = make_double2(0.0, 0.0);
    }
    double gos[NC0], gls[NC0], g1os[NC1], g1ls[NC1], gno = 0.0;
#pragma unroll
    for (int cc = 0; cc < NC0; ++cc) gos[cc] = gls[cc] = 0.0;
#pragma unroll
    for (int k = 0; k < NC1; ++k) g1os[k] = g1ls[k] = 0.0;

    // row -> [lo, hi) of its subject inside a group, from the group's plan entry
    auto row_block = [&](const int* mt_, int t, int R, int& lo, int& hi) {
        lo = t; hi = t;
        if (t < R) {
            lo = 0; hi = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) { const int end = mt_[3 + s]; if (t >= end) lo = end; }
#pragma unroll
            for (int s = 4; s >= 0; --s) { const int end = mt_[3 + s]; if (t < end) hi = end; }
        }
    };
    auto plan_rows = [&](int slot) {                 // threads 0 .. RG-1: row blocks of the group whose plan entry sits in `slot`
        if (tid < RG) {
            const int* mt_ = meta + slot * GT;
            int lo, hi;
            row_block(mt_, tid, mt_[1], lo, hi);
            lo_r[slot * RG + tid] = lo;
            hi_r[slot * RG + tid] = hi;
        }
    };
    auto issue_meta = [&](int gi, int slot) {
        if (tid < 2) cp_async16(meta + slot * GT + 4 * tid, gtab + (size_t)gi * GT + 4 * tid, true, gtab);
    };
    // prefetch of a group: plan entry and row blocks visible in `slot`; covariates go to XC buffer `buf`
    auto issue_data = [&](int slot, int buf) {
        const int* mt_ = meta + slot * GT;
        const int row0 = mt_[0], R = mt_[1];
        const int* lo_ = lo_r + slot * RG;
        const int* hi_ = hi_r + slot * RG;
        double* xc = XC + buf * XCSZ;
#pragma unroll
        for (int rep = 0; rep < (XCSZ + NTHR - 1) / NTHR; ++rep) {
            const int e = tid + NTHR * rep;
            if (e < XCSZ) {
                const int sl = e / RG, t = e - sl * RG, dim = dimtab[sl];
                cp_async8(xc + e, x + (size_t)(row0 + t) * Q + dim, (t < R) && (dim >= 0), x);
            }
        }
        if (tid < RG) cp_async8(mus + tid, mu + (size_t)(row0 + tid) * L + l, tid < R, x);
        // L^-1 and L^-T, block diagonal: Linv[t][lo + k] = Lrows[row0 + t][k] (same for L^-T); 16-byte pieces when the subject
        // starts on an even row.  Both exports are zero outside their triangle, so whole rows of the block are copied.
#pragma unroll
        for (int rep = 0; rep < (RG * RG + NTHR - 1) / NTHR; ++rep) {
            const int e = tid + NTHR * rep;
            if (e < RG * RG) {
                const int which = e / (RG * RG / 2), e2 = e - which * (RG * RG / 2);
                const int t = e2 / (RG / 2), k = 2 * (e2 - t * (RG / 2));
                const int lo = lo_[t], hi = hi_[t];
                const double* src = (which ? Ltrows : Lrows) + (size_t)(row0 + t) * TP + (k - lo);
                double* dst = (which ? LinvT : Linv) + t * LDT + k;
                if ((lo & 1) == 0) {
                    cp_async16(dst, src, (k >= lo) && (k < hi), x);
                } else {
                    cp_async8(dst, src, (k >= lo) && (k < hi), x);
                    cp_async8(dst + 1, src + 1, (k + 1 >= lo) && (k + 1 < hi), x);
                }
            }
        }
    };

    // ---- prologue: plan entries of groups 0 and 1, data of group 0 ---------------------------------------------------
    if (ngroups > 0) {
        issue_meta(0, 0);
        if (ngroups > 1) issue_meta(1, 1);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        plan_rows(0);
        __syncthreads();
        issue_data(0, 0);
        cp_async_commit();
    }

    for (int gi = 0; gi < ngroups; ++gi) {
        const int buf = gi & 1, slot = gi % 3;
        // ---- B0 ------------------------------------------------------------------------------------------------------
        cp_async_wait_all();
        __syncthreads();
        const int* mt_ = meta + slot * GT;
        const int row0 = mt_[0], R = mt_[1];
        const double* xc = XC + buf * XCSZ;
        const double2* xc2 = reinterpret_cast<const double2*>(xc);
        const int* lo_ = lo_r + slot * RG;
        const int* hi_ = hi_r + slot * RG;
        const bool more = gi + 1 < ngroups;
        if (more) plan_rows((gi + 1) % 3);           // next group's row blocks (its plan entry arrived with this group's data)

        // FULL: one subject that reaches into the last row tile — no tile predicates, no block-diagonal trimming
        auto body = [&](auto FULLT) {
            constexpr bool FULL = decltype(FULLT)::value;
            const int nmt = FULL ? NT : ((R + 7) >> 3);
            // ---- J1: Kxz^T from the gathered covariates ; SE-bearing f_c -> FC ; products with a ---------------------
            double kx[NT][2];
#pragma unroll
            for (int tt = 0; tt < NT; ++tt) kx[tt][0] = kx[tt][1] = 0.0;
            {
                int fslot = 0;
                // one component, NM = number of mask factors known at compile time (3 = any, checked at run time)
                auto comp = [&](auto NMT, int cc) {
                    constexpr int NM = decltype(NMT)::value;
                    const double o = osc[cc];
                    const bool rbf = sp.rbf_dim[cc] >= 0;
                    const int nm = sp.n_mask[cc];
                    double zs[NM > 0 ? NM : 1], tg[NM > 0 ? NM : 1];    // x (+/-) z == target  <=>  categorical / binary factor is 1
#pragma unroll
                    for (int i = 0; i < NM; ++i) {
                        const bool cat = sp.mask_type[cc][i] == LVAE_CAT;
                        const double zv = ZC[(cc * CS + 1 + i) * 64 + j];
                        zs[i] = cat ? -zv : zv;
                        tg[i] = cat ? 0.0 : 2.0;
                    }
                    const double zr = ZC[(cc * CS) * 64 + j];
                    const double nh_ = rbf ? -hil2[sp.ls_idx[cc]] : 0.0;
                    double2* fc = FC + (size_t)fslot * NT * NTHR + tid;
#pragma unroll
                    for (int tt = 0; tt < NT; ++tt) {
                        if (FULL || tt < nmt) {
                            const int t0 = 8 * tt + 2 * q;
                            bool on0 = cv, on1 = cv;
                            if (!FULL || tt == NT - 1) { on0 = on0 && (t0 < R); on1 = on1 && (t0 + 1 < R); }
#pragma unroll
                            for (int i = 0; i < NM; ++i) {
                                if (NM < 3 || i < nm) {
                                    const double2 a = xc2[((cc * CS + 1 + i) * RG + t0) >> 1];
                                    on0 = on0 && (a.x + zs[i] == tg[i]);
                                    on1 = on1 && (a.y + zs[i] == tg[i]);
                                }
                            }
                            double f0 = on0 ? 1.0 : 0.0, f1 = on1 ? 1.0 : 0.0;
                            if (rbf) {
                                const double2 a = xc2[((cc * CS) * RG + t0) >> 1];
                                const double d0 = a.x - zr, d1 = a.y - zr;
                                const double e0 = exp_neg((d0 * d0) * nh_, etab), e1 = exp_neg((d1 * d1) * nh_, etab);
                                f0 = on0 ? e0 : 0.0;
                                f1 = on1 ? e1 : 0.0;
                                fc[tt * NTHR] = make_double2(f0, f1);
                            }
                            kx[tt][0] = fma(o, f0, kx[tt][0]);
                            kx[tt][1] = fma(o, f1, kx[tt][1]);
                        }
                    }
                    if (rbf) ++fslot;
                };
#pragma unroll 1
                for (int cc = 0; cc < NC0; ++cc) {
                    const int nm = sp.n_mask[cc];
                    if (nm == 0) comp(ic<0>{}, cc);
                    else if (nm == 1) comp(ic<1>{}, cc);
                    else comp(ic<3>{}, cc);
                }
            }
            // rho = Kxz a: products into this warp's rows of the U^T buffer, column sums of the warp, then warp 7 adds the warps
            {
                double2* P2 = reinterpret_cast<double2*>(UTs + j * LDU);
#pragma unroll
                for (int tt = 0; tt < NT; ++tt) P2[4 * tt + q] = make_double2(kx[tt][0] * aj, kx[tt][1] * aj);
                __syncwarp();
                for (int t = lane; t < RG; t += 32) {
                    double s = 0.0;
#pragma unroll
                    for (int gg = 0; gg < 8; ++gg) s += UTs[(8 * wid + gg) * LDU + t];
                    rpart[wid * RG + t] = s;
                }
                __syncwarp();
            }
            if (wid < NWARP - 1) {
                bar_arrive(1);
            } else {
                bar_sync(1);                         // the other warps' column sums are in rpart
                for (int t = lane; t < RG; t += 32) {
                    double s = -mus[t];
#pragma unroll
                    for (int ww = 0; ww < NWARP; ++ww) s += rpart[ww * RG + t];
                    rs[t] = t < R ? s : 0.0;
                }
                __syncwarp();
                if (g == 6) {                        // column 62 = mu, column 63 = r ride through the products below
                    const double2* m2 = reinterpret_cast<const double2*>(mus);
#pragma unroll
                    for (int tt = 0; tt < NT; ++tt) { const double2 v = m2[4 * tt + q]; kx[tt][0] = v.x; kx[tt][1] = v.y; }
                } else if (g == 7) {
                    const double2* r2 = reinterpret_cast<const double2*>(rs);
#pragma unroll
                    for (int tt = 0; tt < NT; ++tt) { const double2 v = r2[4 * tt + q]; kx[tt][0] = v.x; kx[tt][1] = v.y; }
                }
            }

            // ---- J2: U^T = Kxz^T L^-T   (tile (mt, nt) contributes iff mt <= nt and both touch the same subject) ------
            double ut[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                ut[nt][0] = ut[nt][1] = 0.0;
                if (FULL || nt < nmt) {
                    const int klo = FULL ? 0 : (lo_[8 * nt] >> 3);
                    const double2* Lr = reinterpret_cast<const double2*>(Linv + (8 * nt + g) * LDT) + q;
#pragma unroll
                    for (int mt = 0; mt < NT; ++mt) {
                        if (mt <= nt && (FULL || mt >= klo)) {
                            const double2 b = Lr[4 * mt];
                            dmma(ut[nt][0], ut[nt][1], kx[mt][0], b.x);
                            dmma(ut[nt][0], ut[nt][1], kx[mt][1], b.y);
                        }
                    }
                }
            }
            {
                double2* U2 = reinterpret_cast<double2*>(UTs + j * LDU);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) U2[4 * nt + q] = make_double2(ut[nt][0], ut[nt][1]);
            }
            // ---- J3: V^T = U^T L^-1   (mt >= nt) ----------------------------------------------------------------------
            double vt[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                vt[nt][0] = vt[nt][1] = 0.0;
                if (FULL || nt < nmt) {
                    const int khi = FULL ? NT : ((hi_[min(8 * nt + 7, R - 1)] + 7) >> 3);
                    const double2* Lr = reinterpret_cast<const double2*>(LinvT + (8 * nt + g) * LDT) + q;
#pragma unroll
                    for (int mt = 0; mt < NT; ++mt) {
                        if (mt >= nt && (FULL || mt < khi)) {
                            const double2 b = Lr[4 * mt];
                            dmma(vt[nt][0], vt[nt][1], ut[mt][0], b.x);
                            dmma(vt[nt][0], vt[nt][1], ut[mt][1], b.y);
                        }
                    }
                }
            }
            {
                double2* V2 = reinterpret_cast<double2*>(VTs + j * LDC);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) V2[4 * nt + q] = make_double2(vt[nt][0], vt[nt][1]);
                if (j == COL_R) {                    // u = B^-1 r
                    double2* u2 = reinterpret_cast<double2*>(us);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) u2[4 * nt + q] = make_double2(vt[nt][0], vt[nt][1]);
                }
            }
            __syncthreads();                         // B2: U^T, V^T, u complete ; L^-1 buffers free
            if (more) {
                issue_data((gi + 1) % 3, buf ^ 1);
                if (gi + 2 < ngroups) issue_meta(gi + 2, (gi + 2) % 3);
                cp_async_commit();
            }

            // ---- J4: S += U^T U ; Y^T = W^T V^T ------------------------------------------------------------------------
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                if (d < 4 || wid < 4) {
                    const int tj = (wid + d) & 7;
                    const double2* Ur = reinterpret_cast<const double2*>(UTs + (8 * tj + g) * LDU) + q;
#pragma unroll
                    for (int mt = 0; mt < NT; ++mt) {
                        if (FULL || mt < nmt) {
                            double2 b;
                            if (d == 0) b = make_double2(ut[mt][0], ut[mt][1]);
                            else b = Ur[4 * mt];
                            dmma(sacc[d][0], sacc[d][1], ut[mt][0], b.x);
                            dmma(sacc[d][0], sacc[d][1], ut[mt][1], b.y);
                        }
                    }
                }
            }
            double yacc[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) yacc[nt][0] = yacc[nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                if (ks < nks) {
                    const double* Vc = VTs + (4 * ks + q) * LDC + g;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        if (FULL || nt < nmt) dmma(yacc[nt][0], yacc[nt][1], wf[ks], Vc[8 * nt]);
                    }
                }
            }

            // ---- J5: adjoint of Kxz = 2 (c u a^T + Y) against the component derivatives ; Y^T -> smem ; d_mu ---------
            {
                double gb[NT][2];                    // c u[t] a[j] + Y[t][j]  (the factor 2 is applied once at the end)
                const double2* u2 = reinterpret_cast<const double2*>(us);
#pragma unroll
                for (int tt = 0; tt < NT; ++tt) {
                    const double2 u = u2[4 * tt + q];
                    gb[tt][0] = fma(u.x, caj, yacc[tt][0]);
                    gb[tt][1] = fma(u.y, caj, yacc[tt][1]);
                }
                int fslot = 0;
#pragma unroll
                for (int cc = 0; cc < NC0; ++cc) {
                    const bool rbf = sp.rbf_dim[cc] >= 0;
                    if (rbf) {
                        const double zr = ZC[(cc * CS) * 64 + j];
                        const double2* fc = FC + (size_t)fslot * NT * NTHR + tid;
                        double a1 = 0.0, a2 = 0.0;
#pragma unroll
                        for (int tt = 0; tt < NT; ++tt) {
                            if (FULL || tt < nmt) {
                                const double2 f = fc[tt * NTHR];
                                const double2 a = xc2[((cc * CS) * RG + 8 * tt + 2 * q) >> 1];
                                const double d0 = a.x - zr, d1 = a.y - zr;
                                const double p0 = gb[tt][0] * f.x, p1 = gb[tt][1] * f.y;
                                a1 += p0 + p1;
                                a2 = fma(p0, d0 * d0, fma(p1, d1 * d1, a2));
                            }
                        }
                        gos[cc] += a1;
                        gls[cc] += a2;
                        ++fslot;
                    } else {
                        double zm[LVAE_MAX_MASKS];
#pragma unroll
                        for (int i = 0; i < LVAE_MAX_MASKS; ++i) zm[i] = ZC[(cc * CS + 1 + i) * 64 + j];
#pragma unroll
                        for (int tt = 0; tt < NT; ++tt) {
                            if (FULL || tt < nmt) {
                                const int t0 = 8 * tt + 2 * q;
                                bool on0 = cv && (t0 < R), on1 = cv && (t0 + 1 < R);
#pragma unroll
                                for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
                                    if (i < sp.n_mask[cc]) {
                                        const double2 a = xc2[((cc * CS + 1 + i) * RG + t0) >> 1];
                                        if (sp.mask_type[cc][i] == LVAE_CAT) { on0 = on0 && (a.x - zm[i] == 0.0); on1 = on1 && (a.y - zm[i] == 0.0); }
                                        else { on0 = on0 && (a.x + zm[i] == 2.0); on1 = on1 && (a.y + zm[i] == 2.0); }
                                    }
                                }
                                gos[cc] += (on0 ? gb[tt][0] : 0.0) + (on1 ? gb[tt][1] : 0.0);
                            }
                        }
                    }
                }
                double2* Y2 = reinterpret_cast<double2*>(YTs + j * LDC);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) Y2[4 * nt + q] = make_double2(yacc[nt][0], yacc[nt][1]);
            }
            for (int t = tid; t < R; t += NTHR) d_mu[(size_t)(row0 + t) * L + l] = -2.0 * c * us[t];
            __syncthreads();                         // B3: Y^T complete

            // ---- J6: Q = Y V^T on subject-diagonal upper tiles ; adjoint of B_p against d K1 / d theta ---------------
            for (int tile = wid; tile < NT * (NT + 1) / 2; tile += NWARP) {
                int jt, i;
                tri2(tile, jt, i);                   // i <= jt
                if (FULL || (jt < nmt && 8 * jt < hi_[min(8 * i + 7, R - 1)])) {
                    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;       // two accumulator chains: half the dependency depth
                    const double* Ya = YTs + q * LDC + 8 * i + g;
                    const double* Vb = VTs + q * LDC + 8 * jt + g;
#pragma unroll 4
                    for (int ks = 0; ks + 1 < nks; ks += 2) {
                        dmma(q0, q1, Ya[4 * ks * LDC], Vb[4 * ks * LDC]);
                        dmma(q2, q3, Ya[4 * (ks + 1) * LDC], Vb[4 * (ks + 1) * LDC]);
                    }
                    if (nks & 1) dmma(q0, q1, Ya[4 * (nks - 1) * LDC], Vb[4 * (nks - 1) * LDC]);
                    q0 += q2; q1 += q3;
                    const int t = 8 * i + g;
                    const double wgt = jt > i ? 2.0 : 1.0;
                    const double ut = us[t];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int t2 = 8 * jt + 2 * q + e;
                        if (t < R && t2 < R && (FULL || lo_[t] == lo_[t2])) {
                            const double gB = -wgt * fma(c * ut, us[t2], e ? q1 : q0);
                            if (t == t2) gno += gB;
#pragma unroll
                            for (int k = 0; k < NC1; ++k) {
                                const int cc = NC0 + k;
                                bool on = true;
#pragma unroll
                                for (int i2 = 0; i2 < LVAE_MAX_MASKS; ++i2) {
                                    if (i2 < sp.n_mask[cc]) {
                                        const double a = xc[(cc * CS + 1 + i2) * RG + t], b = xc[(cc * CS + 1 + i2) * RG + t2];
                                        on = on && ((sp.mask_type[cc][i2] == LVAE_CAT) ? (a - b == 0.0) : (a + b == 2.0));
                                    }
                                }
                                double f = on ? 1.0 : 0.0;
                                if (sp.rbf_dim[cc] >= 0) {
                                    const double dd = xc[(cc * CS) * RG + t] - xc[(cc * CS) * RG + t2];
                                    const double d2 = dd * dd;
                                    f = on ? exp_neg(-d2 * hil2[sp.ls_idx[cc]], etab) : 0.0;
                                    g1ls[k] += gB * f * d2;
                                }
                                g1os[k] += gB * f;
                            }
                        }
                    }
                }
            }
        };
        if (mt_[2] == 1 && R > 8 * (NT - 1)) body(std::true_type{});
        else body(std::false_type{});
        if ((gi % FLUSH) == FLUSH - 1) {             // second level of the sum of S (each thread its own slots: no sync needed)
            double2* acc2 = acc2_ptr();
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                double2 v = acc2[d * NTHR];
                v.x += sacc[d][0]; v.y += sacc[d][1];
                acc2[d * NTHR] = v;
                sacc[d][0] = sacc[d][1] = 0.0;
            }
        }
    }
    {
        const double2* acc2 = acc2_ptr();
#pragma unroll
        for (int d = 0; d < 5; ++d) { const double2 v = acc2[d * NTHR]; sacc[d][0] += v.x; sacc[d][1] += v.y; }
    }

    // ---- CTA epilogue: per-thread accumulators -> the partial statistics row of this CTA ------------------------------
    double* part = ws + w.part + ((size_t)chunk * L + l) * w.stride;
    __syncthreads();
    {
        // hyper-gradient partials: [0 .. n_ls) lengthscales, [n_ls .. n_ls + n_comp) outputscales, last = noise
        for (int e = lane; e < nh; e += 32) hyp[wid][e] = 0.0;
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < NC0; ++cc) {
            const double s1 = warp_sum(gos[cc]);
            const double s2 = warp_sum(gls[cc]);
            if (lane == 0) {
                hyp[wid][sp.n_ls + cc] += 2.0 * s1;
                if (sp.rbf_dim[cc] >= 0) hyp[wid][sp.ls_idx[cc]] += 2.0 * s2 * osc[cc] * il3[sp.ls_idx[cc]];
            }
        }
#pragma unroll
        for (int k = 0; k < NC1; ++k) {
            const int cc = NC0 + k;
            const double s1 = warp_sum(g1os[k]);
            const double s2 = warp_sum(g1ls[k]);
            if (lane == 0) {
                hyp[wid][sp.n_ls + cc] += s1;
                if (sp.rbf_dim[cc] >= 0) hyp[wid][sp.ls_idx[cc]] += s2 * osc[cc] * il3[sp.ls_idx[cc]];
            }
        }
        const double n_ = warp_sum(gno);
        if (lane == 0) hyp[wid][nh - 1] += n_;
    }
    // S (symmetric: every tile pair is held once), and from its padding rows / columns: ng1 = S[., 62], da = S[., 63], A = S[63][63]
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        if (d < 4 || wid < 4) {
            const int tj = (wid + d) & 7;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ii = j, jj = 8 * tj + 2 * q + e;
                const double v = sacc[d][e];
                if (ii < M && jj < M) {
                    part[stats_off_S() + (size_t)ii * M + jj] = v;
                    if (d != 0) part[stats_off_S() + (size_t)jj * M + ii] = v;
                }
                if (ii < M && jj == COL_MU) part[stats_off_ng1(M) + ii] = v;
                if (ii < M && jj == COL_R) part[stats_off_da(M) + ii] = v;
                if (d != 0) {                        // the mirrored entry of an off-diagonal tile
                    if (jj < M && ii == COL_MU) part[stats_off_ng1(M) + jj] = v;
                    if (jj < M && ii == COL_R) part[stats_off_da(M) + jj] = v;
                }
                if (ii == COL_R && jj == COL_R) {
                    for (int k = 0; k < LVAE_NSCAL; ++k) part[stats_off_scal(M) + k] = 0.0;
                    part[stats_off_scal(M) + SC_A] = v;
                }
            }
        }
    }
    __syncthreads();
    if (tid < nh) {
        double s = 0.0;
        for (int ww = 0; ww < NWARP; ++ww) s += hyp[ww][tid];
        part[stats_off_hyp(M) + tid] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// group planning: one CTA per chunk; greedy packing of whole subjects into groups of <= rg rows and <= 5 subjects
// entry: row0, R, nsub, end_1..end_5 (local row offsets where each subject ends; unused = R)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_plan_groups3(const int32_t* __restrict__ offsets, int P_b, int per, int gstride, int rg,
                                                      int* __restrict__ gtab, int* __restrict__ gcount) {
    __shared__ int offs[1025];
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const int p_begin = min(P_b, chunk * per), p_end = min(P_b, p_begin + per);
    int* tab = gtab + (size_t)chunk * gstride * GT;
    int ng = 0;
    int cur_row0 = 0, cur_rows = 0, cur_n = 0, ends[5];
    for (int base = p_begin; base < p_end; base += 1024) {
        const int n = min(1024, p_end - base);
        __syncthreads();
        for (int i = tid; i <= n; i += 256) offs[i] = offsets[base + i];
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < n; ++i) {
                const int T = offs[i + 1] - offs[i];
                if (cur_n > 0 && (cur_rows + T > rg || cur_n == 5)) {
                    int* e = tab + (size_t)ng * GT;
                    e[0] = cur_row0; e[1] = cur_rows; e[2] = cur_n;
                    for (int s = 0; s < 5; ++s) e[3 + s] = s < cur_n ? ends[s] : cur_rows;
                    ++ng;
                    cur_n = 0; cur_rows = 0;
                }
                if (cur_n == 0) cur_row0 = offs[i];
                cur_rows += T;
                ends[cur_n++] = cur_rows;
            }
        }
    }
    if (tid == 0) {
        if (cur_n > 0) {
            int* e = tab + (size_t)ng * GT;
            e[0] = cur_row0; e[1] = cur_rows; e[2] = cur_n;
            for (int s = 0; s < 5; ++s) e[3 + s] = s < cur_n ? ends[s] : cur_rows;
            ++ng;
        }
        gcount[chunk] = ng;
    }
}

template <int NC0, int NC1, int NT>
int launch3(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    int nr = 0;
    for (int cc = 0; cc < sp.n0; ++cc) nr += sp.rbf_dim[cc] >= 0;
    const size_t smem = sizeof(double) * Smem<NC0, NC1, NT>::doubles(nr);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_subjects_fused3<NC0, NC1, NT>, smem, attr)) return rc_;
    k_subjects_fused3<NC0, NC1, NT><<<dim3(w.nchunk, p->L), NTHR, smem, st>>>(sp, w, p->L, p->M, p->Q, p->N_b, w.TP, p->x, p->mu,
                                                                            p->z, p->lengthscale, p->outputscale,
                                                                            0.5 * p->scale, p->d_mu, p->workspace);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

template <int NT>
int dispatch3(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int key = sp.n0 * 10 + sp.n1;
    switch (key) {
        case 11: return launch3<1, 1, NT>(p, sp, w, st);
        case 12: return launch3<1, 2, NT>(p, sp, w, st);
        case 21: return launch3<2, 1, NT>(p, sp, w, st);
        case 22: return launch3<2, 2, NT>(p, sp, w, st);
        case 31: return launch3<3, 1, NT>(p, sp, w, st);
        case 32: return launch3<3, 2, NT>(p, sp, w, st);
        case 41: return launch3<4, 1, NT>(p, sp, w, st);
        case 42: return launch3<4, 2, NT>(p, sp, w, st);
    }
    return LVAE_E_BADARG;
}

}  // namespace

int lvae_fused3_rows(const lvae_kld_problem_t* p) { return p->T_max <= 24 ? 24 : 40; }

bool lvae_fused3_supported(const lvae_kld_problem_t* p) {
    if (!(p->M <= COL_MU && p->T_max <= 40 && p->T_max >= 1 && p->ks.n_comp0 >= 1 && p->ks.n_comp0 <= 4 && p->ks.n_comp1 >= 1 &&
          p->ks.n_comp1 <= 2 && p->ks.spec))
        return false;
    int nr = 0;                                  // SE-bearing K0 components keep their values in shared memory
    for (int cc = 0; cc < p->ks.n_comp0; ++cc) nr += p->ks.spec[(size_t)cc * LVAE_SPEC_STRIDE] >= 0;
    return nr <= 3;
}

// CTAs per latent: whole waves of 256-thread CTAs (two per SM for 24-row groups, one for 40-row groups: shared memory);
// k waves cost k * (groups per CTA + set-up)
int lvae_chunks3(int P_b, int L, int T_max) {
    const int rg = T_max <= 24 ? 24 : 40;
    const int per_sm = T_max <= 24 ? 2 : 1;
    const int spg = T_max > 0 ? (rg / T_max > 0 ? rg / T_max : 1) : 1;
    int best = 1;
    long best_cost = -1;
    for (int k = 1; k <= 8; ++k) {
        int n = per_sm * 148 * k / L;
        if (n < 1) continue;
        if (n > P_b) n = P_b > 0 ? P_b : 1;
        const int per = (P_b + n - 1) / n;
        const long cost = (long)k * ((per + spg - 1) / spg + 2);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = n; }
    }
    return best;
}

int lvae_plan_groups3_launch(const lvae_kld_problem_t* p, const KldLayout& w, cudaStream_t st) {
    const int per = (p->P_b + w.nchunk - 1) / w.nchunk;
    k_plan_groups3<<<w.nchunk, 256, 0, st>>>(p->offsets, p->P_b, per, w.gstride, lvae_fused3_rows(p),
                                             reinterpret_cast<int*>(p->workspace + w.gtab),
                                             reinterpret_cast<int*>(p->workspace + w.gcount));
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_subjects_fused3_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    if (!lvae_fused3_supported(p)) return LVAE_E_TOO_LARGE;
    return p->T_max <= 24 ? dispatch3<3>(p, sp, w, st) : dispatch3<5>(p, sp, w, st);
}
