// Per-subject pass of the GP-prior ELBO path for 64 < M <= 256 (at most 24 rows per subject), sm_100a FP64 tensor pipe.
// With M this large the two M^2-sized contractions dominate (4 T M^2 of the ~4 T M^2 + 4 T^2 M flops per subject and
// latent) and their arithmetic intensity against HBM is 2M/8 >= 16 flop/B, so U and V are materialised once in HBM and the
// contractions run as large batched DMMA GEMMs (lvae_gemm.cu) instead of per-subject tiles:
//   k_uv   per (latent, row group): Kxz from covariates, U = L_p^-1 Kxz, V = L_p^-T U (block-diagonal DMMA solves with
//          the L_p^-1 rows of the prep kernel), r = Kxz a - mu, u = V a - B_p^-1 mu, d_mu, A, ng1 = sum V^T mu,
//          da = sum V^T r; writes U, V [L, N_b, MP] and u [L, N_b]                  (elbo_functions.py:171,183,189-190,209-211)
//   GEMM   S = U^T U (lower tiles, mirrored; k split over row ranges -> fixed-order partials)              (184)
//   GEMM   Y = V W   (over U's storage)                                         (adjoint of S)
//   k_adj  per (latent, row group): adjoint of Kxz = 2c u a^T + 2Y against d k_c / d theta; Q = Y V^T on the
//          subject-diagonal tiles (each warp a k-slice, straight from global memory) and the adjoint of B_p
//          = -(c u u^T + Q) against d K1 / d theta and its trace (noise)
//   k_reduce_big  fixed-order sums of all partials into the statistics row (the only buffer exchanged between GPUs)
#include "lvae_blas.h"
#include "lvae_kld.h"

namespace {

constexpr int RG = LVAE_F2_ROWS;   // 24 rows per group
constexpr int NMT = RG / 8;
constexpr int LDL = 28;
constexpr int GT = LVAE_F2_GT;
constexpr int NCB = 8;             // components handled by the register accumulators
#ifndef LVAE_ADJ_CTAS
#define LVAE_ADJ_CTAS 3
#endif

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void row_block(const int* mt_, int t, int R, int& lo, int& hi) {
    lo = t; hi = t;
    if (t < R) {
        lo = 0; hi = 0;
#pragma unroll
        for (int s = 0; s < 5; ++s) { const int end = mt_[3 + s]; if (t >= end) lo = end; }
#pragma unroll
        for (int s = 4; s >= 0; --s) { const int end = mt_[3 + s]; if (t < end) hi = end; }
    }
}

// Per-latent component table in shared memory + covariates GATHERED per component (slot 0: the column of the
// squared-exponential factor, slots 1..3: the columns of the categorical / binary factors), so the evaluation loops only
// see base + constant addresses and warp-uniform branches:  XC[c][slot][row] for the rows of a group, ZC[c][slot][column]
// for the inducing points.
constexpr int CS = 4;
struct CompTab {
    double negh[LVAE_MAXC];      // -1 / (2 l^2) of the component's SE factor
    double osc[LVAE_MAXC];
    double lsw[LVAE_MAXC];       // outputscale / l^3: weight of sum(gbar f d^2) in d/d lengthscale
    double etab[LVAE_EXP_TBL];
    int nmask[LVAE_MAXC], rbf[LVAE_MAXC], lsidx[LVAE_MAXC], mtype[LVAE_MAXC][LVAE_MAX_MASKS];
    int dim[LVAE_MAXC][CS];      // covariate column of every slot (-1: unused)
};
__device__ __forceinline__ void load_comptab(CompTab* ct, const DevSpec& sp, const double* ls, const double* os, int L, int l) {
    const int t = threadIdx.x, nc = sp.n0 + sp.n1;
    if (t < nc) {
        const int li = sp.ls_idx[t], rd = sp.rbf_dim[t];
        double negh = 0.0, lsw = 0.0;
        const double o = os[(size_t)t * L + l];
        if (rd >= 0) { const double v = ls[(size_t)li * L + l]; negh = -0.5 / (v * v); lsw = o / (v * v * v); }
        ct->negh[t] = negh; ct->osc[t] = o; ct->lsw[t] = lsw;
        ct->nmask[t] = sp.n_mask[t]; ct->rbf[t] = rd >= 0; ct->lsidx[t] = li;
        ct->dim[t][0] = rd;
        for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
            ct->mtype[t][i] = sp.mask_type[t][i];
            ct->dim[t][1 + i] = i < sp.n_mask[t] ? sp.mask_dim[t][i] : -1;
        }
    }
    load_exp_table(ct->etab);
}
// gather the covariates of `n` rows (row-major [n][Q] at `src`, rows >= nvalid read as zero) into dst[c][slot][ld]
__device__ __forceinline__ void gather_cov(const CompTab& ct, int c_begin, int c_end, const double* __restrict__ src, int Q,
                                           int n, int nvalid, int ld, double* __restrict__ dst) {
    const int total = (c_end - c_begin) * CS * n;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int r = e % n, sl = (e / n) % CS, c = c_begin + e / (n * CS);
        const int dim = ct.dim[c][sl];
        dst[((size_t)(c - c_begin) * CS + sl) * ld + r] = (dim >= 0 && r < nvalid) ? src[(size_t)r * Q + dim] : 0.0;
    }
}
// Un-scaled values of ONE component on this lane's entries of ONE 8-row tile of a row group: row t = 8*mt + g, column pairs
// (j0, j0 + 1), j0 = colw + 8*nt + 2q (nt < NTW).  NM (number of categorical / binary factors) and RBF are compile-time, so
// the body is straight-line: per factor one FMA + compare per column (cat: a - b == 0, bin: a + b == 2, both written as
// fma(sgn, b, a) == tgt, exactly rounded like the reference's subtraction / addition), per SE factor one exp per column.
// `fn(nt, f0, f1, d20, d21)` consumes the values.  Row / column validity is the caller's business.  The callers keep the loop
// over the row tiles ROLLED: a third of the code and of the live registers of the fully unrolled form.
template <int NM, bool RBF, int NTW, class Fn>
__device__ __forceinline__ void eval_row_block(const CompTab& ct, int c, const double* __restrict__ XCc,
                                               const double* __restrict__ ZCc, int ldz, int t, int q, int colw, Fn&& fn) {
    double sgn[NM > 0 ? NM : 1], tgt[NM > 0 ? NM : 1], am[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        const bool cat = ct.mtype[c][i] == LVAE_CAT;
        sgn[i] = cat ? -1.0 : 1.0;
        tgt[i] = cat ? 0.0 : 2.0;
        am[i] = XCc[(1 + i) * RG + t];
    }
    const double h = ct.negh[c];
    const double ar = RBF ? XCc[t] : 0.0;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
        const int j0 = colw + 8 * nt + 2 * q;
        bool on0 = true, on1 = true;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            const double2 b = *reinterpret_cast<const double2*>(ZCc + (1 + i) * ldz + j0);
            on0 = on0 && (fma(sgn[i], b.x, am[i]) == tgt[i]);
            on1 = on1 && (fma(sgn[i], b.y, am[i]) == tgt[i]);
        }
        double e0 = 1.0, e1 = 1.0, d20 = 0.0, d21 = 0.0;
        if (RBF) {
            const double2 b = *reinterpret_cast<const double2*>(ZCc + j0);
            const double t0 = ar - b.x, t1 = ar - b.y;
            d20 = t0 * t0; d21 = t1 * t1;
            e0 = exp_neg(d20 * h, ct.etab);
            e1 = exp_neg(d21 * h, ct.etab);
        }
        fn(nt, on0 ? e0 : 0.0, on1 ? e1 : 0.0, d20, d21);
    }
}
// generic shape (2-3 factors): run-time factor loop
template <int NTW, class Fn>
__device__ __forceinline__ void eval_row_generic(const CompTab& ct, int c, const double* __restrict__ XCc,
                                                 const double* __restrict__ ZCc, int ldz, int t, int q, int colw, Fn& fn) {
    const int nm = ct.nmask[c];
    const double h = ct.negh[c];
    const bool rbf = ct.rbf[c] != 0;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
        const int j0 = colw + 8 * nt + 2 * q;
        bool on0 = true, on1 = true;
        for (int i = 0; i < nm; ++i) {
            const double a = XCc[(1 + i) * RG + t];
            const double2 b = *reinterpret_cast<const double2*>(ZCc + (1 + i) * ldz + j0);
            if (ct.mtype[c][i] == LVAE_CAT) { on0 = on0 && (a - b.x == 0.0); on1 = on1 && (a - b.y == 0.0); }
            else { on0 = on0 && (a + b.x == 2.0); on1 = on1 && (a + b.y == 2.0); }
        }
        double e0 = 1.0, e1 = 1.0, d20 = 0.0, d21 = 0.0;
        if (rbf) {
            const double a = XCc[t];
            const double2 b = *reinterpret_cast<const double2*>(ZCc + j0);
            const double t0 = a - b.x, t1 = a - b.y;
            d20 = t0 * t0; d21 = t1 * t1;
            e0 = exp_neg(d20 * h, ct.etab);
            e1 = exp_neg(d21 * h, ct.etab);
        }
        fn(nt, on0 ? e0 : 0.0, on1 ? e1 : 0.0, d20, d21);
    }
}
// warp-uniform dispatch on the component's shape: straight-line code for the shapes kernel_gen.py produces without
// missing-value masks (SE, cat/bin x SE, cat/bin), a run-time factor loop for 2-3 factors
template <int NTW, class Fn>
__device__ __forceinline__ void eval_row(const CompTab& ct, int c, const double* __restrict__ XCc, const double* __restrict__ ZCc,
                                         int ldz, int t, int q, int colw, Fn&& fn) {
    const int nm = ct.nmask[c];
    const bool rbf = ct.rbf[c] != 0;
    if (nm == 0) eval_row_block<0, true, NTW>(ct, c, XCc, ZCc, ldz, t, q, colw, fn);
    else if (nm == 1 && rbf) eval_row_block<1, true, NTW>(ct, c, XCc, ZCc, ldz, t, q, colw, fn);
    else if (nm == 1) eval_row_block<1, false, NTW>(ct, c, XCc, ZCc, ldz, t, q, colw, fn);
    else eval_row_generic<NTW>(ct, c, XCc, ZCc, ldz, t, q, colw, fn);
}
// single entry, both sides from the gathered row covariates (K1 components on (X_p, X_p))
__device__ __forceinline__ double eval_rows(const CompTab& ct, int c, const double* __restrict__ XCc, int ldx, int t, int t2,
                                            double& d2) {
    bool on = true;
    const int nm = ct.nmask[c];
    for (int i = 0; i < nm; ++i) {
        const double a = XCc[(1 + i) * ldx + t], b = XCc[(1 + i) * ldx + t2];
        on = on && ((ct.mtype[c][i] == LVAE_CAT) ? (a - b == 0.0) : (a + b == 2.0));
    }
    double e = 1.0;
    d2 = 0.0;
    if (ct.rbf[c]) {
        const double dd = XCc[t] - XCc[t2];
        d2 = dd * dd;
        e = exp_neg(d2 * ct.negh[c], ct.etab);
    }
    return on ? e : 0.0;
}

__host__ __device__ inline size_t uv_doubles(int MP, int Q) {
    (void)Q;
    return (size_t)RG * (MP + 4) + (size_t)RG * LDL + (size_t)4 * CS * MP + (size_t)4 * CS * RG + MP + 4 * RG + 16 * RG + 2 * MP + 2 * RG / 2 + GT;
}

template <int NTW>
__global__ void __launch_bounds__(256, NTW <= 2 ? 3 : 2)
k_uv(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int N_b,
     const double* __restrict__ x, const double* __restrict__ mu, const double* __restrict__ z, const double* __restrict__ ls,
     const double* __restrict__ os, double c, double* __restrict__ d_mu, double* __restrict__ ws) {
    constexpr int MP = 64 * NTW, LDM = MP + 4;
    extern __shared__ double sm[];
    __shared__ CompTab ct;
    __shared__ double red[32];
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, wl = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    double* const K = sm;                       // [RG][LDM] Kxz, then U in place (every warp only touches its own columns)
    double* const Lg = K + RG * LDM;            // [RG][LDL] block-diagonal L^-1 of the group
    double* const ZC = Lg + RG * LDL;           // [4][CS][MP] gathered inducing covariates of the K0 components
    double* const XC = ZC + 4 * CS * MP;        // [4][CS][RG] gathered row covariates
    double* const av = XC + 4 * CS * RG;        // [MP]
    double* const mus = av + MP;                // [RG]
    double* const bmus = mus + RG;
    double* const rs = bmus + RG;
    double* const us = rs + RG;
    double* const rpart = us + RG;              // [8][RG]
    double* const upart = rpart + 8 * RG;       // [8][RG]
    double* const lvl2 = upart + 8 * RG;        // [2][MP] second-level sums of ng1, da
    int* const blo = reinterpret_cast<int*>(lvl2 + 2 * MP);    // [RG]
    int* const bhi = blo + RG;
    int* const meta = bhi + RG;                 // [GT]

    load_comptab(&ct, sp, ls, os, L, l);
    __syncthreads();
    gather_cov(ct, 0, sp.n0, z + (size_t)l * M * Q, Q, MP, M, MP, ZC);
    for (int e = tid; e < MP; e += 256) av[e] = e < M ? ws[w.a + (size_t)l * M + e] : 0.0;
    const int* gtab = reinterpret_cast<const int*>(ws + w.gtab) + (size_t)chunk * w.gstride * GT;
    const int ngroups = reinterpret_cast<const int*>(ws + w.gcount)[chunk];
    const double* Lrows = ws + w.Lrows + (size_t)l * N_b * w.TP;
    const double* bmu_g = ws + w.bmu + (size_t)l * N_b;
    double* const Ug = ws + w.bU + (size_t)l * N_b * MP;
    double* const Vg = ws + w.bV + (size_t)l * N_b * MP;
    double* const ug = ws + w.bu + (size_t)l * N_b;
    const int colw = 8 * NTW * wl;              // first column of this warp

    // two-level sums (a chunk can hold thousands of rows; Kzz^-1 magnifies the last bits of ng1): the register accumulators
    // cover 16 row groups, then they are folded into lvl2 (each warp owns its columns, lanes g == 0 write)
    double ng1acc[NTW][2], daacc[NTW][2], accA = 0.0;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) ng1acc[nt][0] = ng1acc[nt][1] = daacc[nt][0] = daacc[nt][1] = 0.0;
    for (int e = tid; e < 2 * MP; e += 256) lvl2[e] = 0.0;
    auto fold = [&]() {
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double a2 = ng1acc[nt][e], b2 = daacc[nt][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a2 += __shfl_xor_sync(0xffffffffu, a2, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o); }
                if (g == 0) { lvl2[colw + 8 * nt + 2 * q + e] += a2; lvl2[MP + colw + 8 * nt + 2 * q + e] += b2; }
                ng1acc[nt][e] = daacc[nt][e] = 0.0;
            }
        }
    };

    for (int gi = 0; gi < ngroups; ++gi) {
        if ((gi & 15) == 0 && gi) fold();
        __syncthreads();
        if (tid < GT) meta[tid] = gtab[(size_t)gi * GT + tid];
        __syncthreads();
        const int row0 = meta[0], R = meta[1];
        const int R8 = (R + 7) & ~7, nmt = R8 >> 3;
        if (tid < RG) {
            int lo, hi;
            row_block(meta, tid, R, lo, hi);
            blo[tid] = lo; bhi[tid] = hi;
            mus[tid] = tid < R ? mu[(size_t)(row0 + tid) * L + l] : 0.0;
            bmus[tid] = tid < R ? bmu_g[row0 + tid] : 0.0;
        }
        gather_cov(ct, 0, sp.n0, x + (size_t)row0 * Q, Q, RG, R, RG, XC);
        __syncthreads();
        for (int e = tid; e < RG * RG; e += 256) {
            const int t = e / RG, k = e - t * RG, lo = blo[t];
            Lg[t * LDL + k] = (k >= lo && k < bhi[t]) ? Lrows[(size_t)(row0 + t) * w.TP + (k - lo)] : 0.0;
        }
        // ---- Kxz from the gathered covariates (own columns), one 8-row tile at a time ; partial dots of r = Kxz a - mu -----
#pragma unroll 1
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            double kx[NTW][2];
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) kx[nt][0] = kx[nt][1] = 0.0;
            if (mt < nmt) {
                for (int cc = 0; cc < sp.n0; ++cc) {
                    const double o = ct.osc[cc];
                    eval_row<NTW>(ct, cc, XC + cc * CS * RG, ZC + cc * CS * MP, MP, t, q, colw,
                                  [&](int nt, double f0, double f1, double, double) {
                                      kx[nt][0] += o * f0;
                                      kx[nt][1] += o * f1;
                                  });
                }
            }
            double pr = 0.0;
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const int j0 = colw + 8 * nt + 2 * q;
                const double k0_ = (t < R && j0 < M) ? kx[nt][0] : 0.0, k1_ = (t < R && j0 + 1 < M) ? kx[nt][1] : 0.0;
                *reinterpret_cast<double2*>(K + t * LDM + j0) = make_double2(k0_, k1_);
                pr += k0_ * av[j0] + k1_ * av[j0 + 1];
            }
            pr += __shfl_xor_sync(0xffffffffu, pr, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2);
            if (q == 0) rpart[wl * RG + t] = pr;
        }
        __syncthreads();
        if (tid < RG) {
            double s = -mus[tid];
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += rpart[ww * RG + tid];
            rs[tid] = tid < R ? s : 0.0;
        }
        // ---- U = L^-1 Kxz (rows of a tile only see k <= row, inside their subject), in place over this warp's columns of K ----------
        {
            double u[NMT][NTW][2];
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) u[mt][nt][0] = u[mt][nt][1] = 0.0;
                if (mt < nmt) {
                    const int klo = blo[8 * mt] >> 2;
                    const int khi = min(2 * mt + 2, (bhi[min(8 * mt + 7, R - 1)] + 3) >> 2);
                    for (int ks = klo; ks < khi; ++ks) {
                        const double a = Lg[(8 * mt + g) * LDL + 4 * ks + q];
#pragma unroll
                        for (int nt = 0; nt < NTW; ++nt)
                            dmma(u[mt][nt][0], u[mt][nt][1], a, K[(4 * ks + q) * LDM + colw + 8 * nt + g]);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
                const int t = 8 * mt + g;
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) {
                    const int j0 = colw + 8 * nt + 2 * q;
                    *reinterpret_cast<double2*>(K + t * LDM + j0) = make_double2(u[mt][nt][0], u[mt][nt][1]);
                    if (t < R) *reinterpret_cast<double2*>(Ug + (size_t)(row0 + t) * MP + j0) = make_double2(u[mt][nt][0], u[mt][nt][1]);
                }
            }
        }
        __syncthreads();          // rs and U visible
        // ---- V = L^-T U ; ng1 += V^T mu ; da += V^T r ; partial dots of u = V a - B^-1 mu -----------------------------------------
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            double v[NTW][2];
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) v[nt][0] = v[nt][1] = 0.0;
            if (mt < nmt) {
                const int khi = (bhi[min(8 * mt + 7, R - 1)] + 3) >> 2;
                for (int ks = 2 * mt; ks < khi; ++ks) {
                    const double a = Lg[(4 * ks + q) * LDL + 8 * mt + g];
#pragma unroll
                    for (int nt = 0; nt < NTW; ++nt) dmma(v[nt][0], v[nt][1], a, K[(4 * ks + q) * LDM + colw + 8 * nt + g]);
                }
            }
            const double tm = mus[t], tr = rs[t];
            double pu = 0.0;
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const int j0 = colw + 8 * nt + 2 * q;
                if (t < R) *reinterpret_cast<double2*>(Vg + (size_t)(row0 + t) * MP + j0) = make_double2(v[nt][0], v[nt][1]);
                ng1acc[nt][0] += v[nt][0] * tm; ng1acc[nt][1] += v[nt][1] * tm;
                daacc[nt][0] += v[nt][0] * tr; daacc[nt][1] += v[nt][1] * tr;
                pu += v[nt][0] * av[j0] + v[nt][1] * av[j0 + 1];
            }
            pu += __shfl_xor_sync(0xffffffffu, pu, 1);
            pu += __shfl_xor_sync(0xffffffffu, pu, 2);
            if (q == 0) upart[wl * RG + t] = pu;
        }
        __syncthreads();
        if (tid < R) {
            double s = -bmus[tid];
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += upart[ww * RG + tid];
            ug[row0 + tid] = s;
            d_mu[(size_t)(row0 + tid) * L + l] = -2.0 * c * s;
            accA += rs[tid] * s;
        }
    }

    // ---- CTA epilogue: fixed-order partials -------------------------------------------------------------------------------
    double* bp = ws + w.bpart + ((size_t)chunk * L + l) * w.bpstride;
    __syncthreads();
    fold();
    __syncwarp();
    if (g == 0) {
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = colw + 8 * nt + 2 * q + e;
                bp[j] = lvl2[j]; bp[MP + j] = lvl2[MP + j];
            }
    }
    const double a_ = block_sum(accA, red);
    if (tid < LVAE_NSCAL) bp[2 * MP + tid] = (tid == SC_A) ? a_ : 0.0;
}

template <int NTW>
__global__ void __launch_bounds__(256, LVAE_ADJ_CTAS)
k_adj(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int N_b,
      const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ ls, const double* __restrict__ os,
      double c, double* __restrict__ ws) {
    constexpr int MP = 64 * NTW;
    extern __shared__ double sm[];
    __shared__ CompTab ct;
    __shared__ double hypacc[8][2 * LVAE_MAXC + 2];
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, wl = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nh = hyp_count(sp), nc = sp.n0 + sp.n1;
    double* const ZC = sm;                      // [4][CS][MP] gathered inducing covariates of the K0 components
    double* const XC = ZC + 4 * CS * MP;        // [NCB][CS][RG] gathered row covariates of all components
    double* const av = XC + NCB * CS * RG;      // [MP]
    double* const us = av + MP;                 // [RG]
    double* const qred = us + RG;               // [8 warps][6 tiles][32 lanes][2] partial Q of the warps' k-slices
    int* const blo = reinterpret_cast<int*>(qred + 8 * 6 * 64);
    int* const meta = blo + RG;

    load_comptab(&ct, sp, ls, os, L, l);
    __syncthreads();
    gather_cov(ct, 0, sp.n0, z + (size_t)l * M * Q, Q, MP, M, MP, ZC);
    for (int e = tid; e < MP; e += 256) av[e] = e < M ? ws[w.a + (size_t)l * M + e] : 0.0;
    const int* gtab = reinterpret_cast<const int*>(ws + w.gtab) + (size_t)chunk * w.gstride * GT;
    const int ngroups = reinterpret_cast<const int*>(ws + w.gcount)[chunk];
    const double* Yg = ws + w.bU + (size_t)l * N_b * MP;       // Y = V W was written over U
    const double* Vg = ws + w.bV + (size_t)l * N_b * MP;
    const double* ug = ws + w.bu + (size_t)l * N_b;
    const int colw = 8 * NTW * wl;              // this warp's columns = its k-slice of Q = Y V^T

    // hyper-gradient accumulators: one (outputscale, lengthscale) pair per component and warp, kept in shared memory and
    // updated by lane 0 after a warp reduction -> the component loops stay rolled (small code, no register arrays)
    for (int e = tid; e < 8 * (2 * LVAE_MAXC + 2); e += 256) (&hypacc[0][0])[e] = 0.0;
    double gno = 0.0;

    for (int gi = 0; gi < ngroups; ++gi) {
        __syncthreads();
        if (tid < GT) meta[tid] = gtab[(size_t)gi * GT + tid];
        __syncthreads();
        const int row0 = meta[0], R = meta[1];
        const int R8 = (R + 7) & ~7, nmt = R8 >> 3;
        if (tid < RG) {
            int lo, hi;
            row_block(meta, tid, R, lo, hi);
            blo[tid] = lo;
            us[tid] = tid < R ? ug[row0 + tid] : 0.0;
        }
        gather_cov(ct, 0, nc, x + (size_t)row0 * Q, Q, RG, R, RG, XC);
        __syncthreads();
        // ---- adjoint of Kxz = 2c u a^T + 2Y (own columns) against d k_c / d theta of the K0 components: component loop outside,
        // rolled row-tile loop inside (Y re-read per component, from L1), so each component needs ONE warp reduction per group
        for (int cc = 0; cc < sp.n0; ++cc) {
            const double* XCc = XC + cc * CS * RG;
            const double* ZCc = ZC + cc * CS * MP;
            double s1 = 0.0, s2 = 0.0;
#pragma unroll 1
            for (int mt = 0; mt < nmt; ++mt) {
                const int t = 8 * mt + g;
                if (t < R) {
                    const double ut = 2.0 * c * us[t];
                    const double* yrow = Yg + (size_t)(row0 + t) * MP;
                    eval_row<NTW>(ct, cc, XCc, ZCc, MP, t, q, colw, [&](int nt, double f0, double f1, double d20, double d21) {
                        const int j0 = colw + 8 * nt + 2 * q;
                        const double2 y = *reinterpret_cast<const double2*>(yrow + j0);
                        const double w0 = j0 < M ? (ut * av[j0] + 2.0 * y.x) * f0 : 0.0;
                        const double w1 = j0 + 1 < M ? (ut * av[j0 + 1] + 2.0 * y.y) * f1 : 0.0;
                        s1 += w0 + w1;
                        s2 += w0 * d20 + w1 * d21;
                    });
                }
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
                hypacc[wl][sp.n_ls + cc] += s1;
                if (ct.rbf[cc]) hypacc[wl][ct.lsidx[cc]] += s2 * ct.lsw[cc];
            }
        }
        // ---- Q = Y V^T over this warp's k-slice, subject-diagonal upper tiles ; adjoint of B_p = -(c u u^T + Q) ---------------------
        double qa[6][2];
#pragma unroll
        for (int i = 0; i < 6; ++i) qa[i][0] = qa[i][1] = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < 2 * NTW; ++ks) {
            const int kk = colw + 4 * ks + q;
            double ya[NMT], vb[NMT];
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
                const int t = 8 * mt + g;
                const bool ok = t < R;
                ya[mt] = ok ? Yg[(size_t)(row0 + t) * MP + kk] : 0.0;
                vb[mt] = ok ? Vg[(size_t)(row0 + t) * MP + kk] : 0.0;
            }
            dmma(qa[0][0], qa[0][1], ya[0], vb[0]);
            dmma(qa[1][0], qa[1][1], ya[0], vb[1]);
            dmma(qa[2][0], qa[2][1], ya[1], vb[1]);
            dmma(qa[3][0], qa[3][1], ya[0], vb[2]);
            dmma(qa[4][0], qa[4][1], ya[1], vb[2]);
            dmma(qa[5][0], qa[5][1], ya[2], vb[2]);
        }
        // sum the 8 k-slices of Q through shared memory; warp tl < 6 then owns tile tl = (i <= jt): entries (8i + g, 8jt + 2q + e)
#pragma unroll
        for (int tl = 0; tl < 6; ++tl)
            *reinterpret_cast<double2*>(qred + ((wl * 6 + tl) * 32 + lane) * 2) = make_double2(qa[tl][0], qa[tl][1]);
        __syncthreads();
        if (wl < 6) {
            const int tl = wl;
            const int jt = tl < 1 ? 0 : (tl < 3 ? 1 : 2);
            const int i = tl - (jt * (jt + 1)) / 2;
            if (jt < nmt) {
                double qs0 = 0.0, qs1 = 0.0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) {
                    const double2 v = *reinterpret_cast<const double2*>(qred + ((ww * 6 + tl) * 32 + lane) * 2);
                    qs0 += v.x; qs1 += v.y;
                }
                const int t = 8 * i + g;
                const double wgt = jt > i ? -2.0 : -1.0;
                double gB[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int t2 = 8 * jt + 2 * q + e;
                    const bool on = t < R && t2 < R && blo[t] == blo[t2];
                    gB[e] = on ? wgt * (c * us[t] * us[t2] + (e ? qs1 : qs0)) : 0.0;     // adjoint of B_p, zero outside the subject blocks
                    if (t == t2) gno += gB[e];
                }
                for (int cc = sp.n0; cc < nc; ++cc) {
                    const double* XCc = XC + cc * CS * RG;
                    double s1 = 0.0, s2 = 0.0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int t2 = 8 * jt + 2 * q + e;
                        double d2;
                        const double f = eval_rows(ct, cc, XCc, RG, t, t2, d2);
                        const double w_ = gB[e] * f;
                        s1 += w_;
                        s2 += w_ * d2;
                    }
                    s1 = warp_sum(s1);
                    s2 = warp_sum(s2);
                    if (lane == 0) {
                        hypacc[wl][sp.n_ls + cc] += s1;
                        if (ct.rbf[cc]) hypacc[wl][ct.lsidx[cc]] += s2 * ct.lsw[cc];
                    }
                }
            }
        }
    }

    // ---- CTA epilogue: hyper-gradient partials of this CTA, fixed order ------------------------------------------------------------
    {
        const double n_ = warp_sum(gno);
        if (lane == 0) hypacc[wl][nh - 1] += n_;
    }
    __syncthreads();
    double* bp = ws + w.bpart + ((size_t)chunk * L + l) * w.bpstride + 2 * MP + LVAE_NSCAL;
    if (tid < nh) {
        double s = 0.0;
        for (int ww = 0; ww < 8; ++ww) s += hypacc[ww][tid];
        bp[tid] = s;
    }
}

// stats[l][k] = fixed-order sum of the S partials (k < M^2) or of the k_uv / k_adj / prep partials (vectors, scalars, hyper-gradients)
__global__ void __launch_bounds__(256) k_reduce_big(KldLayout w, int L, int M, const double* __restrict__ ws,
                                                    double* __restrict__ stats) {
    const int l = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w.stride) return;
    const int64_t MM = (int64_t)M * M;
    double s = 0.0;
    if (k < MM) {
        for (int ch = 0; ch < w.nsplit; ++ch) s += ws[w.part + ((size_t)ch * L + l) * w.stride + k];
    } else {
        const int64_t kv = k - MM;                          // ng1 [M] | da [M] | scalars | hyp
        const int64_t src = kv < M ? kv : (kv < 2 * M ? w.MP + (kv - M) : 2 * (int64_t)w.MP + (kv - 2 * M));
        for (int ch = 0; ch < w.nchunk; ++ch) s += ws[w.bpart + ((size_t)ch * L + l) * w.bpstride + src];
        const int64_t ks = k - stats_off_scal(M);
        if (ks >= 0) {
            for (int ch = 0; ch < w.nprep; ++ch) s += ws[w.ppart + ((size_t)ch * L + l) * (LVAE_NSCAL + w.nh) + ks];
        }
    }
    stats[(size_t)l * w.stride + k] = s;
}

template <int NTW>
int launch_uv(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const size_t smem = sizeof(double) * uv_doubles(64 * NTW, p->Q);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_uv<NTW>, smem, attr)) return rc_;
    k_uv<NTW><<<dim3(w.nchunk, p->L), 256, smem, st>>>(sp, w, p->L, p->M, p->Q, p->N_b, p->x, p->mu, p->z, p->lengthscale,
                                                        p->outputscale, 0.5 * p->scale, p->d_mu, p->workspace);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

template <int NTW>
int launch_adj(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int MP = 64 * NTW;
    const size_t smem = sizeof(double) * ((size_t)4 * CS * MP + (size_t)NCB * CS * RG + MP + RG + 8 * 6 * 64 + RG / 2 + GT);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_adj<NTW>, smem, attr)) return rc_;
    k_adj<NTW><<<dim3(w.nchunk, p->L), 256, smem, st>>>(sp, w, p->L, p->M, p->Q, p->N_b, p->x, p->z, p->lengthscale,
                                                         p->outputscale, 0.5 * p->scale, p->workspace);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

int lvae_subjects_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int L = p->L, M = p->M, MP = w.MP, N_b = p->N_b;
    double* ws = p->workspace;
    int rc = MP == 128 ? launch_uv<2>(p, sp, w, st) : launch_uv<4>(p, sp, w, st);
    if (rc) return rc;
    GemmDesc s;                                   // S partials: part[split][l] = U_l[rows of the split]^T U_l[same rows]
    int kchunk = (N_b + w.nsplit - 1) / w.nsplit;
    kchunk = (kchunk + 15) & ~15;
    s.A = ws + w.bU; s.lda = MP; s.sA = (int64_t)N_b * MP; s.ta = 1;
    s.B = ws + w.bU; s.ldb = MP; s.sB = (int64_t)N_b * MP;
    s.C = ws + w.part; s.ldc = M; s.sC = w.stride;
    s.m = M; s.n = M; s.k = N_b; s.batch = L;
    s.ksplit = w.nsplit; s.kchunk = kchunk;
    s.kA = (int64_t)kchunk * MP; s.kB = (int64_t)kchunk * MP; s.kC = (int64_t)L * w.stride;
    s.flags = LVAE_GEMM_LOWER | LVAE_GEMM_MIRROR;
    s.flush = 16;                                 // rounding chains of S: 64 DMMA steps, then kchunk / 256 + nsplit additions
    rc = lvae_gemm(s, st);
    if (rc) return rc;
    GemmDesc y;                                   // Y = V W, written over U
    y.A = ws + w.bV; y.lda = MP; y.sA = (int64_t)N_b * MP;
    y.B = ws + w.bWp; y.ldb = MP; y.sB = (int64_t)MP * MP;
    y.C = ws + w.bU; y.ldc = MP; y.sC = (int64_t)N_b * MP;
    y.m = N_b; y.n = MP; y.k = MP; y.batch = L;
    rc = lvae_gemm(y, st);
    if (rc) return rc;
    return MP == 128 ? launch_adj<2>(p, sp, w, st) : launch_adj<4>(p, sp, w, st);
}

int lvae_reduce_big_launch(const lvae_kld_problem_t* p, const KldLayout& w, cudaStream_t st) {
    k_reduce_big<<<dim3((unsigned)((w.stride + 255) / 256), p->L), 256, 0, st>>>(w, p->L, p->M, p->workspace, p->stats);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}
