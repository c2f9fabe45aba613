// Per-subject T x T work of the GP-prior ELBO path, one group of NW warps per (subject, latent) task (NW = 1 for T <= 24:
// no block barriers at all; NW = 4 for 24 < T <= 40, where three T x T scratch matrices per task leave room for only five
// tasks per SM and one warp per task would run the SM at 4-5 warps; the warps of a group meet at a named barrier):
//   B_p = K1(X_p, X_p) + s2 I  ->  Cholesky (in place)  ->  L^-1  ->  B_p^-1 = L^-T L^-1   (elbo_functions.py:174,179-180)
//   K0_p (173), scalars C = 2 sum log L_tt (192), D1 = sum B^-1 o K0 (193), Bt = sum (B^-1)_tt e^{logv} (191), F (196),
//   d_log_v, and the part of the reverse pass that only needs T x T blocks:
//     adjoint of K0_p = c B^-1 ;  local adjoint of B_p = c (B^-1 - B^-1 (diag(v) + K0_p) B^-1)
//   contracted with d k_c / d theta on the fly (component entries are never stored).
// The three T x T products run on the FP64 tensor pipe (DMMA.8x8x4); the three T x T scratch matrices of a warp live in
// shared memory with a row stride = 4 or 12 (mod 16) doubles, which makes every DMMA fragment load conflict-free.
// B_p^-1 is written to the workspace for the fused subject pass.  Partial sums are kept per lane in registers over all
// tasks of the warp and written once (deterministic fixed-order reduction in k_reduce).
#include "lvae_kld.h"

namespace {

constexpr int PWMAX = 8;     // warps per CTA (fewer when the per-warp scratch of a large T does not fit)
constexpr int NCMAX = 8;     // components (K0 + K1) handled by the register accumulators
constexpr int NT8MAX = 5;    // T <= 40

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[i][j] (i,j < T) = sum_k A(i,k) B(k,j) for row-major operands in smem; TA: A(i,k) = A[k][i].  Output row by row of
// 8x8 tiles; `sym`: only tiles tj <= ti are computed and mirrored.
// (i, j) of the e-th element of the lower triangle (row-major, j <= i)
__device__ __forceinline__ void tri_index(int e, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    while (i * (i + 1) / 2 > e) --i;
    j = e - i * (i + 1) / 2;
}

// advance (i, j) of a lower-triangle element by `n` positions (row-major, j <= i)
__device__ __forceinline__ void tri_step(int& i, int& j, int n) {
    j += n;
    while (j > i) { j -= i + 1; ++i; }
}
// barrier of the NW warps working on one task (named barrier `bar`), a plain __syncwarp for NW == 1
template <int NW>
__device__ __forceinline__ void gsync(int bar) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * NW) : "memory");
}

// Blocked Cholesky of the T x T matrix A (row stride ld, zero padded), 8 x 8 blocks, by the NW warps of a group: the
// diagonal block is factored column by column by warp 0, the panel below by one lane per row (forward substitution), the
// trailing matrix by DMMA (tiles dealt to the warps).  dinv[k] receives 1 / L_kk.  `bad` (1 + first non-positive pivot) is
// only meaningful in warp 0.
template <int NW>
__device__ __forceinline__ int grp_cholesky(double* __restrict__ A, int T, int ld, double* __restrict__ dinv, int lane, int wg,
                                            int bar) {
    const int g = lane >> 2, q = lane & 3, gl = lane + 32 * wg;
    const int nb = (T + 7) >> 3;
    int bad = 0;
    // lane -> (ur, uc), 1 <= uc <= ur <= 7: its element of the 7 x 7 lower triangle updated inside a diagonal block
    int ur = 0;
    while ((ur + 1) * (ur + 2) / 2 <= lane) ++ur;
    const int uc = lane - ur * (ur + 1) / 2 + 1;
    ur += 1;
    for (int kb = 0; kb < nb; ++kb) {
        const int k0 = 8 * kb, bs = min(8, T - k0);
        if (wg == 0) {
            for (int k = 0; k < bs; ++k) {
                const double akk = A[(k0 + k) * ld + k0 + k];
                if (!(akk > 0.0) && bad == 0) bad = k0 + k + 1;
                const double ri = rsqrt(akk);
                __syncwarp();
                if (lane == k) { A[(k0 + k) * ld + k0 + k] = akk * ri; dinv[k0 + k] = ri; }
                if (lane > k && lane < bs) A[(k0 + lane) * ld + k0 + k] *= ri;
                __syncwarp();
                if (lane < 28 && uc > k && ur < bs)
                    A[(k0 + ur) * ld + k0 + uc] -= A[(k0 + ur) * ld + k0 + k] * A[(k0 + uc) * ld + k0 + k];
                __syncwarp();
            }
        }
        if (kb + 1 < nb) {
            gsync<NW>(bar);
            for (int r = k0 + 8 + gl; r < T; r += 32 * NW) {        // panel rows: x D^T = a
                double xr[8];
#pragma unroll
                for (int c_ = 0; c_ < 8; ++c_) {
                    double s = A[r * ld + k0 + c_];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < c_) s -= xr[k] * A[(k0 + c_) * ld + k0 + k];
                    xr[c_] = s * dinv[k0 + c_];
                }
#pragma unroll
                for (int c_ = 0; c_ < 8; ++c_) A[r * ld + k0 + c_] = xr[c_];
            }
            gsync<NW>(bar);
            int tl = 0;
            for (int ti = kb + 1; ti < nb; ++ti) {                   // trailing update on the lower tiles
                for (int tj = kb + 1; tj <= ti; ++tj, ++tl) {
                    if (NW > 1 && (tl % NW) != wg) continue;
                    const int i = 8 * ti + g, j = 8 * tj + 2 * q;
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        dmma(c0, c1, A[i * ld + k0 + 4 * ks + q], A[(8 * tj + g) * ld + k0 + 4 * ks + q]);
                    if (i < T && j < T) A[i * ld + j] -= c0;
                    if (i < T && j + 1 < T) A[i * ld + j + 1] -= c1;
                }
            }
            gsync<NW>(bar);
        }
    }
    gsync<NW>(bar);
    return bad;
}

// X = L^-1 (lower, T x T) into a zero-initialised X: 8 x 8 diagonal blocks by one lane per column, then the block
// sub-diagonals X_ij = -X_ii (sum_k L_ik X_kj) on the tensor pipe (blocks of a sub-diagonal dealt to the warps).
// tile: 64 doubles of scratch per warp.
template <int NW>
__device__ __forceinline__ void grp_tri_inverse(const double* __restrict__ Lc, double* __restrict__ X, int T, int ld,
                                                const double* __restrict__ dinv, double* __restrict__ tile, int lane, int wg,
                                                int bar) {
    const int g = lane >> 2, q = lane & 3;
    const int nb = (T + 7) >> 3;
    for (int j = lane + 32 * wg; j < T; j += 32 * NW) {
        const int kend = min(T, (j & ~7) + 8);
        X[j * ld + j] = dinv[j];
        for (int i = j + 1; i < kend; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s += Lc[i * ld + k] * X[k * ld + j];
            X[i * ld + j] = -s * dinv[i];
        }
    }
    gsync<NW>(bar);
    for (int d = 1; d < nb; ++d) {
        for (int bj = wg; bj + d < nb; bj += NW) {
            const int bi = bj + d;
            double t0 = 0.0, t1 = 0.0;
            for (int bk = bj; bk < bi; ++bk) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    dmma(t0, t1, Lc[(8 * bi + g) * ld + 8 * bk + 4 * ks + q], X[(8 * bk + 4 * ks + q) * ld + 8 * bj + g]);
            }
            tile[g * 8 + 2 * q] = t0;
            tile[g * 8 + 2 * q + 1] = t1;
            __syncwarp();
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
                dmma(x0, x1, -X[(8 * bi + g) * ld + 8 * bi + 4 * ks + q], tile[(4 * ks + q) * 8 + g]);
            const int i = 8 * bi + g, j = 8 * bj + 2 * q;
            if (i < T) { X[i * ld + j] = x0; X[i * ld + j + 1] = x1; }
            __syncwarp();
        }
        gsync<NW>(bar);
    }
}

template <bool TA, bool SYM, int NW>
__device__ __forceinline__ void grp_mm(const double* __restrict__ A, const double* __restrict__ B,
                                       double* __restrict__ C, int T, int ld, int nt8, int nk4, int g, int q, int wg) {
    for (int ti = SYM ? nt8 - 1 - wg : wg; SYM ? ti >= 0 : ti < nt8; ti += SYM ? -NW : NW) {   // SYM: longest rows first
        double acc[NT8MAX][2];
#pragma unroll
        for (int tj = 0; tj < NT8MAX; ++tj) acc[tj][0] = acc[tj][1] = 0.0;
        const int tjmax = SYM ? ti + 1 : nt8;
        for (int ks = 0; ks < nk4; ++ks) {
            const double a = TA ? A[(4 * ks + q) * ld + 8 * ti + g] : A[(8 * ti + g) * ld + 4 * ks + q];
#pragma unroll
            for (int tj = 0; tj < NT8MAX; ++tj) {
                if (tj < tjmax) {
                    const double b = B[(4 * ks + q) * ld + 8 * tj + g];
                    dmma(acc[tj][0], acc[tj][1], a, b);
                }
            }
        }
        const int i = 8 * ti + g;
#pragma unroll
        for (int tj = 0; tj < NT8MAX; ++tj) {
            if (tj < tjmax) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * tj + 2 * q + e;
                    if (i < T && j < T) {
                        C[i * ld + j] = acc[tj][e];
                        if (SYM && tj < ti) C[j * ld + i] = acc[tj][e];
                    }
                }
            }
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(NW == 1 ? PWMAX * 32 : 512, NW == 1 ? 2 : 1)
k_prep_warp(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int Q, int P_b, int N_b, int Tmax,
            int ld,
            const double* __restrict__ x, const int32_t* __restrict__ offsets, const double* __restrict__ mu,
            const double* __restrict__ log_v,
            const double* __restrict__ ls, const double* __restrict__ os, const double* __restrict__ noise, double c,
            double* __restrict__ d_log_v, double* __restrict__ ws, int32_t* info) {
    extern __shared__ double sm[];
    __shared__ double hil2[LVAE_MAXC], il3[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    __shared__ double s_noise;
    const int l = blockIdx.y, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int grp = wid / NW, wg = wid % NW, gl = lane + 32 * wg, bar = 1 + grp;      // group of NW warps = one task at a time
    constexpr int NL = 32 * NW;
    const int nc = sp.n0 + sp.n1, nh = hyp_count(sp);
    if (tid < sp.n_ls) { const double v = ls[(size_t)tid * L + l]; hil2[tid] = 0.5 / (v * v); il3[tid] = 1.0 / (v * v * v); }
    if (tid < nc) osc[tid] = os[(size_t)tid * L + l];
    if (tid == 0) s_noise = noise[l];
    load_exp_table(etab);
    const int TP8 = (Tmax + 7) & ~7;
    const int asz = TP8 * ld;
    double* A1 = sm + (size_t)grp * (3 * asz + Tmax * Q + 2 * Tmax + TP8 + 64 * NW);
    double* A2 = A1 + asz;
    double* A3 = A2 + asz;
    double* xs = A3 + asz;
    double* ev = xs + Tmax * Q;
    double* mw = ev + Tmax;
    double* dinv = mw + Tmax;          // [TP8]
    double* tile = dinv + TP8 + 64 * wg;   // [64] per warp
    for (int e = gl; e < 3 * asz; e += NL) A1[e] = 0.0;
    __syncthreads();
    const int64_t* off2 = reinterpret_cast<const int64_t*>(ws + w.off2);   // (block layout only; unused when w.v2)

    double sC = 0.0, sD1 = 0.0, sBt = 0.0, sF = 0.0, gno = 0.0;
    double gos[NCMAX], gls[NCMAX];
#pragma unroll
    for (int k = 0; k < NCMAX; ++k) gos[k] = gls[k] = 0.0;

    const int GPC = (blockDim.x >> 5) / NW;                // groups per CTA
    const int ngrp = gridDim.x * GPC, gg = blockIdx.x * GPC + grp;
    int Tprev = -1;
    for (int p = gg; p < P_b; p += ngrp) {
        const int r0 = offsets[p], T = offsets[p + 1] - r0;
        const int nt8 = (T + 7) >> 3, nk4 = (T + 3) >> 2;
        gsync<NW>(bar);
        if (Tprev != -1 && T != Tprev) {          // ragged batches: restore the zero padding the DMMA tiles rely on
            for (int e = gl; e < 3 * asz; e += NL) A1[e] = 0.0;
        }
        Tprev = T;
        for (int e = gl; e < T * Q; e += NL) xs[e] = x[(size_t)r0 * Q + e];
        for (int t = gl; t < T; t += NL) {
            const double lv = log_v[(size_t)(r0 + t) * L + l];
            ev[t] = exp(lv);
            sF += lv;
            if (w.v2) mw[t] = mu[(size_t)(r0 + t) * L + l];
        }
        gsync<NW>(bar);
        // ---- B_p = K1 + noise I (lower triangle evaluated, mirrored) -------------------------------------------------
        const int ntri = T * (T + 1) / 2;
        {
            int i, j;
            tri_index(gl, i, j);
            for (int e = gl; e < ntri; e += NL) {
                double k1 = 0.0, d2;
                for (int cc = sp.n0; cc < nc; ++cc) k1 += osc[cc] * comp_one(sp, cc, xs + i * Q, xs + j * Q, hil2, etab, d2);
                if (i == j) k1 += s_noise;
                A1[i * ld + j] = k1;
                A1[j * ld + i] = k1;
                tri_step(i, j, NL);
            }
        }
        for (int e = gl; e < asz; e += NL) A2[e] = 0.0;
        gsync<NW>(bar);
        // ---- blocked Cholesky in place, L^-1 into A2 (tensor pipe for the block updates) ------------------------------------
        {
            const int bad = grp_cholesky<NW>(A1, T, ld, dinv, lane, wg, bar);
            if (bad && wg == 0 && lane == 0) atomicCAS(info + 2, 0, l * P_b + p + 1);
        }
        for (int t = gl; t < T; t += NL) sC -= 2.0 * log(dinv[t]);                                            // 192
        grp_tri_inverse<NW>(A1, A2, T, ld, dinv, tile, lane, wg, bar);
        if (w.v2) {   // rows of L^-1 for the fused pass: [row][k'], k' = column inside the subject, zero padded to TP
            double* gl_ = ws + w.Lrows + ((size_t)l * N_b + r0) * w.TP;
            int i = 0, k = gl;
            while (k >= w.TP) { k -= w.TP; ++i; }
            for (int e = gl; e < T * w.TP; e += NL) {
                gl_[e] = (k <= i) ? A2[i * ld + k] : 0.0;
                k += NL;
                while (k >= w.TP) { k -= w.TP; ++i; }
            }
        }
        // ---- B^-1 = L^-T L^-1 into A3 --------------------------------------------------------------------------------
        grp_mm<true, true, NW>(A2, A2, A3, T, ld, nt8, nk4, g, q, wg);
        gsync<NW>(bar);
        // ---- K0_p (+ diag v) into A1 ; D1 ; adjoint of K0 = c B^-1 contracted on the fly (lower triangle, weight 2) ---
        int ti_, tj_;
        tri_index(gl, ti_, tj_);
        for (int e = gl; e < ntri; e += NL) {
            const int i = ti_, j = tj_;
            tri_step(ti_, tj_, NL);
            const double bi = A3[i * ld + j] * (i == j ? 1.0 : 2.0);
            double k0 = 0.0;
#pragma unroll
            for (int cc = 0; cc < NCMAX; ++cc) {
                if (cc < sp.n0) {
                    double d2;
                    const double f = comp_one(sp, cc, xs + i * Q, xs + j * Q, hil2, etab, d2);
                    k0 += osc[cc] * f;
                    gos[cc] += bi * f;
                    if (sp.rbf_dim[cc] >= 0) gls[cc] += bi * f * d2;
                }
            }
            sD1 += bi * k0;
            A1[i * ld + j] = k0 + (i == j ? ev[i] : 0.0);
            A1[j * ld + i] = A1[i * ld + j];
        }
        gsync<NW>(bar);
        // ---- X1 = (diag v + K0) B^-1 -> A2 ; X2 = B^-1 X1 -> A1 -------------------------------------------------------
        grp_mm<false, false, NW>(A1, A3, A2, T, ld, nt8, nk4, g, q, wg);
        gsync<NW>(bar);
        grp_mm<false, false, NW>(A3, A2, A1, T, ld, nt8, nk4, g, q, wg);
        gsync<NW>(bar);
        // ---- B^-1 out ; local adjoint of B_p: (B^-1 - X2) [times c at the end] ; K1 hyper-gradients ; Bt ; d_log_v -------
        if (w.v2) {
            double* gb = ws + w.bmu + (size_t)l * N_b + r0;
            for (int t = gl; t < T; t += NL) {
                double s = 0.0;
                for (int k = 0; k < T; ++k) s += A3[t * ld + k] * mw[k];
                gb[t] = s;
            }
        } else {
            double* gBi = ws + w.Bi + (size_t)l * w.Bi_stride + off2[p];
            int i = 0, j = gl;
            while (j >= T) { j -= T; ++i; }
            for (int e = gl; e < T * T; e += NL) {
                gBi[e] = A3[i * ld + j];
                j += NL;
                while (j >= T) { j -= T; ++i; }
            }
        }
        tri_index(gl, ti_, tj_);
        for (int e = gl; e < ntri; e += NL) {
            const int i = ti_, j = tj_;
            tri_step(ti_, tj_, NL);
            const double bi = A3[i * ld + j];
            const double gB = (i == j) ? bi - A1[i * ld + i] : 2.0 * bi - (A1[i * ld + j] + A1[j * ld + i]);
#pragma unroll
            for (int cc = 0; cc < NCMAX; ++cc) {
                if (cc >= sp.n0 && cc < nc) {
                    double d2;
                    const double f = comp_one(sp, cc, xs + i * Q, xs + j * Q, hil2, etab, d2);
                    gos[cc] += gB * f;
                    if (sp.rbf_dim[cc] >= 0) gls[cc] += gB * f * d2;
                }
            }
            if (i == j) {
                gno += gB;
                const double bt = bi * ev[i];
                sBt += bt;
                d_log_v[(size_t)(r0 + i) * L + l] = c * (bt - 1.0);
            }
        }
    }
    // ---- per-warp partial row ---------------------------------------------------------------------------------------
    const int gw = blockIdx.x * (blockDim.x >> 5) + wid;
    double* out = ws + w.ppart + ((size_t)gw * L + l) * (LVAE_NSCAL + nh);
    sC = warp_sum(sC); sD1 = warp_sum(sD1); sBt = warp_sum(sBt); sF = warp_sum(sF); gno = warp_sum(gno);
    if (lane == 0) {
        for (int k = 0; k < LVAE_NSCAL + nh; ++k) out[k] = 0.0;
        out[SC_C] = sC; out[SC_D1] = sD1; out[SC_BT] = sBt; out[SC_F] = sF;
        out[LVAE_NSCAL + nh - 1] = c * gno;
    }
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < NCMAX; ++cc) {
        if (cc < nc) {
            const double a = warp_sum(gos[cc]);
            const double b = warp_sum(gls[cc]);
            if (lane == 0) {
                out[LVAE_NSCAL + sp.n_ls + cc] = c * a;
                // several components may share no lengthscale row; each SE component owns its row (spec.py)
                if (sp.rbf_dim[cc] >= 0) out[LVAE_NSCAL + sp.ls_idx[cc]] += c * b * osc[cc] * il3[sp.ls_idx[cc]];
            }
        }
    }
}

} // namespace

static int ld_for(int Tmax) {
    int ld = (Tmax + 3) & ~3;
    while (ld % 16 != 4 && ld % 16 != 12) ++ld;
    return ld;
}

static int group_warps(int Tm) { return Tm <= 24 ? 1 : 4; }

static size_t group_doubles(int Tm, int Q) {
    const int ld = ld_for(Tm), TP8 = (Tm + 7) & ~7;
    return 3 * (size_t)TP8 * ld + (size_t)Tm * Q + 2 * Tm + TP8 + 64 * group_warps(Tm);
}

// groups per CTA: as many as fit ~200 KB (NW = 4: at most 4 groups = 16 warps), at most 8 single-warp groups
static int groups_per_cta(int Tm, int Q) {
    const int nw = group_warps(Tm);
    int n = (int)((200 * 1024) / (sizeof(double) * group_doubles(Tm, Q)));
    const int cap = nw == 1 ? PWMAX : 4;
    if (n > cap) n = cap;
    return n < 1 ? 1 : n;
}

int lvae_prep_rows(int P_b, int L, int T_max, int Q) {
    // partial rows per latent = warps per latent: about one wave of CTAs, a few tasks per group
    const int Tm = T_max > 0 ? T_max : 1;
    const int nw = group_warps(Tm), gpc = groups_per_cta(Tm, Q), pw = gpc * nw;
    const size_t cta_bytes = sizeof(double) * gpc * group_doubles(Tm, Q) + 2048;
    int per_sm = (int)((227 * 1024) / cta_bytes);
    if (per_sm < 1) per_sm = 1;
    if (nw > 1) per_sm = 1;
    if (per_sm * pw > 32) per_sm = 32 / pw > 0 ? 32 / pw : 1;
    int ctas = per_sm * 148 / L;
    if (ctas < 1) ctas = 1;
    const int need = (P_b + gpc - 1) / gpc;
    if (ctas > need) ctas = need > 0 ? need : 1;
    return ctas * pw;
}

bool lvae_prep_warp_supported(const lvae_kld_problem_t* p) {
    return p->ks.n_comp0 + p->ks.n_comp1 <= NCMAX && p->T_max <= 8 * NT8MAX;
}

template <int NW>
static int launch_prep(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int Tm = p->T_max > 0 ? p->T_max : 1;
    const int ld = ld_for(Tm);
    const int gpc = groups_per_cta(Tm, p->Q), pw = gpc * NW;
    const size_t smem = sizeof(double) * (size_t)gpc * group_doubles(Tm, p->Q);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_prep_warp<NW>, smem, attr)) return rc_;
    k_prep_warp<NW><<<dim3(w.nprep / pw, p->L), pw * 32, smem, st>>>(sp, w, p->L, p->Q, p->P_b, p->N_b, Tm, ld, p->x, p->offsets,
                                                                    p->mu, p->log_v, p->lengthscale, p->outputscale, p->noise,
                                                                    0.5 * p->scale, p->d_log_v, p->workspace, p->info);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_prep_warp_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int Tm = p->T_max > 0 ? p->T_max : 1;
    return group_warps(Tm) == 1 ? launch_prep<1>(p, sp, w, st) : launch_prep<4>(p, sp, w, st);
}
