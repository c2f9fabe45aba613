// Per-subject T x T work of the GP-prior ELBO path, third generation of the prep kernel (same outputs as k_prep_warp of
// lvae_prep.cu, which stays as the fallback for kernel shapes this one does not cover).  The profile of the previous
// version was flat and 86 % integer / control instructions (ncu, profiles/r01_cfg2_prep_fused2_ncu_summary.txt), so this
// one removes them structurally:
//   * NT8 (8-row tiles per matrix), the row stride LD and NW (warps per task) are compile-time: every shared-memory address
//     is base + constant; the tile loops of the T x T linear algebra are unrolled for 3 x 3 tiles only (for 5 x 5 the code
//     outgrows the instruction cache: measured 2.9 -> 1.8 ms at cfg4 when they were rolled again);
//   * the lower triangle is walked through a (i, j) table built once per CTA instead of incremental index arithmetic;
//   * covariates are gathered per component once per task, and every component is evaluated by straight-line code
//     selected by a warp-uniform switch on its shape (SE | cat/bin x SE | cat/bin | generic), hoisted out of the entry loop;
//     the entry loops themselves stay rolled and accumulate in place in shared memory (a fully unrolled variant with the
//     entries in registers was 10 % slower at fixed T and 3x slower on ragged batches: instruction-fetch stalls);
//   * hyper-gradient partial sums are reduced per component (warp shuffle) into a per-warp shared-memory row instead
//     of per-lane register arrays indexed by a run-time component number.
//   B_p = K1(X_p, X_p) + s2 I -> Cholesky -> L^-1 -> B_p^-1 (elbo_functions.py:174,179-180); K0_p (173); C, D1, Bt, F
//   (191-196); d_log_v; local adjoints c B^-1 (of K0_p) and c (B^-1 - B^-1 (diag v + K0_p) B^-1) (of B_p) contracted
//   with d k_c / d theta; exports the rows of L_p^-1 and B_p^-1 mu_p for the subject pass.
#include <stdlib.h>

#include "lvae_kld.h"

namespace {

constexpr int NCM = 8;          // additive components (K0 + K1)
constexpr int SL = 4;           // slots per component at most (1 SE column + 3 mask columns)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NW>
__device__ __forceinline__ void gsync(int bar) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * NW) : "memory");
}

struct PrepTab {                // per CTA (= one latent), shared memory
    double negh[NCM], osc[NCM], lsw[NCM];      // -1/(2 l^2), outputscale, outputscale / l^3
    double sgn[NCM][3], tgt[NCM][3];           // factor i of component c holds iff fma(sgn, b, a) == tgt
    double noise;
    double etab[LVAE_EXP_TBL];
    int nmask[NCM], rbf[NCM], lsidx[NCM], slot0[NCM];
    int dim[NCM * SL];                         // covariate column of every compact slot
    int nslots;
};

__device__ inline void load_preptab(PrepTab* pt, const DevSpec& sp, const double* ls, const double* os, const double* noise,
                                    int L, int l) {
    const int t = threadIdx.x, nc = sp.n0 + sp.n1;
    if (t < nc) {
        const int rd = sp.rbf_dim[t], li = sp.ls_idx[t];
        const double o = os[(size_t)t * L + l];
        double negh = 0.0, lsw = 0.0;
        if (rd >= 0) { const double v = ls[(size_t)li * L + l]; negh = -0.5 / (v * v); lsw = o / (v * v * v); }
        pt->negh[t] = negh; pt->osc[t] = o; pt->lsw[t] = lsw;
        pt->nmask[t] = sp.n_mask[t]; pt->rbf[t] = rd >= 0; pt->lsidx[t] = li;
        for (int i = 0; i < 3; ++i) {
            const bool cat = sp.mask_type[t][i] == LVAE_CAT;
            pt->sgn[t][i] = cat ? -1.0 : 1.0;
            pt->tgt[t][i] = cat ? 0.0 : 2.0;
        }
    }
    if (t == 0) {
        int s = 0;
        for (int c = 0; c < nc; ++c) {
            pt->slot0[c] = s;
            if (sp.rbf_dim[c] >= 0) pt->dim[s++] = sp.rbf_dim[c];
            for (int i = 0; i < sp.n_mask[c]; ++i) pt->dim[s++] = sp.mask_dim[c][i];
        }
        pt->nslots = s;
        pt->noise = noise[l];
    }
    load_exp_table(pt->etab);
}

// un-scaled value of a component between rows i and j of the gathered covariates X (slot-major, row stride LDX)
template <int NM, bool RBF, int LDX>
__device__ __forceinline__ double eval_rr(const double* __restrict__ X, int i, int j, const double* __restrict__ sgn,
                                          const double* __restrict__ tgt, double negh, const double* __restrict__ etab,
                                          double& d2) {
    bool on = true;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const double a = X[((RBF ? 1 : 0) + m) * LDX + i], b = X[((RBF ? 1 : 0) + m) * LDX + j];
        on = on && (fma(sgn[m], b, a) == tgt[m]);
    }
    double e = 1.0;
    d2 = 0.0;
    if (RBF) {
        const double t = X[i] - X[j];
        d2 = t * t;
        e = exp_neg(d2 * negh, etab);
    }
    return on ? e : 0.0;
}

// fn(i, j, f, d2) for every lower-triangle entry (i, j) of this lane (entries gl, gl + NL, ... < ntri through the (i, j)
// table), component c.  The entry loops stay ROLLED: one body per shape and call site keeps the kernel small — the first,
// fully unrolled version of these loops stalled on instruction fetch (ncu: no_instruction 10 cycles per issue) as soon as the
// warps of an SM worked on subjects of different lengths.
template <int LDX, class Fn>
__device__ __forceinline__ void eval_tri(const PrepTab& pt, int c, const double* __restrict__ XC,
                                         const unsigned short* __restrict__ ijt, int gl, int nl, int ntri, Fn&& fn) {
    const double* X = XC + pt.slot0[c] * LDX;
    const double negh = pt.negh[c];
    const int nm = pt.nmask[c];
    const bool rbf = pt.rbf[c] != 0;
    double sgn[3], tgt[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { sgn[m] = pt.sgn[c][m]; tgt[m] = pt.tgt[c][m]; }
    double d2;
    if (nm == 0) {
#pragma unroll 1
        for (int e = gl; e < ntri; e += nl) {
            const int v = ijt[e], i = v & 255, j = v >> 8;
            const double f = eval_rr<0, true, LDX>(X, i, j, sgn, tgt, negh, pt.etab, d2);
            fn(i, j, f, d2);
        }
    } else if (nm == 1 && rbf) {
#pragma unroll 1
        for (int e = gl; e < ntri; e += nl) {
            const int v = ijt[e], i = v & 255, j = v >> 8;
            const double f = eval_rr<1, true, LDX>(X, i, j, sgn, tgt, negh, pt.etab, d2);
            fn(i, j, f, d2);
        }
    } else if (nm == 1) {
#pragma unroll 1
        for (int e = gl; e < ntri; e += nl) {
            const int v = ijt[e], i = v & 255, j = v >> 8;
            const double f = eval_rr<1, false, LDX>(X, i, j, sgn, tgt, negh, pt.etab, d2);
            fn(i, j, f, d2);
        }
    } else {                                   // 2-3 factors: run-time factor loop
        const int o = rbf ? 1 : 0;
#pragma unroll 1
        for (int e = gl; e < ntri; e += nl) {
            const int v = ijt[e], i = v & 255, j = v >> 8;
            bool on = true;
            for (int m = 0; m < nm; ++m) on = on && (fma(pt.sgn[c][m], X[(o + m) * LDX + j], X[(o + m) * LDX + i]) == pt.tgt[c][m]);
            double ex = 1.0;
            d2 = 0.0;
            if (rbf) {
                const double t = X[i] - X[j];
                d2 = t * t;
                ex = exp_neg(d2 * negh, pt.etab);
            }
            fn(i, j, on ? ex : 0.0, d2);
        }
    }
}

// ---- linear algebra on T x T matrices in shared memory, compile-time stride LD and tile count NT8 ------------------------------
template <int NT8, int LD, int NW, bool TU>
__device__ __forceinline__ int grp_cholesky(double* __restrict__ A, int T, double* __restrict__ dinv, int lane, int wg, int bar) {
    const int g = lane >> 2, q = lane & 3, gl = lane + 32 * wg;
    const int nb = (T + 7) >> 3;
    int bad = 0;
    int ur = 0;
    while ((ur + 1) * (ur + 2) / 2 <= lane) ++ur;
    const int uc = lane - ur * (ur + 1) / 2 + 1;
    ur += 1;
#pragma unroll(TU ? NT8 : 1)
    for (int kb = 0; kb < NT8; ++kb) {
        if (kb < nb) {
            const int k0 = 8 * kb, bs = min(8, T - k0);
            if (wg == 0) {
                for (int k = 0; k < bs; ++k) {
                    const double akk = A[(k0 + k) * LD + k0 + k];
                    if (!(akk > 0.0) && bad == 0) bad = k0 + k + 1;
                    const double ri = rsqrt(akk);
                    __syncwarp();
                    if (lane == k) { A[(k0 + k) * LD + k0 + k] = akk * ri; dinv[k0 + k] = ri; }
                    if (lane > k && lane < bs) A[(k0 + lane) * LD + k0 + k] *= ri;
                    __syncwarp();
                    if (lane < 28 && uc > k && ur < bs)
                        A[(k0 + ur) * LD + k0 + uc] -= A[(k0 + ur) * LD + k0 + k] * A[(k0 + uc) * LD + k0 + k];
                    __syncwarp();
                }
            }
            if (kb + 1 < nb) {
                gsync<NW>(bar);
                for (int r = k0 + 8 + gl; r < T; r += 32 * NW) {
                    double xr[8];
#pragma unroll
                    for (int c_ = 0; c_ < 8; ++c_) {
                        double s = A[r * LD + k0 + c_];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < c_) s -= xr[k] * A[(k0 + c_) * LD + k0 + k];
                        xr[c_] = s * dinv[k0 + c_];
                    }
#pragma unroll
                    for (int c_ = 0; c_ < 8; ++c_) A[r * LD + k0 + c_] = xr[c_];
                }
                gsync<NW>(bar);
                int tl = 0;
#pragma unroll(TU ? NT8 : 1)
                for (int ti = kb + 1; ti < NT8; ++ti) {
#pragma unroll(TU ? NT8 : 1)
                    for (int tj = kb + 1; tj <= ti; ++tj, ++tl) {
                        if (ti < nb && (NW == 1 || (tl % NW) == wg)) {
                            const int i = 8 * ti + g, j = 8 * tj + 2 * q;
                            double c0 = 0.0, c1 = 0.0;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                dmma(c0, c1, A[i * LD + k0 + 4 * ks + q], A[(8 * tj + g) * LD + k0 + 4 * ks + q]);
                            if (i < T && j < T) A[i * LD + j] -= c0;
                            if (i < T && j + 1 < T) A[i * LD + j + 1] -= c1;
                        }
                    }
                }
                gsync<NW>(bar);
            }
        }
    }
    gsync<NW>(bar);
    return bad;
}

template <int NT8, int LD, int NW, bool TU>
__device__ __forceinline__ void grp_tri_inverse(const double* __restrict__ Lc, double* __restrict__ X, int T,
                                                const double* __restrict__ dinv, double* __restrict__ tile, int lane, int wg,
                                                int bar) {
    const int g = lane >> 2, q = lane & 3;
    const int nb = (T + 7) >> 3;
    for (int j = lane + 32 * wg; j < T; j += 32 * NW) {
        const int kend = min(T, (j & ~7) + 8);
        X[j * LD + j] = dinv[j];
        for (int i = j + 1; i < kend; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s += Lc[i * LD + k] * X[k * LD + j];
            X[i * LD + j] = -s * dinv[i];
        }
    }
    gsync<NW>(bar);
#pragma unroll(TU ? NT8 : 1)
    for (int d = 1; d < NT8; ++d) {
        if (d < nb) {
            for (int bj = wg; bj + d < nb; bj += NW) {
                const int bi = bj + d;
                double t0 = 0.0, t1 = 0.0;
                for (int bk = bj; bk < bi; ++bk) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        dmma(t0, t1, Lc[(8 * bi + g) * LD + 8 * bk + 4 * ks + q], X[(8 * bk + 4 * ks + q) * LD + 8 * bj + g]);
                }
                tile[g * 8 + 2 * q] = t0;
                tile[g * 8 + 2 * q + 1] = t1;
                __syncwarp();
                double x0 = 0.0, x1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    dmma(x0, x1, -X[(8 * bi + g) * LD + 8 * bi + 4 * ks + q], tile[(4 * ks + q) * 8 + g]);
                const int i = 8 * bi + g, j = 8 * bj + 2 * q;
                if (i < T) { X[i * LD + j] = x0; X[i * LD + j + 1] = x1; }
                __syncwarp();
            }
            gsync<NW>(bar);
        }
    }
}

// C = op(A) B on the T x T leading blocks (TA: op(A)(i,k) = A[k][i]; SYM: lower tiles only, mirrored); rows of tiles dealt to warps
template <bool TA, bool SYM, int NT8, int LD, int NW, bool TU>
__device__ __forceinline__ void grp_mm(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int T,
                                       int nt8, int nk4, int g, int q, int wg) {
#pragma unroll(TU ? NT8 : 1)
    for (int ti = 0; ti < NT8; ++ti) {
        if (ti < nt8 && (NW == 1 || (ti % NW) == wg)) {
            double acc[NT8][2];
#pragma unroll
            for (int tj = 0; tj < NT8; ++tj) acc[tj][0] = acc[tj][1] = 0.0;
            for (int ks = 0; ks < nk4; ++ks) {
                const double a = TA ? A[(4 * ks + q) * LD + 8 * ti + g] : A[(8 * ti + g) * LD + 4 * ks + q];
#pragma unroll
                for (int tj = 0; tj < NT8; ++tj) {
                    if (tj < nt8 && (!SYM || tj <= ti)) dmma(acc[tj][0], acc[tj][1], a, B[(4 * ks + q) * LD + 8 * tj + g]);
                }
            }
            const int i = 8 * ti + g;
#pragma unroll
            for (int tj = 0; tj < NT8; ++tj) {
                if (tj < nt8 && (!SYM || tj <= ti)) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * tj + 2 * q + e;
                        if (i < T && j < T) {
                            C[i * LD + j] = acc[tj][e];
                            if (SYM && tj < ti) C[j * LD + i] = acc[tj][e];
                        }
                    }
                }
            }
        }
    }
}

// rows allocated per matrix: T_max rounded up to 4 (tile rows beyond that only feed outputs that are never stored)
template <int NT8, int LD>
__host__ __device__ constexpr int rows3() { return LD < 8 * NT8 ? LD : 8 * NT8; }
template <int NT8, int LD, int NW>
__host__ __device__ constexpr int group_doubles3(int nslots_max) {
    return 3 * rows3<NT8, LD>() * LD + nslots_max * LD + 3 * (8 * NT8) + 64 * NW + NW * (2 * NCM + 2);
}

// TU: unroll the tile loops of the T x T linear algebra (pays for 3 x 3 tiles, not for 5 x 5: code size)
template <int NT8, int LD, int NW, bool TU>
__global__ void __launch_bounds__(NW == 1 ? 256 : 512, NW == 1 ? 2 : 1)
k_prep3(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int Q, int P_b, int N_b, int nslots_max,
        const double* __restrict__ x, const int32_t* __restrict__ offsets, const double* __restrict__ mu,
        const double* __restrict__ log_v, const double* __restrict__ ls, const double* __restrict__ os,
        const double* __restrict__ noise, double c, double* __restrict__ d_log_v, double* __restrict__ ws, int32_t* info) {
    constexpr int TP8 = 8 * NT8, NL = 32 * NW, ROWS = rows3<NT8, LD>(), TRI = ROWS * (ROWS + 1) / 2, ASZ = ROWS * LD;
    static_assert(LD >= 8 * NT8 && ROWS == 8 * NT8, "scratch matrices must cover every row / column a tile touches");
    extern __shared__ double sm[];
    __shared__ PrepTab pt;
    __shared__ unsigned short ijt[TRI];
    const int l = blockIdx.y, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int grp = wid / NW, wg = wid % NW, gl = lane + 32 * wg, bar = 1 + grp;
    const int nc = sp.n0 + sp.n1, nh = hyp_count(sp);
    load_preptab(&pt, sp, ls, os, noise, L, l);
    for (int e = tid; e < TRI; e += blockDim.x) {          // (i, j), j <= i, of the e-th element of a lower triangle
        int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        ijt[e] = (unsigned short)(i | ((e - i * (i + 1) / 2) << 8));
    }
    const int gsz = group_doubles3<NT8, LD, NW>(nslots_max);
    double* A1 = sm + (size_t)grp * gsz;
    double* A2 = A1 + ASZ;
    double* A3 = A2 + ASZ;
    double* XC = A3 + ASZ;                     // [nslots][LD] gathered covariates of the task
    double* ev = XC + nslots_max * LD;
    double* mw = ev + TP8;
    double* dinv = mw + TP8;
    double* tile = dinv + TP8 + 64 * wg;
    double* hacc = dinv + TP8 + 64 * NW + (2 * NCM + 2) * wg;   // [2 * NCM + 2] hyper-gradient sums of this WARP (lane 0 adds)
    for (int e = gl; e < 3 * ASZ; e += NL) A1[e] = 0.0;
    if (lane < 2 * NCM + 2) hacc[lane] = 0.0;
    __syncthreads();
    const int nslots = pt.nslots;

    double sC = 0.0, sD1 = 0.0, sBt = 0.0, sF = 0.0, gno = 0.0;
    const int GPC = (blockDim.x >> 5) / NW;
    const int ngrp = gridDim.x * GPC, gg = blockIdx.x * GPC + grp;
    int Tprev = -1;
    for (int p = gg; p < P_b; p += ngrp) {
        const int r0 = offsets[p], T = offsets[p + 1] - r0;
        const int nt8 = (T + 7) >> 3, nk4 = (T + 3) >> 2, ntri = T * (T + 1) / 2;
        gsync<NW>(bar);
        if (Tprev != -1 && T != Tprev) {
            for (int e = gl; e < 3 * ASZ; e += NL) A1[e] = 0.0;
        }
        Tprev = T;
        for (int e = gl; e < nslots * T; e += NL) {
            const int s = e / T, t = e - s * T;
            XC[s * LD + t] = x[(size_t)(r0 + t) * Q + pt.dim[s]];
        }
        for (int t = gl; t < T; t += NL) {
            const double lv = log_v[(size_t)(r0 + t) * L + l];
            ev[t] = exp(lv);
            sF += lv;
            mw[t] = mu[(size_t)(r0 + t) * L + l];
        }
        gsync<NW>(bar);
        // ---- B_p = K1 + noise I: every lane accumulates ITS lower-triangle entries in place over the components, then mirrors ----
        for (int cc = sp.n0; cc < nc; ++cc) {
            const double o = pt.osc[cc];
            const bool first = cc == sp.n0;
            eval_tri<LD>(pt, cc, XC, ijt, gl, NL, ntri, [&](int i, int j, double f, double) {
                A1[i * LD + j] = (first ? 0.0 : A1[i * LD + j]) + o * f;
            });
        }
#pragma unroll 1
        for (int e = gl; e < ntri; e += NL) {
            const int v_ = ijt[e], i = v_ & 255, j = v_ >> 8;
            const double v = A1[i * LD + j] + (i == j ? pt.noise : 0.0);
            A1[i * LD + j] = v;
            A1[j * LD + i] = v;
        }
        for (int e = gl; e < ASZ; e += NL) A2[e] = 0.0;
        gsync<NW>(bar);
        {
            const int bad = grp_cholesky<NT8, LD, NW, TU>(A1, T, dinv, lane, wg, bar);
            if (bad && wg == 0 && lane == 0) atomicCAS(info + 2, 0, l * P_b + p + 1);
        }
        for (int t = gl; t < T; t += NL) sC -= 2.0 * log(dinv[t]);                                            // 192
        grp_tri_inverse<NT8, LD, NW, TU>(A1, A2, T, dinv, tile, lane, wg, bar);
        if (w.v2) {   // rows of L^-1 for the subject pass: [row][k'], k' = column inside the subject, zero padded to TP
            double* gl_ = ws + w.Lrows + ((size_t)l * N_b + r0) * w.TP;
            int i = 0, k = gl;
            while (k >= w.TP) { k -= w.TP; ++i; }
            for (int e = gl; e < T * w.TP; e += NL) {
                gl_[e] = (k <= i) ? A2[i * LD + k] : 0.0;
                k += NL;
                while (k >= w.TP) { k -= w.TP; ++i; }
            }
        }
        grp_mm<true, true, NT8, LD, NW, TU>(A2, A2, A3, T, nt8, nk4, g, q, wg);                                     // B^-1 = L^-T L^-1
        gsync<NW>(bar);
        if (w.v3) {   // third-generation subject pass: also the rows of L^-T ([row a][k'] = L^-1[k'][a], k' >= a), zero padded to TP
            double* gt_ = ws + w.Ltrows + ((size_t)l * N_b + r0) * w.TP;
            int i = 0, k = gl;
            while (k >= w.TP) { k -= w.TP; ++i; }
            for (int e = gl; e < T * w.TP; e += NL) {
                gt_[e] = (k >= i && k < T) ? A2[k * LD + i] : 0.0;
                k += NL;
                while (k >= w.TP) { k -= w.TP; ++i; }
            }
        }
        // ---- K0_p (+ diag v) into A1 (accumulated in place over the components) ; D1 ; adjoint of K0 = c B^-1 against d k_c / d theta
        for (int cc = 0; cc < sp.n0; ++cc) {
            const double o = pt.osc[cc];
            const bool first = cc == 0;
            double s1 = 0.0, s2 = 0.0;
            eval_tri<LD>(pt, cc, XC, ijt, gl, NL, ntri, [&](int i, int j, double f, double d2) {
                A1[i * LD + j] = (first ? 0.0 : A1[i * LD + j]) + o * f;
                const double w_ = A3[i * LD + j] * (i == j ? 1.0 : 2.0) * f;
                s1 += w_;
                s2 += w_ * d2;
            });
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) { hacc[cc] += s1; hacc[NCM + cc] += s2; }
        }
#pragma unroll 1
        for (int e = gl; e < ntri; e += NL) {
            const int v_ = ijt[e], i = v_ & 255, j = v_ >> 8;
            const double k0 = A1[i * LD + j];
            sD1 += A3[i * LD + j] * (i == j ? 1.0 : 2.0) * k0;
            const double v = k0 + (i == j ? ev[i] : 0.0);
            A1[i * LD + j] = v;
            A1[j * LD + i] = v;
        }
        gsync<NW>(bar);
        grp_mm<false, false, NT8, LD, NW, TU>(A1, A3, A2, T, nt8, nk4, g, q, wg);                                   // X1 = (diag v + K0) B^-1
        gsync<NW>(bar);
        grp_mm<false, false, NT8, LD, NW, TU>(A3, A2, A1, T, nt8, nk4, g, q, wg);                                   // X2 = B^-1 X1
        gsync<NW>(bar);
        if (w.v2 && !w.v3) {          // B^-1 mu for the GEMM-based subject pass (the fused pass carries mu through its own products)
            double* gb = ws + w.bmu + (size_t)l * N_b + r0;
            for (int t = gl; t < T; t += NL) {
                double s = 0.0;
                for (int k = 0; k < T; ++k) s += A3[t * LD + k] * mw[k];
                gb[t] = s;
            }
        } else if (!w.v2) {      // generic subject pass: the explicit inverse blocks
            double* gBi = ws + w.Bi + (size_t)l * w.Bi_stride + reinterpret_cast<const int64_t*>(ws + w.off2)[p];
            int i = 0, j = gl;
            while (j >= T) { j -= T; ++i; }
            for (int e = gl; e < T * T; e += NL) {
                gBi[e] = A3[i * LD + j];
                j += NL;
                while (j >= T) { j -= T; ++i; }
            }
        }
        // ---- local adjoint of B_p: (B^-1 - X2) [times c at the end] -> A2 (X1 is dead), then against d K1 / d theta ; noise ; Bt ; d_log_v
#pragma unroll 1
        for (int e = gl; e < ntri; e += NL) {
            const int v_ = ijt[e], i = v_ & 255, j = v_ >> 8;
            const double b_ = A3[i * LD + j];
            double gB;
            if (i == j) {
                gB = b_ - A1[i * LD + i];
                gno += gB;
                const double bt = b_ * ev[i];
                sBt += bt;
                d_log_v[(size_t)(r0 + i) * L + l] = c * (bt - 1.0);
            } else {
                gB = 2.0 * b_ - (A1[i * LD + j] + A1[j * LD + i]);
            }
            A2[i * LD + j] = gB;
        }
        for (int cc = sp.n0; cc < nc; ++cc) {
            double s1 = 0.0, s2 = 0.0;
            eval_tri<LD>(pt, cc, XC, ijt, gl, NL, ntri, [&](int i, int j, double f, double d2) {
                const double w_ = A2[i * LD + j] * f;
                s1 += w_;
                s2 += w_ * d2;
            });
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) { hacc[cc] += s1; hacc[NCM + cc] += s2; }
        }
    }
    // ---- one partial row per warp (summed in fixed order by the reduce kernel) ------------------------------------------------
    __syncwarp();
    const int gw = blockIdx.x * (blockDim.x >> 5) + wid;
    double* out = ws + w.ppart + ((size_t)gw * L + l) * (LVAE_NSCAL + nh);
    sC = warp_sum(sC); sD1 = warp_sum(sD1); sBt = warp_sum(sBt); sF = warp_sum(sF); gno = warp_sum(gno);
    if (lane == 0) {
        for (int k = 0; k < LVAE_NSCAL + nh; ++k) out[k] = 0.0;
        out[SC_C] = sC; out[SC_D1] = sD1; out[SC_BT] = sBt; out[SC_F] = sF;
        out[LVAE_NSCAL + nh - 1] = c * gno;
        for (int cc = 0; cc < nc; ++cc) {
            out[LVAE_NSCAL + sp.n_ls + cc] = c * hacc[cc];
            if (pt.rbf[cc]) out[LVAE_NSCAL + pt.lsidx[cc]] += c * hacc[NCM + cc] * pt.lsw[cc];
        }
    }
}

int slots_of(const DevSpec& sp) {
    int s = 0;
    for (int c = 0; c < sp.n0 + sp.n1; ++c) s += (sp.rbf_dim[c] >= 0) + sp.n_mask[c];
    return s;
}

template <int NT8, int LD, int NW>
int groups_per_cta3(int nslots) {
    const size_t bytes = sizeof(double) * group_doubles3<NT8, LD, NW>(nslots);
    // NW == 1: two CTAs per SM.  (Measured: a sixth task per CTA, 2 x 113 KB, no longer fits two CTAs next to the kernel's
    // static shared memory — prep went from 0.60 to 0.87 ms at cfg2.)
    int n = (int)((NW == 1 ? 108 * 1024 : 200 * 1024) / bytes);
    const int cap = NW == 1 ? 8 : 4;
    if (n > cap) n = cap;
    return n < 1 ? 1 : n;
}

template <int NT8, int LD, int NW, bool TU>
int launch3(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int nslots = slots_of(sp);
    const int gpc = groups_per_cta3<NT8, LD, NW>(nslots), pw = gpc * NW;
    const size_t smem = sizeof(double) * (size_t)gpc * group_doubles3<NT8, LD, NW>(nslots);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_prep3<NT8, LD, NW, TU>, smem, attr)) return rc_;
    if (w.nprep % pw != 0) return LVAE_E_BADARG;
    k_prep3<NT8, LD, NW, TU><<<dim3(w.nprep / pw, p->L), pw * 32, smem, st>>>(sp, w, p->L, p->Q, p->P_b, p->N_b, nslots, p->x, p->offsets,
                                                                          p->mu, p->log_v, p->lengthscale, p->outputscale,
                                                                          p->noise, 0.5 * p->scale, p->d_log_v, p->workspace,
                                                                          p->info);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

// prep kernel: <= 8 components, T <= 40 (one warp per task up to 24 rows, four warps per task beyond)
bool lvae_prep3_supported(const lvae_kld_problem_t* p, const KldLayout& w) {
    (void)w;
    return p->ks.n_comp0 + p->ks.n_comp1 <= NCM && p->T_max <= 40 && p->T_max >= 1;
}

// partial rows per latent (= warps per latent): about one wave of CTAs
int lvae_prep3_rows(const lvae_kld_problem_t* p) {
    int ns = 0;
    for (int c_ = 0; c_ < p->ks.n_comp0 + p->ks.n_comp1; ++c_)
        ns += (p->ks.spec[(size_t)c_ * LVAE_SPEC_STRIDE] >= 0) + p->ks.spec[(size_t)c_ * LVAE_SPEC_STRIDE + 2];
    const int T = p->T_max;
    const int nw = T <= 24 ? 1 : 4;
    const int gpc = T <= 24 ? groups_per_cta3<3, 24, 1>(ns) : groups_per_cta3<5, 44, 4>(ns);
    const int pw = gpc * nw, per_sm = nw == 1 ? 2 : 1;
    int ctas = per_sm * 148 / p->L;
    if (ctas < 1) ctas = 1;
    const int need = (p->P_b + gpc - 1) / gpc;
    if (ctas > need) ctas = need > 0 ? need : 1;
    return ctas * pw;
}

int lvae_prep3_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int T = p->T_max;
    // The scratch matrices must hold every row and column a tile can touch (ROWS = 8 * NT8, LD >= 8 * NT8).  A former
    // <3, 20, 1> instantiation (20 x 20 scratch for T <= 20) let the 8-wide tiles run past column 19 into the next row and past
    // row 19 into the next matrix: correct on zero-initialised scratch, i.e. for the FIRST task of a warp only — the full-size
    // parity check of round 2 caught the later tasks (rows 0..3 of d_log_v, L^-1) being wrong.
    if (T <= 24) return launch3<3, 24, 1, true>(p, sp, w, st);
    return launch3<5, 44, 4, false>(p, sp, w, st);
}
