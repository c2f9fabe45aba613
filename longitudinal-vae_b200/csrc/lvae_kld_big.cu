// Per-latent M x M work of the GP-prior ELBO path for 64 < M <= 256: launch sequences over the batched DMMA GEMM and the
// blocked Cholesky / inverse of lvae_blas.cu, with small element-wise kernels in between.  All M x M operands live in
// matrices padded to MP = 128 or 256 (identity padding while factoring, zero padding in products).  Same outputs and
// statistics layout as the M <= 64 kernels (lvae_kld64.cu) and the generic ones (lvae_kld.cu).
//   head: Kzz + eps I and H -> Cholesky -> inverses (both as ONE batch of 2L matrices), log-dets, a = Kzz^-1 m,
//         G = Kzz^-1 H Kzz^-1, W = c (sym G - Kzz^-1)                      (elbo_functions.py:172,176-178,185-186,194)
//   tail: D, E, KL[q(u)||p(u)], kld, grad_m, grad_H (or d_m, d_H), adjoint of Kzz -> hyper-gradients      (193-214)
//   ng  : natural-gradient update of (m, H)                                               (training.py:129-135)
#include "lvae_blas.h"
#include "lvae_kld.h"

namespace {

struct HypB {
    double hil2[LVAE_MAXC], il3[LVAE_MAXC], os[LVAE_MAXC], etab[LVAE_EXP_TBL];
};
__device__ inline void load_hypb(HypB* h, const DevSpec& sp, const double* ls, const double* os, int L, int l) {
    const int t = threadIdx.x;
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; h->hil2[t] = 0.5 / (v * v); h->il3[t] = 1.0 / (v * v * v); }
    if (t < sp.n0 + sp.n1) h->os[t] = os[(size_t)t * L + l];
    load_exp_table(h->etab);
}

// y[i] = sum_k A[i*ld + k] x[k], i < n: one warp per row (coalesced), x in shared memory.  No barrier inside.
__device__ inline void cta_gemv_rows(const double* __restrict__ A, int ld, const double* __restrict__ x, int n, int ncols,
                                     double* __restrict__ y) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = wid; i < n; i += nw) {
        double s = 0.0;
        for (int k = lane; k < ncols; k += 32) s += A[(size_t)i * ld + k] * x[k];
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
}

// F[l] = Kzz + eps I, F[L + l] = H (identity padded); Hp[l] = H (zero padded).   grid (blocks, L)
__global__ void __launch_bounds__(256) k_big_fill(const __grid_constant__ DevSpec sp, KldLayout w, int L, int M, int Q,
                                                  const double* __restrict__ z, const double* __restrict__ H,
                                                  const double* __restrict__ ls, const double* __restrict__ os, double eps,
                                                  double* __restrict__ ws) {
    __shared__ HypB hyp;
    const int l = blockIdx.y, MP = w.MP;
    load_hypb(&hyp, sp, ls, os, L, l);
    __syncthreads();
    const double* zl = z + (size_t)l * M * Q;
    const double* Hl = H + (size_t)l * M * M;
    double* FK = ws + w.bF + (size_t)l * MP * MP;
    double* FH = ws + w.bF + (size_t)(L + l) * MP * MP;
    double* Hp = ws + w.bHp + (size_t)l * MP * MP;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MP * MP; e += gridDim.x * blockDim.x) {
        const int i = e / MP, j = e % MP;
        double kz = (i == j) ? 1.0 : 0.0, hh = kz, hz = 0.0;
        if (i < M && j < M) {
            double acc = 0.0, d2;
            for (int cc = 0; cc < sp.n0; ++cc) acc += hyp.os[cc] * comp_value(sp, cc, zl + i * Q, zl + j * Q, hyp.hil2, d2, hyp.etab);
            kz = acc + (i == j ? eps : 0.0);                                           // elbo_functions.py:172,176
            hh = hz = Hl[(size_t)i * M + j];
        }
        FK[e] = kz; FH[e] = hh; Hp[e] = hz;
    }
}

// log-dets from the factors; grid 2L
__global__ void __launch_bounds__(256) k_big_logdet(KldLayout w, int L, int M, double* __restrict__ ws) {
    __shared__ double red[32];
    const int b = blockIdx.x, MP = w.MP;
    const double* F = ws + w.bF + (size_t)b * MP * MP;
    double v = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) v += log(F[(size_t)i * MP + i]);
    v = block_sum(v, red);
    if (threadIdx.x == 0) ws[w.logdet + 2 * (b % L) + b / L] = 2.0 * v;
}

// explicit inverses: copy the inverse (bA0, identity padded, symmetric) into bInv with ZERO padding, export the un-padded Kzz^-1 / H^-1.
// grid (blocks, 2L)
__global__ void __launch_bounds__(256) k_big_unpad_inv(KldLayout w, int L, int M, double* __restrict__ ws) {
    const int b = blockIdx.y, MP = w.MP;
    const double* Xr = ws + w.bA0 + (size_t)b * MP * MP;
    double* Inv = ws + w.bInv + (size_t)b * MP * MP;
    double* out = ws + (b < L ? w.Ki + (size_t)b * M * M : w.Hi + (size_t)(b - L) * M * M);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MP * MP; e += gridDim.x * blockDim.x) {
        const int i = e / MP, j = e % MP;
        double v = 0.0;
        if (i < M && j < M) {
            v = 0.5 * (Xr[e] + Xr[(size_t)j * MP + i]);
            out[(size_t)i * M + j] = v;
        }
        Inv[e] = v;
    }
}

// a = Kzz^-1 m ; grid L
__global__ void __launch_bounds__(256) k_big_a(KldLayout w, int M, const double* __restrict__ m, double* __restrict__ ws) {
    __shared__ double ms[LVAE_MAX_M];
    const int l = blockIdx.x, MP = w.MP;
    for (int i = threadIdx.x; i < M; i += blockDim.x) ms[i] = m[(size_t)l * M + i];
    __syncthreads();
    cta_gemv_rows(ws + w.bInv + (size_t)l * MP * MP, MP, ms, M, M, ws + w.a + (size_t)l * M);
}

// G (un-padded) and W = c (sym G - Kzz^-1) (padded, zero outside M x M); Gp = bT2.   grid (blocks, L)
__global__ void __launch_bounds__(256) k_big_W(KldLayout w, int M, double c, double* __restrict__ ws) {
    const int l = blockIdx.y, MP = w.MP;
    const double* Gp = ws + w.bT2 + (size_t)l * MP * MP;
    const double* Ki = ws + w.bInv + (size_t)l * MP * MP;
    double* G = ws + w.G + (size_t)l * M * M;
    double* Wp = ws + w.bWp + (size_t)l * MP * MP;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MP * MP; e += gridDim.x * blockDim.x) {
        const int i = e / MP, j = e % MP;
        double v = 0.0;
        if (i < M && j < M) {
            G[(size_t)i * M + j] = Gp[e];
            v = c * (0.5 * (Gp[e] + Gp[(size_t)j * MP + i]) - Ki[e]);
        }
        Wp[e] = v;
    }
}

// padded copy of the (all-reduced) S statistics.   grid (blocks, L)
__global__ void __launch_bounds__(256) k_big_S_in(KldLayout w, int M, const double* __restrict__ stats, double* __restrict__ ws) {
    const int l = blockIdx.y, MP = w.MP;
    const double* S = stats + (size_t)l * w.stride + stats_off_S();
    double* Sp = ws + w.bS + (size_t)l * MP * MP;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < MP * MP; e += gridDim.x * blockDim.x) {
        const int i = e / MP, j = e % MP;
        Sp[e] = (i < M && j < M) ? S[(size_t)i * M + j] : 0.0;
    }
}

// tail, first half: scalars -> kld ; grad_m / grad_H ; adjoint of Kzz^-1 written over Sp.   grid L, 1024 threads
// inputs: T1 = Ki S, T2 = Ki S Ki, T3 = H Ki S
__global__ void __launch_bounds__(1024) k_big_tail_a(KldLayout w, int L, int M, int natural_gradient, const double* __restrict__ m,
                                                    double c, double const_per_latent, const double* __restrict__ stats,
                                                    double* __restrict__ ws, double* __restrict__ kld,
                                                    double* __restrict__ grad_m, double* __restrict__ grad_H) {
    __shared__ double red[32], ms[LVAE_MAX_M], ga[LVAE_MAX_M], v1[LVAE_MAX_M], v2[LVAE_MAX_M], v3[LVAE_MAX_M];
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, MP = w.MP;
    const size_t mo = (size_t)l * MP * MP;
    const double* st = stats + (size_t)l * w.stride;
    const double* sc = st + stats_off_scal(M);
    const double* Ki = ws + w.bInv + mo;
    const double* Hi = ws + w.bInv + (size_t)(L + l) * MP * MP;
    const double* Hp = ws + w.bHp + mo;
    const double* G = ws + w.G + (size_t)l * M * M;
    const double* T2 = ws + w.bT2 + mo;
    const double* T3 = ws + w.bT3 + mo;
    const double* a = ws + w.a + (size_t)l * M;
    double* Sp = ws + w.bS + mo;
    for (int i = tid; i < M; i += nt) {
        ms[i] = m[(size_t)l * M + i];
        ga[i] = 2.0 * c * st[stats_off_da(M) + i];
        v3[i] = st[stats_off_ng1(M) + i];
    }
    __syncthreads();
    double d2 = 0.0, ee = 0.0, tr = 0.0, qf = 0.0;
    for (int e = tid; e < M * M; e += nt) {
        const int i = e / M, j = e % M;
        const double s = Sp[(size_t)i * MP + j];
        d2 += s * Ki[(size_t)i * MP + j];                                              // 193
        ee += G[(size_t)j * M + i] * s;                                                // 195 / 282
        tr += Ki[(size_t)i * MP + j] * Hp[(size_t)j * MP + i];                         // 199
    }
    for (int i = tid; i < M; i += nt) qf += ms[i] * a[i];                              // 200
    d2 = block_sum(d2, red);
    ee = block_sum(ee, red);
    tr = block_sum(tr, red);
    qf = block_sum(qf, red);
    if (tid == 0) {
        const double ldK = ws[w.logdet + 2 * l], ldH = ws[w.logdet + 2 * l + 1];
        const double kl_qp = 0.5 * (tr + qf - M + ldK - ldH);                                          // 199-203
        kld[l] = c * (sc[SC_A] + sc[SC_BT] + sc[SC_C] + sc[SC_D1] - d2 + ee - sc[SC_F]) + kl_qp - const_per_latent;   // 204
    }
    double* gm = grad_m + (size_t)l * M;
    double* gH = grad_H + (size_t)l * M * M;
    if (natural_gradient) {                                                            // 208-214, 301-305
        cta_gemv_rows(Ki, MP, v3, M, M, v1);                                           // Ki ng1
        cta_gemv_rows(T2, MP, ms, M, M, v2);                                           // Ki S Ki m
        __syncthreads();
        for (int i = tid; i < M; i += nt) gm[i] = -v1[i] + v2[i] + a[i];
        for (int e = tid; e < M * M; e += nt) {
            const int i = e / M, j = e % M;
            gH[e] = 0.5 * (T2[(size_t)i * MP + j] + Ki[(size_t)i * MP + j] - Hi[(size_t)i * MP + j]);
        }
    } else {                                                                           // autograd of kld_total
        cta_gemv_rows(Ki, MP, ga, M, M, v1);
        __syncthreads();
        for (int i = tid; i < M; i += nt) gm[i] = v1[i] + a[i];
        for (int e = tid; e < M * M; e += nt) {
            const int i = e / M, j = e % M;
            gH[e] = c * 0.5 * (T2[(size_t)i * MP + j] + T2[(size_t)j * MP + i]) + 0.5 * Ki[(size_t)i * MP + j] -
                    0.5 * Hi[(size_t)i * MP + j];
        }
    }
    // adjoint of Kzz^-1, in place over Sp:  -c S + c (HP + HP^T) + (H^T + m m^T)/2 + ga m^T
    for (int e = tid; e < M * M; e += nt) {
        const int i = e / M, j = e % M;
        const size_t ij = (size_t)i * MP + j, ji = (size_t)j * MP + i;
        Sp[ij] = -c * Sp[ij] + c * (T3[ij] + T3[ji]) + 0.5 * (Hp[ji] + ms[i] * ms[j]) + ga[i] * ms[j];
    }
}

// tail, second half: adjoint of Kzz = -Ki gKi Ki + Ki/2 (symmetrised; T3 = Ki gKi Ki) against d k_c / d theta on (Z, Z)
__global__ void __launch_bounds__(1024) k_big_tail_b(const __grid_constant__ DevSpec sp, KldLayout w, int L, int M, int Q,
                                                    const double* __restrict__ z, const double* __restrict__ ls,
                                                    const double* __restrict__ os, const double* __restrict__ stats,
                                                    const double* __restrict__ ws, double* __restrict__ d_ls,
                                                    double* __restrict__ d_os, double* __restrict__ d_noise) {
    __shared__ HypB hyp;
    __shared__ double red[32];
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, MP = w.MP, nh = hyp_count(sp);
    load_hypb(&hyp, sp, ls, os, L, l);
    __syncthreads();
    const double* T3 = ws + w.bT3 + (size_t)l * MP * MP;
    const double* Ki = ws + w.bInv + (size_t)l * MP * MP;
    const double* hy = stats + (size_t)l * w.stride + stats_off_hyp(M);
    const double* zl = z + (size_t)l * M * Q;
    double acc[2 * LVAE_MAXC + 1];
    for (int k = 0; k < nh; ++k) acc[k] = 0.0;
    for (int e = tid; e < M * M; e += nt) {
        const int i = e / M, j = e % M;
        const double gK = -0.5 * (T3[(size_t)i * MP + j] + T3[(size_t)j * MP + i]) + 0.5 * Ki[(size_t)i * MP + j];
        for (int cc = 0; cc < sp.n0; ++cc) {
            double dd;
            const double f = comp_value(sp, cc, zl + i * Q, zl + j * Q, hyp.hil2, dd, hyp.etab);
            acc[sp.n_ls + cc] += gK * f;
            if (sp.rbf_dim[cc] >= 0) acc[sp.ls_idx[cc]] += gK * hyp.os[cc] * f * dd * hyp.il3[sp.ls_idx[cc]];
        }
    }
    const int ncmp = sp.n0 + sp.n1;
    for (int k = 0; k < nh; ++k) {
        const double t = block_sum(acc[k], red) + hy[k];
        if (tid == 0) {
            if (k < sp.n_ls) d_ls[(size_t)k * L + l] = t;
            else if (k < sp.n_ls + ncmp) d_os[(size_t)(k - sp.n_ls) * L + l] = t;
            else d_noise[l] = t;
        }
    }
}

// natural-gradient step, first half: F = iH + lr (gH + gH^T) (identity padded) ; v1 = iH m - lr (g_m - 2 gH m).  grid L
__global__ void __launch_bounds__(1024) k_big_ng_a(int M, int MP, const double* __restrict__ m, const double* __restrict__ grad_m,
                                                  const double* __restrict__ grad_H, const double* __restrict__ iH, int ldi,
                                                  int64_t istride, double lr, double* __restrict__ F, double* __restrict__ v1g) {
    __shared__ double ms[LVAE_MAX_M], t1[LVAE_MAX_M], t2[LVAE_MAX_M];
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* iHl = iH + (size_t)l * istride;
    const double* gH = grad_H + (size_t)l * M * M;
    double* Fl = F + (size_t)l * MP * MP;
    for (int i = tid; i < M; i += nt) ms[i] = m[(size_t)l * M + i];
    __syncthreads();
    cta_gemv_rows(iHl, ldi, ms, M, M, t1);
    cta_gemv_rows(gH, M, ms, M, M, t2);
    for (int e = tid; e < MP * MP; e += nt) {
        const int i = e / MP, j = e % MP;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < M && j < M) v = iHl[(size_t)i * ldi + j] + lr * (gH[(size_t)i * M + j] + gH[(size_t)j * M + i]);   // training.py:132
        Fl[e] = v;
    }
    __syncthreads();
    for (int i = tid; i < M; i += nt) v1g[(size_t)l * M + i] = t1[i] - lr * (grad_m[(size_t)l * M + i] - 2.0 * t2[i]);   // 135
}

// un-padded symmetrised copy: dst[l] (M x M) = (src[l] + src[l]^T) / 2 (src padded np x np).  grid (blocks, L)
__global__ void __launch_bounds__(256) k_big_sym_out(double* __restrict__ dst, const double* __restrict__ src, int M, int MP) {
    const int l = blockIdx.y;
    const double* S = src + (size_t)l * MP * MP;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * M; e += gridDim.x * blockDim.x) {
        const int i = e / M, j = e % M;
        dst[(size_t)l * M * M + e] = 0.5 * (S[(size_t)i * MP + j] + S[(size_t)j * MP + i]);
    }
}

// second half: H <- sym(Xr) (un-padded), m <- H v1.  grid L
__global__ void __launch_bounds__(1024) k_big_ng_b(int M, int MP, const double* __restrict__ Xr, const double* __restrict__ v1g,
                                                  double* __restrict__ m, double* __restrict__ H) {
    __shared__ double vs[LVAE_MAX_M];
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* In = Xr + (size_t)l * MP * MP;
    double* Hl = H + (size_t)l * M * M;
    for (int i = tid; i < M; i += nt) vs[i] = v1g[(size_t)l * M + i];
    for (int e = tid; e < M * M; e += nt) {
        const int i = e / M, j = e % M;
        Hl[e] = 0.5 * (In[(size_t)i * MP + j] + In[(size_t)j * MP + i]);                                           // 134
    }
    __syncthreads();
    cta_gemv_rows(Hl, M, vs, M, M, m + (size_t)l * M);
}

GemmDesc mm(const double* A, const double* B, double* C, int MP, int L) {
    GemmDesc d;
    d.A = A; d.B = B; d.C = C;
    d.m = d.n = d.k = MP; d.lda = d.ldb = d.ldc = MP;
    d.batch = L; d.sA = d.sB = d.sC = (int64_t)MP * MP;
    return d;
}

}  // namespace

bool lvae_big_supported(const lvae_kld_problem_t* p) {
    return p->M > 62 && p->M <= LVAE_MAX_M && p->T_max <= LVAE_F2_ROWS && lvae_prep_warp_supported(p) && p->ks.n_comp0 >= 1 &&
           p->ks.n_comp0 <= 4 && p->ks.n_comp1 >= 1 && p->ks.n_comp0 + p->ks.n_comp1 <= 8;
}

#define RUN(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)
#define KCHECK() do { LVAE_COUNT_LAUNCH(); cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return lvae_cuda_rc(e_); } while (0)

int lvae_head_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int L = p->L, M = p->M, MP = w.MP;
    double* ws = p->workspace;
    k_big_fill<<<dim3(32, L), 256, 0, st>>>(sp, w, L, M, p->Q, p->z, p->H, p->lengthscale, p->outputscale, p->eps, ws);
    KCHECK();
    // 177-178 / 185-186: Cholesky and explicit inverses of Kzz + eps I and H (info[0]: Kzz, info[1]: H)
    RUN(lvae_spd_inverse_big(ws + w.bF, ws + w.bX, ws + w.bT, ws + w.bA0, ws + w.bDinv, MP, 2 * L, p->info, L, st));
    k_big_logdet<<<2 * L, 256, 0, st>>>(w, L, M, ws);
    KCHECK();
    k_big_unpad_inv<<<dim3(32, 2 * L), 256, 0, st>>>(w, L, M, ws);
    KCHECK();
    k_big_a<<<L, 256, 0, st>>>(w, M, p->m, ws);
    KCHECK();
    RUN(lvae_gemm(mm(ws + w.bInv, ws + w.bHp, ws + w.bT1, MP, L), st));               // Ki H
    RUN(lvae_gemm(mm(ws + w.bT1, ws + w.bInv, ws + w.bT2, MP, L), st));               // G = Ki H Ki (194)
    k_big_W<<<dim3(32, L), 256, 0, st>>>(w, M, 0.5 * p->scale, ws);
    KCHECK();
    return 0;
}

int lvae_tail_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const int L = p->L, M = p->M, MP = w.MP;
    double* ws = p->workspace;
    k_big_S_in<<<dim3(32, L), 256, 0, st>>>(w, M, p->stats, ws);
    KCHECK();
    RUN(lvae_gemm(mm(ws + w.bInv, ws + w.bS, ws + w.bT1, MP, L), st));                // T1 = Ki S
    RUN(lvae_gemm(mm(ws + w.bT1, ws + w.bInv, ws + w.bT2, MP, L), st));               // T2 = Ki S Ki
    RUN(lvae_gemm(mm(ws + w.bHp, ws + w.bT1, ws + w.bT3, MP, L), st));                // T3 = H Ki S
    k_big_tail_a<<<L, 1024, 0, st>>>(w, L, M, p->natural_gradient, p->m, 0.5 * p->scale, p->const_term / p->L, p->stats, ws,
                                     p->kld_per_latent, p->grad_m, p->grad_H);
    KCHECK();
    RUN(lvae_gemm(mm(ws + w.bInv, ws + w.bS, ws + w.bT1, MP, L), st));                // Ki gKi
    RUN(lvae_gemm(mm(ws + w.bT1, ws + w.bInv, ws + w.bT3, MP, L), st));               // Ki gKi Ki
    k_big_tail_b<<<L, 1024, 0, st>>>(sp, w, L, M, p->Q, p->z, p->lengthscale, p->outputscale, p->stats, ws, p->d_lengthscale,
                                     p->d_outputscale, p->d_noise);
    KCHECK();
    return 0;
}

// workspace: F | X | T | Inv | A0 (L*MP*MP each) | dinv (L * MP/64 * 4096) | v1 (L*M)
int64_t lvae_ng_big_workspace(int L, int M) {
    const int64_t MP = M <= 128 ? 128 : 256;
    return 5 * (int64_t)L * MP * MP + (int64_t)L * (MP / 64) * 4096 + (int64_t)L * M + 2;
}

int lvae_ng_big_launch(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hi, double lr, int L,
                       int M, double* ws, int32_t* info, cudaStream_t st) {
    const int MP = M <= 128 ? 128 : 256;
    const int64_t MP2 = (int64_t)MP * MP;
    double* F = ws;
    double* X = F + L * MP2;
    double* T = X + L * MP2;
    double* Inv = T + L * MP2;
    double* A0 = Inv + L * MP2;
    double* dinv = A0 + L * MP2;
    double* v1 = dinv + (int64_t)L * (MP / 64) * 4096;
    const double* iH = Hi;
    int ldi = M;
    int64_t istride = (int64_t)M * M;
    if (!iH) {                                                                         // training.py:130-131
        RUN(lvae_pad_in(F, H, M, MP, (int64_t)M * M, L, 1.0, 0, st));
        RUN(lvae_spd_inverse_big(F, X, T, A0, dinv, MP, L, info + 3, 0, st));
        // A0 holds H^-1 (identity padded); un-padded copy into Inv for k_big_ng_a
        k_big_sym_out<<<dim3(32, L), 256, 0, st>>>(Inv, A0, M, MP);
        KCHECK();
        iH = Inv;
    }
    k_big_ng_a<<<L, 1024, 0, st>>>(M, MP, m, grad_m, grad_H, iH, ldi, istride, lr, F, v1);
    KCHECK();
    RUN(lvae_spd_inverse_big(F, X, T, A0, dinv, MP, L, info + 3, 0, st));              // 133-134
    k_big_ng_b<<<L, 1024, 0, st>>>(M, MP, A0, v1, m, H);
    KCHECK();
    return 0;
}
