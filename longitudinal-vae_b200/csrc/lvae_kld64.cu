// Per-latent M x M work of the GP-prior ELBO path for M <= 64, entirely in shared memory with DMMA GEMMs
// (lvae_small.cuh).  Same outputs and workspace layout as the generic kernels of lvae_kld.cu.
//   k_head64 (grid 2 x L):  role 0: Kzz + eps I -> chol -> Kzz^-1, a = Kzz^-1 m, G = Kzz^-1 H Kzz^-1, W = c(G - Kzz^-1)
//                           role 1: H -> chol -> H^-1, log det H                (elbo_functions.py:172,176-178,185-186,194)
//   k_tail64 (grid L):      D, E, KL[q(u)||p(u)], kld, grad_m, grad_H (or d_m, d_H), Kzz adjoint and its hyper-gradients
//   k_ng64   (grid L):      natural-gradient update (training.py:129-135); reuses H^-1 of the head when supplied
#include "lvae_kld.h"
#include "lvae_small.cuh"

namespace {

struct Hyp64 {
    double hil2[LVAE_MAXC], il3[LVAE_MAXC], os[LVAE_MAXC], etab[LVAE_EXP_TBL];
};

__device__ inline void load_hyp64(Hyp64* h, const DevSpec& sp, const double* ls, const double* os, int L, int l) {
    const int t = threadIdx.x;
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; h->hil2[t] = 0.5 / (v * v); h->il3[t] = 1.0 / (v * v * v); }
    if (t < sp.n0 + sp.n1) h->os[t] = os[(size_t)t * L + l];
    load_exp_table(h->etab);
}

__global__ void __launch_bounds__(512, 1)
k_head64(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q,
         const double* __restrict__ z, const double* __restrict__ m, const double* __restrict__ H,
         const double* __restrict__ ls, const double* __restrict__ os, double eps, double c, double* __restrict__ ws,
         int32_t* info) {
    extern __shared__ double sm[];
    __shared__ Hyp64 hyp;
    __shared__ double dinv[64], red[32], vec[64];
    __shared__ int flag;
    const int role = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, MM = M * M;
    const int n8 = (M + 7) & ~7;
    double* A = sm;                 // matrix being factored
    double* X = A + SMAT;           // its triangular inverse, later a product
    double* Inv = X + SMAT;         // the explicit inverse
    double* Hs = Inv + SMAT;        // H (role 0 only)
    double* scratch = Hs + SMAT;    // 16 x 64
    if (role == 0) {
        load_hyp64(&hyp, sp, ls, os, L, l);
        __syncthreads();
        const double* zl = z + (size_t)l * M * Q;
        for (int e = tid; e < 64 * 64; e += 512) {
            const int i = e >> 6, j = e & 63;
            double v = (i == j) ? 1.0 : 0.0;
            if (i < M && j < M) {
                double acc = 0.0, d2;
                for (int cc = 0; cc < sp.n0; ++cc) acc += hyp.os[cc] * comp_value(sp, cc, zl + i * Q, zl + j * Q, hyp.hil2, d2, hyp.etab);
                v = acc + (i == j ? eps : 0.0);                                        // elbo_functions.py:172,176
            }
            A[i * SLD + j] = v;
        }
        s_load(Hs, H + (size_t)l * MM, M, 0.0);
        __syncthreads();
    } else {
        s_load(A, H + (size_t)l * MM, M, 1.0);
        __syncthreads();
    }
    const int rc = s_cholesky(A, n8, dinv, &flag);                                     // 177 / 185
    if (rc && tid == 0) atomicCAS(info + role, 0, l + 1);
    {
        double v = 0.0;
        if (tid < M) v = log(A[tid * SLD + tid]);
        const double ld = 2.0 * block_sum(v, red);
        if (tid == 0) ws[w.logdet + 2 * l + role] = ld;
    }
    s_tri_inverse(A, X, n8, dinv, scratch);
    s_potrs_identity(A, X, Inv, n8, dinv, scratch);                                    // 178 / 186 (explicit inverse)
    if (role == 1) {
        s_store(ws + w.Hi + (size_t)l * MM, Inv, M);
        return;
    }
    s_store(ws + w.Ki + (size_t)l * MM, Inv, M);
    if (tid < 64) vec[tid] = tid < M ? m[(size_t)l * M + tid] : 0.0;
    __syncthreads();
    if (tid < M) {                                                                      // a = Kzz^-1 m
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += Inv[tid * SLD + k] * vec[k];
        ws[w.a + (size_t)l * M + tid] = s;
    }
    // G = Ki H Ki (194): X <- Ki H ; A <- X Ki
    s_gemm<false, false>(Inv, Hs, n8, [&](int i, int j, double v) { X[i * SLD + j] = v; });
    __syncthreads();
    s_gemm<false, false>(X, Inv, n8, [&](int i, int j, double v) { A[i * SLD + j] = v; });
    __syncthreads();
    double* G = ws + w.G + (size_t)l * MM;
    double* W = ws + w.W + (size_t)l * MM;
    for (int e = tid; e < MM; e += 512) {
        const int i = e / M, j = e % M;
        G[e] = A[i * SLD + j];
        W[e] = c * (0.5 * (A[i * SLD + j] + A[j * SLD + i]) - Inv[i * SLD + j]);       // adjoint of S, symmetrised
    }
}

__global__ void __launch_bounds__(512, 1)
k_tail64(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int natural_gradient,
         const double* __restrict__ z, const double* __restrict__ m, const double* __restrict__ H,
         const double* __restrict__ ls, const double* __restrict__ os, double c, double const_per_latent,
         const double* __restrict__ stats, double* __restrict__ ws, double* __restrict__ kld, double* __restrict__ grad_m,
         double* __restrict__ grad_H, double* __restrict__ d_ls, double* __restrict__ d_os, double* __restrict__ d_noise) {
    extern __shared__ double sm[];
    __shared__ Hyp64 hyp;
    __shared__ double red[32], ms[64], ga[64], ng1s[64], as[64];
    const int l = blockIdx.x, tid = threadIdx.x, MM = M * M, nh = hyp_count(sp);
    const int n8 = (M + 7) & ~7;
    double* Ki = sm;
    double* S = Ki + SMAT;      // later the adjoint of Ki
    double* Hs = S + SMAT;
    double* P1 = Hs + SMAT;     // Ki S, later Ki gKi
    double* T2 = P1 + SMAT;     // Ki S Ki, later (Ki gKi) Ki
    double* HP = T2 + SMAT;     // H Ki S
    load_hyp64(&hyp, sp, ls, os, L, l);
    const double* st = stats + (size_t)l * w.stride;
    const double* sc = st + stats_off_scal(M);
    const double* hy = st + stats_off_hyp(M);
    s_load(Ki, ws + w.Ki + (size_t)l * MM, M, 0.0);
    s_load(S, st + stats_off_S(), M, 0.0);
    s_load(Hs, H + (size_t)l * MM, M, 0.0);
    if (tid < 64) {
        ms[tid] = tid < M ? m[(size_t)l * M + tid] : 0.0;
        as[tid] = tid < M ? ws[w.a + (size_t)l * M + tid] : 0.0;
        ga[tid] = tid < M ? 2.0 * c * st[stats_off_da(M) + tid] : 0.0;
        ng1s[tid] = tid < M ? st[stats_off_ng1(M) + tid] : 0.0;
    }
    __syncthreads();
    {   // D2 = sum S o Ki (193), E = sum G^T o S (195 / 282), tr = sum Ki o H^T (199), qf = m.a (200)
        const double* G = ws + w.G + (size_t)l * MM;
        double d2 = 0.0, ee = 0.0, tr = 0.0, qf = 0.0;
        for (int e = tid; e < MM; e += 512) {
            const int i = e / M, j = e % M;
            d2 += S[i * SLD + j] * Ki[i * SLD + j];
            ee += G[e] * S[j * SLD + i];
            tr += Ki[i * SLD + j] * Hs[j * SLD + i];
        }
        if (tid < M) qf = ms[tid] * as[tid];
        d2 = block_sum(d2, red);
        ee = block_sum(ee, red);
        tr = block_sum(tr, red);
        qf = block_sum(qf, red);
        if (tid == 0) {
            const double ldK = ws[w.logdet + 2 * l], ldH = ws[w.logdet + 2 * l + 1];
            const double kl_qp = 0.5 * (tr + qf - M + ldK - ldH);                                   // 199-203
            kld[l] = c * (sc[SC_A] + sc[SC_BT] + sc[SC_C] + sc[SC_D1] - d2 + ee - sc[SC_F]) + kl_qp - const_per_latent;
        }
    }
    s_gemm<false, false>(Ki, S, n8, [&](int i, int j, double v) { P1[i * SLD + j] = v; });
    __syncthreads();
    s_gemm<false, false>(P1, Ki, n8, [&](int i, int j, double v) { T2[i * SLD + j] = v; });         // Ki S Ki
    s_gemm<false, false>(Hs, P1, n8, [&](int i, int j, double v) { HP[i * SLD + j] = v; });         // H Ki S
    __syncthreads();
    {
        double* gm = grad_m + (size_t)l * M;
        double* gH = grad_H + (size_t)l * MM;
        const double* Hi = ws + w.Hi + (size_t)l * MM;
        if (natural_gradient) {                                                                    // 208-214, 301-305
            if (tid < M) {
                double s = 0.0;
                for (int k = 0; k < M; ++k) s += -Ki[tid * SLD + k] * ng1s[k] + (T2[tid * SLD + k] + Ki[tid * SLD + k]) * ms[k];
                gm[tid] = s;
            }
            for (int e = tid; e < MM; e += 512) {
                const int i = e / M, j = e % M;
                gH[e] = 0.5 * (T2[i * SLD + j] + Ki[i * SLD + j] - Hi[e]);
            }
        } else {                                                                                   // autograd of kld_total
            if (tid < M) {
                double s = 0.0;
                for (int k = 0; k < M; ++k) s += Ki[tid * SLD + k] * ga[k];
                gm[tid] = s + as[tid];
            }
            for (int e = tid; e < MM; e += 512) {
                const int i = e / M, j = e % M;
                gH[e] = c * 0.5 * (T2[i * SLD + j] + T2[j * SLD + i]) + 0.5 * Ki[i * SLD + j] - 0.5 * Hi[e];
            }
        }
    }
    // adjoint of Ki, in place over S:  -c S + c (HP + HP^T) + (H^T + m m^T)/2 + ga m^T
    for (int e = tid; e < 64 * 64; e += 512) {
        const int i = e >> 6, j = e & 63;
        double v = 0.0;
        if (i < M && j < M)
            v = -c * S[i * SLD + j] + c * (HP[i * SLD + j] + HP[j * SLD + i]) + 0.5 * (Hs[j * SLD + i] + ms[i] * ms[j]) + ga[i] * ms[j];
        S[i * SLD + j] = v;
    }
    __syncthreads();
    s_gemm<false, false>(Ki, S, n8, [&](int i, int j, double v) { P1[i * SLD + j] = v; });
    __syncthreads();
    s_gemm<false, false>(P1, Ki, n8, [&](int i, int j, double v) { T2[i * SLD + j] = v; });
    __syncthreads();
    // adjoint of Kzz = -Ki gKi Ki + Ki/2 (symmetrised), contracted with d k_c / d theta on (Z, Z)
    const double* zl = z + (size_t)l * M * Q;
    double acc[2 * LVAE_MAXC + 1];
    for (int k = 0; k < nh; ++k) acc[k] = 0.0;
    for (int e = tid; e < MM; e += 512) {
        const int i = e / M, j = e % M;
        const double gK = -0.5 * (T2[i * SLD + j] + T2[j * SLD + i]) + 0.5 * Ki[i * SLD + j];
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

__global__ void __launch_bounds__(512, 1)
k_ng64(double* __restrict__ m, double* __restrict__ H, const double* __restrict__ grad_m, const double* __restrict__ grad_H,
       const double* __restrict__ Hi_in, double lr, int M, int32_t* info) {
    extern __shared__ double sm[];
    __shared__ double dinv[64], ms[64], v1[64];
    __shared__ int flag;
    const int l = blockIdx.x, tid = threadIdx.x, MM = M * M;
    const int n8 = (M + 7) & ~7;
    double* iH = sm;
    double* A = iH + SMAT;
    double* X = A + SMAT;
    double* Hn = X + SMAT;
    double* scratch = Hn + SMAT;
    double* Hl = H + (size_t)l * MM;
    const double* gH = grad_H + (size_t)l * MM;
    if (tid < 64) ms[tid] = tid < M ? m[(size_t)l * M + tid] : 0.0;
    if (Hi_in) {
        s_load(iH, Hi_in + (size_t)l * MM, M, 0.0);
        __syncthreads();
    } else {
        s_load(A, Hl, M, 1.0);
        __syncthreads();
        const int rc = s_cholesky(A, n8, dinv, &flag);                                 // training.py:130
        if (rc && tid == 0) atomicCAS(info + 3, 0, l + 1);
        s_tri_inverse(A, X, n8, dinv, scratch);
        s_potrs_identity(A, X, iH, n8, dinv, scratch);                                 // 131
    }
    for (int e = tid; e < 64 * 64; e += 512) {                                          // 132
        const int i = e >> 6, j = e & 63;
        A[i * SLD + j] = (i < M && j < M) ? iH[i * SLD + j] + lr * (gH[i * M + j] + gH[j * M + i]) : (i == j ? 1.0 : 0.0);
    }
    if (tid < M) {                                                                      // 135, old m and iH
        double s = 0.0, g2 = 0.0;
        for (int k = 0; k < M; ++k) { s += iH[tid * SLD + k] * ms[k]; g2 += gH[tid * M + k] * ms[k]; }
        v1[tid] = s - lr * (grad_m[(size_t)l * M + tid] - 2.0 * g2);
    }
    __syncthreads();
    const int rc = s_cholesky(A, n8, dinv, &flag);                                     // 133
    if (rc && tid == 0) atomicCAS(info + 3, 0, l + 1);
    s_tri_inverse(A, X, n8, dinv, scratch);
    s_potrs_identity(A, X, Hn, n8, dinv, scratch);                                     // 134
    s_store(Hl, Hn, M);
    if (tid < M) {
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += Hn[tid * SLD + k] * v1[k];
        m[(size_t)l * M + tid] = s;
    }
}

}  // namespace

int lvae_head64_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * (4 * SMAT + 16 * 64);
    int rc = lvae_ensure_smem(k_head64, smem, attr);
    if (rc) return rc;
    k_head64<<<dim3(2, p->L), 512, smem, st>>>(sp, w, p->L, p->M, p->Q, p->z, p->m, p->H, p->lengthscale, p->outputscale,
                                               p->eps, 0.5 * p->scale, p->workspace, p->info);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_tail64_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * (6 * SMAT);
    int rc = lvae_ensure_smem(k_tail64, smem, attr);
    if (rc) return rc;
    k_tail64<<<p->L, 512, smem, st>>>(sp, w, p->L, p->M, p->Q, p->natural_gradient, p->z, p->m, p->H, p->lengthscale,
                                      p->outputscale, 0.5 * p->scale, p->const_term / p->L, p->stats, p->workspace,
                                      p->kld_per_latent, p->grad_m, p->grad_H, p->d_lengthscale, p->d_outputscale,
                                      p->d_noise);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_ng64_launch(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hi, double lr, int L,
                     int M, int32_t* info, cudaStream_t st) {
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * (4 * SMAT + 16 * 64);
    int rc = lvae_ensure_smem(k_ng64, smem, attr);
    if (rc) return rc;
    k_ng64<<<L, 512, smem, st>>>(m, H, grad_m, grad_H, Hi, lr, M, info);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}
