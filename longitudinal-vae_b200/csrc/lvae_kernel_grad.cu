// Hyper-parameter adjoints of the materialising kernels (lvae_kernel_dense_f64 / lvae_kernel_blocks_f64): given the
// adjoint G of a kernel matrix, reduce
//     d/d outputscale[c][l] = sum G * f_c                       f_c = masks * exp(-d^2 / (2 l^2))
//     d/d lengthscale[r][l] = sum G * outputscale * f_c * d^2 / l^3
//     d/d diag_add[l]       = sum_i G_ii
// per latent.  These make the non-minibatch bounds (deviance_upper_bound / elbo, elbo_functions.py:36-142) trainable
// through autograd the way gpytorch's lazy kernels are in the reference; the Hensman step has its own fused adjoints
// (lvae_subjects_fused*.cu) and never comes here.  Per entry and component: one exp (the table-driven one of
// lvae_common.cuh) and 8 bytes of G (re-read per component, rolled component loop); covariates come from L1.
// Two launches, no atomics: per-CTA partial rows, then a fixed-order sum per latent (bitwise reproducible).
#include "lvae_host.h"

#define GRAD_ROW (2 * LVAE_MAXC + 1)          // [sum G f_c (MAXC) | sum G f_c d^2 (MAXC) | sum_i G_ii]
#define GRAD_THREADS 256

struct GradSetup {
    double hil2[LVAE_MAXC];
    double etab[LVAE_EXP_TBL];
    double red[GRAD_THREADS / 32];
};

__device__ __forceinline__ void grad_setup(GradSetup& s, const DevSpec& sp, const double* __restrict__ ls, int L, int l) {
    const int t = threadIdx.x;
    if (t < LVAE_EXP_TBL) s.etab[t] = c_exp2_tbl[t];
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; s.hil2[t] = 0.5 / (v * v); }
    __syncthreads();
}

// matrix b of the batch uses latent b % L; grid (ceil(n2/64), n_chunk, L), 256 threads as 64 columns x 4 rows (the layout
// of the forward kernel: coalesced on j, no integer division per entry); CTA (cx, ch, l) walks rows ch*4+ty, +4*n_chunk, ...
__global__ void __launch_bounds__(GRAD_THREADS) k_dense_bwd(DevSpec sp, int c0, int c1, int Q, const double* __restrict__ x1,
                                                            int64_t s1, int n1, const double* __restrict__ x2, int64_t s2,
                                                            int n2, const double* __restrict__ ls, int L, int n_rep,
                                                            const double* __restrict__ G, int want_diag,
                                                            double* __restrict__ part) {
    __shared__ GradSetup s;
    const int l = blockIdx.z;
    grad_setup(s, sp, ls, L, l);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = blockIdx.x * 64 + tx, istep = 4 * gridDim.y;
    const int64_t per = (int64_t)n1 * n2;
    double* row = part + ((size_t)l * gridDim.y * gridDim.x + (size_t)blockIdx.y * gridDim.x + blockIdx.x) * GRAD_ROW;
    for (int c = c0; c < c1; ++c) {
        double a = 0.0, a2 = 0.0;
        if (j < n2) {
            for (int rep = 0; rep < n_rep; ++rep) {
                const int64_t b = (int64_t)rep * L + l;
                const double* xb = x2 + b * s2 + (size_t)j * Q;
                const double* g = G + b * per + j;
                for (int i = blockIdx.y * 4 + ty; i < n1; i += istep) {
                    double d2;
                    const double f = comp_value(sp, c, x1 + b * s1 + (size_t)i * Q, xb, s.hil2, d2, s.etab);
                    const double gf = g[(size_t)i * n2] * f;
                    a += gf;
                    a2 = fma(gf, d2, a2);
                }
            }
        }
        a = block_sum(a, s.red);
        a2 = block_sum(a2, s.red);
        if (threadIdx.x == 0) { row[c] = a; row[LVAE_MAXC + c] = a2; }
    }
    double dg = 0.0;
    if (want_diag) {
        if (j < n2 && j < n1)
            for (int rep = 0; rep < n_rep; ++rep)
                for (int i = blockIdx.y * 4 + ty; i < n1; i += istep)
                    if (i == j) dg += G[((int64_t)rep * L + l) * per + (int64_t)i * n2 + i];
        dg = block_sum(dg, s.red);
    }
    if (threadIdx.x == 0) row[2 * LVAE_MAXC] = dg;
}

// per-subject blocks: grid (n_chunk, L); CTA (ch, l) takes subjects ch, ch + n_chunk, ... of latent l
__global__ void __launch_bounds__(GRAD_THREADS) k_blocks_bwd(DevSpec sp, int c0, int c1, int Q, const double* __restrict__ x,
                                                             const int32_t* __restrict__ offsets,
                                                             const int64_t* __restrict__ off2, int P_b, int64_t block_stride,
                                                             const double* __restrict__ ls, int L,
                                                             const double* __restrict__ G, int want_diag,
                                                             double* __restrict__ part) {
    __shared__ GradSetup s;
    const int l = blockIdx.y;
    grad_setup(s, sp, ls, L, l);
    const double* Gl = G + (size_t)l * block_stride;
    double* row = part + ((size_t)l * gridDim.x + blockIdx.x) * GRAD_ROW;
    for (int c = c0; c < c1; ++c) {
        double a = 0.0, a2 = 0.0;
        for (int p = blockIdx.x; p < P_b; p += gridDim.x) {
            const int r0 = offsets[p], T = offsets[p + 1] - r0;
            const double* g = Gl + off2[p];
            for (int e = threadIdx.x; e < T * T; e += GRAD_THREADS) {
                const int i = e / T, j = e - i * T;
                double d2;
                const double f = comp_value(sp, c, x + (size_t)(r0 + i) * Q, x + (size_t)(r0 + j) * Q, s.hil2, d2, s.etab);
                const double gf = g[e] * f;
                a += gf;
                a2 = fma(gf, d2, a2);
            }
        }
        a = block_sum(a, s.red);
        a2 = block_sum(a2, s.red);
        if (threadIdx.x == 0) { row[c] = a; row[LVAE_MAXC + c] = a2; }
    }
    double dg = 0.0;
    if (want_diag) {
        for (int p = blockIdx.x; p < P_b; p += gridDim.x) {
            const int T = offsets[p + 1] - offsets[p];
            const double* g = Gl + off2[p];
            for (int i = threadIdx.x; i < T; i += GRAD_THREADS) dg += g[(size_t)i * T + i];
        }
        dg = block_sum(dg, s.red);
    }
    if (threadIdx.x == 0) row[2 * LVAE_MAXC] = dg;
}

// grid L, 64 threads: fixed-order sum over the chunk rows, then the chain rule into the three outputs
__global__ void __launch_bounds__(64) k_grad_finish(DevSpec sp, int c0, int c1, int n_chunk, const double* __restrict__ part,
                                                    const double* __restrict__ ls, const double* __restrict__ os, int L,
                                                    double* __restrict__ d_ls, double* __restrict__ d_os,
                                                    double* __restrict__ d_diag) {
    __shared__ double tot[GRAD_ROW];
    const int l = blockIdx.x, t = threadIdx.x;
    if (t < GRAD_ROW) {
        double a = 0.0;
        const double* p = part + (size_t)l * n_chunk * GRAD_ROW + t;
        for (int ch = 0; ch < n_chunk; ++ch) a += p[(size_t)ch * GRAD_ROW];
        tot[t] = a;
    }
    __syncthreads();
    const int nc = sp.n0 + sp.n1;
    if (t < nc) d_os[(size_t)t * L + l] = (t >= c0 && t < c1) ? tot[t] : 0.0;
    if (t < sp.n_ls) {
        const double v = ls[(size_t)t * L + l];
        double a = 0.0;
        for (int c = c0; c < c1; ++c)
            if (sp.rbf_dim[c] >= 0 && sp.ls_idx[c] == t) a += os[(size_t)c * L + l] * tot[LVAE_MAXC + c];
        d_ls[(size_t)t * L + l] = a / (v * v * v);
    }
    if (t == 0 && d_diag) d_diag[l] = tot[2 * LVAE_MAXC];
}

static int grad_chunks(int64_t work_per_latent, int L) {
    // about two waves of CTAs over the 148 SMs, at least ~4k entries per CTA
    int64_t want = (2 * 148 * 4 + L - 1) / L;
    int64_t cap = (work_per_latent + 4095) / 4096;
    int64_t n = want < cap ? want : cap;
    return (int)(n < 1 ? 1 : n);
}

static int grad_finish(const DevSpec& sp, int c0, int c1, int n_chunk, double* part, const double* ls, const double* os, int L,
                       double* d_ls, double* d_os, double* d_diag, cudaStream_t st) {
    k_grad_finish<<<L, 64, 0, st>>>(sp, c0, c1, n_chunk, part, ls, os, L, d_ls, d_os, d_diag);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int lvae_kernel_dense_bwd_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                                         int32_t n_batch, int32_t Q, const double* x1, int64_t s1, int32_t n1,
                                         const double* x2, int64_t s2, int32_t n2, const double* lengthscale,
                                         const double* outputscale, const double* grad_out, double* d_lengthscale,
                                         double* d_outputscale, double* d_diag_add, void* stream) {
    DevSpec sp;
    int rc = lvae_make_devspec(ks, Q, &sp);
    if (rc) return rc;
    if (comp_begin < 0 || comp_end > sp.n0 + sp.n1 || comp_begin > comp_end || L <= 0 || L > 65535) return LVAE_E_BADARG;
    if (n_batch <= 0 || n_batch % L != 0 || n1 < 0 || n2 < 0 || !grad_out || !d_lengthscale || !d_outputscale)
        return LVAE_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_rep = n_batch / L;
    const int ncx = n2 > 0 ? (n2 + 63) / 64 : 1;
    int ncy = grad_chunks((int64_t)n1 * n2 * n_rep, L) / ncx;          // row chunks per column block
    if (ncy > (n1 + 3) / 4) ncy = (n1 + 3) / 4;
    if (ncy < 1) ncy = 1;
    if (ncy > 65535 || ncx > 65535) return LVAE_E_TOO_LARGE;
    const int n_chunk = ncx * ncy;
    double* part = nullptr;
    cudaError_t e = lvae_scratch_alloc((void**)&part, sizeof(double) * (size_t)L * n_chunk * GRAD_ROW, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    k_dense_bwd<<<dim3(ncx, ncy, L), GRAD_THREADS, 0, st>>>(sp, comp_begin, comp_end, Q, x1, s1, n1, x2, s2, n2, lengthscale, L,
                                                            n_rep, grad_out, d_diag_add != nullptr, part);
    LVAE_COUNT_LAUNCH();
    rc = lvae_cuda_rc(cudaGetLastError());
    if (!rc) rc = grad_finish(sp, comp_begin, comp_end, n_chunk, part, lengthscale, outputscale, L, d_lengthscale,
                              d_outputscale, d_diag_add, st);
    cudaFreeAsync(part, st);
    return rc;
}

extern "C" int lvae_kernel_blocks_bwd_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                                          int32_t Q, const double* x, const int32_t* offsets, int32_t P_b,
                                          int64_t block_stride, const double* lengthscale, const double* outputscale,
                                          const double* grad_out, double* d_lengthscale, double* d_outputscale,
                                          double* d_diag_add, void* stream) {
    DevSpec sp;
    int rc = lvae_make_devspec(ks, Q, &sp);
    if (rc) return rc;
    if (comp_begin < 0 || comp_end > sp.n0 + sp.n1 || comp_begin > comp_end || L <= 0 || L > 65535) return LVAE_E_BADARG;
    if (P_b < 0 || !grad_out || !d_lengthscale || !d_outputscale) return LVAE_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int n_chunk = grad_chunks(block_stride, L);
    if (n_chunk > P_b) n_chunk = P_b > 0 ? P_b : 1;
    int64_t* off2 = nullptr;
    double* part = nullptr;
    cudaError_t e = lvae_scratch_alloc((void**)&off2, sizeof(int64_t) * ((size_t)P_b + 1), st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    e = lvae_scratch_alloc((void**)&part, sizeof(double) * (size_t)L * n_chunk * GRAD_ROW, st);
    if (e != cudaSuccess) { cudaFreeAsync(off2, st); return lvae_cuda_rc(e); }
    rc = lvae_block_offsets(offsets, P_b, off2, st);
    if (!rc) {
        k_blocks_bwd<<<dim3(n_chunk, L), GRAD_THREADS, 0, st>>>(sp, comp_begin, comp_end, Q, x, offsets, off2, P_b,
                                                                block_stride, lengthscale, L, grad_out,
                                                                d_diag_add != nullptr, part);
        LVAE_COUNT_LAUNCH();
        rc = lvae_cuda_rc(cudaGetLastError());
    }
    if (!rc) rc = grad_finish(sp, comp_begin, comp_end, n_chunk, part, lengthscale, outputscale, L, d_lengthscale,
                              d_outputscale, d_diag_add, st);
    cudaFreeAsync(off2, st);
    cudaFreeAsync(part, st);
    return rc;
}
