// Blocked FP64 Cholesky / triangular inverse / SPD inverse for identity-padded matrices of order np = 128 or 256
// (64 x 64 blocks).  The 64 x 64 diagonal blocks are factored and inverted in shared memory by one CTA per matrix
// (lvae_small.cuh: blocked-by-8 Cholesky and triangular inverse with DMMA trailing updates); panels, trailing updates
// and the off-diagonal blocks of the inverse are batched DMMA GEMMs (lvae_gemm.cu).
// Replaces torch.cholesky / torch.cholesky_solve(I, L) of elbo_functions.py:177-178,185-186 and training.py:130-134 for
// 64 < M <= 256.
#include "lvae_blas.h"
#include "lvae_small.cuh"

namespace {

// Cholesky of a tall panel (nr rows x 64 columns, row stride SLD, nr a multiple of 8): the 64 x 64 top block is factored
// and every row below is solved against it, 8 columns at a time: warp 0 factors the 8 x 8 diagonal block, one thread per
// row does the forward substitution, DMMA updates the remaining columns of the panel.  Same arithmetic as s_cholesky
// (lvae_small.cuh), so the blocked factorisation rounds like an un-blocked one (no explicit block inverses in the solves:
// they cost a factor cond(L_kk) in accuracy on the nearly singular Kzz of the reference).
__device__ inline int s_cholesky_tall(double* __restrict__ A, int nr, double* __restrict__ dinv, int* flag) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nrb = nr >> 3;
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int kb = 0; kb < 8; ++kb) {
        const int k0 = 8 * kb;
        if (wid == 0) {
            for (int k = 0; k < 8; ++k) {
                const double akk = A[(k0 + k) * SLD + k0 + k];
                if (!(akk > 0.0) && lane == 0 && *flag == 0) *flag = k0 + k + 1;
                double ri = rsqrt(akk);
                ri = ri * (1.5 - 0.5 * akk * ri * ri);
                const double d = akk * ri;
                __syncwarp();
                if (lane == k) { A[(k0 + k) * SLD + k0 + k] = d; dinv[k0 + k] = ri; }
                if (lane > k && lane < 8) A[(k0 + lane) * SLD + k0 + k] *= ri;
                __syncwarp();
                if (lane < 28) {
                    int r, c_;
                    tri_ij(lane, r, c_);
                    r += 1; c_ += 1;
                    if (c_ > k && r < 8) A[(k0 + r) * SLD + k0 + c_] -= A[(k0 + r) * SLD + k0 + k] * A[(k0 + c_) * SLD + k0 + k];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int r = k0 + 8 + tid; r < nr; r += blockDim.x) {
            double xr[8];
#pragma unroll
            for (int c_ = 0; c_ < 8; ++c_) {
                double s = A[r * SLD + k0 + c_];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < c_) s -= xr[k] * A[(k0 + c_) * SLD + k0 + k];
                xr[c_] = s * dinv[k0 + c_];
            }
#pragma unroll
            for (int c_ = 0; c_ < 8; ++c_) A[r * SLD + k0 + c_] = xr[c_];
        }
        __syncthreads();
        // trailing update inside the panel: tiles (ti, tj), kb < tj <= min(ti, 7), kb < ti < nrb
        const int ncol = 7 - kb;                                   // remaining column blocks
        const int nrow = nrb - kb - 1;                             // remaining row blocks
        for (int tl = wid; tl < nrow * ncol; tl += (blockDim.x >> 5)) {
            const int ti = kb + 1 + tl / ncol, tj = kb + 1 + tl % ncol;
            if (tj > ti) continue;
            double* Ct = A + (8 * ti + g) * SLD + 8 * tj + 2 * q;
            double c0 = Ct[0], c1 = Ct[1];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const double a = -A[(8 * ti + g) * SLD + k0 + 4 * ks + q];
                const double b = A[(8 * tj + g) * SLD + k0 + 4 * ks + q];
                dmma884(c0, c1, a, b);
            }
            Ct[0] = c0; Ct[1] = c1;
        }
        __syncthreads();
    }
    for (int e = tid; e < 64 * 64; e += blockDim.x) {
        const int i = e >> 6, j = e & 63;
        if (j > i) A[i * SLD + j] = 0.0;
    }
    __syncthreads();
    return *flag;
}

// mode 0: factor the panel below and including diagonal block kb in place (rows 64 kb .. np-1, columns of the block) and
// invert the 64 x 64 factor block; mode 1: the block already holds the factor (inverse only).
__global__ void __launch_bounds__(512, 1)
k_diag_block(double* __restrict__ F, int np, int kb, double* __restrict__ dinv_out, int mode, int32_t* info, int info_mod) {
    extern __shared__ double sm[];
    __shared__ double dinv[64];
    __shared__ int flag;
    const int b = blockIdx.x, tid = threadIdx.x, nb = np >> 6;
    const int nr = mode == 0 ? np - 64 * kb : 64;
    double* A = sm;                              // [nr][SLD]
    double* X = A + (size_t)nr * SLD;
    double* scratch = X + SMAT;
    double* blk = F + (size_t)b * np * np + (size_t)(64 * kb) * np + 64 * kb;
    for (int e = tid; e < nr * 64; e += 512) {
        const int i = e >> 6, j = e & 63;
        A[i * SLD + j] = (mode == 0 || j <= i) ? blk[(size_t)i * np + j] : 0.0;
    }
    __syncthreads();
    if (mode == 0) {
        const int rc = s_cholesky_tall(A, nr, dinv, &flag);
        if (rc && tid == 0 && info) atomicCAS(info + (info_mod > 0 ? b / info_mod : 0), 0, (info_mod > 0 ? b % info_mod : b) + 1);
        for (int e = tid; e < nr * 64; e += 512) {
            const int i = e >> 6, j = e & 63;
            blk[(size_t)i * np + j] = A[i * SLD + j];               // strict upper of the top block is zero now
        }
    } else {
        if (tid < 64) dinv[tid] = 1.0 / A[tid * SLD + tid];
        __syncthreads();
    }
    s_tri_inverse(A, X, 64, dinv, scratch);
    double* out = dinv_out + ((size_t)b * nb + kb) * 4096;
    for (int e = tid; e < 64 * 64; e += 512) out[e] = X[(e >> 6) * SLD + (e & 63)];
}

// X (zero-initialised) <- diagonal blocks from dinv
__global__ void k_place_diag(double* __restrict__ X, const double* __restrict__ dinv, int np) {
    const int b = blockIdx.y, nb = np >> 6, kb = blockIdx.x;
    const double* src = dinv + ((size_t)b * nb + kb) * 4096;
    double* dst = X + (size_t)b * np * np + (size_t)(64 * kb) * np + 64 * kb;
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) dst[(size_t)(e >> 6) * np + (e & 63)] = src[e];
}

// padded copy-in: dst[b] (np x np) = src[b] (n x n, row stride n) with `diag` on the padded diagonal
__global__ void k_pad_in(double* __restrict__ dst, const double* __restrict__ src, int n, int np, int64_t sstride,
                         double diag, int lower_only) {
    const int b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < np * np; e += gridDim.x * blockDim.x) {
        const int i = e / np, j = e % np;
        double v = (i == j) ? diag : 0.0;
        if (i < n && j < n && (!lower_only || j <= i)) v = src[(size_t)b * sstride + (size_t)i * n + j];
        dst[(size_t)b * np * np + e] = v;
    }
}
// copy-out: dst[b] (n x n) = src[b] (np x np); lower_only zeroes the strict upper triangle
__global__ void k_pad_out(double* __restrict__ dst, const double* __restrict__ src, int n, int np, int64_t dstride,
                          int lower_only) {
    const int b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int i = e / n, j = e % n;
        dst[(size_t)b * dstride + e] = (lower_only && j > i) ? 0.0 : src[(size_t)b * np * np + (size_t)i * np + j];
    }
}

int diag_launch(double* F, int np, int kb, int batch, double* dinv, int mode, int32_t* info, int info_mod, cudaStream_t st) {
    static SmemAttrCache attr;
    const size_t smem_max = sizeof(double) * ((size_t)256 * SLD + SMAT + 16 * 64);
    const int nr = mode == 0 ? np - 64 * kb : 64;
    const size_t smem = sizeof(double) * ((size_t)nr * SLD + SMAT + 16 * 64);
    if (int rc_ = lvae_ensure_smem(k_diag_block, smem_max, attr)) return rc_;
    k_diag_block<<<batch, 512, smem, st>>>(F, np, kb, dinv, mode, info, info_mod);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

int lvae_potrf_big(double* F, int np, int batch, double* dinv, int32_t* info_slot, int info_mod, cudaStream_t st) {
    const int nb = np >> 6;
    const int64_t ms = (int64_t)np * np;
    for (int kb = 0; kb < nb; ++kb) {
        int rc = diag_launch(F, np, kb, batch, dinv, 0, info_slot, info_mod, st);
        if (rc) return rc;
        const int k0 = 64 * kb, rem = np - k0 - 64;
        if (rem <= 0) break;
        const double* P = F + (size_t)(k0 + 64) * np + k0;   // panel rows below the diagonal block (solved by k_diag_block)
        GemmDesc t;                                   // trailing: A22 -= P P^T on the lower tiles
        t.A = P; t.lda = np; t.sA = ms;
        t.B = P; t.ldb = np; t.sB = ms; t.tb = 1;
        t.C = F + (size_t)(k0 + 64) * np + k0 + 64; t.ldc = np; t.sC = ms;
        t.m = rem; t.n = rem; t.k = 64; t.batch = batch;
        t.alpha = -1.0; t.beta = 1.0; t.flags = LVAE_GEMM_LOWER;
        rc = lvae_gemm(t, st);
        if (rc) return rc;
    }
    return 0;
}

int lvae_trtri_big(const double* F, const double* dinv, double* X, double* T, int np, int batch, cudaStream_t st) {
    const int nb = np >> 6;
    const int64_t ms = (int64_t)np * np;
    cudaError_t e = cudaMemsetAsync(X, 0, sizeof(double) * (size_t)batch * ms, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    k_place_diag<<<dim3(nb, batch), 256, 0, st>>>(X, dinv, np);
    LVAE_COUNT_LAUNCH();
    // [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]] : first on the 64-blocks inside every 128-block, then on the 128-blocks
    for (int bs = 64; bs < np; bs *= 2) {
        const int pairs = np / (2 * bs);
        const int64_t s2 = (int64_t)(2 * bs) * np + 2 * bs;
        GemmDesc a;                                   // T = B A^-1
        a.A = F + (size_t)bs * np; a.lda = np; a.sA = ms; a.sA2 = s2;
        a.B = X; a.ldb = np; a.sB = ms; a.sB2 = s2;
        a.C = T + (size_t)bs * np; a.ldc = np; a.sC = ms; a.sC2 = s2;
        a.m = bs; a.n = bs; a.k = bs; a.batch = batch; a.batch2 = pairs;
        int rc = lvae_gemm(a, st);
        if (rc) return rc;
        GemmDesc b;                                   // X21 = -C^-1 T
        b.A = X + (size_t)bs * np + bs; b.lda = np; b.sA = ms; b.sA2 = s2;
        b.B = T + (size_t)bs * np; b.ldb = np; b.sB = ms; b.sB2 = s2;
        b.C = X + (size_t)bs * np; b.ldc = np; b.sC = ms; b.sC2 = s2;
        b.m = bs; b.n = bs; b.k = bs; b.batch = batch; b.batch2 = pairs;
        b.alpha = -1.0;
        rc = lvae_gemm(b, st);
        if (rc) return rc;
    }
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_gram_big(const double* X, double* Inv, int np, int batch, cudaStream_t st) {
    GemmDesc g;
    g.A = X; g.lda = np; g.sA = (int64_t)np * np; g.ta = 1;
    g.B = X; g.ldb = np; g.sB = (int64_t)np * np;
    g.C = Inv; g.ldc = np; g.sC = (int64_t)np * np;
    g.m = np; g.n = np; g.k = np; g.batch = batch;
    g.flags = LVAE_GEMM_LOWER | LVAE_GEMM_MIRROR;
    return lvae_gemm(g, st);
}

// SPD inverse: blocked Cholesky of F (in place), X = F^-1, Inv = X^T X (symmetric, identity padded like F).  T scratch.
int lvae_spd_inverse_big(double* F, double* X, double* T, double* Inv, double* dinv, int np, int batch, int32_t* info,
                         int info_mod, cudaStream_t st) {
    int rc = lvae_potrf_big(F, np, batch, dinv, info, info_mod, st);
    if (rc) return rc;
    rc = lvae_trtri_big(F, dinv, X, T, np, batch, st);
    if (rc) return rc;
    return lvae_gram_big(X, Inv, np, batch, st);
}

int lvae_pad_in(double* dst, const double* src, int n, int np, int64_t sstride, int batch, double diag, int lower_only,
                cudaStream_t st) {
    k_pad_in<<<dim3(32, batch), 256, 0, st>>>(dst, src, n, np, sstride, diag, lower_only);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}
int lvae_pad_out(double* dst, const double* src, int n, int np, int64_t dstride, int batch, int lower_only, cudaStream_t st) {
    k_pad_out<<<dim3(32, batch), 256, 0, st>>>(dst, src, n, np, dstride, lower_only);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// batched potrf / potri ABI for 64 < n <= 256 (include/lvae_b200.h); scratch comes from the stream-ordered allocator
// ---------------------------------------------------------------------------------------------------------------
int lvae_potrf_big_abi(double* A, int n, int64_t stride, int batch, int32_t* info, cudaStream_t st) {
    const int np = lvae_pad_order(n), nb = np >> 6;
    double* buf = nullptr;
    const size_t doubles = (size_t)batch * np * np + (size_t)batch * nb * 4096;
    cudaError_t e = lvae_scratch_alloc((void**)&buf, sizeof(double) * doubles, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    double* F = buf;
    double* dinv = F + (size_t)batch * np * np;
    int rc = lvae_pad_in(F, A, n, np, stride, batch, 1.0, 0, st);
    if (!rc) rc = lvae_potrf_big(F, np, batch, dinv, info, 0, st);
    if (!rc) rc = lvae_pad_out(A, F, n, np, stride, batch, 1, st);
    cudaFreeAsync(buf, st);
    return rc;
}

int lvae_potri_big_abi(const double* Lc, double* Ainv, int n, int64_t stride, int batch, cudaStream_t st) {
    const int np = lvae_pad_order(n), nb = np >> 6;
    double* buf = nullptr;
    const size_t ms = (size_t)batch * np * np;
    cudaError_t e = lvae_scratch_alloc((void**)&buf, sizeof(double) * (3 * ms + (size_t)batch * nb * 4096), st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    double* F = buf;
    double* X = F + ms;
    double* T = X + ms;
    double* dinv = T + ms;
    int rc = lvae_pad_in(F, Lc, n, np, stride, batch, 1.0, 1, st);
    for (int kb = 0; kb < nb && !rc; ++kb) rc = diag_launch(F, np, kb, batch, dinv, 1, nullptr, 0, st);
    if (!rc) rc = lvae_trtri_big(F, dinv, X, T, np, batch, st);
    if (!rc) rc = lvae_gram_big(X, T, np, batch, st);
    if (!rc) rc = lvae_pad_out(Ainv, T, n, np, stride, batch, 0, st);
    cudaFreeAsync(buf, st);
    return rc;
}
