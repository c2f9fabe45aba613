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

// Block row i of X = F^-1: the columns 64 c .. 64 c + 63 of the row hold R = -sum_{k<i} L_ik X_k (c < i) or stand for the
// identity (c == i); the CTA overwrites them with L_ii^-1 R by forward substitution, one thread per column.
__global__ void __launch_bounds__(64) k_trsm_block_row(const double* __restrict__ F, double* __restrict__ X, int np, int i) {
    extern __shared__ double sm_trsm[];
    double (*Ls)[65] = reinterpret_cast<double (*)[65]>(sm_trsm);
    double (*xs)[64] = reinterpret_cast<double (*)[64]>(sm_trsm + 64 * 65);
    const int c = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const double* Lb = F + (size_t)b * np * np + (size_t)(64 * i) * np + 64 * i;
    double* Xb = X + (size_t)b * np * np + (size_t)(64 * i) * np + 64 * c;
    for (int e = t; e < 64 * 64; e += 64) Ls[e >> 6][e & 63] = Lb[(size_t)(e >> 6) * np + (e & 63)];
    __syncthreads();
    const bool diag = c == i;
    for (int r = 0; r < 64; ++r) {
        double s0 = diag ? (r == t ? 1.0 : 0.0) : Xb[(size_t)r * np + t], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const int k0 = diag ? t : 0;              // x_k = 0 for k < t in the identity columns
        int k = k0;
        for (; k + 3 < r; k += 4) {
            s0 -= Ls[r][k] * xs[k][t];
            s1 -= Ls[r][k + 1] * xs[k + 1][t];
            s2 -= Ls[r][k + 2] * xs[k + 2][t];
            s3 -= Ls[r][k + 3] * xs[k + 3][t];
        }
        for (; k < r; ++k) s0 -= Ls[r][k] * xs[k][t];
        xs[r][t] = (diag && r < t) ? 0.0 : ((s0 + s1) + (s2 + s3)) / Ls[r][r];
    }
    for (int r = 0; r < 64; ++r) Xb[(size_t)r * np + t] = xs[r][t];
}

// X = F^-1 by block forward substitution on 64-row blocks (Higham's Method 1B: the diagonal blocks are applied by
// substitution, never through their explicit inverses).  The recursive [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1,C^-1]]
// of round 1 multiplied by explicit block inverses: on the reference's Kzz (cond 1e8 .. 1e9) that cost a factor ~50 in the
// accuracy of grad_m / grad_H against LAPACK (measured against an extended-precision evaluation).
int lvae_trtri_big(const double* F, const double* dinv, double* X, double* T, int np, int batch, cudaStream_t st) {
    (void)dinv; (void)T;
    const int nb = np >> 6;
    const int64_t ms = (int64_t)np * np;
    cudaError_t e = cudaMemsetAsync(X, 0, sizeof(double) * (size_t)batch * ms, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * (64 * 65 + 64 * 64);
    if (int rc_ = lvae_ensure_smem(k_trsm_block_row, smem, attr)) return rc_;
    for (int i = 0; i < nb; ++i) {
        if (i > 0) {                                  // R_i = -L[i, 0:i] X[0:i, 0:i]
            GemmDesc a;
            a.A = F + (size_t)(64 * i) * np; a.lda = np; a.sA = ms;
            a.B = X; a.ldb = np; a.sB = ms;
            a.C = X + (size_t)(64 * i) * np; a.ldc = np; a.sC = ms;
            a.m = 64; a.n = 64 * i; a.k = 64 * i; a.batch = batch;
            a.alpha = -1.0;
            const int rc = lvae_gemm(a, st);
            if (rc) return rc;
        }
        k_trsm_block_row<<<dim3(i + 1, batch), 64, smem, st>>>(F, X, np, i);
        LVAE_COUNT_LAUNCH();
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

// Block row i of Inv = F^-T X: the row holds R = X_i - sum_{k>i} L_ki^T Inv_k; the CTA overwrites 64 columns of it with
// L_ii^-T R by BACK substitution, one thread per column.
__global__ void __launch_bounds__(64) k_trsmT_block_row(const double* __restrict__ F, double* __restrict__ Inv, int np, int i) {
    extern __shared__ double sm_trsm[];
    double (*Ls)[65] = reinterpret_cast<double (*)[65]>(sm_trsm);
    double (*ys)[64] = reinterpret_cast<double (*)[64]>(sm_trsm + 64 * 65);
    const int c = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const double* Lb = F + (size_t)b * np * np + (size_t)(64 * i) * np + 64 * i;
    double* Rb = Inv + (size_t)b * np * np + (size_t)(64 * i) * np + 64 * c;
    for (int e = t; e < 64 * 64; e += 64) Ls[e >> 6][e & 63] = Lb[(size_t)(e >> 6) * np + (e & 63)];
    __syncthreads();
    for (int r = 63; r >= 0; --r) {
        double s0 = Rb[(size_t)r * np + t], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = r + 1;
        for (; k + 3 < 64; k += 4) {
            s0 -= Ls[k][r] * ys[k][t];
            s1 -= Ls[k + 1][r] * ys[k + 1][t];
            s2 -= Ls[k + 2][r] * ys[k + 2][t];
            s3 -= Ls[k + 3][r] * ys[k + 3][t];
        }
        for (; k < 64; ++k) s0 -= Ls[k][r] * ys[k][t];
        ys[r][t] = ((s0 + s1) + (s2 + s3)) / Ls[r][r];
    }
    for (int r = 0; r < 64; ++r) Rb[(size_t)r * np + t] = ys[r][t];
}

// Inv = F^-T X for the lower factor F and X = F^-1: the second triangular solve of LAPACK's potrs with the identity as
// right-hand side (what torch.cholesky_solve(I, L) runs, elbo_functions.py:178,186), as a block back substitution.  The Gram
// product X^T X gives the same inverse to the same FORWARD accuracy, but a residual |A Inv - I| that grows with cond(L): on
// the reference's Kzz (cond 1e8 .. 1e9) that residual is what Kxz Kzz^-1 and Kzz^-1 S Kzz^-1 cancel against, and it cost a
// factor ~50 in grad_m / grad_H against LAPACK (measured against an extended-precision evaluation).
int lvae_potrs_identity_big(const double* F, const double* X, double* Inv, int np, int batch, cudaStream_t st) {
    const int nb = np >> 6;
    const int64_t ms = (int64_t)np * np;
    cudaError_t e = cudaMemcpyAsync(Inv, X, sizeof(double) * (size_t)batch * ms, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * (64 * 65 + 64 * 64);
    if (int rc_ = lvae_ensure_smem(k_trsmT_block_row, smem, attr)) return rc_;
    for (int i = nb - 1; i >= 0; --i) {
        if (i < nb - 1) {                             // R_i = X_i - L[i+1:, i]^T Inv[i+1:, :]
            GemmDesc a;
            a.A = F + (size_t)(64 * (i + 1)) * np + 64 * i; a.lda = np; a.sA = ms; a.ta = 1;
            a.B = Inv + (size_t)(64 * (i + 1)) * np; a.ldb = np; a.sB = ms;
            a.C = Inv + (size_t)(64 * i) * np; a.ldc = np; a.sC = ms;
            a.m = 64; a.n = np; a.k = np - 64 * (i + 1); a.batch = batch;
            a.alpha = -1.0; a.beta = 1.0;
            const int rc = lvae_gemm(a, st);
            if (rc) return rc;
        }
        k_trsmT_block_row<<<dim3(nb, batch), 64, smem, st>>>(F, Inv, np, i);
        LVAE_COUNT_LAUNCH();
    }
    return lvae_cuda_rc(cudaGetLastError());
}

// SPD inverse: blocked Cholesky of F (in place), X = F^-1, Inv = X^T X (symmetric, identity padded like F).  T scratch.
int lvae_spd_inverse_big(double* F, double* X, double* T, double* Inv, double* dinv, int np, int batch, int32_t* info,
                         int info_mod, cudaStream_t st) {
    int rc = lvae_potrf_big(F, np, batch, dinv, info, info_mod, st);
    if (rc) return rc;
    rc = lvae_trtri_big(F, dinv, X, T, np, batch, st);
    if (rc) return rc;
    return lvae_potrs_identity_big(F, X, Inv, np, batch, st);
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
    if (!rc) rc = lvae_potrs_identity_big(F, X, T, np, batch, st);
    if (!rc) rc = lvae_pad_out(Ainv, T, n, np, stride, batch, 0, st);
    cudaFreeAsync(buf, st);
    return rc;
}
