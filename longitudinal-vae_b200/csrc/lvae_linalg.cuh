// CTA-cooperative dense FP64 linear algebra on matrices that live in shared or global (L2-resident) memory.
// Generic path: used by the head/tail/natural-gradient kernels for any M <= 256 and by the batched potrf/potri ABI.
// All functions must be called by every thread of the CTA; they end with __syncthreads().
#pragma once
#include "lvae_common.cuh"

// In-place lower Cholesky, row-major, right-looking.  Returns (to all threads) 0 or 1 + failing column.
// `flag` is one int in shared memory.
__device__ inline int cta_cholesky(double* __restrict__ A, int n, int ld, int* flag) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = A[k * ld + k];
            if (!(d > 0.0)) { if (*flag == 0) *flag = k + 1; A[k * ld + k] = nan(""); }
            else A[k * ld + k] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / A[k * ld + k];
        for (int i = k + 1 + tid; i < n; i += nt) A[i * ld + k] *= inv;
        __syncthreads();
        const int rem = n - k - 1;
        // trailing lower triangle (rows i>k, cols k<j<=i): flatten rem x rem, keep j<=i
        for (int e = tid; e < rem * rem; e += nt) {
            const int i = k + 1 + e / rem, j = k + 1 + e % rem;
            if (j <= i) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
        }
        __syncthreads();
    }
    // zero the strict upper triangle so the factor can be used as a dense matrix
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        if (j > i) A[i * ld + j] = 0.0;
    }
    __syncthreads();
    return *flag;
}

// X = Lc^{-1} (lower), one thread per column (forward substitution), row-major.
__device__ inline void cta_tri_inverse(const double* __restrict__ Lc, double* __restrict__ X, int n, int ld) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < n; j += nt) {
        for (int i = 0; i < j; ++i) X[i * ld + j] = 0.0;
        X[j * ld + j] = 1.0 / Lc[j * ld + j];
        for (int i = j + 1; i < n; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s += Lc[i * ld + k] * X[k * ld + j];
            X[i * ld + j] = -s / Lc[i * ld + i];
        }
    }
    __syncthreads();
}

// Ainv = X^T X with X lower triangular (so A^{-1} = L^{-T} L^{-1}); symmetric output written in full.
__device__ inline void cta_gram_lower(const double* __restrict__ X, double* __restrict__ Ainv, int n, int ld) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        if (j > i) continue;
        double s = 0.0;
        for (int k = i; k < n; ++k) s += X[k * ld + i] * X[k * ld + j];
        Ainv[i * ld + j] = s;
        Ainv[j * ld + i] = s;
    }
    __syncthreads();
}

// C = alpha * op(A) * op(B) (+ beta * C), n x n, naive (each thread owns output elements).
template <bool TA, bool TB>
__device__ inline void cta_gemm_nn(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                   int n, int ld, double alpha, double beta) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        double s = 0.0;
        for (int k = 0; k < n; ++k) {
            const double a = TA ? A[k * ld + i] : A[i * ld + k];
            const double b = TB ? B[j * ld + k] : B[k * ld + j];
            s += a * b;
        }
        C[i * ld + j] = alpha * s + (beta != 0.0 ? beta * C[i * ld + j] : 0.0);
    }
    __syncthreads();
}

// y = A x  (n x n times n)
__device__ inline void cta_gemv(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y,
                                int n, int ld) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += A[i * ld + k] * x[k];
        y[i] = s;
    }
    __syncthreads();
}

// sum_i log(L_ii) * 2
__device__ inline double cta_logdet_from_chol(const double* __restrict__ Lc, int n, int ld, double* red) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += log(Lc[i * ld + i]);
    return 2.0 * block_sum(v, red);
}
