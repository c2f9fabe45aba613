// Internal batched FP64 building blocks on the sm_100a FP64 tensor pipe (DMMA.8x8x4): a tiled GEMM with cp.async
// staging and, on top of it, blocked Cholesky / triangular inverse / SPD inverse for matrices of order 64 < n <= 256
// that do not fit one CTA's shared memory.  Used by the M > 64 head / tail / natural-gradient paths, by the big
// subject pass (S = U^T U, Y = V W) and by the batched potrf / potri ABI.
#pragma once
#include "lvae_host.h"

// C[b2][b1][s] (m x n) = alpha * op(A) op(B) + beta * C, row-major.
//   op(A) is m x k: ta == 0 -> A[i*lda + kk], ta == 1 -> A[kk*lda + i];  op(B) is k x n: tb == 0 -> B[kk*ldb + j],
//   tb == 1 -> B[j*ldb + kk].  Two batch dimensions (b1 < batch, b2 < batch2) with independent strides, and an optional
//   split of the k range into `ksplit` parts of `kchunk` (part s reads A + s*kA, B + s*kB and writes C + s*kC).
//   flags: LVAE_GEMM_LOWER  skip tiles strictly above the diagonal and entries j > i;
//          LVAE_GEMM_MIRROR also store C[j][i] for j < i (symmetric result; needs beta == 0).
// Pointers must be 16-byte aligned with even lda / ldb for the 16-byte cp.async path; otherwise the kernel falls back
// to 8-byte copies (detected per call).

struct GemmDesc {
    const double* A = nullptr;
    const double* B = nullptr;
    double* C = nullptr;
    int m = 0, n = 0, k = 0, lda = 0, ldb = 0, ldc = 0;
    int ta = 0, tb = 0;
    int batch = 1;
    int64_t sA = 0, sB = 0, sC = 0;
    int batch2 = 1;
    int64_t sA2 = 0, sB2 = 0, sC2 = 0;
    int ksplit = 1, kchunk = 0;
    int64_t kA = 0, kB = 0, kC = 0;
    double alpha = 1.0, beta = 0.0;
    int flags = 0;
    int flush = 0;              // > 0 (beta == 0 only): two-level sum, accumulators folded into C every `flush` k-tiles of 16
};
int lvae_gemm(const GemmDesc& d, cudaStream_t st);

// Padded order used by the blocked factorisations: 128 or 256 (matrices are identity-padded).
static inline int lvae_pad_order(int n) { return n <= 128 ? 128 : 256; }

// In-place blocked lower Cholesky of `batch` identity-padded np x np matrices (row stride np, matrix stride np*np).
// The strict upper triangle is NOT cleared.  dinv: [batch][np/64][64*64] receives the inverses of the diagonal blocks of
// the factor.  info_slot[index / info_mod] (info_mod > 0; else info_slot[0]) is set to 1 + (index % info_mod) of the first
// failing matrix.
int lvae_potrf_big(double* F, int np, int batch, double* dinv, int32_t* info_slot, int info_mod, cudaStream_t st);
// X = F^-1 for the lower factor computed by lvae_potrf_big (X fully written, upper triangle zero); T: scratch batch*np*np.
int lvae_trtri_big(const double* F, const double* dinv, double* X, double* T, int np, int batch, cudaStream_t st);
// Inv = X^T X (symmetric, full).
int lvae_gram_big(const double* X, double* Inv, int np, int batch, cudaStream_t st);
// Inv = F^-T X (block back substitution; with X = F^-1 this is the SPD inverse with LAPACK potrs' small residual).
int lvae_potrs_identity_big(const double* F, const double* X, double* Inv, int np, int batch, cudaStream_t st);

// SPD inverse: blocked Cholesky of F in place, X = F^-1, Inv = X^T X (symmetric, full); T scratch.
int lvae_spd_inverse_big(double* F, double* X, double* T, double* Inv, double* dinv, int np, int batch, int32_t* info,
                         int info_mod, cudaStream_t st);
int lvae_pad_in(double* dst, const double* src, int n, int np, int64_t sstride, int batch, double diag, int lower_only,
                cudaStream_t st);
int lvae_pad_out(double* dst, const double* src, int n, int np, int64_t dstride, int batch, int lower_only, cudaStream_t st);
int lvae_potrf_big_abi(double* A, int n, int64_t stride, int batch, int32_t* info, cudaStream_t st);
int lvae_potri_big_abi(const double* Lc, double* Ainv, int n, int64_t stride, int batch, cudaStream_t st);
