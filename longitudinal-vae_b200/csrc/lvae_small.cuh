// CTA-cooperative FP64 linear algebra for matrices of order <= 64 held in shared memory (row stride SLD = 68 doubles,
// = 4 mod 16, so DMMA fragment loads in either orientation are bank-conflict-free).  512 threads (16 warps) per CTA.
// GEMMs and the Cholesky trailing updates run on the FP64 tensor pipe (DMMA.8x8x4).  Matrices are padded to a multiple of
// 8 with the identity (factorisations) or zeros (products); n8 = 8 * ceil(n / 8) is the padded order actually touched.
#pragma once
#include "lvae_common.cuh"

#define SLD 68
#define SMAT (64 * SLD)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// (i, j), j <= i, of the e-th element of a lower triangle stored row-major
__device__ __forceinline__ void tri_ij(int e, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    while (i * (i + 1) / 2 > e) --i;
    j = e - i * (i + 1) / 2;
}

// C(i,j) = sum_k opA(i,k) opB(k,j) over the padded order n8; every warp owns a 16 x 16 block of C (2 x 2 tiles) and hands
// each element to `epi(i, j, value)`.  TA: opA(i,k) = A[k][i];  TB: opB(k,j) = B[j][k].  No barrier inside.
template <bool TA, bool TB, class Epi>
__device__ __forceinline__ void s_gemm(const double* __restrict__ A, const double* __restrict__ B, int n8, Epi epi) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int i0 = 16 * (wid >> 2), j0 = 16 * (wid & 3);
    if (i0 >= n8 || j0 >= n8) return;
    double acc[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int nks = n8 >> 2;
#pragma unroll 2
    for (int ks = 0; ks < nks; ++ks) {
        const int k = 4 * ks + q;
        const double a0 = TA ? A[k * SLD + i0 + g] : A[(i0 + g) * SLD + k];
        const double a1 = TA ? A[k * SLD + i0 + 8 + g] : A[(i0 + 8 + g) * SLD + k];
        const double b0 = TB ? B[(j0 + g) * SLD + k] : B[k * SLD + j0 + g];
        const double b1 = TB ? B[(j0 + 8 + g) * SLD + k] : B[k * SLD + j0 + 8 + g];
        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) epi(i0 + 8 * a + g, j0 + 8 * b + 2 * q + e, acc[a][b][e]);
}

// In-place lower Cholesky of the n8 x n8 leading block of A (identity-padded beyond the true order), blocked by 8:
// warp 0 factors the 8 x 8 diagonal block, one thread per row solves the panel, DMMA updates the trailing matrix.
// dinv[64] receives 1 / L_kk.  Returns (to all threads) 0 or 1 + failing column.  Ends with __syncthreads().
__device__ inline int s_cholesky(double* __restrict__ A, int n8, double* __restrict__ dinv, int* flag) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nb = n8 >> 3;
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int kb = 0; kb < nb; ++kb) {
        const int k0 = 8 * kb;
        if (wid == 0) {
            for (int k = 0; k < 8; ++k) {
                const double akk = A[(k0 + k) * SLD + k0 + k];
                if (!(akk > 0.0) && lane == 0 && *flag == 0) *flag = k0 + k + 1;
                double ri = rsqrt(akk);
                ri = ri * (1.5 - 0.5 * akk * ri * ri);          // one Newton step on the FP64 rsqrt (guards the last ulps)
                const double d = akk * ri;
                __syncwarp();
                if (lane == k) { A[(k0 + k) * SLD + k0 + k] = d; dinv[k0 + k] = ri; }
                if (lane > k && lane < 8) A[(k0 + lane) * SLD + k0 + k] *= ri;
                __syncwarp();
                if (lane < 28) {                                   // 7 x 7 lower triangle at most; rows/cols > k only
                    int r, c_;
                    tri_ij(lane, r, c_);
                    r += 1; c_ += 1;                               // (r, c) in 1..7, c <= r
                    if (c_ > k && r < 8) {
                        A[(k0 + r) * SLD + k0 + c_] -= A[(k0 + r) * SLD + k0 + k] * A[(k0 + c_) * SLD + k0 + k];
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // panel: rows below the diagonal block, x D^T = a  (forward substitution along the row)
        for (int r = k0 + 8 + tid; r < n8; r += blockDim.x) {
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
        // trailing update: A22 -= L21 L21^T on the lower-triangular tiles
        const int rem = nb - kb - 1, ntile = rem * (rem + 1) / 2;
        for (int tl = wid; tl < ntile; tl += (blockDim.x >> 5)) {
            int ti, tj;
            tri_ij(tl, ti, tj);
            ti += kb + 1; tj += kb + 1;
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
    // clear the strict upper triangle (the factor is used as a dense operand afterwards)
    for (int e = tid; e < n8 * n8; e += blockDim.x) {
        const int i = e / n8, j = e % n8;
        if (j > i) A[i * SLD + j] = 0.0;
    }
    __syncthreads();
    return *flag;
}

// X = Lc^-1 for a lower-triangular n8 x n8 factor, blocked by 8: the 8 x 8 diagonal blocks by forward substitution (one
// thread per column), then block sub-diagonal by sub-diagonal  X_ij = -L_ii^-1 (sum_{k=j}^{i-1} L_ik X_kj): the block sum on
// the tensor pipe (one warp per block, a per-warp 8 x 8 scratch tile), L_ii^-1 applied by forward SUBSTITUTION (eight lanes,
// one column each) — Higham's Method 1B.  Multiplying by the explicit inverse X_ii instead (Method 1C, what round 1 did)
// measured 3-4 x LAPACK's error in grad_m / grad_H on the reference's nearly singular Kzz (cond 1e8 .. 1e9).
// scratch: 16 * 64 doubles.  Ends with __syncthreads().
__device__ inline void s_tri_inverse(const double* __restrict__ Lc, double* __restrict__ X, int n8,
                                     const double* __restrict__ dinv, double* __restrict__ scratch) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nb = n8 >> 3;
    for (int e = tid; e < n8 * n8; e += blockDim.x) X[(e / n8) * SLD + e % n8] = 0.0;
    __syncthreads();
    if (tid < n8) {
        const int j = tid, k0 = j & ~7;
        X[j * SLD + j] = dinv[j];
        for (int i = j + 1; i < k0 + 8; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s += Lc[i * SLD + k] * X[k * SLD + j];
            X[i * SLD + j] = -s * dinv[i];
        }
    }
    __syncthreads();
    double* tile = scratch + wid * 64;
    for (int d = 1; d < nb; ++d) {
        for (int bj = wid; bj + d < nb; bj += (blockDim.x >> 5)) {
            const int bi = bj + d;
            double t0 = 0.0, t1 = 0.0;
            for (int bk = bj; bk < bi; ++bk) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double a = Lc[(8 * bi + g) * SLD + 8 * bk + 4 * ks + q];
                    const double b = X[(8 * bk + 4 * ks + q) * SLD + 8 * bj + g];
                    dmma884(t0, t1, a, b);
                }
            }
            tile[g * 8 + 2 * q] = t0;
            tile[g * 8 + 2 * q + 1] = t1;
            __syncwarp();
            if (lane < 8) {                       // column `lane` of the block: solve L_ii x = -tile[:, lane]
                double xr[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double s_ = -tile[r * 8 + lane];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < r) s_ -= Lc[(8 * bi + r) * SLD + 8 * bi + k] * xr[k];
                    xr[r] = s_ * dinv[8 * bi + r];
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) X[(8 * bi + r) * SLD + 8 * bj + lane] = xr[r];
            }
            __syncwarp();
        }
        __syncthreads();
    }
}

// Ainv = Lc^-T X for X = Lc^-1: the second triangular solve of LAPACK's potrs with the identity as right-hand side (what
// torch.cholesky_solve(I, L) runs, elbo_functions.py:178,186).  Blocked by 8: one warp per block COLUMN bj walks the block
// rows bottom-up, T = X_ij - sum_{k>i} L_ki^T Ainv_kj on the tensor pipe, then L_ii^T Y = T by back substitution (eight lanes,
// one column each).  A column only depends on itself, so the warps never synchronise with each other.  Unlike the Gram
// product X^T X (s_gram) this keeps the residual |A Ainv - I| at LAPACK's level when A is nearly singular — the bound's
// Kxz Kzz^-1 and Kzz^-1 S Kzz^-1 cancel against exactly that residual.  Ends with __syncthreads().
__device__ inline void s_potrs_identity(const double* __restrict__ Lc, const double* __restrict__ X, double* __restrict__ Ainv,
                                        int n8, const double* __restrict__ dinv, double* __restrict__ scratch) {
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nb = n8 >> 3;
    __syncthreads();
    double* tile = scratch + wid * 64;
    for (int bj = wid; bj < nb; bj += (blockDim.x >> 5)) {
        for (int bi = nb - 1; bi >= 0; --bi) {
            double t0 = X[(8 * bi + g) * SLD + 8 * bj + 2 * q], t1 = X[(8 * bi + g) * SLD + 8 * bj + 2 * q + 1];
            for (int bk = bi + 1; bk < nb; ++bk) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const double a = -Lc[(8 * bk + 4 * ks + q) * SLD + 8 * bi + g];           // (L_ki^T)[g][k]
                    const double b = Ainv[(8 * bk + 4 * ks + q) * SLD + 8 * bj + g];
                    dmma884(t0, t1, a, b);
                }
            }
            tile[g * 8 + 2 * q] = t0;
            tile[g * 8 + 2 * q + 1] = t1;
            __syncwarp();
            if (lane < 8) {                       // column `lane` of the block: solve L_ii^T y = tile[:, lane], bottom-up
                double yr[8];
#pragma unroll
                for (int r = 7; r >= 0; --r) {
                    double s_ = tile[r * 8 + lane];
#pragma unroll
                    for (int k = 7; k >= 0; --k)
                        if (k > r) s_ -= Lc[(8 * bi + k) * SLD + 8 * bi + r] * yr[k];
                    yr[r] = s_ * dinv[8 * bi + r];
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) Ainv[(8 * bi + r) * SLD + 8 * bj + lane] = yr[r];
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

// Ainv = X^T X (A^-1 = L^-T L^-1), full symmetric output.  Ends with __syncthreads().
__device__ inline void s_gram(const double* __restrict__ X, double* __restrict__ Ainv, int n8) {
    s_gemm<true, false>(X, X, n8, [&](int i, int j, double v) { Ainv[i * SLD + j] = v; });
    __syncthreads();
}

// load an n x n row-major global matrix into a padded smem matrix; pad_diag on the padded diagonal, zeros elsewhere
__device__ inline void s_load(double* __restrict__ dst, const double* __restrict__ src, int n, double pad_diag) {
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
        const int i = e >> 6, j = e & 63;
        dst[i * SLD + j] = (i < n && j < n) ? src[i * n + j] : (i == j ? pad_diag : 0.0);
    }
}

__device__ inline void s_store(double* __restrict__ dst, const double* __restrict__ src, int n) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) dst[e] = src[(e / n) * SLD + e % n];
}
