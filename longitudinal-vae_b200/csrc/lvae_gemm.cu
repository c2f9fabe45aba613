// Batched FP64 GEMM on the sm_100a FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; FP64 has no tcgen05 kind).
// CTA tile 128 x 64, k-panel 16, 3-stage cp.async pipeline, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 (16 DMMA per
// 8 fragment loads).  Shared-memory tiles keep the operand's own orientation (no transposition on the way in); every
// row stride is = 4 (mod 16) doubles so the fragment loads of either orientation are bank-conflict-free.
// Replaces the torch.matmul / einsum contractions of elbo_functions.py:183-184, 189, 194, 208-214 for M > 64 and carries
// the trailing updates of the blocked Cholesky / inverse (lvae_blas.cu).
#include "lvae_blas.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NST = 3;
constexpr int LDA_N = BK + 4;      // A tile [BM][20]   (op(A) = A)
constexpr int LDA_T = BM + 4;      // A tile [BK][132]  (op(A) = A^T)
constexpr int LDB_N = BN + 4;      // B tile [BK][68]   (op(B) = B)
constexpr int LDB_T = BK + 4;      // B tile [BN][20]   (op(B) = B^T)
constexpr int A_ST = BM * LDA_N;   // 2560 doubles (>= BK * LDA_T = 2112)
constexpr int B_ST = BN * LDB_T;   // 1280 doubles (>= BK * LDB_N = 1088)
constexpr int STAGE = A_ST + B_ST;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(void* smem, const void* g, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp8(void* smem, const void* g, int bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
// copy two consecutive doubles (nv of them valid, the rest zero-filled); never forms an out-of-range source address
__device__ __forceinline__ void copy2(double* dst, const double* src, int nv, int al16, const double* safe) {
    if (al16) {
        cp16(dst, nv > 0 ? src : safe, 8 * nv);
    } else {
        cp8(dst, nv > 0 ? src : safe, nv > 0 ? 8 : 0);
        cp8(dst + 1, nv > 1 ? src + 1 : safe, nv > 1 ? 8 : 0);
    }
}
__device__ __forceinline__ int clamp2(int v) { return v < 0 ? 0 : (v > 2 ? 2 : v); }

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) k_gemm(const __grid_constant__ GemmDesc d, int al16) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    if ((d.flags & LVAE_GEMM_LOWER) && col0 > row0 + BM - 1) return;
    int zz = blockIdx.z;
    const int b1 = zz % d.batch;
    zz /= d.batch;
    const int b2 = zz % d.batch2, s = zz / d.batch2;
    const int klen = min(d.kchunk, d.k - s * d.kchunk);
    const double* __restrict__ A = d.A + (size_t)b1 * d.sA + (size_t)b2 * d.sA2 + (size_t)s * d.kA;
    const double* __restrict__ B = d.B + (size_t)b1 * d.sB + (size_t)b2 * d.sB2 + (size_t)s * d.kB;
    double* __restrict__ C = d.C + (size_t)b1 * d.sC + (size_t)b2 * d.sC2 + (size_t)s * d.kC;
    const int nkt = klen > 0 ? (klen + BK - 1) / BK : 0;

    auto load_stage = [&](int stg, int kt) {
        double* As = sm + stg * STAGE;
        double* Bs = As + A_ST;
        const int k0 = kt * BK, krem = klen - k0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int ch = tid + 256 * it;
            if (!TA) {
                const int r = ch >> 3, c = ch & 7, row = row0 + r;
                const int nv = row < d.m ? clamp2(krem - 2 * c) : 0;
                copy2(As + r * LDA_N + 2 * c, A + (size_t)row * d.lda + k0 + 2 * c, nv, al16, d.A);
            } else {
                const int kk = ch >> 6, c = ch & 63;
                const int nv = kk < krem ? clamp2(d.m - row0 - 2 * c) : 0;
                copy2(As + kk * LDA_T + 2 * c, A + (size_t)(k0 + kk) * d.lda + row0 + 2 * c, nv, al16, d.A);
            }
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int ch = tid + 256 * it;
            if (!TB) {
                const int kk = ch >> 5, c = ch & 31;
                const int nv = kk < krem ? clamp2(d.n - col0 - 2 * c) : 0;
                copy2(Bs + kk * LDB_N + 2 * c, B + (size_t)(k0 + kk) * d.ldb + col0 + 2 * c, nv, al16, d.B);
            } else {
                const int r = ch >> 3, c = ch & 7, col = col0 + r;
                const int nv = col < d.n ? clamp2(krem - 2 * c) : 0;
                copy2(Bs + r * LDB_T + 2 * c, B + (size_t)col * d.ldb + k0 + 2 * c, nv, al16, d.B);
            }
        }
    };

    const int wm = wid & 3, wn = wid >> 2;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int st = 0; st < NST - 1; ++st) {
        if (st < nkt) load_stage(st, st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < nkt; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
        __syncthreads();
        if (kt + NST - 1 < nkt) load_stage((kt + NST - 1) % NST, kt + NST - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* As = sm + (kt % NST) * STAGE;
        const double* Bs = As + A_ST;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                a[mt] = TA ? As[(4 * ks + q) * LDA_T + wm * 32 + mt * 8 + g] : As[(wm * 32 + mt * 8 + g) * LDA_N + 4 * ks + q];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                b[nt] = TB ? Bs[(wn * 32 + nt * 8 + g) * LDB_T + 4 * ks + q] : Bs[(4 * ks + q) * LDB_N + wn * 32 + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        // Two-level sum over a long k range (d.flush > 0, beta == 0): every `flush` k-tiles the accumulators are added to
        // this thread's own elements of C (L2-resident, no other thread touches them) and restart from zero, so no rounding
        // chain is longer than 4 * flush DMMA steps + (k-tiles / flush) additions.
        if (d.flush > 0 && kt + 1 < nkt && (kt + 1) % d.flush == 0) {
            const bool first = kt + 1 == d.flush;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const int i = row0 + wm * 32 + mt * 8 + g;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int j = col0 + wn * 32 + nt * 8 + 2 * q;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (i < d.m && j + e < d.n && !((d.flags & LVAE_GEMM_LOWER) && j + e > i)) {
                            double* c = C + (size_t)i * d.ldc + j + e;
                            *c = first ? acc[mt][nt][e] : *c + acc[mt][nt][e];
                        }
                        acc[mt][nt][e] = 0.0;
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    const bool lower = d.flags & LVAE_GEMM_LOWER, mirror = d.flags & LVAE_GEMM_MIRROR;
    const bool flushed = d.flush > 0 && nkt > d.flush;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int i = row0 + wm * 32 + mt * 8 + g;
        if (i >= d.m) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = col0 + wn * 32 + nt * 8 + 2 * q + e;
                if (j >= d.n || (lower && j > i)) continue;
                double v = acc[mt][nt][e];
                if (flushed) v += C[(size_t)i * d.ldc + j];
                v *= d.alpha;
                if (d.beta != 0.0) v += d.beta * C[(size_t)i * d.ldc + j];
                C[(size_t)i * d.ldc + j] = v;
                if (mirror && j < i) C[(size_t)j * d.ldc + i] = v;
            }
        }
    }
}

template <bool TA, bool TB>
int launch(const GemmDesc& d, int al16, cudaStream_t st) {
    static SmemAttrCache attr;
    const size_t smem = sizeof(double) * NST * STAGE;
    if (int rc_ = lvae_ensure_smem(k_gemm<TA, TB>, smem, attr)) return rc_;
    const dim3 grid((d.n + BN - 1) / BN, (d.m + BM - 1) / BM, (unsigned)(d.batch * d.batch2 * d.ksplit));
    k_gemm<TA, TB><<<grid, 256, smem, st>>>(d, al16);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

int lvae_gemm(const GemmDesc& din, cudaStream_t st) {
    GemmDesc d = din;
    if (d.m <= 0 || d.n <= 0 || d.batch <= 0 || d.batch2 <= 0) return 0;
    if (d.ksplit <= 1) { d.ksplit = 1; d.kchunk = d.k; }
    if (d.beta != 0.0) d.flush = 0;               // the intermediate sums live in C itself
    if ((int64_t)d.batch * d.batch2 * d.ksplit > 65535) return LVAE_E_TOO_LARGE;
    const auto even = [](int64_t v) { return (v & 1) == 0; };
    const int al16 = ((reinterpret_cast<uintptr_t>(d.A) | reinterpret_cast<uintptr_t>(d.B)) & 15) == 0 && even(d.lda) &&
                     even(d.ldb) && even(d.sA) && even(d.sB) && even(d.sA2) && even(d.sB2) && even(d.kA) && even(d.kB);
    if (d.ta) return d.tb ? launch<true, true>(d, al16, st) : launch<true, false>(d, al16, st);
    return d.tb ? launch<false, true>(d, al16, st) : launch<false, false>(d, al16, st);
}

// C[b][i][j] = alpha * sum_s part[s][b][i][j] + beta * C[b][i][j]; fixed summation order.  lower_only: entries j > i untouched.
__global__ void __launch_bounds__(256) k_splitk_reduce(const double* __restrict__ part, int ksplit, int64_t per_split,
                                                       int m, int n, double alpha, double beta, double* __restrict__ C,
                                                       int ldc, int64_t sC, int64_t total, int lower_only) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / ((int64_t)m * n), r = e - b * (int64_t)m * n;
        const int i = (int)(r / n), j = (int)(r - (int64_t)i * n);
        if (lower_only && j > i) continue;
        double a = 0.0;
        for (int s = 0; s < ksplit; ++s) a += part[(size_t)s * per_split + e];
        double* c = C + (size_t)b * sC + (size_t)i * ldc + j;
        *c = beta == 0.0 ? alpha * a : fma(alpha, a, beta * *c);
    }
}

// C ABI (include/lvae_b200.h): one batch dimension.  A product with few output tiles and a long k (S = Kxz^T B^-1 Kxz over
// all rows of a data set: 60 x 60 x 20 000) would leave most of the 148 SMs idle, so the k range is split over ~2 waves of
// CTAs into a scratch buffer and summed in a fixed order.
// Stacks of thousands of tiny products (the per-subject T x T blocks of the non-minibatch bounds and of the predictors,
// elbo_functions.py:61-63,113-115, utils.py:165-190: B_p^-1 K0xz_p, B_p^-1 mu_p and their adjoints): one 128-thread CTA per
// matrix, 8 x 8 output tiles dealt to its four warps, DMMA fragments read straight from global memory (the operands of one
// product are a few KB and stay in L1).  A 128 x 64 tile of the large-matrix kernel above would be > 90 % padding here.
__global__ void __launch_bounds__(128) k_gemm_small(int ta, int tb, int m, int n, int k, double alpha,
                                                    const double* __restrict__ A, int lda, int64_t sA,
                                                    const double* __restrict__ B, int ldb, int64_t sB, double* __restrict__ C,
                                                    int ldc, int64_t sC) {
    const int b = blockIdx.x, wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const double* Ab = A + (size_t)b * sA;
    const double* Bb = B + (size_t)b * sB;
    double* Cb = C + (size_t)b * sC;
    const int mt = (m + 7) >> 3, nt = (n + 7) >> 3, nk = (k + 3) >> 2;
    for (int tile = wid; tile < mt * nt; tile += 4) {
        const int ti = tile / nt, tj = tile - ti * nt;
        const int i = 8 * ti + g, j = 8 * tj + g;
        double c0 = 0.0, c1 = 0.0;
        for (int ks = 0; ks < nk; ++ks) {
            const int kk = 4 * ks + q;
            const bool kv = kk < k;
            const double a = (kv && i < m) ? (ta ? Ab[(size_t)kk * lda + i] : Ab[(size_t)i * lda + kk]) : 0.0;
            const double bb = (kv && j < n) ? (tb ? Bb[(size_t)j * ldb + kk] : Bb[(size_t)kk * ldb + j]) : 0.0;
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0), "+d"(c1) : "d"(a), "d"(bb));
        }
        const int jo = 8 * tj + 2 * q;
        if (i < m && jo < n) Cb[(size_t)i * ldc + jo] = alpha * c0;
        if (i < m && jo + 1 < n) Cb[(size_t)i * ldc + jo + 1] = alpha * c1;
    }
}

extern "C" int lvae_gemm_batched_f64(int32_t trans_a, int32_t trans_b, int32_t m, int32_t n, int32_t k, double alpha,
                                     const double* A, int32_t lda, int64_t stride_a, const double* B, int32_t ldb,
                                     int64_t stride_b, double beta, double* C, int32_t ldc, int64_t stride_c,
                                     int32_t batch, int32_t flags, void* stream) {
    if (m < 0 || n < 0 || k < 0 || batch < 0 || !A || !B || !C) return LVAE_E_BADARG;
    if ((flags & LVAE_GEMM_MIRROR) && beta != 0.0) return LVAE_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (batch >= 256 && beta == 0.0 && flags == 0 && m <= 256 && n <= 256 && k <= 256 &&
        (int64_t)m * n * k <= (int64_t)48 * 256 * 64) {         // many tiny products: one CTA per matrix
        if (batch == 0 || m == 0 || n == 0) return 0;
        k_gemm_small<<<batch, 128, 0, st>>>(trans_a, trans_b, m, n, k, alpha, A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c);
        LVAE_COUNT_LAUNCH();
        return lvae_cuda_rc(cudaGetLastError());
    }
    GemmDesc d;
    d.A = A; d.B = B; d.C = C;
    d.m = m; d.n = n; d.k = k; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.ta = trans_a; d.tb = trans_b;
    d.batch = batch; d.sA = stride_a; d.sB = stride_b; d.sC = stride_c;
    d.alpha = alpha; d.beta = beta; d.flags = flags;
    const int64_t ctas = (int64_t)((m + BM - 1) / BM) * ((n + BN - 1) / BN) * batch;
    int ksplit = 1;
    if (ctas > 0 && ctas < 148 && k >= 2048) {
        int64_t want = (2 * 148 + ctas - 1) / ctas;
        if (want > k / 512) want = k / 512;
        if (want * batch > 65535) want = 65535 / batch;
        ksplit = (int)want;
    }
    if (ksplit <= 1) return lvae_gemm(d, st);
    int kchunk = (k + ksplit - 1) / ksplit;
    kchunk = (kchunk + BK - 1) / BK * BK;
    ksplit = (k + kchunk - 1) / kchunk;
    const int64_t per_split = (int64_t)batch * m * n;
    double* part = nullptr;
    cudaError_t e = lvae_scratch_alloc((void**)&part, sizeof(double) * (size_t)per_split * ksplit, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    d.C = part; d.ldc = n; d.sC = (int64_t)m * n;
    d.ksplit = ksplit; d.kchunk = kchunk;
    d.kA = trans_a ? (int64_t)kchunk * lda : kchunk;
    d.kB = trans_b ? kchunk : (int64_t)kchunk * ldb;
    d.kC = per_split;
    d.alpha = 1.0; d.beta = 0.0;
    int rc = lvae_gemm(d, st);
    if (!rc) {
        const int lower_only = (flags & LVAE_GEMM_LOWER) && !(flags & LVAE_GEMM_MIRROR);
        const int blocks = (int)((per_split + 255) / 256 < 148 * 8 ? (per_split + 255) / 256 : 148 * 8);
        k_splitk_reduce<<<blocks, 256, 0, st>>>(part, ksplit, per_split, m, n, alpha, beta, C, ldc, stride_c, per_split,
                                                lower_only);
        LVAE_COUNT_LAUNCH();
        rc = lvae_cuda_rc(cudaGetLastError());
    }
    cudaFreeAsync(part, st);
    return rc;
}
