// Materialising kernels: dense additive kernel matrices and per-subject blocks from covariates, batched
// Cholesky / explicit inverse.  These back the drop-in `covar_module(x1, x2).evaluate()` of the host mirror and the
// parity tests on kernel matrices and Cholesky factors; the training hot path never materialises Kxz (see
// lvae_subjects_fused.cu).  The dense kernel is HBM-write-bound: 8*L*n1*n2 bytes out, covariates stay in L1/L2.
#include "lvae_host.h"
#include "lvae_blas.h"
#include "lvae_linalg.cuh"

int64_t& lvae_launch_counter() {
    static int64_t c = 0;
    return c;
}

cudaError_t lvae_scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    static bool tuned[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !tuned[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = 2ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned[dev] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 8, st);
}

int lvae_make_devspec(const lvae_kernel_spec_t* ks, int Q, DevSpec* out) {
    if (!ks || !ks->spec) return LVAE_E_BADARG;
    const int nc = ks->n_comp0 + ks->n_comp1;
    if (nc < 0 || nc > LVAE_MAXC || ks->n_ls < 0 || ks->n_ls > LVAE_MAXC) return LVAE_E_SPEC;
    DevSpec s;
    memset(&s, 0, sizeof(s));
    s.n0 = ks->n_comp0;
    s.n1 = ks->n_comp1;
    s.n_ls = ks->n_ls;
    for (int c = 0; c < nc; ++c) {
        const int32_t* r = ks->spec + (size_t)c * LVAE_SPEC_STRIDE;
        if (r[0] >= Q || r[0] < -1) return LVAE_E_SPEC;
        if (r[0] >= 0 && (r[1] < 0 || r[1] >= ks->n_ls)) return LVAE_E_SPEC;
        if (r[2] < 0 || r[2] > LVAE_MAX_MASKS) return LVAE_E_SPEC;
        s.rbf_dim[c] = (signed char)r[0];
        s.ls_idx[c] = (signed char)(r[0] >= 0 ? r[1] : 0);
        s.n_mask[c] = (signed char)r[2];
        for (int i = 0; i < r[2]; ++i) {
            const int ty = r[3 + 2 * i], dm = r[4 + 2 * i];
            if ((ty != LVAE_CAT && ty != LVAE_BIN) || dm < 0 || dm >= Q) return LVAE_E_SPEC;
            s.mask_type[c][i] = (signed char)ty;
            s.mask_dim[c][i] = (signed char)dm;
        }
    }
    *out = s;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// dense kernel: grid (ceil(n2/64), ceil(n1/4), L), block (64, 4); each thread one output element, coalesced on j.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dense(DevSpec sp, int c0, int c1, int Q, const double* __restrict__ x1,
                                               int64_t s1, int n1, const double* __restrict__ x2, int64_t s2, int n2,
                                               const double* __restrict__ ls, const double* __restrict__ os, int L,
                                               const double* __restrict__ diag_add, double* __restrict__ out) {
    __shared__ double hil2[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    const int b = blockIdx.z, l = b % L;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < LVAE_EXP_TBL) etab[t] = c_exp2_tbl[t];
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; hil2[t] = 0.5 / (v * v); }
    if (t < sp.n0 + sp.n1) osc[t] = os[(size_t)t * L + l];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n1 || j >= n2) return;
    const double* xa = x1 + (size_t)b * s1 + (size_t)i * Q;
    const double* xb = x2 + (size_t)b * s2 + (size_t)j * Q;
    double acc = 0.0, d2;
    for (int c = c0; c < c1; ++c) acc += osc[c] * comp_value(sp, c, xa, xb, hil2, d2, etab);
    if (diag_add && i == j) acc += diag_add[l];
    out[((size_t)b * n1 + i) * n2 + j] = acc;
}

extern "C" int lvae_kernel_dense_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                                     int32_t n_batch, int32_t Q, const double* x1, int64_t s1, int32_t n1, const double* x2, int64_t s2,
                                     int32_t n2, const double* lengthscale, const double* outputscale,
                                     const double* diag_add, double* out, void* stream) {
    DevSpec sp;
    int rc = lvae_make_devspec(ks, Q, &sp);
    if (rc) return rc;
    if (comp_begin < 0 || comp_end > sp.n0 + sp.n1 || comp_begin > comp_end || L <= 0) return LVAE_E_BADARG;
    if (n_batch <= 0 || n_batch % L != 0) return LVAE_E_BADARG;
    if (n_batch > 65535) return LVAE_E_TOO_LARGE;
    if (n1 == 0 || n2 == 0) return 0;
    dim3 block(64, 4), grid((n2 + 63) / 64, (n1 + 3) / 4, n_batch);
    if (grid.y > 65535) return LVAE_E_TOO_LARGE;
    k_dense<<<grid, block, 0, (cudaStream_t)stream>>>(sp, comp_begin, comp_end, Q, x1, s1, n1, x2, s2, n2, lengthscale,
                                                      outputscale, L, diag_add, out);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// per-subject blocks: grid (P_b, L); CTA loops over the T_p x T_p block.  off2[p] computed by a prefix over T_q^2
// (serial per CTA over p' < p would be O(P^2); instead thread 0 of each CTA reads a precomputed table).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_block_offsets(const int32_t* __restrict__ offsets, int P_b, int64_t* __restrict__ off2) {
    // single CTA exclusive scan of T_p^2 (P_b up to ~1e6: chunked serial per thread + block scan)
    extern __shared__ int64_t part[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (P_b + nt - 1) / nt;
    const int b = tid * per, e = min(P_b, b + per);
    int64_t s = 0;
    for (int p = b; p < e; ++p) { const int64_t T = offsets[p + 1] - offsets[p]; s += T * T; }
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int64_t run = 0;
        for (int i = 0; i < nt; ++i) { const int64_t v = part[i]; part[i] = run; run += v; }
        off2[P_b] = run;
    }
    __syncthreads();
    s = part[tid];
    for (int p = b; p < e; ++p) { off2[p] = s; const int64_t T = offsets[p + 1] - offsets[p]; s += T * T; }
}

__global__ void __launch_bounds__(128) k_blocks(DevSpec sp, int c0, int c1, int Q, const double* __restrict__ x,
                                                const int32_t* __restrict__ offsets, const int64_t* __restrict__ off2,
                                                int64_t block_stride, const double* __restrict__ ls,
                                                const double* __restrict__ os, int L,
                                                const double* __restrict__ diag_add, double* __restrict__ out) {
    __shared__ double hil2[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    const int p = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
    if (t < LVAE_EXP_TBL) etab[t] = c_exp2_tbl[t];
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; hil2[t] = 0.5 / (v * v); }
    if (t < sp.n0 + sp.n1) osc[t] = os[(size_t)t * L + l];
    __syncthreads();
    const int r0 = offsets[p], T = offsets[p + 1] - r0;
    double* o = out + (size_t)l * block_stride + off2[p];
    for (int e = t; e < T * T; e += blockDim.x) {
        const int i = e / T, j = e % T;
        const double* xa = x + (size_t)(r0 + i) * Q;
        const double* xb = x + (size_t)(r0 + j) * Q;
        double acc = 0.0, d2;
        for (int c = c0; c < c1; ++c) acc += osc[c] * comp_value(sp, c, xa, xb, hil2, d2, etab);
        if (diag_add && i == j) acc += diag_add[l];
        o[e] = acc;
    }
}

int lvae_block_offsets(const int32_t* offsets, int P_b, int64_t* off2, cudaStream_t st) {
    // simple and robust: one thread per subject is not a scan; do the scan in one CTA of 256 threads
    k_block_offsets<<<1, 256, 256 * sizeof(int64_t), st>>>(offsets, P_b, off2);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int lvae_kernel_blocks_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                                      int32_t Q, const double* x, const int32_t* offsets, int32_t P_b,
                                      int64_t block_stride, const double* lengthscale, const double* outputscale,
                                      const double* diag_add, double* out, void* stream) {
    DevSpec sp;
    int rc = lvae_make_devspec(ks, Q, &sp);
    if (rc) return rc;
    if (comp_begin < 0 || comp_end > sp.n0 + sp.n1 || comp_begin > comp_end || L <= 0 || L > 65535) return LVAE_E_BADARG;
    if (P_b == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t* off2 = nullptr;
    cudaError_t e = lvae_scratch_alloc((void**)&off2, sizeof(int64_t) * ((size_t)P_b + 1), st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    rc = lvae_block_offsets(offsets, P_b, off2, st);
    if (!rc) {
        k_blocks<<<dim3(P_b, L), 128, 0, st>>>(sp, comp_begin, comp_end, Q, x, offsets, off2, block_stride, lengthscale,
                                               outputscale, L, diag_add, out);
        LVAE_COUNT_LAUNCH();
        rc = lvae_cuda_rc(cudaGetLastError());
    }
    cudaFreeAsync(off2, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// batched Cholesky / inverse (generic, one CTA per matrix, operating in place in global/L2 memory)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_potrf(double* __restrict__ A, int n, int64_t stride, int32_t* info) {
    __shared__ int flag;
    const int rc = cta_cholesky(A + (size_t)blockIdx.x * stride, n, n, &flag);
    if (rc && threadIdx.x == 0) atomicCAS(info, 0, (int)blockIdx.x + 1);
}

__global__ void __launch_bounds__(256) k_potri(const double* __restrict__ Lc, double* __restrict__ Ainv,
                                               double* __restrict__ tmp, int n, int64_t stride) {
    const double* Lb = Lc + (size_t)blockIdx.x * stride;
    double* X = tmp + (size_t)blockIdx.x * n * n;
    cta_tri_inverse(Lb, X, n, n);
    cta_gram_lower(X, Ainv + (size_t)blockIdx.x * stride, n, n);
}

// ---------------------------------------------------------------------------------------------------------------
// n <= 32 (the per-subject T x T blocks of deviance_upper_bound / elbo / batch_predict, thousands per call): one WARP per
// matrix in shared memory instead of one 256-thread CTA per matrix in global memory.  Lane i owns row i of the factor;
// odd row stride -> conflict-free column walks.  Same operation order per entry as cta_cholesky.
// ---------------------------------------------------------------------------------------------------------------
#define SMALL_WARPS 4
__global__ void __launch_bounds__(32 * SMALL_WARPS) k_potrf_warp(double* __restrict__ A, int n, int64_t stride, int batch,
                                                                 int32_t* info) {
    extern __shared__ double sm_small[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, ld = n | 1;
    const int64_t b = (int64_t)blockIdx.x * SMALL_WARPS + w;
    if (b >= batch) return;
    double* S = sm_small + (size_t)w * n * ld;
    double* Ab = A + b * stride;
    for (int i = 0; i < n; ++i)
        if (lane < n) S[i * ld + lane] = Ab[(size_t)i * n + lane];
    __syncwarp();
    int fail = 0;
    for (int k = 0; k < n; ++k) {
        double d = S[k * ld + k];
        if (!(d > 0.0)) { if (!fail) fail = k + 1; d = nan(""); }
        else d = sqrt(d);
        __syncwarp();
        if (lane == k) S[k * ld + k] = d;
        const double inv = 1.0 / d;
        const bool below = lane > k && lane < n;
        double lik = 0.0;
        if (below) { lik = S[lane * ld + k] * inv; S[lane * ld + k] = lik; }
        __syncwarp();
        if (below)
            for (int j = k + 1; j <= lane; ++j) S[lane * ld + j] -= lik * S[j * ld + k];
        __syncwarp();
    }
    for (int i = 0; i < n; ++i)
        if (lane < n) Ab[(size_t)i * n + lane] = lane <= i ? S[i * ld + lane] : 0.0;
    if (fail && lane == 0) atomicCAS(info, 0, (int)b + 1);
}

// Ainv = L^-T L^-1 from the lower factor: lane c solves L x = e_c (column c of X = L^-1), then lane b forms column b of X^T X.
__global__ void __launch_bounds__(32 * SMALL_WARPS) k_potri_warp(const double* __restrict__ Lc, double* __restrict__ Ainv,
                                                                 int n, int64_t stride, int batch) {
    extern __shared__ double sm_small[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, ld = n | 1;
    const int64_t b = (int64_t)blockIdx.x * SMALL_WARPS + w;
    if (b >= batch) return;
    double* S = sm_small + (size_t)w * 2 * n * ld;
    double* X = S + (size_t)n * ld;
    const double* Lb = Lc + b * stride;
    for (int i = 0; i < n; ++i)
        if (lane < n) S[i * ld + lane] = Lb[(size_t)i * n + lane];
    __syncwarp();
    if (lane < n) {
        for (int i = 0; i < n; ++i) {                      // uniform trip counts; the predicate k >= lane selects the terms
            double s = 0.0;
            for (int k = 0; k < i; ++k) {
                const double lik = S[i * ld + k];          // broadcast
                if (k >= lane) s = fma(lik, X[k * ld + lane], s);
            }
            const double dii = S[i * ld + i];
            X[i * ld + lane] = i < lane ? 0.0 : (i == lane ? 1.0 / dii : -s / dii);
        }
    }
    __syncwarp();
    double* Ob = Ainv + b * stride;
    if (lane < n) {
        for (int a = 0; a < n; ++a) {
            double s = 0.0;
            for (int i = a; i < n; ++i) s = fma(X[i * ld + a], X[i * ld + lane], s);     // X[i][lane] = 0 for i < lane
            Ob[(size_t)a * n + lane] = s;
        }
    }
}

extern "C" int lvae_potrf_batched_f64(double* A, int32_t n, int64_t batch_stride, int32_t batch, int32_t* info,
                                      void* stream) {
    if (n <= 0 || n > LVAE_MAX_M || batch < 0 || batch_stride < (int64_t)n * n) return LVAE_E_BADARG;
    if (batch == 0) return 0;
    if (n > 64) return lvae_potrf_big_abi(A, n, batch_stride, batch, info, (cudaStream_t)stream);
    if (n <= 32 && batch >= 64) {
        const size_t smem = sizeof(double) * SMALL_WARPS * n * (n | 1);
        k_potrf_warp<<<(batch + SMALL_WARPS - 1) / SMALL_WARPS, 32 * SMALL_WARPS, smem, (cudaStream_t)stream>>>(
            A, n, batch_stride, batch, info);
        LVAE_COUNT_LAUNCH();
        return lvae_cuda_rc(cudaGetLastError());
    }
    k_potrf<<<batch, 256, 0, (cudaStream_t)stream>>>(A, n, batch_stride, info);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int lvae_potri_batched_f64(const double* Lc, double* Ainv, int32_t n, int64_t batch_stride, int32_t batch,
                                      void* stream) {
    if (n <= 0 || n > LVAE_MAX_M || batch < 0 || batch_stride < (int64_t)n * n) return LVAE_E_BADARG;
    if (batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 64) return lvae_potri_big_abi(Lc, Ainv, n, batch_stride, batch, st);
    if (n <= 32 && batch >= 64) {
        static SmemAttrCache attr;
        const size_t smem = sizeof(double) * SMALL_WARPS * 2 * n * (n | 1);
        // n = 32 needs 66 KB: above the 48 KB default
        if (int rc_ = lvae_ensure_smem(k_potri_warp, sizeof(double) * SMALL_WARPS * 2 * 32 * 33, attr)) return rc_;
        k_potri_warp<<<(batch + SMALL_WARPS - 1) / SMALL_WARPS, 32 * SMALL_WARPS, smem, st>>>(Lc, Ainv, n, batch_stride, batch);
        LVAE_COUNT_LAUNCH();
        return lvae_cuda_rc(cudaGetLastError());
    }
    double* tmp = nullptr;
    cudaError_t e = lvae_scratch_alloc((void**)&tmp, sizeof(double) * (size_t)batch * n * n, st);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    k_potri<<<batch, 256, 0, st>>>(Lc, Ainv, tmp, n, batch_stride);
    LVAE_COUNT_LAUNCH();
    int rc = lvae_cuda_rc(cudaGetLastError());
    cudaFreeAsync(tmp, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// per-phase timing
// ---------------------------------------------------------------------------------------------------------------
static int g_prof = 0;
static cudaEvent_t g_ev[LVAE_NPHASE][2];
static bool g_ev_init = false, g_ev_used[LVAE_NPHASE];
void lvae_prof_begin(int ph, cudaStream_t st) {
    if (!g_prof) return;
    if (!g_ev_init) {
        for (int i = 0; i < LVAE_NPHASE; ++i) { cudaEventCreate(&g_ev[i][0]); cudaEventCreate(&g_ev[i][1]); g_ev_used[i] = false; }
        g_ev_init = true;
    }
    cudaEventRecord(g_ev[ph][0], st);
}
void lvae_prof_end(int ph, cudaStream_t st) {
    if (!g_prof) return;
    cudaEventRecord(g_ev[ph][1], st);
    g_ev_used[ph] = true;
}
extern "C" int lvae_profile_enable(int on) { g_prof = on; return 0; }
// milliseconds of the most recent launch of `phase` (synchronises on its end event); <0 if never recorded
extern "C" float lvae_profile_last_ms(int ph) {
    if (ph < 0 || ph >= LVAE_NPHASE || !g_ev_init || !g_ev_used[ph]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(g_ev[ph][1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, g_ev[ph][0], g_ev[ph][1]) != cudaSuccess) return -1.f;
    return ms;
}

// test hook: the device exp used by every squared-exponential factor
__global__ void k_exp_neg(const double* __restrict__ x, double* __restrict__ out, int n) {
    __shared__ double etab[LVAE_EXP_TBL];
    load_exp_table(etab);
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = exp_neg(x[i], etab);
}
extern "C" int lvae_debug_exp_neg_f64(const double* x, double* out, int32_t n, void* stream) {
    if (n <= 0) return 0;
    k_exp_neg<<<(n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, out, n);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// one-shot all-reduce of the SVGP statistics over NVLink peer memory: every rank sums the peers' buffers (mapped into
// this process, e.g. torch symmetric memory) in RANK ORDER, so all ranks obtain bit-identical sums.  The caller provides
// the inter-GPU barrier before (peers' buffers complete) — see distributed.py.
// ---------------------------------------------------------------------------------------------------------------
struct PeerPtrs { const double* p[16]; };
__global__ void __launch_bounds__(256) k_peer_sum(PeerPtrs pp, int world, int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i < n; i += stride) {
        if (2 * i + 1 < n) {
            double2 acc = make_double2(0.0, 0.0);
            for (int r = 0; r < world; ++r) {
                const double2 v = *reinterpret_cast<const double2*>(pp.p[r] + 2 * i);
                acc.x += v.x; acc.y += v.y;
            }
            *reinterpret_cast<double2*>(out + 2 * i) = acc;
        } else {
            double acc = 0.0;
            for (int r = 0; r < world; ++r) acc += pp.p[r][2 * i];
            out[2 * i] = acc;
        }
    }
}
extern "C" int lvae_peer_sum_f64(const uint64_t* peer_ptrs, int32_t world, int64_t n, double* out, void* stream) {
    if (!peer_ptrs || world <= 0 || world > 16 || n < 0 || !out) return LVAE_E_BADARG;
    if (n == 0) return 0;
    PeerPtrs pp;
    for (int r = 0; r < 16; ++r) pp.p[r] = r < world ? reinterpret_cast<const double*>(peer_ptrs[r]) : nullptr;
    for (int r = 0; r < world; ++r)
        if (peer_ptrs[r] & 15) return LVAE_E_BADARG;
    if (reinterpret_cast<uintptr_t>(out) & 15) return LVAE_E_BADARG;
    int blocks = (int)((n / 2 + 255) / 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    k_peer_sum<<<blocks, 256, 0, (cudaStream_t)stream>>>(pp, world, n, out);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int64_t lvae_launch_count(void) { return lvae_launch_counter(); }
extern "C" const char* lvae_version(void) { return "lvae_b200 0.1 (sm_100a, fp64)"; }
