// Shared device-side pieces of the L-VAE GP-prior ELBO path: the flattened additive-kernel spec, component
// evaluation from covariates (kernel_spec.py:22-32, GP_model.py:31-144) and small block reductions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lvae_b200.h"

#define LVAE_MAXC LVAE_MAX_COMPONENTS

// One additive component: outputscale * prod(masks) * SE(rbf_dim).  Passed by value (kernel parameter space).
struct DevSpec {
    int n0, n1, n_ls;
    signed char rbf_dim[LVAE_MAXC];
    signed char ls_idx[LVAE_MAXC];
    signed char n_mask[LVAE_MAXC];
    signed char mask_type[LVAE_MAXC][LVAE_MAX_MASKS];
    signed char mask_dim[LVAE_MAXC][LVAE_MAX_MASKS];
};

// layout of the per-latent hyper-gradient vector: [d_lengthscale (n_ls) | d_outputscale (n_comp) | d_noise]
__host__ __device__ inline int hyp_count(const DevSpec& s) { return s.n_ls + s.n0 + s.n1 + 1; }

// layout of the per-latent statistics row (see lvae_kld_stats_stride)
#define LVAE_NSCAL 8
enum { SC_A = 0, SC_BT = 1, SC_C = 2, SC_D1 = 3, SC_F = 4 };
__host__ __device__ inline int64_t stats_off_S() { return 0; }
__host__ __device__ inline int64_t stats_off_ng1(int M) { return (int64_t)M * M; }
__host__ __device__ inline int64_t stats_off_da(int M) { return (int64_t)M * M + M; }
__host__ __device__ inline int64_t stats_off_scal(int M) { return (int64_t)M * M + 2 * M; }
__host__ __device__ inline int64_t stats_off_hyp(int M) { return (int64_t)M * M + 2 * M + LVAE_NSCAL; }
__host__ __device__ inline int64_t stats_stride(int M, int nh) { return (int64_t)M * M + 2 * M + LVAE_NSCAL + nh; }

// ---------------------------------------------------------------------------------------------------------------
// exp(x) for x <= 0 (squared-exponential kernels), ~16 instructions instead of libdevice's ~45 with its special-case
// branches: x = (64 k_hi + j) ln2/64 + r, |r| <= ln2/128, exp(x) = 2^k_hi * 2^(j/64) * (1 + r + ... + r^5/120).
// Truncation error r^6/720 <= 3.6e-17 relative; the 64-entry table of 2^(j/64) lives in shared memory (a divergent
// index would serialise on the constant cache).  Arguments below -700 (result < 1e-304) return 0.
// ---------------------------------------------------------------------------------------------------------------
#define LVAE_EXP_TBL 64
__device__ __constant__ double c_exp2_tbl[LVAE_EXP_TBL] = {   // 2^(j/64), correctly rounded
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};
__device__ __forceinline__ void load_exp_table(double* tbl) {
    for (int j = threadIdx.x; j < LVAE_EXP_TBL; j += blockDim.x) tbl[j] = c_exp2_tbl[j];
}
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tbl) {
    const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52: round-to-nearest-integer trick
    const double t = fma(x, 92.33248261689366, MAGIC);             // 64 / ln 2
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -0.010830424493178725, x);                  // ln2/64, high part (27 trailing zero bits)
    r = fma(kd, -2.030704202170295e-10, r);                        // ln2/64, low part
    // exp(r) - 1 = r + r^2 ((1/2 + r/6) + r^2 (1/24 + r/120)), evaluated Estrin-style: dependency depth 4 instead of Horner's 6
    // (the kernels that call this run 4 warps per scheduler and are bound by exactly such dependency chains)
    const double r2 = r * r;
    const double hi_ = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    const double lo_ = fma(r, 1.6666666666666666e-01, 0.5);
    const double p = fma(fma(hi_, r2, lo_), r2, r);
    const double tj = tbl[k & (LVAE_EXP_TBL - 1)];
    const double v = fma(tj, p, tj);
    const int hi = __double2hiint(v) + ((k >> 6) << 20);
    const double res = __hiloint2double(hi, __double2loint(v));
    return x < -700.0 ? 0.0 : res;
}

// mask product of component c between covariate rows xa, xb (exact 0/1 arithmetic on float equality, as the reference)
__device__ __forceinline__ double comp_mask(const DevSpec& s, int c, const double* __restrict__ xa,
                                            const double* __restrict__ xb) {
    double f = 1.0;
#pragma unroll
    for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
        if (i < s.n_mask[c]) {
            const double a = xa[s.mask_dim[c][i]], b = xb[s.mask_dim[c][i]];
            const bool on = (s.mask_type[c][i] == LVAE_CAT) ? (a - b == 0.0) : (a + b == 2.0);
            f = on ? f : 0.0;
        }
    }
    return f;
}

// unscaled component value f_c = masks * exp(-d^2 * h) with h = 1/(2 l^2) given per lengthscale row in `half_inv_l2`;
// d2 receives the squared distance of the SE factor (0 if none).
__device__ __forceinline__ double comp_value(const DevSpec& s, int c, const double* __restrict__ xa,
                                             const double* __restrict__ xb, const double* __restrict__ half_inv_l2,
                                             double& d2, const double* __restrict__ etab) {
    double f = comp_mask(s, c, xa, xb);
    d2 = 0.0;
    const int rd = s.rbf_dim[c];
    if (rd >= 0) {
        const double d = xa[rd] - xb[rd];
        d2 = d * d;
        if (f != 0.0) f *= exp_neg(-d2 * half_inv_l2[s.ls_idx[c]], etab);
    }
    return f;
}

// single-element branch-free variant
__device__ __forceinline__ double comp_one(const DevSpec& s, int c, const double* __restrict__ xa,
                                           const double* __restrict__ xb, const double* __restrict__ half_inv_l2,
                                           const double* __restrict__ etab, double& d2) {
    bool on = true;
#pragma unroll
    for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
        if (i < s.n_mask[c]) {
            const int dm = s.mask_dim[c][i];
            const double a = xa[dm], b = xb[dm];
            on = on && ((s.mask_type[c][i] == LVAE_CAT) ? (a - b == 0.0) : (a + b == 2.0));
        }
    }
    double e = 1.0;
    d2 = 0.0;
    const int rd = s.rbf_dim[c];
    if (rd >= 0) {
        const double t = xa[rd] - xb[rd];
        d2 = t * t;
        e = exp_neg(-d2 * half_inv_l2[s.ls_idx[c]], etab);
    }
    return on ? e : 0.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum; result valid in thread 0 (and broadcast through smem to all); `red` >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

int64_t& lvae_launch_counter();
#define LVAE_COUNT_LAUNCH() (++lvae_launch_counter())
