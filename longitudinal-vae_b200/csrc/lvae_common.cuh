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
                                             double& d2) {
    double f = comp_mask(s, c, xa, xb);
    d2 = 0.0;
    const int rd = s.rbf_dim[c];
    if (rd >= 0) {
        const double d = xa[rd] - xb[rd];
        d2 = d * d;
        if (f != 0.0) f *= exp(-d2 * half_inv_l2[s.ls_idx[c]]);
    }
    return f;
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
