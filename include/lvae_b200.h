/*
 * lvae_b200 — C ABI of the B200-native GP-prior ELBO path of L-VAE (SidRama/Longitudinal-VAE).
 *
 * The reference is pure Python/PyTorch and has no FFI layer of its own; the drop-in boundary is its Python call
 * signatures.  Each entry point below names the reference lines whose arithmetic it replaces; the
 * Python host side (longitudinal-vae_b200/*.py) mirrors the reference's functions over these calls through ctypes.
 *
 * Conventions: plain pointers and sizes only (no torch types).  Unless marked HOST, every pointer is a device pointer
 * to FP64 data owned by the caller for the duration of the call.  All work is enqueued on `stream` (a cudaStream_t
 * passed as void*); no call synchronises the device.  Return value: 0 = enqueued, <0 = -(cudaError_t), >0 = invalid
 * argument (LVAE_E_*).  Cholesky failures (non-PD block) are reported asynchronously in `info` (device int32[4]:
 * info[0] = 1 + index of the first failing latent for Kzz, info[1] for H, info[2] = 1 + flat (latent*P_b + subject)
 * for a per-subject block B_p, info[3] for the natural-gradient update); the Python wrapper raises RuntimeError like
 * torch.cholesky does.
 */
#ifndef LVAE_B200_H
#define LVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVAE_MAX_COMPONENTS 16
#define LVAE_MAX_MASKS 3
#define LVAE_MAX_M 256
#define LVAE_MAX_T 40 /* rows per subject in lvae_kld_*; the Python mirror composes longer subjects from the ops below */
#define LVAE_SPEC_STRIDE 10

/* factor types inside a component */
#define LVAE_CAT 0 /* 1[x1 == x2]      kernel_spec.py:31-32, GP_model.py:43-53 */
#define LVAE_BIN 1 /* 1[x1 + x2 == 2]  kernel_spec.py:22-23, GP_model.py:31-41 */

#define LVAE_E_BADARG 1
#define LVAE_E_TOO_LARGE 2
#define LVAE_E_SPEC 3

/*
 * Additive-kernel structure (kernel_gen.py:199-310 / GP_model.py:146-236), flattened.
 * HOST int32 table, one row of LVAE_SPEC_STRIDE ints per component, K0 components first, then K1:
 *   [0] rbf_dim   covariate column of the squared-exponential factor, or -1 if the component has none
 *   [1] ls_index  row of `lengthscale` used by that factor (ignored if rbf_dim < 0)
 *   [2] n_masks   number of categorical/binary factors (0..3)
 *   [3+2i] type (LVAE_CAT | LVAE_BIN), [4+2i] covariate column, i < n_masks
 * component value:  outputscale[c][l] * prod(masks) * exp(-(x1[rbf_dim]-x2[rbf_dim])^2 / (2 lengthscale[ls][l]^2))
 */
typedef struct {
    int32_t n_comp0;     /* components without the id covariate (K0) */
    int32_t n_comp1;     /* components with the id covariate (K1)    */
    int32_t n_ls;        /* rows of `lengthscale`                    */
    const int32_t* spec; /* HOST [(n_comp0+n_comp1) * LVAE_SPEC_STRIDE] */
} lvae_kernel_spec_t;

/* Dense additive kernel  out[l, i, j] = sum_c k_c(x1[i], x2[j]) (+ diag_add[l] on i == j)   — replaces
 * `covar_module(x1, x2).evaluate()` (elbo_functions.py:171-172; gpytorch Kernel.__call__) and GP_model.AdditiveKernel.
 * x1: [n1,Q] shared by all latents (x1_latent_stride = 0) or [n_batch,n1,Q] (stride n1*Q); same for x2.
 * n_batch matrices are produced (n_batch a multiple of L); matrix b uses the hyper-parameters of latent b % L, which
 * covers the reference's [P_b,L,T,Q] stacking (elbo_functions.py:168-174) with n_batch = P_b*L.
 * comp_begin/comp_end select K0 (0..n_comp0) or K1 (n_comp0..n_comp0+n_comp1). */
int lvae_kernel_dense_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                          int32_t n_batch, int32_t Q, const double* x1, int64_t x1_latent_stride, int32_t n1, const double* x2,
                          int64_t x2_latent_stride, int32_t n2, const double* lengthscale, const double* outputscale,
                          const double* diag_add, double* out, void* stream);

/* Per-subject blocks  out[l, off2[p] + t*T_p + t'] = K(x_p[t], x_p[t'])  — replaces the [L,P_b,T,T] stacks of
 * elbo_functions.py:173-174 and the per-subject evaluations of 269-271.  offsets: device int32 [P_b+1] row offsets;
 * off2[p] = sum_{q<p} T_q^2 is computed on the device; block_stride = sum_p T_p^2 (elements per latent). */
int lvae_kernel_blocks_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L, int32_t Q,
                           const double* x, const int32_t* offsets, int32_t P_b, int64_t block_stride,
                           const double* lengthscale, const double* outputscale, const double* diag_add, double* out,
                           void* stream);

/* Hyper-parameter adjoints of the two functions above — what autograd through gpytorch's ScaleKernel / RBFKernel
 * (kernel_spec.py:58-69, GP_model.py:55-110) delivers in the reference when deviance_upper_bound / elbo
 * (elbo_functions.py:36-142) are trained on (training.py:326-343, 533-548, 654, 730).  grad_out has the layout of `out`.
 *   d_outputscale[c][l] = sum grad_out * k_c / outputscale[c][l]            (rows outside comp_begin..comp_end: 0)
 *   d_lengthscale[r][l] = sum grad_out * k_c * d^2 / lengthscale[r][l]^3     over the components that use row r
 *   d_diag_add[l]       = sum_i grad_out[., i, i]                            (skipped when NULL)
 * All three are overwritten, [rows, L] like the inputs; sums run in a fixed order (bitwise reproducible). */
int lvae_kernel_dense_bwd_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L,
                              int32_t n_batch, int32_t Q, const double* x1, int64_t x1_latent_stride, int32_t n1,
                              const double* x2, int64_t x2_latent_stride, int32_t n2, const double* lengthscale,
                              const double* outputscale, const double* grad_out, double* d_lengthscale,
                              double* d_outputscale, double* d_diag_add, void* stream);
int lvae_kernel_blocks_bwd_f64(const lvae_kernel_spec_t* ks, int32_t comp_begin, int32_t comp_end, int32_t L, int32_t Q,
                               const double* x, const int32_t* offsets, int32_t P_b, int64_t block_stride,
                               const double* lengthscale, const double* outputscale, const double* grad_out,
                               double* d_lengthscale, double* d_outputscale, double* d_diag_add, void* stream);

/* Batched GEMM on the FP64 tensor pipe: C[b] (m x n) = alpha * op(A[b]) op(B[b]) + beta * C[b], row-major — replaces the
 * torch.matmul / einsum contractions of elbo_functions.py:183-184,189,194,208-214 when M > 64.
 * op(A) is m x k: trans_a == 0 reads A[i*lda + kk], else A[kk*lda + i]; likewise op(B) (k x n).
 * flags: LVAE_GEMM_LOWER computes only entries j <= i; LVAE_GEMM_MIRROR also stores C[j][i] (needs beta == 0). */
#define LVAE_GEMM_LOWER 1
#define LVAE_GEMM_MIRROR 2
int lvae_gemm_batched_f64(int32_t trans_a, int32_t trans_b, int32_t m, int32_t n, int32_t k, double alpha, const double* A,
                          int32_t lda, int64_t stride_a, const double* B, int32_t ldb, int64_t stride_b, double beta,
                          double* C, int32_t ldc, int64_t stride_c, int32_t batch, int32_t flags, void* stream);

/* Batched Cholesky, in place, lower, n <= 256: replaces torch.cholesky (elbo_functions.py:177,179,185).
 * n <= 32 and at least 64 matrices: one warp per matrix in shared memory; n <= 64: one CTA per matrix; n > 64: blocked by 64 (diagonal blocks in shared memory, panels and trailing updates as
 * batched DMMA GEMMs; scratch from the stream-ordered allocator).
 * info: device int32[1], set to 1 + index of the first non-PD matrix (never cleared). */
int lvae_potrf_batched_f64(double* A, int32_t n, int64_t batch_stride, int32_t batch, int32_t* info, void* stream);
/* Explicit inverse from the Cholesky factor: replaces cholesky_solve(I, L) (elbo_functions.py:178,180,186). */
int lvae_potri_batched_f64(const double* Lc, double* Ainv, int32_t n, int64_t batch_stride, int32_t batch,
                           void* stream);

/* One minibatch of the Hensman-style KL upper bound and ALL its gradients (elbo_functions.py:144-216 / 219-307 plus the
 * reverse pass autograd would run, SURVEY 8a adjoints). */
typedef struct {
    /* sizes */
    int32_t L, M, Q, P_b, N_b, T_max; /* P_b, N_b: subjects / rows held by THIS GPU; T_max = max rows per subject */
    int64_t sum_T2;                   /* sum_p T_p^2 over this GPU's subjects */
    int32_t natural_gradient; /* 1: emit grad_m/grad_H (elbo_functions.py:208-214); 0: emit d_m/d_H (autograd) */
    int32_t path;             /* 0 = auto (fused DMMA kernel for M <= 64, GEMM-based path for 64 < M <= 256), 1 = generic kernels,
                                 2 = force the fused kernel (M <= 64 only) */
    double scale;             /* P_tot / P_b                                           (elbo_functions.py:204) */
    double const_term;        /* L*P_tot*T/2 (204) or L*N/2 (299), subtracted once     */
    double eps;               /* jitter on Kzz only (176)                               */
    lvae_kernel_spec_t ks;
    /* inputs */
    const double* x;           /* [N_b,Q]  rows grouped by subject, subject p = rows offsets[p]..offsets[p+1]-1 */
    const int32_t* offsets;    /* device int32 [P_b+1] */
    const double* mu;          /* [N_b,L] */
    const double* log_v;       /* [N_b,L] */
    const double* z;           /* [L,M,Q] */
    const double* m;           /* [L,M]   */
    const double* H;           /* [L,M,M] SPD */
    const double* lengthscale; /* [n_ls,L] */
    const double* outputscale; /* [n_comp0+n_comp1,L] */
    const double* noise;       /* [L] */
    /* outputs */
    double* kld_per_latent; /* [L]; kld_total = sum                                   */
    double* grad_m;         /* [L,M]   natural-gradient "gradients" (NOT scaled by P_tot/P_b, 208-214) or d kld/d m */
    double* grad_H;         /* [L,M,M]                                                                or d kld/d H */
    double* d_mu;           /* [N_b,L] d kld_total / d mu     */
    double* d_log_v;        /* [N_b,L] d kld_total / d log_v  */
    double* d_lengthscale;  /* [n_ls,L]   */
    double* d_outputscale;  /* [n_comp,L] */
    double* d_noise;        /* [L]        */
    /* scratch, sized by lvae_kld_workspace_doubles(); stats are the SVGP sufficient statistics that are summed across
     * GPUs between lvae_kld_subjects_f64 and lvae_kld_tail_f64 */
    double* stats;     /* [L, lvae_kld_stats_stride()] */
    double* workspace; /* lvae_kld_workspace_doubles() doubles */
    int32_t* info;     /* device int32[4] */
} lvae_kld_problem_t;

int64_t lvae_kld_stats_stride(int32_t M, int32_t n_ls, int32_t n_comp);
int64_t lvae_kld_workspace_doubles(const lvae_kld_problem_t* p);

/* Per-latent M x M work that does not depend on the minibatch rows: Kzz, chol, Kzz^-1, chol H, H^-1, a = Kzz^-1 m,
 * G = Kzz^-1 H Kzz^-1 (elbo_functions.py:172,176-178,185-186,194). */
int lvae_kld_head_f64(const lvae_kld_problem_t* p, void* stream);
/* Per-subject work, sharded by subject across GPUs: kernel blocks, chol/inverse of B_p, S, A..F partial sums, ng1,
 * d_mu, d_log_v and the subject part of the hyper-parameter gradients (173-174,179-184,189-196, 264-288). */
int lvae_kld_subjects_f64(const lvae_kld_problem_t* p, void* stream);
/* Per-latent tail on the (all-reduced) statistics: D, E, KL[q(u)||p(u)], kld, grad_m, grad_H, Kzz adjoint (193-214). */
int lvae_kld_tail_f64(const lvae_kld_problem_t* p, void* stream);
/* head + subjects + tail on one GPU */
int lvae_kld_minibatch_f64(const lvae_kld_problem_t* p, void* stream);

/* Natural-gradient update of (m, H), in place (training.py:129-135).  Hinv: optional [L,M,M] H^-1 already computed by
 * lvae_kld_head_f64 for the same H (at workspace + lvae_kld_hinv_offset()); NULL recomputes it (training.py:130-131).
 * workspace: lvae_ng_workspace_doubles(L, M) doubles (scratch of the blocked factorisation when M > 64); info[3]. */
int lvae_ng_step_f64(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hinv, double lr,
                     int32_t L, int32_t M, double* workspace, int32_t* info, void* stream);
int64_t lvae_kld_hinv_offset(const lvae_kld_problem_t* p);
int64_t lvae_ng_workspace_doubles(int32_t L, int32_t M);
/* Where the head leaves what the subject pass reads, as offsets (in doubles) into `workspace`: out[0] = W = c (G - Kzz^-1),
 * out[1] = its per-latent stride (M*M, or MP*MP zero padded for 64 < M <= 256), out[2] = a = Kzz^-1 m, out[3] = its per-latent
 * stride (M).  Used when head / tail / natural-gradient step are sharded by LATENT across GPUs while the subject pass is
 * sharded by subject (reduce-scatter of the statistics by latent + all-gather of W, a): a rank runs the head on its
 * latents and the gathered W, a are written at these offsets of the all-latent problem's workspace. */
int lvae_kld_head_offsets(const lvae_kld_problem_t* p, int64_t* out4);

/* One-shot all-reduce of the statistics row over NVLink peer memory (the exchange step of the subject-sharded path):
 * out[i] = sum_r peer_r[i], summed in rank order so every GPU obtains bit-identical results.  peer_ptrs: HOST array of
 * `world` device pointers (16-byte aligned) to the peers' buffers mapped into this process (e.g. torch symmetric memory);
 * the caller issues the inter-GPU barrier that makes those buffers complete before this call (distributed.py). */
int lvae_peer_sum_f64(const uint64_t* peer_ptrs, int32_t world, int64_t n, double* out, void* stream);

/* Optional per-phase device timing with CUDA events on the launch stream (bench.py's roofline leg).
 * phase: 0 head, 1 prep, 2 subjects, 3 reduce, 4 tail, 5 ng_step.  lvae_profile_last_ms synchronises on the event. */
int lvae_profile_enable(int on);
float lvae_profile_last_ms(int phase);

/* Test hook: out[i] = the library's exp(x[i]) for x <= 0 (the squared-exponential factor's exp; see lvae_common.cuh). */
int lvae_debug_exp_neg_f64(const double* x, double* out, int32_t n, void* stream);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t lvae_launch_count(void);
const char* lvae_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LVAE_B200_H */
