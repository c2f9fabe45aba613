// Internal: workspace layout shared by the KLD kernels (offsets in doubles into lvae_kld_problem_t::workspace).
#pragma once
#include "lvae_host.h"

#define LVAE_F2_ROWS 24
#define LVAE_F2_GT 8
#define LVAE_F3_ACC2 (5 * 256 * 2)      // doubles per CTA of the fused pass: 5 S tiles x 256 threads x 2 accumulators

struct KldLayout {
    int64_t Ki, Hi, G, W, T1, T2, T3;   // [L, M*M] each
    int64_t a;                          // [L, M]
    int64_t logdet;                     // [L, 2]  (log det Kzz, log det H)
    int64_t Bi;                         // [L, sum_T2]  explicit inverses of the per-subject blocks
    int64_t Bi_stride;                  // = sum_T2
    int64_t off2;                       // int64 [P_b+1] prefix of T_p^2
    int64_t part;                       // [nchunk, L, stride] per-CTA partial statistics of the subject pass
    int64_t ppart;                      // [nchunk, L, NSCAL+nh] per-CTA partials of the prep pass
    int64_t acc2;                       // [nchunk, L, LVAE_F3_ACC2] second-level accumulators of S (fused pass, two-level sum)
    int64_t total;
    int64_t stride;                     // statistics row length
    // second-generation fused pass (lvae_subjects_fused2.cu)
    int64_t Lrows;                      // [L, N_b, TP]  rows of the per-subject L^-1 (lower triangular, zero padded)
    int64_t Ltrows;                     // [L, N_b, TP]  rows of the per-subject L^-T (upper triangular; third-generation pass only)
    int64_t bmu;                        // [L, N_b]      B_p^-1 mu_p
    int64_t gtab;                       // int32 [nchunk, gstride, LVAE_F2_GT] row-group plan
    int64_t gcount;                     // int32 [nchunk]
    int TP, gstride;
    int v2;                             // row-group plan and L^-1 rows are part of the workspace (fused pass or GEMM-based path)
    int v3;                             // fused subject pass (lvae_subjects_fused3.cu); implies v2
    int nh, nchunk;                     // nchunk: CTAs per latent of the subject pass
    int nprep;                          // partial rows per latent of the prep pass
    int prep3;                          // 1: third-generation prep kernel (lvae_prep3.cu)
    // 64 < M <= 256: GEMM-based path (lvae_kld_big.cu, lvae_subjects_big.cu); matrices padded to MP = 128 or 256
    int big, MP, nsplit;                // nsplit: k-splits of S = U^T U (rows of `part`); nchunk: CTAs per latent of k_uv / k_adj
    int64_t bF, bX, bInv, bT, bA0;      // [2L, MP*MP]: factors, triangular inverses, explicit inverses (Kzz | H), scratch, originals
    int64_t bDinv;                      // [2L, MP/64, 64*64] inverses of the diagonal blocks of the factors
    int64_t bHp, bWp, bS, bT1, bT2, bT3;   // [L, MP*MP]: zero-padded H, W, S and products
    int64_t bU, bV;                     // [L, N_b, MP]: U = L_p^-1 Kxz (later Y = V W), V = B_p^-1 Kxz
    int64_t bu;                         // [L, N_b]: u = B_p^-1 r
    int64_t bpart;                      // [nchunk, L, bpstride]: ng1 | da (MP each) | scalars | hyper-gradients
    int64_t bpstride;
};

KldLayout lvae_layout(const lvae_kld_problem_t* p);
int lvae_chunks(int P_b, int L, int T_max);
int lvae_prep_rows(int P_b, int L, int T_max, int Q);
bool lvae_prep_warp_supported(const lvae_kld_problem_t* p);
int lvae_prep_warp_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);

// M <= 64 shared-memory kernels (lvae_kld64.cu)
int lvae_head64_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
int lvae_tail64_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
int lvae_ng64_launch(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hi, double lr, int L,
                     int M, int32_t* info, cudaStream_t st);

// fused subject pass (third generation: transposed register layout, 2 CTAs of 8 warps per SM; M <= 62, T <= 40) and the
// row-group planner it shares with the GEMM-based path
bool lvae_fused3_supported(const lvae_kld_problem_t* p);
int lvae_chunks3(int P_b, int L, int T_max);
int lvae_plan_groups3_launch(const lvae_kld_problem_t* p, const KldLayout& w, cudaStream_t st);
int lvae_subjects_fused3_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);

// 64 < M <= 256 (lvae_kld_big.cu / lvae_subjects_big.cu)
bool lvae_big_supported(const lvae_kld_problem_t* p);
int lvae_head_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
int lvae_tail_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
int lvae_subjects_big_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
int lvae_reduce_big_launch(const lvae_kld_problem_t* p, const KldLayout& w, cudaStream_t st);
int64_t lvae_ng_big_workspace(int L, int M);
int lvae_ng_big_launch(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hi, double lr, int L,
                       int M, double* ws, int32_t* info, cudaStream_t st);

// third-generation prep kernel (lvae_prep3.cu)
bool lvae_prep3_supported(const lvae_kld_problem_t* p, const KldLayout& w);
int lvae_prep3_rows(const lvae_kld_problem_t* p);
int lvae_prep3_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st);
