// GP-prior ELBO path (Hensman minibatch KL upper bound + all gradients), generic kernels for any M <= 256, T <= 40:
//   head     per latent : Kzz, chol, Kzz^-1, chol(H), H^-1, a = Kzz^-1 m, G = Kzz^-1 H Kzz^-1, W = c(G - Kzz^-1)
//   prep     per (subject, latent): K0_p, B_p = K1_p + s2 I, chol, B_p^-1, scalars C, D1, Bt, F, d_log_v, local adjoints
//   subjects per (subject, latent): Kxz_p, V = B^-1 Kxz, S, r, u, A, ng1, da, d_mu, Y = V W, hyper-parameter adjoints
//   reduce   fixed-order sum of per-chunk partials into the statistics row (deterministic, no atomics)
//   tail     per latent : D, E, KL[q(u)||p(u)], kld, grad_m, grad_H (or d_m, d_H), Kzz adjoint -> hyper-gradients
// Formulas and line references: SURVEY.md 8(a); elbo_functions.py:144-216, 219-307.
// The fast paths replace prep / subjects: lvae_prep3.cu + lvae_subjects_fused3.cu (M <= 62, T <= 40), lvae_subjects_big.cu (62 < M <= 256).
#include <stdlib.h>

#include "lvae_host.h"
#include <mutex>
#include <vector>

#include "lvae_kld.h"
#include "lvae_linalg.cuh"

// ---------------------------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------------------------
int lvae_chunks(int P_b, int L, int T_max) {
    // CTAs per latent of the subject pass (one 512-thread CTA per SM): pick the number of waves k that minimises
    // k * (row groups per CTA + 1 group-equivalent of per-CTA set-up), with L * nchunk <= 148 * k.
    const int spg = T_max > 0 ? (48 / T_max > 0 ? 48 / T_max : 1) : 1;
    int best = 1;
    long best_cost = -1;
    for (int k = 1; k <= 8; ++k) {
        int n = 148 * k / L;
        if (n < 1) continue;
        if (n > P_b) n = P_b > 0 ? P_b : 1;
        const int per = (P_b + n - 1) / n;
        const long cost = (long)k * ((per + spg - 1) / spg + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = n; }
    }
    return best;
}

KldLayout lvae_layout(const lvae_kld_problem_t* p) {
    KldLayout w;
    const int64_t L = p->L, M = p->M, MM = M * M;
    w.nh = p->ks.n_ls + p->ks.n_comp0 + p->ks.n_comp1 + 1;
    w.stride = stats_stride((int)M, w.nh);
    w.nchunk = lvae_chunks(p->P_b, p->L, p->T_max);
    int64_t o = 0;
    w.Ki = o; o += L * MM;
    w.Hi = o; o += L * MM;
    w.G = o; o += L * MM;
    w.W = o; o += L * MM;
    w.T1 = o; o += L * MM;
    w.T2 = o; o += L * MM;
    w.T3 = o; o += L * MM;
    w.a = o; o += L * M;
    w.logdet = o; o += L * 2;
    w.big = (p->path != 1) && lvae_big_supported(p) ? 1 : 0;
    w.MP = w.big ? (p->M <= 128 ? 128 : 256) : 0;
    // path: 0 auto, 1 generic kernels (cross-check of the fast paths), 2 = 0
    w.prep3 = (p->path != 1 && p->ks.spec && lvae_prep3_supported(p, w)) ? 1 : 0;
    // the fused subject pass (M <= 62, T <= 40) reads the L^-1 / L^-T rows that only k_prep3 exports
    w.v3 = (!w.big && w.prep3 && p->path != 1 && lvae_fused3_supported(p)) ? 1 : 0;
    w.v2 = (w.v3 || w.big) ? 1 : 0;               // row-group plan + L^-1 rows in the workspace
    if (w.v3) w.nchunk = lvae_chunks3(p->P_b, p->L, p->T_max);
    w.nprep = w.prep3 ? lvae_prep3_rows(p) : lvae_prep_rows(p->P_b, p->L, p->T_max, p->Q);
    w.nsplit = 1;
    if (w.big) {
        // CTAs per latent of the U/V and adjoint kernels: about two waves at 2 CTAs per SM
        int n = (4 * 148 + p->L - 1) / p->L;
        if (n > p->P_b) n = p->P_b > 0 ? p->P_b : 1;
        w.nchunk = n < 1 ? 1 : n;
        // k-splits of S = U^T U: enough (tile, latent, split) CTAs for two waves, at least 512 rows per split
        const int tiles = w.MP == 128 ? 2 : 6;
        int ns = (2 * 148 + tiles * p->L - 1) / (tiles * p->L);
        const int cap = p->N_b / 512 > 0 ? p->N_b / 512 : 1;
        if (ns > cap) ns = cap;
        if (ns > 16) ns = 16;
        // ... and for ACCURACY at most ~1024 rows per split: a GEMM accumulator is one sequential chain over its k range, and
        // the rounding error of such a chain grows linearly with its length.  Kzz^-1 S Kzz^-1 magnifies the last bits of S by
        // ~1e10 (cond(Kzz) ~ 1e8): with one 20 000-row chain grad_m was 1.7e-4 from the exact value at cfg3, 50 x the
        // reference's own LAPACK error (measured against an extended-precision evaluation).  The partial S of
        // the splits are then summed in fixed order: a two-level sum.
        int ns_acc = (p->N_b + 1023) / 1024;
        if (ns_acc > 64) ns_acc = 64;
        if (ns < ns_acc) ns = ns_acc;
        w.nsplit = ns < 1 ? 1 : ns;
    }
    w.TP = (p->T_max + 3) & ~3;
    if (w.TP < 4) w.TP = 4;
    w.gstride = (p->P_b + w.nchunk - 1) / w.nchunk + 1;
    w.Bi = o; o += w.v2 ? 0 : L * p->sum_T2;
    w.off2 = o; o += (int64_t)p->P_b + 1;
    o += o & 1;
    w.Lrows = o; o += w.v2 ? L * (int64_t)p->N_b * w.TP : 0;
    w.Ltrows = o; o += w.v3 ? L * (int64_t)p->N_b * w.TP : 0;
    w.bmu = o; o += w.v2 ? L * (int64_t)p->N_b : 0;
    o += o & 1;
    w.gtab = o; o += w.v2 ? ((int64_t)w.nchunk * w.gstride * LVAE_F2_GT + 1) / 2 : 0;
    o += o & 1;
    w.gcount = o; o += w.v2 ? (w.nchunk + 1) / 2 : 0;
    w.part = o; o += (int64_t)(w.big ? w.nsplit : w.nchunk) * L * w.stride;   // subject partials (S, ng1, da, A, hyp)
    o += o & 1;                                                               // 16-byte aligned (double2 accesses)
    w.acc2 = o; o += w.v3 ? (int64_t)w.nchunk * L * LVAE_F3_ACC2 : 0;         // second-level S accumulators of the fused pass
    w.ppart = o; o += (int64_t)w.nprep * L * (LVAE_NSCAL + w.nh);   // prep partials (scalars, hyp)
    w.bpstride = 2 * (int64_t)w.MP + LVAE_NSCAL + w.nh;
    if (w.big) {
        const int64_t MP2 = (int64_t)w.MP * w.MP;
        o += o & 1;
        w.bF = o; o += 2 * L * MP2;
        w.bX = o; o += 2 * L * MP2;
        w.bInv = o; o += 2 * L * MP2;
        w.bT = o; o += 2 * L * MP2;
        w.bA0 = o; o += 2 * L * MP2;
        w.bDinv = o; o += 2 * L * (w.MP / 64) * 4096;
        w.bHp = o; o += L * MP2;
        w.bWp = o; o += L * MP2;
        w.bS = o; o += L * MP2;
        w.bT1 = o; o += L * MP2;
        w.bT2 = o; o += L * MP2;
        w.bT3 = o; o += L * MP2;
        w.bU = o; o += L * (int64_t)p->N_b * w.MP;
        w.bV = o; o += L * (int64_t)p->N_b * w.MP;
        w.bu = o; o += L * (int64_t)p->N_b;
        o += o & 1;
        w.bpart = o; o += (int64_t)w.nchunk * L * w.bpstride;
    }
    w.total = o;
    return w;
}

extern "C" int64_t lvae_kld_stats_stride(int32_t M, int32_t n_ls, int32_t n_comp) {
    return stats_stride(M, n_ls + n_comp + 1);
}
extern "C" int64_t lvae_kld_workspace_doubles(const lvae_kld_problem_t* p) { return lvae_layout(p).total; }

static int check_problem(const lvae_kld_problem_t* p, DevSpec* sp) {
    if (!p) return LVAE_E_BADARG;
    if (p->L <= 0 || p->L > 65535 || p->M <= 0 || p->Q <= 0 || p->P_b < 0 || p->N_b < 0) return LVAE_E_BADARG;
    if (p->M > LVAE_MAX_M || p->T_max > LVAE_MAX_T) return LVAE_E_TOO_LARGE;
    return lvae_make_devspec(&p->ks, p->Q, sp);
}

// per-latent hyper-parameters in shared memory
struct LatentHyp {
    double hil2[LVAE_MAXC];   // 1 / (2 l^2)
    double il3[LVAE_MAXC];    // 1 / l^3
    double os[LVAE_MAXC];
    double noise;
    double etab[LVAE_EXP_TBL];
};
__device__ inline void load_hyp(LatentHyp* h, const DevSpec& sp, const double* ls, const double* os, const double* noise,
                                int L, int l) {
    const int t = threadIdx.x;
    if (t < sp.n_ls) { const double v = ls[(size_t)t * L + l]; h->hil2[t] = 0.5 / (v * v); h->il3[t] = 1.0 / (v * v * v); }
    if (t < sp.n0 + sp.n1) h->os[t] = os[(size_t)t * L + l];
    if (t == 0) h->noise = noise ? noise[l] : 0.0;
    load_exp_table(h->etab);
}

// ---------------------------------------------------------------------------------------------------------------
// head
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_head(DevSpec sp, KldLayout w, int L, int M, int Q, const double* __restrict__ z,
                                              const double* __restrict__ m, const double* __restrict__ H,
                                              const double* __restrict__ ls, const double* __restrict__ os, double eps,
                                              double c, double* __restrict__ ws, int32_t* info) {
    __shared__ LatentHyp hyp;
    __shared__ double red[32];
    __shared__ int flag;
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, MM = M * M;
    load_hyp(&hyp, sp, ls, os, nullptr, L, l);
    __syncthreads();
    double* Ki = ws + w.Ki + (size_t)l * MM;
    double* Hi = ws + w.Hi + (size_t)l * MM;
    double* G = ws + w.G + (size_t)l * MM;
    double* W = ws + w.W + (size_t)l * MM;
    double* T1 = ws + w.T1 + (size_t)l * MM;
    double* T2 = ws + w.T2 + (size_t)l * MM;
    double* a = ws + w.a + (size_t)l * M;
    const double* zl = z + (size_t)l * M * Q;
    // Kzz + eps I   (elbo_functions.py:172,176)
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        double acc = 0.0, d2;
        for (int cc = 0; cc < sp.n0; ++cc) acc += hyp.os[cc] * comp_value(sp, cc, zl + i * Q, zl + j * Q, hyp.hil2, d2, hyp.etab);
        T1[e] = acc + (i == j ? eps : 0.0);
    }
    __syncthreads();
    int rc = cta_cholesky(T1, M, M, &flag);                       // 177
    if (rc && tid == 0) atomicCAS(info + 0, 0, l + 1);
    const double ldK = cta_logdet_from_chol(T1, M, M, red);
    cta_tri_inverse(T1, T2, M, M);
    cta_gram_lower(T2, Ki, M, M);                                  // 178 (explicit inverse)
    for (int e = tid; e < MM; e += nt) T1[e] = H[(size_t)l * MM + e];
    __syncthreads();
    rc = cta_cholesky(T1, M, M, &flag);                            // 185
    if (rc && tid == 0) atomicCAS(info + 1, 0, l + 1);
    const double ldH = cta_logdet_from_chol(T1, M, M, red);
    cta_tri_inverse(T1, T2, M, M);
    cta_gram_lower(T2, Hi, M, M);                                  // 186
    if (tid == 0) { ws[w.logdet + 2 * l] = ldK; ws[w.logdet + 2 * l + 1] = ldH; }
    cta_gemv(Ki, m + (size_t)l * M, a, M, M);                      // a = Kzz^-1 m
    cta_gemm_nn<false, false>(Ki, H + (size_t)l * MM, T1, M, M, 1.0, 0.0);
    cta_gemm_nn<false, false>(T1, Ki, G, M, M, 1.0, 0.0);          // 194
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        W[e] = c * (0.5 * (G[i * M + j] + G[j * M + i]) - Ki[e]);  // dS adjoint, symmetrised
    }
}

// ---------------------------------------------------------------------------------------------------------------
// prep: everything that only needs the T_p x T_p blocks
// ---------------------------------------------------------------------------------------------------------------


__global__ void __launch_bounds__(128) k_prep(DevSpec sp, KldLayout w, int L, int Q, int P_b, int Tmax,
                                              const double* __restrict__ x, const int32_t* __restrict__ offsets,
                                              const double* __restrict__ log_v, const double* __restrict__ ls,
                                              const double* __restrict__ os, const double* __restrict__ noise, double c,
                                              double* __restrict__ d_log_v, double* __restrict__ ws, int32_t* info) {
    extern __shared__ double sm[];
    __shared__ LatentHyp hyp;
    __shared__ double red[32];
    __shared__ int flag;
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const int nc = sp.n0 + sp.n1, nh = hyp_count(sp), TT = Tmax * Tmax;
    double* K0 = sm;            // [T,T]
    double* Bm = K0 + TT;       // B -> chol factor
    double* Li = Bm + TT;       // L^-1, later X2
    double* Bi = Li + TT;
    double* X1 = Bi + TT;
    double* fc = X1 + TT;       // [nc][T,T]
    double* xs = fc + (size_t)nc * TT;   // [T,Q]
    double* ev = xs + Tmax * Q;          // exp(log_v) [T]
    load_hyp(&hyp, sp, ls, os, noise, L, l);
    const int64_t* off2 = reinterpret_cast<const int64_t*>(ws + w.off2);
    double acc[LVAE_NSCAL + 2 * LVAE_MAXC + 1];
    for (int k = 0; k < LVAE_NSCAL + nh; ++k) acc[k] = 0.0;
    const int per = (P_b + gridDim.x - 1) / gridDim.x;
    const int p0 = chunk * per, p1 = min(P_b, p0 + per);
    for (int p = p0; p < p1; ++p) {
        const int r0 = offsets[p], T = offsets[p + 1] - r0;
        __syncthreads();
        for (int e = tid; e < T * Q; e += nt) xs[e] = x[(size_t)r0 * Q + e];
        for (int t = tid; t < T; t += nt) {
            const double lv = log_v[(size_t)(r0 + t) * L + l];
            ev[t] = exp(lv);
            acc[SC_F] += lv;
        }
        __syncthreads();
        for (int e = tid; e < T * T; e += nt) {
            const int i = e / T, j = e % T;
            double k0 = 0.0, k1 = 0.0, d2;
            for (int cc = 0; cc < nc; ++cc) {
                const double f = comp_value(sp, cc, xs + i * Q, xs + j * Q, hyp.hil2, d2, hyp.etab);
                fc[(size_t)cc * TT + e] = f;
                if (cc < sp.n0) k0 += hyp.os[cc] * f; else k1 += hyp.os[cc] * f;
            }
            K0[e] = k0;                                           // elbo_functions.py:173
            Bm[e] = k1 + (i == j ? hyp.noise : 0.0);              // 174
        }
        __syncthreads();
        const int rc = cta_cholesky(Bm, T, T, &flag);             // 179
        if (rc && tid == 0) atomicCAS(info + 2, 0, l * P_b + p + 1);
        cta_tri_inverse(Bm, Li, T, T);
        cta_gram_lower(Li, Bi, T, T);                             // 180
        double* gBi = ws + w.Bi + (size_t)l * (w.Bi_stride) + off2[p];
        for (int e = tid; e < T * T; e += nt) gBi[e] = Bi[e];
        for (int t = tid; t < T; t += nt) {
            acc[SC_C] += 2.0 * log(Bm[t * T + t]);                // 192
            const double bt = Bi[t * T + t] * ev[t];
            acc[SC_BT] += bt;                                     // 191
            d_log_v[(size_t)(r0 + t) * L + l] = c * (bt - 1.0);
        }
        // X1 = (diag(v) + K0) Bi ; X2 = Bi X1 (into Li) ; local adjoint of B: GB = c (Bi - X2)
        for (int e = tid; e < T * T; e += nt) {
            const int i = e / T, j = e % T;
            acc[SC_D1] += Bi[e] * K0[e];                          // 193 first term
            double s = ev[i] * Bi[e];
            for (int k = 0; k < T; ++k) s += K0[i * T + k] * Bi[k * T + j];
            X1[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < T * T; e += nt) {
            const int i = e / T, j = e % T;
            double s = 0.0;
            for (int k = 0; k < T; ++k) s += Bi[i * T + k] * X1[k * T + j];
            const double gB = c * (Bi[e] - s);
            const double gK0 = c * Bi[e];
            const double dd0 = xs[i * Q] - xs[j * Q];
            (void)dd0;
            for (int cc = 0; cc < nc; ++cc) {
                const double f = fc[(size_t)cc * TT + e];
                const double gbar = cc < sp.n0 ? gK0 : gB;
                acc[LVAE_NSCAL + sp.n_ls + cc] += gbar * f;
                const int rd = sp.rbf_dim[cc];
                if (rd >= 0) {
                    const double d = xs[i * Q + rd] - xs[j * Q + rd];
                    acc[LVAE_NSCAL + sp.ls_idx[cc]] += gbar * hyp.os[cc] * f * d * d * hyp.il3[sp.ls_idx[cc]];
                }
            }
            if (i == j) acc[LVAE_NSCAL + nh - 1] += gB;
        }
    }
    // deterministic block reduction of every accumulator
    double* out = ws + w.ppart + ((size_t)chunk * L + l) * (LVAE_NSCAL + nh);
    for (int k = 0; k < LVAE_NSCAL + nh; ++k) {
        const double t = block_sum(acc[k], red);
        if (tid == 0) out[k] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// subjects (generic): one subject at a time per CTA, S partial kept in global (L2-resident) memory
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_subjects_generic(DevSpec sp, KldLayout w, int L, int M, int Q, int P_b, int Tmax,
                                                          const double* __restrict__ x,
                                                          const int32_t* __restrict__ offsets,
                                                          const double* __restrict__ mu, const double* __restrict__ z,
                                                          const double* __restrict__ ls, const double* __restrict__ os,
                                                          double c, double* __restrict__ d_mu, double* __restrict__ ws) {
    extern __shared__ double sm[];
    __shared__ LatentHyp hyp;
    __shared__ double red[32];
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const int nh = hyp_count(sp), MM = M * M, TT = Tmax * Tmax;
    double* Kx = sm;                       // [T,M]  Kxz_p, later Y
    double* V = Kx + (size_t)Tmax * M;     // [T,M]
    double* Bi = V + (size_t)Tmax * M;     // [T,T]
    double* Qm = Bi + TT;                  // [T,T]
    double* xs = Qm + TT;                  // [T,Q]
    double* zs = xs + Tmax * Q;            // [M,Q]
    double* av = zs + (size_t)M * Q;       // [M]
    double* ng1 = av + M;                  // [M]
    double* da = ng1 + M;                  // [M]
    double* r = da + M;                    // [T]
    double* u = r + Tmax;                  // [T]
    double* mus = u + Tmax;                // [T]
    load_hyp(&hyp, sp, ls, os, nullptr, L, l);
    const int64_t* off2 = reinterpret_cast<const int64_t*>(ws + w.off2);
    const double* Wl = ws + w.W + (size_t)l * MM;
    double* part = ws + w.part + ((size_t)chunk * L + l) * w.stride;
    double* Sp = part + stats_off_S();
    for (int e = tid; e < M * Q; e += nt) zs[e] = z[(size_t)l * M * Q + e];
    for (int e = tid; e < M; e += nt) { av[e] = ws[w.a + (size_t)l * M + e]; ng1[e] = 0.0; da[e] = 0.0; }
    for (int e = tid; e < MM; e += nt) Sp[e] = 0.0;
    double acc[1 + 2 * LVAE_MAXC + 1];     // [0] = A, then hyper-gradient vector
    for (int k = 0; k < 1 + nh; ++k) acc[k] = 0.0;
    const int per = (P_b + gridDim.x - 1) / gridDim.x;
    const int p0 = chunk * per, p1 = min(P_b, p0 + per);
    for (int p = p0; p < p1; ++p) {
        const int r0 = offsets[p], T = offsets[p + 1] - r0;
        __syncthreads();
        const double* gBi = ws + w.Bi + (size_t)l * w.Bi_stride + off2[p];
        for (int e = tid; e < T * T; e += nt) Bi[e] = gBi[e];
        for (int e = tid; e < T * Q; e += nt) xs[e] = x[(size_t)r0 * Q + e];
        for (int t = tid; t < T; t += nt) mus[t] = mu[(size_t)(r0 + t) * L + l];
        __syncthreads();
        for (int e = tid; e < T * M; e += nt) {                    // Kxz_p (elbo_functions.py:171)
            const int t = e / M, j = e % M;
            double k0 = 0.0, d2;
            for (int cc = 0; cc < sp.n0; ++cc) k0 += hyp.os[cc] * comp_value(sp, cc, xs + t * Q, zs + j * Q, hyp.hil2, d2, hyp.etab);
            Kx[e] = k0;
        }
        __syncthreads();
        for (int t = tid; t < T; t += nt) {                        // r = Kxz a - mu (189, re-associated)
            double s = 0.0;
            for (int j = 0; j < M; ++j) s += Kx[t * M + j] * av[j];
            r[t] = s - mus[t];
        }
        for (int e = tid; e < T * M; e += nt) {                    // V = Bi Kxz (183)
            const int t = e / M, j = e % M;
            double s = 0.0;
            for (int k = 0; k < T; ++k) s += Bi[t * T + k] * Kx[k * M + j];
            V[e] = s;
        }
        __syncthreads();
        for (int t = tid; t < T; t += nt) {                        // u = Bi r ; A += r.u (190) ; d_mu
            double s = 0.0;
            for (int k = 0; k < T; ++k) s += Bi[t * T + k] * r[k];
            u[t] = s;
            acc[0] += r[t] * s;
            d_mu[(size_t)(r0 + t) * L + l] = -2.0 * c * s;
        }
        for (int e = tid; e < MM; e += nt) {                       // S += Kxz^T V (184)
            const int i = e / M, j = e % M;
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += Kx[t * M + i] * V[t * M + j];
            Sp[e] += s;
        }
        for (int j = tid; j < M; j += nt) {                        // ng1 += V^T mu (209-211) ; da += V^T r
            double s1 = 0.0, s2 = 0.0;
            for (int t = 0; t < T; ++t) { s1 += V[t * M + j] * mus[t]; s2 += V[t * M + j] * r[t]; }
            ng1[j] += s1;
            da[j] += s2;
        }
        __syncthreads();
        for (int e = tid; e < T * M; e += nt) {                    // Y = V W ; adjoint of Kxz ; K0 hyper-gradients
            const int t = e / M, j = e % M;
            double y = 0.0;
            for (int k = 0; k < M; ++k) y += V[t * M + k] * Wl[k * M + j];
            Kx[e] = y;
            const double gbar = 2.0 * c * u[t] * av[j] + 2.0 * y;
            for (int cc = 0; cc < sp.n0; ++cc) {
                double d2;
                const double f = comp_value(sp, cc, xs + t * Q, zs + j * Q, hyp.hil2, d2, hyp.etab);
                acc[1 + sp.n_ls + cc] += gbar * f;
                if (sp.rbf_dim[cc] >= 0) acc[1 + sp.ls_idx[cc]] += gbar * hyp.os[cc] * f * d2 * hyp.il3[sp.ls_idx[cc]];
            }
        }
        __syncthreads();
        for (int e = tid; e < T * T; e += nt) {                    // adjoint of B_p: -(c u u^T + Y V^T) ; K1 hyper-gradients
            const int i = e / T, j = e % T;
            double q = 0.0;
            for (int k = 0; k < M; ++k) q += Kx[i * M + k] * V[j * M + k];
            const double gB = -(c * u[i] * u[j] + q);
            for (int cc = sp.n0; cc < sp.n0 + sp.n1; ++cc) {
                double d2;
                const double f = comp_value(sp, cc, xs + i * Q, xs + j * Q, hyp.hil2, d2, hyp.etab);
                acc[1 + sp.n_ls + cc] += gB * f;
                if (sp.rbf_dim[cc] >= 0) acc[1 + sp.ls_idx[cc]] += gB * hyp.os[cc] * f * d2 * hyp.il3[sp.ls_idx[cc]];
            }
            if (i == j) acc[1 + nh - 1] += gB;
        }
    }
    __syncthreads();
    for (int j = tid; j < M; j += nt) { part[stats_off_ng1(M) + j] = ng1[j]; part[stats_off_da(M) + j] = da[j]; }
    {
        const double t = block_sum(acc[0], red);
        if (tid == 0) {
            for (int k = 0; k < LVAE_NSCAL; ++k) part[stats_off_scal(M) + k] = 0.0;
            part[stats_off_scal(M) + SC_A] = t;
        }
    }
    for (int k = 0; k < nh; ++k) {
        const double t = block_sum(acc[1 + k], red);
        if (tid == 0) part[stats_off_hyp(M) + k] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// reduce: stats[l][k] = sum over chunks, fixed order
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(KldLayout w, int L, int M, const double* __restrict__ ws,
                                                double* __restrict__ stats) {
    const int l = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= w.stride) return;
    double s = 0.0;
    for (int ch = 0; ch < w.nchunk; ++ch) s += ws[w.part + ((size_t)ch * L + l) * w.stride + k];
    const int64_t ks = k - stats_off_scal(M);
    if (ks >= 0) {   // scalars and hyper-gradients also receive the prep partials
        for (int ch = 0; ch < w.nprep; ++ch) s += ws[w.ppart + ((size_t)ch * L + l) * (LVAE_NSCAL + w.nh) + ks];
    }
    stats[(size_t)l * w.stride + k] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// tail
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tail(DevSpec sp, KldLayout w, int L, int M, int Q, int natural_gradient,
                                              const double* __restrict__ z, const double* __restrict__ m,
                                              const double* __restrict__ H, const double* __restrict__ ls,
                                              const double* __restrict__ os, double c, double const_per_latent,
                                              const double* __restrict__ stats, double* __restrict__ ws,
                                              double* __restrict__ kld, double* __restrict__ grad_m,
                                              double* __restrict__ grad_H, double* __restrict__ d_ls,
                                              double* __restrict__ d_os, double* __restrict__ d_noise) {
    __shared__ LatentHyp hyp;
    __shared__ double red[32];
    __shared__ double vec[3 * LVAE_MAX_M];   // ga (adjoint of a), tmp vectors
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, MM = M * M, nh = hyp_count(sp);
    load_hyp(&hyp, sp, ls, os, nullptr, L, l);
    __syncthreads();
    const double* st = stats + (size_t)l * w.stride;
    const double* S = st + stats_off_S();
    const double* ng1 = st + stats_off_ng1(M);
    const double* da = st + stats_off_da(M);
    const double* sc = st + stats_off_scal(M);
    const double* hy = st + stats_off_hyp(M);
    const double* Ki = ws + w.Ki + (size_t)l * MM;
    const double* Hi = ws + w.Hi + (size_t)l * MM;
    const double* G = ws + w.G + (size_t)l * MM;
    const double* a = ws + w.a + (size_t)l * M;
    const double* Hl = H + (size_t)l * MM;
    const double* ml = m + (size_t)l * M;
    double* T1 = ws + w.T1 + (size_t)l * MM;
    double* T2 = ws + w.T2 + (size_t)l * MM;
    double* T3 = ws + w.T3 + (size_t)l * MM;
    // scalar reductions: D2 = sum S*Ki (193), E = sum G*S (195 / 282), tr = sum Ki*H^T (199), qf = m.a (200)
    double d2 = 0.0, ee = 0.0, tr = 0.0, qf = 0.0;
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        d2 += S[e] * Ki[e];
        ee += G[j * M + i] * S[e];
        tr += Ki[e] * Hl[j * M + i];
    }
    for (int i = tid; i < M; i += nt) qf += ml[i] * a[i];
    d2 = block_sum(d2, red);
    ee = block_sum(ee, red);
    tr = block_sum(tr, red);
    qf = block_sum(qf, red);
    if (tid == 0) {
        const double ldK = ws[w.logdet + 2 * l], ldH = ws[w.logdet + 2 * l + 1];
        const double kl_qp = 0.5 * (tr + qf - M + ldK - ldH);                                   // 199-203
        kld[l] = c * (sc[SC_A] + sc[SC_BT] + sc[SC_C] + sc[SC_D1] - d2 + ee - sc[SC_F]) + kl_qp - const_per_latent;  // 204
    }
    // P1 = Ki S ; KSK = P1 Ki
    cta_gemm_nn<false, false>(Ki, S, T1, M, M, 1.0, 0.0);      // T1 = P1
    cta_gemm_nn<false, false>(T1, Ki, T2, M, M, 1.0, 0.0);     // T2 = Ki S Ki
    double* gm = grad_m + (size_t)l * M;
    double* gH = grad_H + (size_t)l * MM;
    double* ga = vec;          // adjoint of a = 2c da
    for (int i = tid; i < M; i += nt) ga[i] = 2.0 * c * da[i];
    __syncthreads();
    if (natural_gradient) {
        // Bm = Ki S Ki + Ki ; grad_m = -Ki ng1 + Bm m ; grad_H = (Bm - Hi)/2      (208-214, 301-305)
        for (int i = tid; i < M; i += nt) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s += -Ki[i * M + k] * ng1[k] + (T2[i * M + k] + Ki[i * M + k]) * ml[k];
            gm[i] = s;
        }
        for (int e = tid; e < MM; e += nt) gH[e] = 0.5 * (T2[e] + Ki[e] - Hi[e]);
    } else {
        // autograd of kld_total: d/dm = Ki ga + a ; d/dH = c Ki S Ki + Ki/2 - Hi/2
        for (int i = tid; i < M; i += nt) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s += Ki[i * M + k] * ga[k];
            gm[i] = s + a[i];
        }
        for (int e = tid; e < MM; e += nt) {
            const int i = e / M, j = e % M;
            gH[e] = c * 0.5 * (T2[i * M + j] + T2[j * M + i]) + 0.5 * Ki[e] - 0.5 * Hi[e];
        }
    }
    __syncthreads();
    // adjoint of Ki: -c S + c (H P1 + (H P1)^T) + (H^T + m m^T)/2 + ga m^T
    cta_gemm_nn<false, false>(Hl, T1, T3, M, M, 1.0, 0.0);     // T3 = H Ki S
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        T2[e] = -c * S[e] + c * (T3[i * M + j] + T3[j * M + i]) + 0.5 * (Hl[j * M + i] + ml[i] * ml[j]) + ga[i] * ml[j];
    }
    __syncthreads();
    // adjoint of Kzz: -Ki gKi Ki + Ki/2, symmetrised
    cta_gemm_nn<false, false>(Ki, T2, T1, M, M, 1.0, 0.0);
    cta_gemm_nn<false, false>(T1, Ki, T3, M, M, 1.0, 0.0);
    const double* zl = z + (size_t)l * M * Q;
    double acc[2 * LVAE_MAXC + 1];
    for (int k = 0; k < nh; ++k) acc[k] = 0.0;
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        const double gK = -0.5 * (T3[i * M + j] + T3[j * M + i]) + 0.5 * Ki[e];
        for (int cc = 0; cc < sp.n0; ++cc) {
            double dd;
            const double f = comp_value(sp, cc, zl + i * Q, zl + j * Q, hyp.hil2, dd, hyp.etab);
            acc[sp.n_ls + cc] += gK * f;
            if (sp.rbf_dim[cc] >= 0) acc[sp.ls_idx[cc]] += gK * hyp.os[cc] * f * dd * hyp.il3[sp.ls_idx[cc]];
        }
    }
    const int ncmp = sp.n0 + sp.n1;
    for (int k = 0; k < nh; ++k) {
        const double t = block_sum(acc[k], red) + hy[k];
        if (tid == 0) {
            if (k < sp.n_ls) d_ls[(size_t)k * L + l] = t;
            else if (k < sp.n_ls + ncmp) d_os[(size_t)(k - sp.n_ls) * L + l] = t;
            else d_noise[l] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// natural-gradient step (training.py:129-135)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ng_step(double* __restrict__ m, double* __restrict__ H,
                                                 const double* __restrict__ grad_m, const double* __restrict__ grad_H,
                                                 double lr, int M, double* __restrict__ ws, int L, int32_t* info) {
    __shared__ int flag;
    __shared__ double v1[LVAE_MAX_M], v2[LVAE_MAX_M];
    const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, MM = M * M;
    double* Hl = H + (size_t)l * MM;
    double* ml = m + (size_t)l * M;
    const double* gH = grad_H + (size_t)l * MM;
    const double* gm = grad_m + (size_t)l * M;
    double* T1 = ws + ((size_t)0 * L + l) * MM;
    double* T2 = ws + ((size_t)1 * L + l) * MM;
    double* iH = ws + ((size_t)2 * L + l) * MM;
    double* iHn = ws + ((size_t)3 * L + l) * MM;
    for (int e = tid; e < MM; e += nt) T1[e] = Hl[e];
    __syncthreads();
    int rc = cta_cholesky(T1, M, M, &flag);                        // 130
    if (rc && tid == 0) atomicCAS(info + 3, 0, l + 1);
    cta_tri_inverse(T1, T2, M, M);
    cta_gram_lower(T2, iH, M, M);                                  // 131
    for (int e = tid; e < MM; e += nt) {
        const int i = e / M, j = e % M;
        T1[e] = iH[e] + lr * (gH[i * M + j] + gH[j * M + i]);      // 132
    }
    // v1 = iH m - lr (grad_m - 2 grad_H m)   (135, uses the OLD m and iH)
    for (int i = tid; i < M; i += nt) {
        double s = 0.0, g = 0.0;
        for (int k = 0; k < M; ++k) { s += iH[i * M + k] * ml[k]; g += gH[i * M + k] * ml[k]; }
        v1[i] = s - lr * (gm[i] - 2.0 * g);
    }
    __syncthreads();
    for (int e = tid; e < MM; e += nt) iHn[e] = T1[e];
    __syncthreads();
    rc = cta_cholesky(T1, M, M, &flag);                            // 133
    if (rc && tid == 0) atomicCAS(info + 3, 0, l + 1);
    cta_tri_inverse(T1, T2, M, M);
    cta_gram_lower(T2, Hl, M, M);                                  // 134: H <- iH_new^-1
    for (int i = tid; i < M; i += nt) {
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += Hl[i * M + k] * v1[k];
        v2[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < M; i += nt) ml[i] = v2[i];
}

// ---------------------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------------------
// The head (per-latent M x M work) does not depend on the minibatch rows and the prep kernel does not depend on the
// head, so lvae_kld_head_f64 forks onto a per-device side stream (event fork / join, capture-safe) and the subject kernel
// joins it: head and prep overlap.  LVAE_NO_OVERLAP=1 keeps everything on the caller's stream.
namespace {
constexpr int MAXDEV = 16;
// `waiting` lists the user streams whose head was forked and not joined yet.  The side stream is in-order and `join` is
// re-recorded after every forked head, so waiting on it covers every head forked before; the list only says WHO still has
// to wait (two user streams or threads may interleave head / subjects / tail calls on one device).
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    std::vector<cudaStream_t> waiting;
};
SideStream g_side[MAXDEV];
std::mutex g_side_mu;
int g_overlap = -1;

SideStream* side_for_current_device() {
    if (g_overlap < 0) {
        const char* e = getenv("LVAE_NO_OVERLAP");
        g_overlap = (e && e[0] == '1') ? 0 : 1;
    }
    if (!g_overlap) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    SideStream* s = &g_side[dev];
    if (!s->st) {
        if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) { s->st = nullptr; return nullptr; }
        cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming);
    }
    return s;
}
// make `user` wait for a head that was forked onto the side stream
void join_head(cudaStream_t user) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return;
    std::lock_guard<std::mutex> lock(g_side_mu);
    SideStream* s = &g_side[dev];
    for (size_t i = 0; i < s->waiting.size(); ++i) {
        if (s->waiting[i] == user) {
            cudaStreamWaitEvent(user, s->join, 0);
            s->waiting.erase(s->waiting.begin() + i);
            break;
        }
    }
}
}  // namespace
static size_t prep_smem(const DevSpec& sp, int Tmax, int Q) {
    return sizeof(double) * ((size_t)(5 + sp.n0 + sp.n1) * Tmax * Tmax + (size_t)Tmax * Q + Tmax);
}
static size_t subj_smem(int Tmax, int M, int Q) {
    return sizeof(double) * ((size_t)2 * Tmax * M + 2 * (size_t)Tmax * Tmax + (size_t)Tmax * Q + (size_t)M * Q + 3 * M + 3 * Tmax);
}

extern "C" int lvae_kld_head_f64(const lvae_kld_problem_t* p, void* stream) {
    DevSpec sp;
    int rc = check_problem(p, &sp);
    if (rc) return rc;
    KldLayout w = lvae_layout(p);
    w.Bi_stride = p->sum_T2;
    cudaStream_t user = (cudaStream_t)stream, hs = user;
    join_head(user);                               // a previous head nobody joined yet (defensive)
    // overlap only pays when the prep kernel leaves SMs idle (small minibatches, the reference's default of 20 subjects per
    // batch); at throughput batch sizes the head's 512-thread CTAs just take SMs away from prep
    // (P_b == 0: a head-only problem of the latent-sharded tail — its caller reads W, a right after this call, keep it in order)
    SideStream* side = (p->P_b > 0 && (int64_t)p->P_b * p->L < 148 * 32) ? side_for_current_device() : nullptr;
    std::unique_lock<std::mutex> lock(g_side_mu, std::defer_lock);
    if (side) {
        lock.lock();                               // fork .. record(join) is one critical section per device table
        cudaEventRecord(side->fork, user);         // everything enqueued so far (previous step's tail / NG update) comes first
        cudaStreamWaitEvent(side->st, side->fork, 0);
        hs = side->st;
    }
    lvae_prof_begin(0, hs);
    if (w.big) {
        rc = lvae_head_big_launch(p, sp, w, hs);
    } else if (p->M <= 64 && p->path != 1) {
        rc = lvae_head64_launch(p, sp, w, hs);
    } else {
        k_head<<<p->L, 256, 0, hs>>>(sp, w, p->L, p->M, p->Q, p->z, p->m, p->H, p->lengthscale, p->outputscale, p->eps,
                                     0.5 * p->scale, p->workspace, p->info);
        LVAE_COUNT_LAUNCH();
        rc = lvae_cuda_rc(cudaGetLastError());
    }
    lvae_prof_end(0, hs);
    if (side) {
        cudaEventRecord(side->join, side->st);
        side->waiting.push_back(user);
    }
    return rc;
}

extern "C" int lvae_kld_subjects_f64(const lvae_kld_problem_t* p, void* stream) {
    DevSpec sp;
    int rc = check_problem(p, &sp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    KldLayout w = lvae_layout(p);
    w.Bi_stride = p->sum_T2;
    const double c = 0.5 * p->scale;
    if (p->P_b > 0) {
        if (w.v2) rc = lvae_plan_groups3_launch(p, w, st);
        else rc = lvae_block_offsets(p->offsets, p->P_b, reinterpret_cast<int64_t*>(p->workspace + w.off2), st);
        if (rc) return rc;
        const int Tm = p->T_max > 0 ? p->T_max : 1;
        const size_t s1 = prep_smem(sp, Tm, p->Q);
        static SmemAttrCache prep_attr;
        if (int rc_ = lvae_ensure_smem(k_prep, s1, prep_attr)) return rc_;
        lvae_prof_begin(1, st);
        if (w.prep3) {
            rc = lvae_prep3_launch(p, sp, w, st);
            if (rc) return rc;
        } else if (p->path != 1 && lvae_prep_warp_supported(p)) {
            rc = lvae_prep_warp_launch(p, sp, w, st);
            if (rc) return rc;
        } else {
            k_prep<<<dim3(w.nprep, p->L), 128, s1, st>>>(sp, w, p->L, p->Q, p->P_b, Tm, p->x, p->offsets, p->log_v,
                                                         p->lengthscale, p->outputscale, p->noise, c, p->d_log_v,
                                                         p->workspace, p->info);
            LVAE_COUNT_LAUNCH();
            rc = lvae_cuda_rc(cudaGetLastError());
            if (rc) return rc;
        }
        lvae_prof_end(1, st);
        join_head(st);                             // W, a (head) are needed from here on
        if (w.big) {
            lvae_prof_begin(2, st);
            rc = lvae_subjects_big_launch(p, sp, w, st);
            lvae_prof_end(2, st);
            if (rc) return rc;
            lvae_prof_begin(3, st);
            rc = lvae_reduce_big_launch(p, w, st);
            lvae_prof_end(3, st);
            return rc;
        }
        if (w.v3) {
            lvae_prof_begin(2, st);
            rc = lvae_subjects_fused3_launch(p, sp, w, st);
            lvae_prof_end(2, st);
            if (rc) return rc;
        } else {
            const size_t s2 = subj_smem(Tm, p->M, p->Q);
            static SmemAttrCache subj_attr;
            if (int rc_ = lvae_ensure_smem(k_subjects_generic, s2, subj_attr)) return rc_;
            lvae_prof_begin(2, st);
            k_subjects_generic<<<dim3(w.nchunk, p->L), 256, s2, st>>>(sp, w, p->L, p->M, p->Q, p->P_b, Tm, p->x,
                                                                      p->offsets, p->mu, p->z, p->lengthscale,
                                                                      p->outputscale, c, p->d_mu, p->workspace);
            lvae_prof_end(2, st);
            LVAE_COUNT_LAUNCH();
        }
    } else {
        join_head(st);
        cudaError_t e = cudaMemsetAsync(p->workspace + w.part, 0, sizeof(double) * ((size_t)w.nchunk * p->L * w.stride), st);
        if (e != cudaSuccess) return lvae_cuda_rc(e);
        e = cudaMemsetAsync(p->workspace + w.ppart, 0, sizeof(double) * ((size_t)w.nprep * p->L * (LVAE_NSCAL + w.nh)), st);
        if (e != cudaSuccess) return lvae_cuda_rc(e);
        if (w.big) {
            e = cudaMemsetAsync(p->workspace + w.part, 0, sizeof(double) * ((size_t)w.nsplit * p->L * w.stride), st);
            if (e != cudaSuccess) return lvae_cuda_rc(e);
            e = cudaMemsetAsync(p->workspace + w.bpart, 0, sizeof(double) * ((size_t)w.nchunk * p->L * w.bpstride), st);
            if (e != cudaSuccess) return lvae_cuda_rc(e);
            return lvae_reduce_big_launch(p, w, st);
        }
    }
    lvae_prof_begin(3, st);
    k_reduce<<<dim3((unsigned)((w.stride + 255) / 256), p->L), 256, 0, st>>>(w, p->L, p->M, p->workspace, p->stats);
    lvae_prof_end(3, st);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int lvae_kld_tail_f64(const lvae_kld_problem_t* p, void* stream) {
    DevSpec sp;
    int rc = check_problem(p, &sp);
    if (rc) return rc;
    KldLayout w = lvae_layout(p);
    w.Bi_stride = p->sum_T2;
    join_head((cudaStream_t)stream);
    lvae_prof_begin(4, (cudaStream_t)stream);
    if (w.big) {
        rc = lvae_tail_big_launch(p, sp, w, (cudaStream_t)stream);
        lvae_prof_end(4, (cudaStream_t)stream);
        return rc;
    }
    if (p->M <= 64 && p->path != 1) {
        rc = lvae_tail64_launch(p, sp, w, (cudaStream_t)stream);
        lvae_prof_end(4, (cudaStream_t)stream);
        return rc;
    }
    k_tail<<<p->L, 256, 0, (cudaStream_t)stream>>>(sp, w, p->L, p->M, p->Q, p->natural_gradient, p->z, p->m, p->H,
                                                   p->lengthscale, p->outputscale, 0.5 * p->scale,
                                                   p->const_term / p->L, p->stats, p->workspace, p->kld_per_latent,
                                                   p->grad_m, p->grad_H, p->d_lengthscale, p->d_outputscale, p->d_noise);
    lvae_prof_end(4, (cudaStream_t)stream);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

extern "C" int lvae_kld_minibatch_f64(const lvae_kld_problem_t* p, void* stream) {
    int rc = lvae_kld_head_f64(p, stream);
    if (rc) return rc;
    rc = lvae_kld_subjects_f64(p, stream);
    if (rc) return rc;
    return lvae_kld_tail_f64(p, stream);
}

extern "C" int64_t lvae_kld_hinv_offset(const lvae_kld_problem_t* p) { return lvae_layout(p).Hi; }
extern "C" int lvae_kld_head_offsets(const lvae_kld_problem_t* p, int64_t* out4) {
    if (!p || !out4) return LVAE_E_BADARG;
    const KldLayout w = lvae_layout(p);
    out4[0] = w.big ? w.bWp : w.W;
    out4[1] = w.big ? (int64_t)w.MP * w.MP : (int64_t)p->M * p->M;
    out4[2] = w.a;
    out4[3] = p->M;
    return 0;
}
extern "C" int64_t lvae_ng_workspace_doubles(int32_t L, int32_t M) { return M <= 64 ? 2 : lvae_ng_big_workspace(L, M); }

extern "C" int lvae_ng_step_f64(double* m, double* H, const double* grad_m, const double* grad_H, const double* Hinv,
                                double lr, int32_t L, int32_t M, double* workspace, int32_t* info, void* stream) {
    if (L <= 0 || M <= 0 || M > LVAE_MAX_M) return LVAE_E_BADARG;
    lvae_prof_begin(5, (cudaStream_t)stream);
    if (M <= 64) {
        const int rc = lvae_ng64_launch(m, H, grad_m, grad_H, Hinv, lr, L, M, info, (cudaStream_t)stream);
        lvae_prof_end(5, (cudaStream_t)stream);
        return rc;
    }
    const int rc = lvae_ng_big_launch(m, H, grad_m, grad_H, Hinv, lr, L, M, workspace, info, (cudaStream_t)stream);
    lvae_prof_end(5, (cudaStream_t)stream);
    return rc;
}
