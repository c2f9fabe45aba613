// Fused per-subject pass of the GP-prior ELBO path for M <= 64 (sm_100a, FP64 tensor pipe: DMMA.8x8x4 via mma.sync).
//
// One CTA (16 warps) owns one latent l and a contiguous range of subjects.  W = c(G - Kzz^-1), a = Kzz^-1 m, the inducing
// covariates Z_l and the hyper-parameters stay in shared memory for the whole CTA; the 64 x 64 accumulator of
// S = sum_p Kxz_p^T B_p^-1 Kxz_p stays in registers (one 16 x 16 block per warp) until the CTA retires.  Subjects are
// processed in row groups of up to 48 rows (whole subjects, e.g. two subjects of T = 20); per group, with a
// __syncthreads() between intervals:
//   I0  load covariate rows, mu rows and the (block-diagonal) B^-1 of the group
//   I1  build Kxz (R x 64) from covariates in registers (one exp per SE component entry), keep the un-scaled component
//       values f_c in registers in the accumulator layout they are needed in later, store Kxz to shared memory
//   I2  V = B^-1 Kxz (DMMA, zero blocks skipped) -> smem ;  ng1 += V^T mu ;  r = Kxz a - mu
//   I3  S += Kxz^T V (DMMA) ; Y = V W (DMMA, accumulators stay in registers) ; u = B^-1 r
//   I4  adjoint of Kxz = 2c u a^T + 2Y contracted with d k_c / d theta using the f_c still in registers ; da += Kxz^T u ;
//       A += r.u ; d_mu ; Y -> smem (over Kxz)
//   I5  Q = Y V^T on the subject-diagonal tiles (DMMA, upper triangle, weight 2) ; adjoint of B_p = -(c u u^T + Q)
//       contracted with d K1 / d theta (K1 entries recomputed) and its trace (noise)
// Kxz, V, Y, Q never touch HBM.  Shared-memory strides are = 4 (mod 16) doubles so every DMMA fragment load is
// bank-conflict-free.  Partials (S, ng1, da, A, hyper-gradients) go to this CTA's slot of the `part` workspace and are
// summed in fixed order by k_reduce (deterministic).
#include "lvae_kld.h"

namespace {

constexpr int NT = 6;             // m-tiles per row group
constexpr int RMAX = 8 * NT;      // 48 rows
constexpr int LD = 68;            // stride of 64-wide matrices (Kxz, V, W)
constexpr int LDB = 52;           // stride of the R x R block-diagonal B^-1
constexpr int NWARP = 16;
constexpr int NC1MAX = 4;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct FusedSmem {
    double* Wp;    // [64][LD]
    double* Kx;    // [RMAX][LD]   Kxz, later Y
    double* Vs;    // [RMAX][LD]
    double* BiG;   // [RMAX][LDB]
    double* zs;    // [64][Q]
    double* xs;    // [RMAX][Q]
    double* av;    // [64]
    double* mus;   // [RMAX]
    double* r;     // [RMAX]
    double* u;     // [RMAX]
    double* hyp;   // [NWARP][nh + 1]
    double* cols;  // [2][NWARP][8]
    int* blk_lo;   // [RMAX]
    int* blk_hi;   // [RMAX]
    int* ginfo;    // [4]: row0, R, next subject, nsub
};

__host__ __device__ inline size_t fused_smem_doubles(int Q, int nh) {
    return (size_t)64 * LD + 2 * (size_t)RMAX * LD + (size_t)RMAX * LDB + (size_t)64 * Q + (size_t)RMAX * Q + 64 +
           3 * RMAX + (size_t)NWARP * (nh + 1) + 2 * NWARP * 8;
}

template <int NC0>
__global__ void __launch_bounds__(512, 1)
k_subjects_fused(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int P_b,
                 const double* __restrict__ x, const int32_t* __restrict__ offsets, const double* __restrict__ mu,
                 const double* __restrict__ z, const double* __restrict__ ls, const double* __restrict__ os, double c,
                 double* __restrict__ d_mu, double* __restrict__ ws) {
    extern __shared__ double sm[];
    __shared__ double hil2[LVAE_MAXC], il3[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    const int wid = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nh = hyp_count(sp), MM = M * M;
    FusedSmem S_;
    {
        double* p = sm;
        S_.Wp = p; p += 64 * LD;
        S_.Kx = p; p += RMAX * LD;
        S_.Vs = p; p += RMAX * LD;
        S_.BiG = p; p += RMAX * LDB;
        S_.zs = p; p += 64 * Q;
        S_.xs = p; p += RMAX * Q;
        S_.av = p; p += 64;
        S_.mus = p; p += RMAX;
        S_.r = p; p += RMAX;
        S_.u = p; p += RMAX;
        S_.hyp = p; p += NWARP * (nh + 1);
        S_.cols = p; p += 2 * NWARP * 8;
        S_.blk_lo = reinterpret_cast<int*>(p);
        S_.blk_hi = S_.blk_lo + RMAX;
        S_.ginfo = S_.blk_hi + RMAX;
    }
    double* const Wp = S_.Wp; double* const Kx = S_.Kx; double* const Vs = S_.Vs; double* const BiG = S_.BiG;
    double* const zs = S_.zs; double* const xs = S_.xs; double* const av = S_.av; double* const mus = S_.mus;
    double* const rr = S_.r; double* const uu = S_.u; double* const hyp = S_.hyp;
    int* const blk_lo = S_.blk_lo; int* const blk_hi = S_.blk_hi; int* const ginfo = S_.ginfo;

    // ---- per-CTA constants ---------------------------------------------------------------------------------------
    if (tid < sp.n_ls) { const double v = ls[(size_t)tid * L + l]; hil2[tid] = 0.5 / (v * v); il3[tid] = 1.0 / (v * v * v); }
    if (tid < sp.n0 + sp.n1) osc[tid] = os[(size_t)tid * L + l];
    load_exp_table(etab);
    {
        const double* Wl = ws + w.W + (size_t)l * MM;
        for (int e = tid; e < 64 * LD; e += 512) {
            const int i = e / LD, j = e % LD;
            Wp[e] = (i < M && j < M) ? Wl[i * M + j] : 0.0;
        }
        for (int e = tid; e < 64 * Q; e += 512) zs[e] = (e < M * Q) ? z[(size_t)l * M * Q + e] : 0.0;
        if (tid < 64) av[tid] = tid < M ? ws[w.a + (size_t)l * M + tid] : 0.0;
        for (int e = tid; e < NWARP * (nh + 1); e += 512) hyp[e] = 0.0;
    }
    const int64_t* off2 = reinterpret_cast<const int64_t*>(ws + w.off2);
    const double* gBi_l = ws + w.Bi + (size_t)l * w.Bi_stride;

    // layout YL: this thread's Kxz / V / Y elements: rows 8*(h + 2i) + g, columns 8*nt + 2q + {0,1}
    const int nt = wid & 7, hh = wid >> 3;
    const int j0 = 8 * nt + 2 * q;
    // S accumulators: warp block rows 16*wi.., cols 16*wj..
    const int wi = wid >> 2, wj = wid & 3;
    double sacc[2][2][2];
#pragma unroll
    for (int a_ = 0; a_ < 2; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < 2; ++b_) sacc[a_][b_][0] = sacc[a_][b_][1] = 0.0;
    double ng1acc[2] = {0.0, 0.0}, daacc[2] = {0.0, 0.0};
    double accA = 0.0;

    const int per = (P_b + gridDim.x - 1) / gridDim.x;
    const int p_begin = chunk * per, p_end = min(P_b, p_begin + per);
    int p_next = p_begin;

    while (p_next < p_end) {            // uniform across the CTA
        // ---- I0: form the group and load it --------------------------------------------------------------------
        __syncthreads();
        if (tid == 0) {
            int p = p_next, rows = 0;
            const int row0 = offsets[p];
            while (p < p_end) {
                const int T = offsets[p + 1] - offsets[p];
                if (rows + T > RMAX && rows > 0) break;
                for (int t = 0; t < T && rows + t < RMAX; ++t) { blk_lo[rows + t] = rows; blk_hi[rows + t] = rows + T; }
                rows += T;
                ++p;
            }
            ginfo[0] = row0; ginfo[1] = rows; ginfo[2] = p; ginfo[3] = p - p_next;
        }
        for (int e = tid; e < RMAX * LDB; e += 512) BiG[e] = 0.0;
        __syncthreads();
        const int row0 = ginfo[0], R = ginfo[1], p_hi = ginfo[2];
        const int R8 = (R + 7) & ~7, nmt = R8 >> 3;
        for (int e = tid; e < RMAX * Q; e += 512) xs[e] = (e < R * Q) ? x[(size_t)row0 * Q + e] : 0.0;
        if (tid < RMAX) {
            mus[tid] = tid < R ? mu[(size_t)(row0 + tid) * L + l] : 0.0;
            if (tid >= R) { blk_lo[tid] = tid; blk_hi[tid] = tid; }
        }
        for (int p = p_next; p < p_hi; ++p) {
            const int o = offsets[p] - row0, T = offsets[p + 1] - offsets[p];
            const double* src = gBi_l + off2[p];
            for (int e = tid; e < T * T; e += 512) BiG[(o + e / T) * LDB + o + e % T] = src[e];
        }
        p_next = p_hi;
        __syncthreads();

        // ---- I1: Kxz from covariates; f_c kept in registers ---------------------------------------------------------
        double fc[NT / 2][2][NC0];
#pragma unroll
        for (int i = 0; i < NT / 2; ++i) {
            const int mt = hh + 2 * i, t = 8 * mt + g;
            double kx[2] = {0.0, 0.0};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + e;
                const bool valid = (t < R) && (j < M);
#pragma unroll
                for (int cc = 0; cc < NC0; ++cc) {
                    double d2, f = 0.0;
                    if (valid) f = comp_value(sp, cc, xs + t * Q, zs + j * Q, hil2, d2, etab);
                    fc[i][e][cc] = f;
                    kx[e] += osc[cc] * f;
                }
            }
            if (mt < nmt) *reinterpret_cast<double2*>(Kx + t * LD + j0) = make_double2(kx[0], kx[1]);
        }
        __syncthreads();

        // ---- I2: r = Kxz a - mu (3 rows per warp) ; V = BiG Kxz ; ng1 += V^T mu ----------------------------------
        for (int t = wid; t < R; t += NWARP) {
            double s = Kx[t * LD + lane] * av[lane] + Kx[t * LD + 32 + lane] * av[32 + lane];
            s = warp_sum(s);
            if (lane == 0) rr[t] = s - mus[t];
        }
#pragma unroll
        for (int i = 0; i < NT / 2; ++i) {
            const int mt = hh + 2 * i;
            if (mt < nmt) {
                const int t = 8 * mt + g;
                const int klo = blk_lo[8 * mt] >> 2, khi = (blk_hi[min(8 * mt + 7, R - 1)] + 3) >> 2;
                double v0 = 0.0, v1 = 0.0;
                for (int ks = klo; ks < khi; ++ks) {
                    const double a_ = BiG[t * LDB + 4 * ks + q];
                    const double b_ = Kx[(4 * ks + q) * LD + 8 * nt + g];
                    dmma(v0, v1, a_, b_);
                }
                *reinterpret_cast<double2*>(Vs + t * LD + j0) = make_double2(v0, v1);
                const double mt_ = mus[t];
                ng1acc[0] += v0 * mt_;
                ng1acc[1] += v1 * mt_;
            }
        }
        __syncthreads();

        // ---- I3: u = BiG r ; S += Kxz^T V ; Y = V W ----------------------------------------------------------------
        for (int t = wid; t < R; t += NWARP) {
            double s = 0.0;
            for (int k = blk_lo[t] + lane; k < blk_hi[t]; k += 32) s += BiG[t * LDB + k] * rr[k];
            s = warp_sum(s);
            if (lane == 0) uu[t] = s;
        }
        {
            const int nks = R8 >> 2;
            const double* Ka = Kx + 16 * wi + g;
            const double* Vb = Vs + 16 * wj + g;
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks) {
                const int ro = (4 * ks + q) * LD;
                const double a0 = Ka[ro], a1 = Ka[ro + 8];
                const double b0 = Vb[ro], b1 = Vb[ro + 8];
                dmma(sacc[0][0][0], sacc[0][0][1], a0, b0);
                dmma(sacc[0][1][0], sacc[0][1][1], a0, b1);
                dmma(sacc[1][0][0], sacc[1][0][1], a1, b0);
                dmma(sacc[1][1][0], sacc[1][1][1], a1, b1);
            }
        }
        double yacc[NT / 2][2];
#pragma unroll
        for (int i = 0; i < NT / 2; ++i) yacc[i][0] = yacc[i][1] = 0.0;
        {
#pragma unroll 4
            for (int ks = 0; ks < 16; ++ks) {
                const double b_ = Wp[(4 * ks + q) * LD + 8 * nt + g];
#pragma unroll
                for (int i = 0; i < NT / 2; ++i) {
                    const int mt = hh + 2 * i;
                    if (mt < nmt) {
                        const double a_ = Vs[(8 * mt + g) * LD + 4 * ks + q];
                        dmma(yacc[i][0], yacc[i][1], a_, b_);
                    }
                }
            }
        }
        __syncthreads();

        // ---- I4: adjoint of Kxz contracted with the component derivatives ; da ; A ; d_mu ; Y -> smem -------------
        {
            double gos[NC0], gls[NC0];
#pragma unroll
            for (int cc = 0; cc < NC0; ++cc) gos[cc] = gls[cc] = 0.0;
#pragma unroll
            for (int i = 0; i < NT / 2; ++i) {
                const int mt = hh + 2 * i, t = 8 * mt + g;
                if (mt < nmt) {
                    const double ut = uu[t];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = j0 + e;
                        const double gbar = 2.0 * c * ut * av[j] + 2.0 * yacc[i][e];
                        double kx = 0.0;
#pragma unroll
                        for (int cc = 0; cc < NC0; ++cc) {
                            const double f = fc[i][e][cc];
                            kx += osc[cc] * f;
                            gos[cc] += gbar * f;
                            const int rd = sp.rbf_dim[cc];
                            if (rd >= 0) {
                                const double d = xs[t * Q + rd] - zs[j * Q + rd];
                                gls[cc] += gbar * f * (d * d);
                            }
                        }
                        daacc[e] += kx * ut;
                    }
                    *reinterpret_cast<double2*>(Kx + t * LD + j0) = make_double2(yacc[i][0], yacc[i][1]);
                }
            }
#pragma unroll
            for (int cc = 0; cc < NC0; ++cc) {
                const double a_ = warp_sum(gos[cc]);
                const int rd = sp.rbf_dim[cc];
                double b_ = 0.0;
                if (rd >= 0) b_ = warp_sum(gls[cc]);
                if (lane == 0) {
                    hyp[wid * (nh + 1) + 1 + sp.n_ls + cc] += a_;
                    if (rd >= 0) hyp[wid * (nh + 1) + 1 + sp.ls_idx[cc]] += b_ * osc[cc] * il3[sp.ls_idx[cc]];
                }
            }
            if (tid < R) {
                accA += rr[tid] * uu[tid];
                d_mu[(size_t)(row0 + tid) * L + l] = -2.0 * c * uu[tid];
            }
        }
        __syncthreads();

        // ---- I5: Q = Y V^T on subject-diagonal tiles (upper triangle) ; adjoint of B_p ; K1 hyper-gradients ---------
        {
            double g1os[NC1MAX], g1ls[NC1MAX], gno = 0.0;
#pragma unroll
            for (int k = 0; k < NC1MAX; ++k) g1os[k] = g1ls[k] = 0.0;
            int pair = 0;
            for (int i = 0; i < nmt; ++i) {
                const int khi_i = blk_hi[min(8 * i + 7, R - 1)];
                for (int jt = i; jt < nmt && 8 * jt < khi_i; ++jt, ++pair) {
                    if ((pair & (NWARP - 1)) != wid) continue;
                    double q0 = 0.0, q1 = 0.0;
                    const double* Ya = Kx + (8 * i + g) * LD + q;
                    const double* Vb = Vs + (8 * jt + g) * LD + q;
#pragma unroll 4
                    for (int ks = 0; ks < 16; ++ks) dmma(q0, q1, Ya[4 * ks], Vb[4 * ks]);
                    const int t = 8 * i + g;
                    const double wgt = jt > i ? 2.0 : 1.0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int t2 = 8 * jt + 2 * q + e;
                        if (t < R && t2 < R && blk_lo[t] == blk_lo[t2]) {
                            const double gB = -wgt * (c * uu[t] * uu[t2] + (e ? q1 : q0));
                            if (t == t2) gno += gB;
#pragma unroll
                            for (int k = 0; k < NC1MAX; ++k) {
                                if (k < sp.n1) {
                                    const int cc = sp.n0 + k;
                                    double d2;
                                    const double f = comp_value(sp, cc, xs + t * Q, xs + t2 * Q, hil2, d2, etab);
                                    g1os[k] += gB * f;
                                    if (sp.rbf_dim[cc] >= 0) g1ls[k] += gB * f * d2;
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NC1MAX; ++k) {
                if (k < sp.n1) {
                    const int cc = sp.n0 + k;
                    const double a_ = warp_sum(g1os[k]);
                    double b_ = 0.0;
                    if (sp.rbf_dim[cc] >= 0) b_ = warp_sum(g1ls[k]);
                    if (lane == 0) {
                        hyp[wid * (nh + 1) + 1 + sp.n_ls + cc] += a_;
                        if (sp.rbf_dim[cc] >= 0) hyp[wid * (nh + 1) + 1 + sp.ls_idx[cc]] += b_ * osc[cc] * il3[sp.ls_idx[cc]];
                    }
                }
            }
            gno = warp_sum(gno);
            if (lane == 0) hyp[wid * (nh + 1) + nh] += gno;
        }
    }

    // ---- CTA epilogue: partials to the workspace --------------------------------------------------------------------
    double* part = ws + w.part + ((size_t)chunk * L + l) * w.stride;
    {
        accA = warp_sum(accA);
        if (lane == 0) hyp[wid * (nh + 1)] += accA;
        // column sums held per thread (columns j0, j0+1): reduce over g (lanes with the same q), then over the two halves
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double a_ = ng1acc[e], b_ = daacc[e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a_ += __shfl_xor_sync(0xffffffffu, a_, o); b_ += __shfl_xor_sync(0xffffffffu, b_, o); }
            if (g == 0) { S_.cols[(0 * NWARP + wid) * 8 + 2 * q + e] = a_; S_.cols[(1 * NWARP + wid) * 8 + 2 * q + e] = b_; }
        }
    }
    __syncthreads();
    if (tid < 64) {
        const int ntile = tid >> 3, cidx = tid & 7;
        if (tid < M) {
            part[stats_off_ng1(M) + tid] = S_.cols[(0 * NWARP + ntile) * 8 + cidx] + S_.cols[(0 * NWARP + ntile + 8) * 8 + cidx];
            part[stats_off_da(M) + tid] = S_.cols[(1 * NWARP + ntile) * 8 + cidx] + S_.cols[(1 * NWARP + ntile + 8) * 8 + cidx];
        }
    }
    if (tid <= nh) {
        double s = 0.0;
        for (int ww = 0; ww < NWARP; ++ww) s += hyp[ww * (nh + 1) + tid];
        if (tid == 0) {
            for (int k = 0; k < LVAE_NSCAL; ++k) part[stats_off_scal(M) + k] = 0.0;
            part[stats_off_scal(M) + SC_A] = s;
        } else {
            part[stats_off_hyp(M) + tid - 1] = s;
        }
    }
#pragma unroll
    for (int a_ = 0; a_ < 2; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < 2; ++b_)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 16 * wi + 8 * a_ + g, j = 16 * wj + 8 * b_ + 2 * q + e;
                if (i < M && j < M) part[stats_off_S() + (size_t)i * M + j] = sacc[a_][b_][e];
            }
}

template <int NC0>
int launch_nc0(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    const size_t smem = sizeof(double) * fused_smem_doubles(p->Q, w.nh) + sizeof(int) * (2 * RMAX + 4);
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_subjects_fused<NC0>, smem, attr)) return rc_;
    k_subjects_fused<NC0><<<dim3(w.nchunk, p->L), 512, smem, st>>>(sp, w, p->L, p->M, p->Q, p->P_b, p->x, p->offsets,
                                                                   p->mu, p->z, p->lengthscale, p->outputscale,
                                                                   0.5 * p->scale, p->d_mu, p->workspace);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

bool lvae_fused_supported(const lvae_kld_problem_t* p) {
    return p->M <= 64 && p->T_max <= RMAX && p->ks.n_comp0 >= 1 && p->ks.n_comp0 <= 6 && p->ks.n_comp1 <= NC1MAX &&
           p->Q <= 16;
}

int lvae_subjects_fused_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    if (!lvae_fused_supported(p)) return LVAE_E_TOO_LARGE;
    switch (sp.n0) {
        case 1: return launch_nc0<1>(p, sp, w, st);
        case 2: return launch_nc0<2>(p, sp, w, st);
        case 3: return launch_nc0<3>(p, sp, w, st);
        case 4: return launch_nc0<4>(p, sp, w, st);
        case 5: return launch_nc0<5>(p, sp, w, st);
        case 6: return launch_nc0<6>(p, sp, w, st);
    }
    return LVAE_E_BADARG;
}
