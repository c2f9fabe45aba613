// Fused per-subject pass, second generation (M <= 64, at most 24 rows per subject): sm_100a, FP64 tensor pipe (DMMA.8x8x4).
//
// One CTA owns a latent l and a contiguous range of subjects and runs TWO independent 8-warp sets.  The sets share W,
// a = Kzz^-1 m, Z_l and the hyper-parameters in shared memory, each works on its own row groups (whole subjects, <= 24
// rows) with named barriers, so one set's latency-bound intervals overlap the other's tensor-pipe intervals.
// Per group (set-local barrier between intervals):
//   J0  wait for the cp.async prefetch of this group (covariates, mu, B^-1 mu, block-diagonal L^-1) ; issue the next one
//   J1  Kxz (R x 64) from covariates in registers ; f_c stay in registers ; partial dots of r = Kxz a - mu
//   J2  U = L^-1 Kxz            (DMMA, lower-triangular k-range)
//   J3  V = L^-T U              (DMMA, upper-triangular k-range) ; S += U^T U (SYRK: 36 lower tiles, accumulators in
//       registers for the whole kernel) ; ng1 += V^T mu ; partial dots of u = V a - B^-1 mu
//   J4  Y = V W                 (DMMA, accumulators stay in registers)
//   J5  adjoint of Kxz = 2c u a^T + 2Y contracted with d k_c / d theta (f_c from registers) ; da ; A ; d_mu ; Y -> smem
//   J6  Q = Y V^T on the subject-diagonal upper tiles (DMMA) ; adjoint of B_p = -(c u u^T + Q) contracted with d K1
// Row groups are planned once per step by k_plan_groups (they do not depend on the latent).  L^-1 and B^-1 mu come
// from the prep kernel in row-major per-row layout.  Nothing of size T x M touches HBM.
#include "lvae_kld.h"

namespace {

constexpr int RG = LVAE_F2_ROWS;   // 24 rows per group
constexpr int NMT = RG / 8;        // 3 m-tiles
constexpr int LD = 68;
constexpr int LDL = 28;
constexpr int SETW = 8;            // warps per set
constexpr int GT = LVAE_F2_GT;     // ints per group-table entry

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 8-byte asynchronous copy; !valid writes zeros (src-size 0) and never forms an out-of-range address (falls back to `safe`)
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid, const void* safe) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    const void* src = valid ? gmem : safe;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void set_barrier(int set) {
    asm volatile("bar.sync %0, %1;" ::"r"(set + 1), "r"(SETW * 32) : "memory");
}
__device__ __forceinline__ void tri2(int e, int& i, int& j) {   // e-th element of a lower triangle, j <= i
    i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    j = e - i * (i + 1) / 2;
}

// Covariates are kept GATHERED per component: slot 4*c + 0 holds the column of component c's squared-exponential factor,
// slots 4*c + 1..3 the columns of its categorical/binary factors (unused slots are zero).  XC[slot][row] for the rows of a
// group (all components, per set, double-buffered, filled by cp.async), ZC[slot][column] for the inducing points (K0
// components, per CTA).  Every inner-loop address is then base + compile-time offset.
constexpr int CS = 4;

// Per-set shared memory, all at compile-time offsets from the set base; the size-dependent blocks (XC, FC) come last.
struct SetSmem {
    double* sb;
    __device__ __forceinline__ double* B1() const { return sb; }                          // [RG][LD]  Kxz, then V
    __device__ __forceinline__ double* B2() const { return sb + RG * LD; }                // [RG][LD]  U, then Y
    __device__ __forceinline__ double* Lg() const { return sb + 2 * RG * LD; }            // [2][RG][LDL]
    __device__ __forceinline__ double* mus() const { return Lg() + 2 * RG * LDL; }        // [2][RG]
    __device__ __forceinline__ double* bmu() const { return mus() + 2 * RG; }             // [2][RG]
    __device__ __forceinline__ double* rpart() const { return bmu() + 2 * RG; }           // [SETW][RG]
    __device__ __forceinline__ double* upart() const { return rpart() + SETW * RG; }      // [SETW][RG]
    __device__ __forceinline__ double* rs() const { return upart() + SETW * RG; }         // [RG]
    __device__ __forceinline__ double* us() const { return rs() + RG; }                   // [RG]
    __device__ __forceinline__ int* meta() const { return reinterpret_cast<int*>(us() + RG); }   // [3][GT], 16-byte aligned
    __device__ __forceinline__ int* blo() const { return meta() + 3 * GT; }               // [RG]
    __device__ __forceinline__ int* bhi() const { return blo() + RG; }
    __device__ __forceinline__ int* nlo() const { return bhi() + RG; }                    // same for the NEXT group
    __device__ __forceinline__ int* nhi() const { return nlo() + RG; }
    __device__ __forceinline__ double* XC() const { return us() + RG + (3 * GT + 4 * RG) / 2; }   // [2][NCT*CS][RG]
};
constexpr int SET_FIXED = 2 * RG * LD + 2 * RG * LDL + 4 * RG + 2 * SETW * RG + 2 * RG + (3 * GT + 4 * RG) / 2;

__host__ __device__ inline size_t set_doubles(int nct, int nr) {
    return (size_t)SET_FIXED + 2 * (size_t)nct * CS * RG + (size_t)nr * RG * LD;
}
__host__ __device__ inline size_t common_doubles(int nc0, int nct, int nh) {
    return (size_t)64 * LD + 64 + 2 * 16 * 8 + (size_t)16 * (nh + 2) + (size_t)nc0 * CS * 64 + (size_t)(nct * CS + 2) / 2 + 2;
}

template <int NC0, int NC1>
__global__ void __launch_bounds__(512, 1)
k_subjects_fused2(const __grid_constant__ DevSpec sp, const __grid_constant__ KldLayout w, int L, int M, int Q, int N_b, int TP,
                  int NR, const double* __restrict__ x, const double* __restrict__ mu, const double* __restrict__ z,
                  const double* __restrict__ ls, const double* __restrict__ os, double c, double* __restrict__ d_mu,
                  double* __restrict__ ws) {
    extern __shared__ double sm[];
    __shared__ double hil2[LVAE_MAXC], il3[LVAE_MAXC], osc[LVAE_MAXC], etab[LVAE_EXP_TBL];
    constexpr int NCT = NC0 + NC1;
    constexpr int XCSZ = NCT * CS * RG;
    const int chunk = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
    const int set = tid >> 8, lt = tid & 255, wid = tid >> 5, wl = wid & 7, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int nh = hyp_count(sp), MM = M * M;

    double* const Wp = sm;                           // [64][LD]
    double* const av = Wp + 64 * LD;                 // [64]
    double* const cols = av + 64;                    // [2][16][8]
    double* const hyp = cols + 2 * 16 * 8;           // [16][nh + 1]
    double* const ZC = hyp + ((16 * (nh + 2)) & ~1); // [NC0*CS][64]
    int* const dimtab = reinterpret_cast<int*>(ZC + NC0 * CS * 64);   // [NCT*CS]
    double* const sets = ZC + NC0 * CS * 64 + ((NCT * CS + 2) / 2) + ((((NCT * CS + 2) / 2) & 1));
    SetSmem S;
    S.sb = sets + (size_t)set * set_doubles(NCT, NR);
    double* const FC = S.XC() + 2 * XCSZ;            // [NR][RG][LD]  un-scaled values of the SE-bearing K0 components
    double* const sxch = sets + set_doubles(NCT, NR);   // set 1's B1|B2 (36*64 doubles), reused at the very end

    // ---- per-CTA constants -------------------------------------------------------------------------------------------
    if (tid < sp.n_ls) { const double v = ls[(size_t)tid * L + l]; hil2[tid] = 0.5 / (v * v); il3[tid] = 1.0 / (v * v * v); }
    if (tid < sp.n0 + sp.n1) osc[tid] = os[(size_t)tid * L + l];
    load_exp_table(etab);
    if (tid < NCT * CS) {
        const int cc = tid / CS, sl = tid % CS;
        int dim = -1;
        if (sl == 0) dim = sp.rbf_dim[cc];
        else if (sl - 1 < sp.n_mask[cc]) dim = sp.mask_dim[cc][sl - 1];
        dimtab[tid] = dim;
    }
    __syncthreads();
    {
        const double* Wl = ws + w.W + (size_t)l * MM;
        for (int e = tid; e < 64 * LD; e += 512) {
            const int i = e / LD, j = e % LD;
            Wp[e] = (i < M && j < M) ? Wl[i * M + j] : 0.0;
        }
        for (int e = tid; e < NC0 * CS * 64; e += 512) {
            const int sl = e >> 6, j = e & 63, dim = dimtab[sl];
            ZC[e] = (dim >= 0 && j < M) ? z[((size_t)l * M + j) * Q + dim] : 0.0;
        }
        if (tid < 64) av[tid] = tid < M ? ws[w.a + (size_t)l * M + tid] : 0.0;
        for (int e = tid; e < 16 * (nh + 1); e += 512) hyp[e] = 0.0;
    }
    const int* gtab = reinterpret_cast<const int*>(ws + w.gtab) + (size_t)chunk * w.gstride * GT;
    const int ngroups = reinterpret_cast<const int*>(ws + w.gcount)[chunk];
    const double* Lrows = ws + w.Lrows + (size_t)l * N_b * TP;
    const double* bmu_g = ws + w.bmu + (size_t)l * N_b;

    // this thread's Kxz / U / V / Y elements: rows 8*mt + g (mt = 0..2), columns j0 = 8*wl + 2q and j0 + 1
    const int j0 = 8 * wl + 2 * q;
    const bool cv0 = j0 < M, cv1 = j0 + 1 < M;
    // SYRK tiles of S owned by this warp (lower triangle of the 8 x 8 tile grid, 36 tiles over 8 warps)
    double sacc[5][2];
    int so_i[5], so_j[5];                            // column offsets (8 * tile index) of the A and B fragments in U
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int e = wl + 8 * i;
        int ti = 0, tj = 0;
        if (e < 36) tri2(e, ti, tj);
        so_i[i] = 8 * ti; so_j[i] = 8 * tj;
        sacc[i][0] = sacc[i][1] = 0.0;
    }
    double ng1acc[2] = {0.0, 0.0}, daacc[2] = {0.0, 0.0}, accA = 0.0;
    double gos[NC0], gls[NC0], g1os[NC1], g1ls[NC1], gno = 0.0;
#pragma unroll
    for (int cc = 0; cc < NC0; ++cc) gos[cc] = gls[cc] = 0.0;
#pragma unroll
    for (int k = 0; k < NC1; ++k) g1os[k] = g1ls[k] = 0.0;

    // row -> [lo, hi) of its subject inside a group, from the group's meta entry
    auto row_block = [&](const int* mt_, int t, int R, int& lo, int& hi) {
        lo = t; hi = t;
        if (t < R) {
            lo = 0; hi = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) { const int end = mt_[3 + s]; if (t >= end) lo = end; }
#pragma unroll
            for (int s = 4; s >= 0; --s) { const int end = mt_[3 + s]; if (t < end) hi = end; }
        }
    };
    // issue the prefetch of a group (meta entry visible in smem slot `slot`; nlo/nhi hold its row blocks) into buffer `buf`
    auto issue_data = [&](int slot, int buf) {
        const int* mt_ = S.meta() + slot * GT;
        const int row0 = mt_[0], R = mt_[1];
        double* xc = S.XC() + buf * XCSZ;
#pragma unroll
        for (int rep = 0; rep < (XCSZ + 255) / 256; ++rep) {
            const int e = lt + 256 * rep;
            if (e < XCSZ) {
                const int sl = e / RG, t = e - sl * RG, dim = dimtab[sl];
                cp_async8(xc + e, x + (size_t)(row0 + t) * Q + dim, (t < R) && (dim >= 0), x);
            }
        }
        if (lt < RG) {
            cp_async8(S.mus() + buf * RG + lt, mu + (size_t)(row0 + lt) * L + l, lt < R, x);
            cp_async8(S.bmu() + buf * RG + lt, bmu_g + row0 + lt, lt < R, x);
        }
        double* dst = S.Lg() + buf * RG * LDL;
#pragma unroll
        for (int rep = 0; rep < 3; ++rep) {
            const int e = lt + 256 * rep;
            if (e < RG * RG) {
                const int t = e / RG, k = e - t * RG;
                const int lo = S.nlo()[t];
                cp_async8(dst + t * LDL + k, Lrows + (size_t)(row0 + t) * TP + (k - lo), (k >= lo) && (k < S.nhi()[t]), x);
            }
        }
    };
    auto issue_meta = [&](int gi, int slot) {
        if (lt < 2) cp_async16(S.meta() + slot * GT + 4 * lt, gtab + (size_t)gi * GT + 4 * lt);
    };

    __syncthreads();
    // ---- prologue of the 3-stage prefetch pipeline -------------------------------------------------------------------------
    const int first = set;
    if (first < ngroups) {
        issue_meta(first, 0);
        if (first + 2 < ngroups) issue_meta(first + 2, 1);
        cp_async_commit();
        cp_async_wait_all();
        set_barrier(set);
        if (lt < RG) { int lo, hi; row_block(S.meta(), lt, S.meta()[1], lo, hi); S.nlo()[lt] = lo; S.nhi()[lt] = hi; }
        set_barrier(set);
        issue_data(0, 0);
        cp_async_commit();
    }

    int it = 0;
    for (int gi = first; gi < ngroups; gi += 2, ++it) {
        const int buf = it & 1, slot = it % 3;
        // ---- J0 ----------------------------------------------------------------------------------------------------------
        cp_async_wait_all();
        set_barrier(set);
        const int* mt_ = S.meta() + slot * GT;
        const int row0 = mt_[0], R = mt_[1];
        const int R8 = (R + 7) & ~7, nmt = R8 >> 3, nk4 = (R + 3) >> 2;
        const double* xc = S.XC() + buf * XCSZ;
        const double* mus = S.mus() + buf * RG;
        const double* Lg = S.Lg() + buf * RG * LDL;
        const bool more = gi + 2 < ngroups;
        if (lt < RG) {
            S.blo()[lt] = S.nlo()[lt]; S.bhi()[lt] = S.nhi()[lt];          // planned when this group was prefetched
        }

        // ---- J1: Kxz from the gathered covariates ; SE-bearing f_c -> FC (smem) ; partial dots of r ---------------------------
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            const bool rv = t < R;
            double kx0 = 0.0, kx1 = 0.0;
            int fslot = 0;
#pragma unroll
            for (int cc = 0; cc < NC0; ++cc) {
                bool on0 = rv && cv0, on1 = rv && cv1;
#pragma unroll
                for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
                    if (i < sp.n_mask[cc]) {
                        const double a = xc[(cc * CS + 1 + i) * RG + t];
                        const double2 b = *reinterpret_cast<const double2*>(ZC + (cc * CS + 1 + i) * 64 + j0);
                        if (sp.mask_type[cc][i] == LVAE_CAT) { on0 = on0 && (a - b.x == 0.0); on1 = on1 && (a - b.y == 0.0); }
                        else { on0 = on0 && (a + b.x == 2.0); on1 = on1 && (a + b.y == 2.0); }
                    }
                }
                double f0 = on0 ? 1.0 : 0.0, f1 = on1 ? 1.0 : 0.0;
                if (sp.rbf_dim[cc] >= 0) {
                    const double a = xc[(cc * CS) * RG + t], h = hil2[sp.ls_idx[cc]];
                    const double2 b = *reinterpret_cast<const double2*>(ZC + (cc * CS) * 64 + j0);
                    const double t0 = a - b.x, t1 = a - b.y;
                    const double e0 = exp_neg(-(t0 * t0) * h, etab), e1 = exp_neg(-(t1 * t1) * h, etab);
                    f0 = on0 ? e0 : 0.0;
                    f1 = on1 ? e1 : 0.0;
                    *reinterpret_cast<double2*>(FC + fslot * RG * LD + t * LD + j0) = make_double2(f0, f1);
                    ++fslot;
                }
                kx0 += osc[cc] * f0;
                kx1 += osc[cc] * f1;
            }
            *reinterpret_cast<double2*>(S.B1() + t * LD + j0) = make_double2(kx0, kx1);
            double pr = kx0 * av[j0] + kx1 * av[j0 + 1];
            pr += __shfl_xor_sync(0xffffffffu, pr, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2);
            if (q == 0) S.rpart()[wl * RG + t] = pr;
        }
        set_barrier(set);
        // row blocks of the next group of this set (its meta entry arrived with this group's data)
        if (more) {
            const int* nm = S.meta() + ((it + 1) % 3) * GT;
            if (lt < RG) { int lo, hi; row_block(nm, lt, nm[1], lo, hi); S.nlo()[lt] = lo; S.nhi()[lt] = hi; }
        }

        // ---- J2: U = L^-1 Kxz (rows of a tile only see k <= row, inside their subject) ; r -----------------------------------------
        if (lt < RG) {
            double s = -mus[lt];
#pragma unroll
            for (int ww = 0; ww < SETW; ++ww) s += S.rpart()[ww * RG + lt];
            S.rs()[lt] = s;
        }
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            double u0 = 0.0, u1 = 0.0;
            if (mt < nmt) {
                const int klo = S.blo()[8 * mt] >> 2;
                const int khi = min(2 * mt + 2, (S.bhi()[min(8 * mt + 7, R - 1)] + 3) >> 2);
                for (int ks = klo; ks < khi; ++ks)
                    dmma(u0, u1, Lg[t * LDL + 4 * ks + q], S.B1()[(4 * ks + q) * LD + 8 * wl + g]);
            }
            *reinterpret_cast<double2*>(S.B2() + t * LD + j0) = make_double2(u0, u1);
        }
        set_barrier(set);
        if (more) {
            issue_data((it + 1) % 3, buf ^ 1);
            if (gi + 4 < ngroups) issue_meta(gi + 4, (it + 2) % 3);
            cp_async_commit();
        }

        // ---- J3: V = L^-T U -> B1 ; ng1 ; partial dots of u ; S += U^T U ------------------------------------------------------
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            double v0 = 0.0, v1 = 0.0;
            if (mt < nmt) {
                const int khi = (S.bhi()[min(8 * mt + 7, R - 1)] + 3) >> 2;
                for (int ks = 2 * mt; ks < khi; ++ks)
                    dmma(v0, v1, Lg[(4 * ks + q) * LDL + 8 * mt + g], S.B2()[(4 * ks + q) * LD + 8 * wl + g]);
            }
            *reinterpret_cast<double2*>(S.B1() + t * LD + j0) = make_double2(v0, v1);
            const double mt_mu = mus[t];
            ng1acc[0] += v0 * mt_mu;
            ng1acc[1] += v1 * mt_mu;
            double pu = v0 * av[j0] + v1 * av[j0 + 1];
            pu += __shfl_xor_sync(0xffffffffu, pu, 1);
            pu += __shfl_xor_sync(0xffffffffu, pu, 2);
            if (q == 0) S.upart()[wl * RG + t] = pu;
        }
        for (int ks = 0; ks < nk4; ++ks) {
            const double* Urow = S.B2() + (4 * ks + q) * LD + g;
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(sacc[i][0], sacc[i][1], Urow[so_i[i]], Urow[so_j[i]]);
            if (wl < 4) dmma(sacc[4][0], sacc[4][1], Urow[so_i[4]], Urow[so_j[4]]);
        }
        set_barrier(set);

        // ---- J4: u = V a - B^-1 mu ; Y = V W ---------------------------------------------------------------------------------
        if (lt < RG) {
            double s = -S.bmu()[buf * RG + lt];
#pragma unroll
            for (int ww = 0; ww < SETW; ++ww) s += S.upart()[ww * RG + lt];
            S.us()[lt] = lt < R ? s : 0.0;
        }
        double yacc[NMT][2];
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) yacc[mt][0] = yacc[mt][1] = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            const double b_ = Wp[(4 * ks + q) * LD + 8 * wl + g];
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
                if (mt < nmt) dmma(yacc[mt][0], yacc[mt][1], S.B1()[(8 * mt + g) * LD + 4 * ks + q], b_);
            }
        }
        set_barrier(set);

        // ---- J5: adjoint of Kxz against the component derivatives ; da ; A ; d_mu ; Y -> B2 ------------------------------------------
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int t = 8 * mt + g;
            const bool rv = t < R;
            const double ut = S.us()[t];
            const double gb0 = 2.0 * c * ut * av[j0] + 2.0 * yacc[mt][0];
            const double gb1 = 2.0 * c * ut * av[j0 + 1] + 2.0 * yacc[mt][1];
            double kx0 = 0.0, kx1 = 0.0;
            int fslot = 0;
#pragma unroll
            for (int cc = 0; cc < NC0; ++cc) {
                double f0, f1;
                if (sp.rbf_dim[cc] >= 0) {
                    const double2 f = *reinterpret_cast<const double2*>(FC + fslot * RG * LD + t * LD + j0);
                    ++fslot;
                    f0 = f.x; f1 = f.y;
                    const double a = xc[(cc * CS) * RG + t];
                    const double2 b = *reinterpret_cast<const double2*>(ZC + (cc * CS) * 64 + j0);
                    const double t0 = a - b.x, t1 = a - b.y;
                    gls[cc] += gb0 * f0 * (t0 * t0) + gb1 * f1 * (t1 * t1);
                } else {
                    bool on0 = rv && cv0, on1 = rv && cv1;
#pragma unroll
                    for (int i = 0; i < LVAE_MAX_MASKS; ++i) {
                        if (i < sp.n_mask[cc]) {
                            const double a = xc[(cc * CS + 1 + i) * RG + t];
                            const double2 b = *reinterpret_cast<const double2*>(ZC + (cc * CS + 1 + i) * 64 + j0);
                            if (sp.mask_type[cc][i] == LVAE_CAT) { on0 = on0 && (a - b.x == 0.0); on1 = on1 && (a - b.y == 0.0); }
                            else { on0 = on0 && (a + b.x == 2.0); on1 = on1 && (a + b.y == 2.0); }
                        }
                    }
                    f0 = on0 ? 1.0 : 0.0; f1 = on1 ? 1.0 : 0.0;
                }
                gos[cc] += gb0 * f0 + gb1 * f1;
                kx0 += osc[cc] * f0;
                kx1 += osc[cc] * f1;
            }
            daacc[0] += kx0 * ut;
            daacc[1] += kx1 * ut;
            *reinterpret_cast<double2*>(S.B2() + t * LD + j0) = make_double2(yacc[mt][0], yacc[mt][1]);
        }
        if (lt < R) {
            accA += S.rs()[lt] * S.us()[lt];
            d_mu[(size_t)(row0 + lt) * L + l] = -2.0 * c * S.us()[lt];
        }
        set_barrier(set);

        // ---- J6: Q = Y V^T on subject-diagonal upper tiles ; adjoint of B_p against d K1 / d theta -----------------------------------
        if (wl < 6) {
            int jt, i;
            tri2(wl, jt, i);                                 // i <= jt
            if (jt < nmt && 8 * jt < S.bhi()[min(8 * i + 7, R - 1)]) {
                double q0 = 0.0, q1 = 0.0;
                const double* Ya = S.B2() + (8 * i + g) * LD + q;
                const double* Vb = S.B1() + (8 * jt + g) * LD + q;
#pragma unroll 4
                for (int ks = 0; ks < 16; ++ks) dmma(q0, q1, Ya[4 * ks], Vb[4 * ks]);
                const int t = 8 * i + g;
                const double wgt = jt > i ? 2.0 : 1.0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int t2 = 8 * jt + 2 * q + e;
                    if (t < R && t2 < R && S.blo()[t] == S.blo()[t2]) {
                        const double gB = -wgt * (c * S.us()[t] * S.us()[t2] + (e ? q1 : q0));
                        if (t == t2) gno += gB;
#pragma unroll
                        for (int k = 0; k < NC1; ++k) {
                            const int cc = NC0 + k;
                            bool on = true;
#pragma unroll
                            for (int i2 = 0; i2 < LVAE_MAX_MASKS; ++i2) {
                                if (i2 < sp.n_mask[cc]) {
                                    const double a = xc[(cc * CS + 1 + i2) * RG + t], b = xc[(cc * CS + 1 + i2) * RG + t2];
                                    on = on && ((sp.mask_type[cc][i2] == LVAE_CAT) ? (a - b == 0.0) : (a + b == 2.0));
                                }
                            }
                            double f = on ? 1.0 : 0.0;
                            if (sp.rbf_dim[cc] >= 0) {
                                const double dd = xc[(cc * CS) * RG + t] - xc[(cc * CS) * RG + t2];
                                const double d2 = dd * dd;
                                f = on ? exp_neg(-d2 * hil2[sp.ls_idx[cc]], etab) : 0.0;
                                g1ls[k] += gB * f * d2;
                            }
                            g1os[k] += gB * f;
                        }
                    }
                }
            }
        }
    }

    // ---- CTA epilogue: per-thread accumulators -> fixed-order partials of this CTA ------------------------------------------------------
    double* part = ws + w.part + ((size_t)chunk * L + l) * w.stride;
    {
        const double a_ = warp_sum(accA);
        if (lane == 0) hyp[wid * (nh + 1)] = a_;
#pragma unroll
        for (int cc = 0; cc < NC0; ++cc) {
            const double s1 = warp_sum(gos[cc]);
            const double s2 = warp_sum(gls[cc]);
            if (lane == 0) {
                hyp[wid * (nh + 1) + 1 + sp.n_ls + cc] += s1;
                if (sp.rbf_dim[cc] >= 0) hyp[wid * (nh + 1) + 1 + sp.ls_idx[cc]] += s2 * osc[cc] * il3[sp.ls_idx[cc]];
            }
        }
#pragma unroll
        for (int k = 0; k < NC1; ++k) {
            const int cc = NC0 + k;
            const double s1 = warp_sum(g1os[k]);
            const double s2 = warp_sum(g1ls[k]);
            if (lane == 0) {
                hyp[wid * (nh + 1) + 1 + sp.n_ls + cc] += s1;
                if (sp.rbf_dim[cc] >= 0) hyp[wid * (nh + 1) + 1 + sp.ls_idx[cc]] += s2 * osc[cc] * il3[sp.ls_idx[cc]];
            }
        }
        const double n_ = warp_sum(gno);
        if (lane == 0) hyp[wid * (nh + 1) + nh] += n_;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double a2 = ng1acc[e], b2 = daacc[e];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) { a2 += __shfl_xor_sync(0xffffffffu, a2, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o); }
            if (g == 0) { cols[(0 * 16 + wid) * 8 + 2 * q + e] = a2; cols[(1 * 16 + wid) * 8 + 2 * q + e] = b2; }
        }
        if (set == 1) {
            set_barrier(set);                                // every warp of set 1 is done with its B1 | B2
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (wl + 8 * i < 36) {
                    double* d = sxch + (wl + 8 * i) * 64 + g * 8 + 2 * q;
                    d[0] = sacc[i][0]; d[1] = sacc[i][1];
                }
            }
        }
    }
    __syncthreads();
    if (set == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (wl + 8 * i < 36) {
                const double* d = sxch + (wl + 8 * i) * 64 + g * 8 + 2 * q;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ii = so_i[i] + g, jj = so_j[i] + 2 * q + e;
                    if (ii < M && jj < M) {
                        const double v = sacc[i][e] + d[e];
                        part[stats_off_S() + (size_t)ii * M + jj] = v;
                        if (so_i[i] != so_j[i]) part[stats_off_S() + (size_t)jj * M + ii] = v;
                    }
                }
            }
        }
    }
    if (tid < 64 && tid < M) {
        const int ntile = tid >> 3, cidx = tid & 7;
        part[stats_off_ng1(M) + tid] = cols[(0 * 16 + ntile) * 8 + cidx] + cols[(0 * 16 + ntile + 8) * 8 + cidx];
        part[stats_off_da(M) + tid] = cols[(1 * 16 + ntile) * 8 + cidx] + cols[(1 * 16 + ntile + 8) * 8 + cidx];
    }
    if (tid <= nh) {
        double s = 0.0;
        for (int ww = 0; ww < 16; ++ww) s += hyp[ww * (nh + 1) + tid];
        if (tid == 0) {
            for (int k = 0; k < LVAE_NSCAL; ++k) part[stats_off_scal(M) + k] = 0.0;
            part[stats_off_scal(M) + SC_A] = s;
        } else {
            part[stats_off_hyp(M) + tid - 1] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// group planning: one CTA per chunk; greedy packing of whole subjects into groups of <= RG rows and <= 5 subjects
// entry: row0, R, nsub, end_1..end_5 (local row offsets where each subject ends; unused = R)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_plan_groups(const int32_t* __restrict__ offsets, int P_b, int per, int gstride,
                                                     int* __restrict__ gtab, int* __restrict__ gcount) {
    __shared__ int offs[1025];
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const int p_begin = min(P_b, chunk * per), p_end = min(P_b, p_begin + per);
    int* tab = gtab + (size_t)chunk * gstride * GT;
    int ng = 0;
    int cur_row0 = 0, cur_rows = 0, cur_n = 0, ends[5];
    for (int base = p_begin; base < p_end; base += 1024) {
        const int n = min(1024, p_end - base);
        __syncthreads();
        for (int i = tid; i <= n; i += 256) offs[i] = offsets[base + i];
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < n; ++i) {
                const int T = offs[i + 1] - offs[i];
                if (cur_n > 0 && (cur_rows + T > RG || cur_n == 5)) {
                    int* e = tab + (size_t)ng * GT;
                    e[0] = cur_row0; e[1] = cur_rows; e[2] = cur_n;
                    for (int s = 0; s < 5; ++s) e[3 + s] = s < cur_n ? ends[s] : cur_rows;
                    ++ng;
                    cur_n = 0; cur_rows = 0;
                }
                if (cur_n == 0) cur_row0 = offs[i];
                cur_rows += T;
                ends[cur_n++] = cur_rows;
            }
        }
    }
    if (tid == 0) {
        if (cur_n > 0) {
            int* e = tab + (size_t)ng * GT;
            e[0] = cur_row0; e[1] = cur_rows; e[2] = cur_n;
            for (int s = 0; s < 5; ++s) e[3 + s] = s < cur_n ? ends[s] : cur_rows;
            ++ng;
        }
        gcount[chunk] = ng;
    }
}

template <int NC0, int NC1>
int launch2(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    int nr = 0;
    for (int cc = 0; cc < sp.n0; ++cc) nr += sp.rbf_dim[cc] >= 0;
    const size_t smem = sizeof(double) * (common_doubles(NC0, NC0 + NC1, w.nh) + 2 * set_doubles(NC0 + NC1, nr));
    static SmemAttrCache attr;
    if (int rc_ = lvae_ensure_smem(k_subjects_fused2<NC0, NC1>, smem, attr)) return rc_;
    k_subjects_fused2<NC0, NC1><<<dim3(w.nchunk, p->L), 512, smem, st>>>(sp, w, p->L, p->M, p->Q, p->N_b, w.TP, nr, p->x, p->mu,
                                                                        p->z, p->lengthscale, p->outputscale,
                                                                        0.5 * p->scale, p->d_mu, p->workspace);
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

}  // namespace

bool lvae_fused2_supported(const lvae_kld_problem_t* p) {
    if (!(p->M <= 64 && p->T_max <= RG && p->ks.n_comp0 >= 1 && p->ks.n_comp0 <= 4 && p->ks.n_comp1 >= 1 &&
          p->ks.n_comp1 <= 2 && p->ks.spec))
        return false;
    int nr = 0;                                  // SE-bearing K0 components keep their values in shared memory
    for (int cc = 0; cc < p->ks.n_comp0; ++cc) nr += p->ks.spec[(size_t)cc * LVAE_SPEC_STRIDE] >= 0;
    return nr <= 3;
}

int lvae_plan_groups_launch(const lvae_kld_problem_t* p, const KldLayout& w, cudaStream_t st) {
    const int per = (p->P_b + w.nchunk - 1) / w.nchunk;
    k_plan_groups<<<w.nchunk, 256, 0, st>>>(p->offsets, p->P_b, per, w.gstride,
                                            reinterpret_cast<int*>(p->workspace + w.gtab),
                                            reinterpret_cast<int*>(p->workspace + w.gcount));
    LVAE_COUNT_LAUNCH();
    return lvae_cuda_rc(cudaGetLastError());
}

int lvae_subjects_fused2_launch(const lvae_kld_problem_t* p, const DevSpec& sp, const KldLayout& w, cudaStream_t st) {
    if (!lvae_fused2_supported(p)) return LVAE_E_TOO_LARGE;
    const int key = sp.n0 * 10 + sp.n1;
    switch (key) {
        case 11: return launch2<1, 1>(p, sp, w, st);
        case 12: return launch2<1, 2>(p, sp, w, st);
        case 21: return launch2<2, 1>(p, sp, w, st);
        case 22: return launch2<2, 2>(p, sp, w, st);
        case 31: return launch2<3, 1>(p, sp, w, st);
        case 32: return launch2<3, 2>(p, sp, w, st);
        case 41: return launch2<4, 1>(p, sp, w, st);
        case 42: return launch2<4, 2>(p, sp, w, st);
    }
    return LVAE_E_BADARG;
}
