// Micro-benchmark (sm_100a): do DMMA (mma.sync m8n8k4 f64) and DFMA share one FP64 pipe, and what does a DMMA cost when its
// operands come from shared memory?  Decides the layout of the fused subject kernel (lvae_subjects_fused3.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ND DMMA + NF DFMA per iteration, independent accumulators
template <int ND, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b) {
    double c[8][2], f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; f[j] = j + 0.5; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ND; ++j) dmma(c[j & 7][0], c[j & 7][1], a, b);
#pragma unroll
        for (int j = 0; j < NF; ++j) f[j & 7] = fma(f[j & 7], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8 DMMA per iteration with NA A-fragments and NB B-fragments loaded from shared memory per iteration (LDS.64 each),
// the register tile is NA x NB (NA * NB = 8)
template <int NA, int NB>
__global__ void __launch_bounds__(256) k_dmma_lds(double* out, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = 1e-3 * i;
    __syncthreads();
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        double a[NA], b[NB];
#pragma unroll
        for (int j = 0; j < NA; ++j) a[j] = sm[(lane + 32 * j + i) & 2047];
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = sm[(lane + 32 * (j + NA) + 7 * i) & 2047];
#pragma unroll
        for (int ja = 0; ja < NA; ++ja)
#pragma unroll
            for (int jb = 0; jb < NB; ++jb) dmma(c[ja * NB + jb][0], c[ja * NB + jb][1], a[ja], b[jb]);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// table-driven exp for x <= 0 (the library's exp_neg), table in shared memory
__global__ void __launch_bounds__(256) k_expneg(double* out, int iters, double x0) {
    __shared__ double tbl[64];
    if (threadIdx.x < 64) tbl[threadIdx.x] = exp2(threadIdx.x / 64.0);
    __syncthreads();
    double x = x0 - 1e-3 * threadIdx.x, s = 0;
    for (int i = 0; i < iters; ++i) {
        const double MAGIC = 6755399441055744.0;
        const double t = fma(x, 92.33248261689366, MAGIC);
        const int k = __double2loint(t);
        const double kd = t - MAGIC;
        double r = fma(kd, -0.010830424493178725, x);
        r = fma(kd, -2.030704202170295e-10, r);
        double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
        q = fma(q, r, 1.6666666666666666e-01);
        q = fma(q, r, 0.5);
        q = fma(q, r, 1.0);
        const double tj = tbl[k & 63];
        const double v = fma(tj, q * r, tj);
        const int hi = __double2hiint(v) + ((k >> 6) << 20);
        s += __hiloint2double(hi, __double2loint(v));
        x -= 1e-6;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s sms %d\n", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 4000;
    for (int bps : {2, 4}) {
        const int grid = sms * bps;
        const double thr = double(grid) * 256, warps = double(grid) * 8;
        auto rep = [&](const char* name, float ms, int nd, int nf) {
            const double dm = warps * iters * nd * 512.0 / ms * 1e-9, df = thr * iters * nf * 2.0 / ms * 1e-9;
            printf("%-28s blocks/SM %d : %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", name, bps, ms, dm, df, dm + df);
        };
        rep("8 dmma", time_ms([&] { k_mix<8, 0><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 8, 0);
        rep("8 dfma", time_ms([&] { k_mix<0, 8><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 0, 8);
        rep("8 dmma + 8 dfma", time_ms([&] { k_mix<8, 8><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 8, 8);
        rep("8 dmma + 16 dfma", time_ms([&] { k_mix<8, 16><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 8, 16);
        rep("8 dmma + 32 dfma", time_ms([&] { k_mix<8, 32><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 8, 32);
        rep("4 dmma + 32 dfma", time_ms([&] { k_mix<4, 32><<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5), 4, 32);
        rep("8 dmma, lds 8A+1B (9/8)", time_ms([&] { k_dmma_lds<8, 1><<<grid, 256>>>(out, iters); }, 5), 8, 0);
        rep("8 dmma, lds 4A+2B (6/8)", time_ms([&] { k_dmma_lds<4, 2><<<grid, 256>>>(out, iters); }, 5), 8, 0);
        rep("8 dmma, lds 2A+4B (6/8)", time_ms([&] { k_dmma_lds<2, 4><<<grid, 256>>>(out, iters); }, 5), 8, 0);
        rep("8 dmma, lds 1A+8B (9/8)", time_ms([&] { k_dmma_lds<1, 8><<<grid, 256>>>(out, iters); }, 5), 8, 0);
        const float ms = time_ms([&] { k_expneg<<<grid, 256>>>(out, iters, -0.5); }, 5);
        printf("%-28s blocks/SM %d : %.3f ms  %.1f Gexp/s\n", "exp_neg (table)", bps, ms, thr * iters / ms * 1e-6);
    }
    CK(cudaDeviceSynchronize());
    cudaFree(out);
    return 0;
}
