// Micro-benchmark: FP64 issue-rate ceilings on sm_100a (B200).
// Measures (a) DFMA vector peak, (b) DMMA mma.sync f64 peak for the shapes ptxas accepts,
// (c) FP64 exp() throughput.  Used once to choose the inner-product engine for the
// L-VAE subject-pass kernel; numbers are recorded in profiles/.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef HAVE_M16
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double av, double bv) {
    double c[4][4]; double a[8], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = av + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = bv + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma16816(c[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double av, double bv) {
    double c[8][4]; double a[2] = {av, av + 1};
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma1684(c[j], a, bv);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double x0) {
    double x = x0 - 1e-3 * threadIdx.x, s = 0;
    for (int i = 0; i < iters; ++i) { s += exp(x); x -= 1e-6; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DFMA with one LDS.64 per 2 DFMA (does LDS issue steal FP64 slots?)
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters, double a) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i * 1e-3;
    __syncthreads();
    double c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    int idx = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        double b0 = sm[(idx + i) & 1023], b1 = sm[(idx + i + 32) & 1023], b2 = sm[(idx + i + 64) & 1023], b3 = sm[(idx + i + 96) & 1023];
        c0 = fma(c0, a, b0); c1 = fma(c1, a, b1); c2 = fma(c2, a, b2); c3 = fma(c3, a, b3);
        c4 = fma(c4, a, b0); c5 = fma(c5, a, b1); c6 = fma(c6, a, b2); c7 = fma(c7, a, b3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
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
    int sms = p.multiProcessorCount;
    printf("device %s sms %d cc %d.%d\n", p.name, sms, p.major, p.minor);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 20000;
    for (int bps : {1, 2, 4, 8}) {
        int grid = sms * bps; double thr = double(grid) * 256;
        float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 0.999, 1e-3); }, 5);
        printf("dfma        blocks/SM %d : %.2f TFLOP/s\n", bps, thr * iters * 8 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<<<grid, 256>>>(out, iters / 4, 0.999, 1e-3); }, 5);
        printf("dmma m8n8k4 blocks/SM %d : %.2f TFLOP/s\n", bps, double(grid) * 8 * (iters / 4) * 8 * (8 * 8 * 4 * 2.0) / ms * 1e-9);
#ifdef HAVE_M16
        ms = time_ms([&] { k_dmma1684<<<grid, 256>>>(out, iters / 4, 0.999, 1e-3); }, 5);
        printf("dmma m16n8k4 blocks/SM %d : %.2f TFLOP/s\n", bps, double(grid) * 8 * (iters / 4) * 8 * (16 * 8 * 4 * 2.0) / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<<<grid, 256>>>(out, iters / 8, 0.999, 1e-3); }, 5);
        printf("dmma m16n8k16 blocks/SM %d : %.2f TFLOP/s\n", bps, double(grid) * 8 * (iters / 8) * 4 * (16 * 8 * 16 * 2.0) / ms * 1e-9);
#endif
        ms = time_ms([&] { k_exp<<<grid, 256>>>(out, iters / 8, -0.5); }, 5);
        printf("exp(double) blocks/SM %d : %.2f Gexp/s\n", bps, thr * (iters / 8) / ms * 1e-6);
        ms = time_ms([&] { k_dfma_lds<<<grid, 256>>>(out, iters, 0.999); }, 5);
        printf("dfma+lds    blocks/SM %d : %.2f TFLOP/s\n", bps, thr * iters * 8 * 2 / ms * 1e-9);
    }
    CK(cudaDeviceSynchronize());
    cudaFree(out);
    return 0;
}
