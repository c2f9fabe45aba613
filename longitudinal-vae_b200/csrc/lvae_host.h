// Host-side helpers shared by the translation units of liblvae_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "lvae_common.cuh"

static inline int lvae_cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : -(int)e; }
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: the largest value set so far is cached
// per device, so a process that drives several GPUs sets it on each of them.
#define LVAE_MAX_DEVICES 16
struct SmemAttrCache { size_t v[LVAE_MAX_DEVICES] = {}; };
template <class K>
static inline int lvae_ensure_smem(K kernel, size_t bytes, SmemAttrCache& c) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return lvae_cuda_rc(e);
    const bool cached = dev >= 0 && dev < LVAE_MAX_DEVICES;
    if (!cached || bytes > c.v[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return lvae_cuda_rc(e);
        if (cached) c.v[dev] = bytes;
    }
    return 0;
}
int lvae_make_devspec(const lvae_kernel_spec_t* ks, int Q, DevSpec* out);
int lvae_block_offsets(const int32_t* offsets, int P_b, int64_t* off2, cudaStream_t st);
// Stream-ordered scratch for the stand-alone ABI entry points (cudaMallocAsync from the device's default pool).  On first
// use per device the pool's release threshold is raised to 2 GiB: with the default of 0 every synchronisation (e.g. the
// Cholesky info check) hands the scratch back to the driver and the next call pays a fresh allocation — milliseconds for
// the 100 MB a batched inverse of 32 000 blocks needs.  Free with cudaFreeAsync.
cudaError_t lvae_scratch_alloc(void** p, size_t bytes, cudaStream_t st);

// optional per-phase device timing (bench.py's roofline: duration of the dominant kernel measured live with CUDA events
// on the launch stream).  Phases: 0 head, 1 prep, 2 subjects, 3 reduce, 4 tail, 5 ng_step.
#define LVAE_NPHASE 6
void lvae_prof_begin(int phase, cudaStream_t st);
void lvae_prof_end(int phase, cudaStream_t st);
