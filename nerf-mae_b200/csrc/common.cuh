// Shared helpers for the nmae sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define NMAE_OK 0
#define NMAE_ERR_ARG -1
#define NMAE_ERR_CUDA -2

void nmae_set_error(const char* fmt, ...);
// Bottleneck-experiment mask for the tensor-core kernels (skip copies / MMAs / stores: results are garbage under it).  Always 0
// in the product build; only a library compiled with -DNMAE_DBG (tools/ experiments) reads the NMAE_DBG environment variable.
int nmae_debug_mask(void);

#define NMAE_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            nmae_set_error(__VA_ARGS__);          \
            return NMAE_ERR_ARG;                  \
        }                                         \
    } while (0)

#define NMAE_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            nmae_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return NMAE_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

extern unsigned long long g_nmae_launches;  // kernels launched by this library (bench.py reports it)
#define NMAE_LAUNCH_CHECK()                   \
    do {                                      \
        ++g_nmae_launches;                    \
        NMAE_CUDA(cudaPeekAtLastError());     \
    } while (0)

// every entry point takes the device explicitly: autograd runs backward on its own
// worker thread, so thread-local "current device" state cannot be relied on.
#define NMAE_SET_DEVICE(dev) NMAE_CUDA(cudaSetDevice(dev))

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum, result valid in every thread; `sh` must hold >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (w == 0) {
        r = warp_sum(r);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    return sh[0];
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float c = 0.39894228040143267794f;  // 1/sqrt(2 pi)
    return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) + x * c * expf(-0.5f * x * x);
}
