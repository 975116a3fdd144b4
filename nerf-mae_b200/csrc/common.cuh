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

// Exact-erf GELU (S:352-358 nn.GELU()) and its derivative, branch-free: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) from Abramowitz-Stegun 26.2.17
// (|error| < 7.5e-8, the fp32 rounding level of the erff form) - one reciprocal, one exp2 (shared with the density term of the
// derivative) and six FMAs instead of erff's two divergent branches plus expf: the GELU epilogues of the tcgen05 linears are bound by
// the arithmetic of their four epilogue warps.
__device__ __forceinline__ void gelu_phi(float x, float& Phi, float& E) {
    const float ax = fabsf(x);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(x * x * -0.72134752044448170368f));      // exp(-x^2 / 2)
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2316419f, ax, 1.f)));
    float p = 1.061405429f;
    p = fmaf(p, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float y = 0.5f * p * t * E;          // 1 - Phi(|x|)
    Phi = x >= 0.f ? 1.f - y : y;
}
__device__ __forceinline__ float gelu_erf(float x) {
    float Phi, E;
    gelu_phi(x, Phi, E);
    return x * Phi;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float Phi, E;
    gelu_phi(x, Phi, E);
    return fmaf(x * 0.39894228040143267794f, E, Phi);     // Phi(x) + x phi(x)
}
