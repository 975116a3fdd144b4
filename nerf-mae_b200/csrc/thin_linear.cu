// Linear layers with a 4-wide output (the 1x1x1 output convolution C/2 -> 4, unetr_block.py:96-116): HBM-bound streaming
// kernels instead of a 128x64 GEMM tile that would be 94 % padding.  K (input features) is a multiple of 4, <= 128.
#include "kernels.cuh"

#define TL_ROWS 256

// y[m][0..3] = x[m][:] . w[0..3][:] + b        one thread per row, weights broadcast from shared memory
__global__ void __launch_bounds__(256) thin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, long long M, int K, float* __restrict__ y) {
    __shared__ float4 sw[128];  // sw[k] = {w[0][k], w[1][k], w[2][k], w[3][k]}
    for (int k = threadIdx.x; k < K; k += blockDim.x) sw[k] = make_float4(w[k], w[K + k], w[2 * K + k], w[3 * K + k]);
    __syncthreads();
    const float4 b4 = bias ? make_float4(bias[0], bias[1], bias[2], bias[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const float4* xr = reinterpret_cast<const float4*>(x + m * K);
        float4 acc = b4;
        for (int j = 0; j < K / 4; j++) {
            const float4 v = __ldg(xr + j);
            const float4 w0 = sw[4 * j], w1 = sw[4 * j + 1], w2 = sw[4 * j + 2], w3 = sw[4 * j + 3];
            acc.x += v.x * w0.x + v.y * w1.x + v.z * w2.x + v.w * w3.x;
            acc.y += v.x * w0.y + v.y * w1.y + v.z * w2.y + v.w * w3.y;
            acc.z += v.x * w0.z + v.y * w1.z + v.z * w2.z + v.w * w3.z;
            acc.w += v.x * w0.w + v.y * w1.w + v.z * w2.w + v.w * w3.w;
        }
        reinterpret_cast<float4*>(y)[m] = acc;
    }
}

// dx[m][k] (+)= sum_n dy[m][n] * w[n][k]      flat float4 units, coalesced stores
__global__ void __launch_bounds__(256) thin_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, long long M, int K,
                                                         int accumulate, float* __restrict__ dx) {
    __shared__ float sw[4 * 128];
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int upr = K / 4;  // float4 units per row
    const long long units = M * upr;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
        const long long m = u / upr;
        const int j = (int)(u - m * upr);
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy) + m);
        float4 o;
        o.x = g.x * sw[4 * j] + g.y * sw[K + 4 * j] + g.z * sw[2 * K + 4 * j] + g.w * sw[3 * K + 4 * j];
        o.y = g.x * sw[4 * j + 1] + g.y * sw[K + 4 * j + 1] + g.z * sw[2 * K + 4 * j + 1] + g.w * sw[3 * K + 4 * j + 1];
        o.z = g.x * sw[4 * j + 2] + g.y * sw[K + 4 * j + 2] + g.z * sw[2 * K + 4 * j + 2] + g.w * sw[3 * K + 4 * j + 2];
        o.w = g.x * sw[4 * j + 3] + g.y * sw[K + 4 * j + 3] + g.z * sw[2 * K + 4 * j + 3] + g.w * sw[3 * K + 4 * j + 3];
        float4* d = reinterpret_cast<float4*>(dx) + u;
        if (accumulate) {
            const float4 old = *d;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *d = o;
    }
}

// dw[n][k] += sum_m dy[m][n] * x[m][k] ; db[n] += sum_m dy[m][n].  Tiles of 256 rows are staged in shared memory with
// coalesced float4 loads; thread t < 4K owns output (n = t / K, k = t % K) and keeps it in a register across tiles.
__global__ void __launch_bounds__(512) thin_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long M, int K,
                                                         float* __restrict__ dw, float* __restrict__ db) {
    extern __shared__ float sm[];
    float* xs = sm;                    // [TL_ROWS][K]
    float* ds = sm + TL_ROWS * K;      // [TL_ROWS][4]
    const int t = threadIdx.x;
    const int n = t / K, k = t - n * K;
    const bool owner = t < 4 * K;
    float acc = 0.f, accb = 0.f;
    const long long tiles = (M + TL_ROWS - 1) / TL_ROWS;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long m0 = tile * TL_ROWS;
        const int rows = (int)min((long long)TL_ROWS, M - m0);
        const float4* xsrc = reinterpret_cast<const float4*>(x + m0 * K);
        for (int u = t; u < rows * (K / 4); u += blockDim.x) reinterpret_cast<float4*>(xs)[u] = __ldg(xsrc + u);
        const float4* dsrc = reinterpret_cast<const float4*>(dy + m0 * 4);
        for (int u = t; u < rows; u += blockDim.x) reinterpret_cast<float4*>(ds)[u] = __ldg(dsrc + u);
        __syncthreads();
        if (owner) {
#pragma unroll 8
            for (int r = 0; r < rows; r++) acc = fmaf(ds[r * 4 + n], xs[r * K + k], acc);
            if (k == 0)
                for (int r = 0; r < rows; r++) accb += ds[r * 4 + n];
        }
        __syncthreads();
    }
    if (owner) {
        atomicAdd(dw + n * K + k, acc);
        if (k == 0 && db) atomicAdd(db + n, accb);
    }
}

bool k_thin_supported(int N, int K) { return N == 4 && K % 4 == 0 && K >= 4 && K <= 128; }

int k_thin_fwd(const float* x, const float* w, const float* bias, long long M, int K, float* y, cudaStream_t st) {
    int g = (int)min((long long)148 * 8, (M + 255) / 256);
    thin_fwd_kernel<<<g, 256, 0, st>>>(x, w, bias, M, K, y);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_thin_dgrad(const float* dy, const float* w, long long M, int K, int accumulate, float* dx, cudaStream_t st) {
    long long units = M * (K / 4);
    int g = (int)min((long long)148 * 16, (units + 255) / 256);
    thin_dgrad_kernel<<<g, 256, 0, st>>>(dy, w, M, K, accumulate, dx);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// dw [4][K] and db [4] (may be NULL) are overwritten
int k_thin_wgrad(const float* x, const float* dy, long long M, int K, float* dw, float* db, cudaStream_t st) {
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 4 * K, st));
    if (db) NMAE_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 4, st));
    size_t smem = sizeof(float) * TL_ROWS * (K + 4);
    static bool attr_set[64] = {false};
    int dev;
    NMAE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(thin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * TL_ROWS * 132)));
        attr_set[dev] = true;
    }
    long long tiles = (M + TL_ROWS - 1) / TL_ROWS;
    int g = (int)min((long long)148 * 3, tiles);
    thin_wgrad_kernel<<<g, 512, smem, st>>>(x, dy, M, K, dw, db);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
