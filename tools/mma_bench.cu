// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, both operands in shared memory) as a function of N and of
// the operand layouts (K-major / MN-major, no swizzle / 128B swizzle).  One CTA per SM, one issuing thread, batches of MMAs
// committed to alternating mbarriers so the tensor pipe never drains.  Operand data is zeros: only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu -I nerf-mae_b200/csrc
#include <cstdio>
#include <cstdlib>

#include "tc.cuh"

using namespace tc;

void nmae_set_error(const char*, ...) {}

struct Cfg {
    int N, a_mn, b_mn, swz;   // swz: 0 none, 1 = 128B swizzle (K-major only)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_step16, b_step16;   // per-MMA start-address step (16-byte units), cycling over 8 positions
    int batch, iters;
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(Cfg c, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16(128, c.N, c.a_mn, c.b_mn);
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 96 * 1024) >> 4;
        const uint64_t swz = c.swz ? ((uint64_t)2 << 61) : 0;
        long long t0 = 0;
        int ph[2] = {0, 0};
        for (int it = 0; it < c.iters; it++) {
            if (it == 2) t0 = clock64();
            if (elect_one()) {
                for (int j = 0; j < c.batch; j++) {
                    const uint32_t k = (uint32_t)(j & (c.swz ? 3 : 7));
                    const uint64_t da = smem_desc((a0 + k * c.a_step16) << 4, c.a_lbo, c.a_sbo) | swz;
                    const uint64_t db = smem_desc((b0 + k * c.b_step16) << 4, c.b_lbo, c.b_sbo) | swz;
                    mma_bf16(tmem_base + (uint32_t)((j & 1) * 256), da, db, idesc, 1);
                }
                mma_commit(smem_u32(&bars[it & 1]));
            }
            __syncwarp();
            if (it >= 1) {
                const int b = (it - 1) & 1;
                mbar_wait(smem_u32(&bars[b]), ph[b]);
                ph[b] ^= 1;
            }
        }
        {
            const int b = (c.iters - 1) & 1;
            mbar_wait(smem_u32(&bars[b]), ph[b]);
        }
        long long t1 = clock64();
        if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Named { const char* name; Cfg c; };
    const int Ns[] = {48, 96, 144, 256};
    for (int layout = 0; layout < 4; layout++) {
        for (int n : Ns) {
            Cfg c;
            memset(&c, 0, sizeof(c));
            c.N = n; c.batch = 64; c.iters = 200;
            const char* name;
            if (layout == 0) {          // K-major, no swizzle: [chunk][row][8]; A 198-row images, tap = +16 B
                name = "K-major  none ";
                c.a_lbo = 198 * 16; c.a_sbo = 128; c.b_lbo = n * 16; c.b_sbo = 128; c.a_step16 = 1; c.b_step16 = (n * 16 * 2) >> 4;
            } else if (layout == 1) {   // MN-major, no swizzle: SBO = chunk stride, LBO = 128 B; k-step = +256 B
                name = "MN-major none ";
                c.a_mn = c.b_mn = 1;
                c.a_lbo = 128; c.a_sbo = 130 * 16; c.b_lbo = 128; c.b_sbo = 128 * 16; c.a_step16 = 16; c.b_step16 = 16;
            } else if (layout == 2) {   // K-major, 128B swizzle: rows of 128 B, 8-row groups of 1024 B; k-step = +32 B
                name = "K-major  sw128";
                c.swz = 1;
                c.a_lbo = 16; c.a_sbo = 1024; c.b_lbo = 16; c.b_sbo = 1024; c.a_step16 = 2; c.b_step16 = 2;
                c.batch = 64;
            } else {                    // A MN-major (M=128 from 16 chunks), B MN-major with N chunks: the wgrad "swap" shape
                name = "MN-major A=Y  ";
                c.a_mn = c.b_mn = 1;
                c.a_lbo = 128; c.a_sbo = 128 * 16; c.b_lbo = 128; c.b_sbo = 130 * 16; c.a_step16 = 16; c.b_step16 = 16;
            }
            if (layout == 2) { /* only 4 distinct 32-byte k-steps inside a 128 B swizzle row */ }
            mma_bench_kernel<<<148, 128, 200 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, n, cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; i++) avg += (double)h[i];
            avg /= 148;
            const double per = avg / ((double)(c.iters - 2) * c.batch);
            const double bytes = 128 * 16 * 2 + n * 16 * 2;
            printf("%s N=%3d: %7.1f cycles/MMA  (tensor floor %5.1f, smem operand bytes %5.0f -> %5.1f B/clk)\n", name, n, per, n / 2.0, bytes,
                   bytes / per);
        }
    }
    return 0;
}
