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
    int nacc;     // number of TMEM accumulators the MMAs rotate over (1 = every MMA depends on the previous one)
    int random;   // 1: random operand data instead of zeros
    int mixed;    // 1: the conv3 forward pattern, N then N/2 into the same accumulator
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(Cfg c, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) {
        // random bf16 pairs in (-2, 2) (exponent field <= 127), or zeros: data toggling decides the power draw
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = c.random ? (h & 0xBFFFBFFFu) : 0u;
    }
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
        const uint32_t idesc = idesc_bf16(128, c.N, c.a_mn, c.b_mn), idesc_half = idesc_bf16(128, c.N / 2, c.a_mn, c.b_mn);
        const uint32_t a0 = smem_u32(smem) >> 4, b0 = (smem_u32(smem) + 96 * 1024) >> 4;
        const uint64_t swz = c.swz ? ((uint64_t)2 << 61) : 0;
        long long t0 = 0;
        int ph[2] = {0, 0};
        const int mixed = c.mixed, amask = c.nacc - 1, dstride = c.mixed ? 128 : 512 / c.nacc;
        for (int it = 0; it < c.iters; it++) {
            if (it == 2) t0 = clock64();
            if (elect_one()) {
                for (int j = 0; j < c.batch; j++) {
                    const uint32_t k = (uint32_t)(j & (c.swz ? 3 : 7));
                    const uint64_t da = smem_desc((a0 + k * c.a_step16) << 4, c.a_lbo, c.a_sbo) | swz;
                    const uint64_t db = smem_desc((b0 + k * c.b_step16) << 4, c.b_lbo, c.b_sbo) | swz;
                    const uint32_t dcol = (uint32_t)(((j >> mixed) & amask) * dstride);
                    mma_bf16(tmem_base + dcol, da, db, (c.mixed && (j & 1)) ? idesc_half : idesc, 1);
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
    for (int mode = 0; mode < 5; mode++) {
        for (int n : Ns) {
            Cfg c;
            memset(&c, 0, sizeof(c));
            c.N = n; c.batch = 64; c.iters = 200;
            c.a_lbo = 198 * 16; c.a_sbo = 128; c.b_lbo = n * 16; c.b_sbo = 128; c.a_step16 = 1; c.b_step16 = (n * 16 * 2) >> 4;
            const char* name;
            if (mode == 0) { name = "K-major, 1 accumulator (dependent) "; c.nacc = 1; }
            else if (mode == 1) { name = "K-major, 2 accumulators            "; c.nacc = 2; }
            else if (mode == 2) { name = "K-major, 4 accumulators            "; c.nacc = 4; if (n > 128) continue; }
            else if (mode == 3) { name = "mixed N,N/2 -> 1 accumulator       "; c.nacc = 1; c.mixed = 1; if (n != 96) continue; }
            else { name = "mixed N,N/2 -> 2 accumulators      "; c.nacc = 2; c.mixed = 1; if (n != 96) continue; }
            mma_bench_kernel<<<148, 128, 200 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, n, cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; i++) avg += (double)h[i];
            avg /= 148;
            const double per = avg / ((double)(c.iters - 2) * c.batch);
            printf("%s N=%3d: %7.1f cycles/MMA  (tensor floor %5.1f)\n", name, n, per, c.mixed ? 0.75 * n / 2.0 : n / 2.0);
        }
    }
    // effective SM clock under sustained tensor load: cycles (clock64) / wall time (CUDA events) of a ~50 ms run
    for (int n : {96, 144, 256}) {
        Cfg c;
        memset(&c, 0, sizeof(c));
        c.N = n; c.batch = 64; c.iters = 12000; c.nacc = 1; c.random = 1;
        c.a_lbo = 198 * 16; c.a_sbo = 128; c.b_lbo = n * 16; c.b_sbo = 128; c.a_step16 = 1; c.b_step16 = (n * 16 * 2) >> 4;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            mma_bench_kernel<<<148, 128, 200 * 1024>>>(c, d);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; i++) avg += (double)h[i];
            avg /= 148;
            printf("sustained N=%3d rep %d: %.1f ms, %.1f Mcycles -> %.0f MHz effective, %.1f cycles/MMA, %.0f TFLOP/s\n", n, rep, ms, avg / 1e6,
                   avg / ms / 1e3, avg / ((double)(c.iters - 2) * c.batch), 148.0 * c.iters * c.batch * 2.0 * 128 * n * 16 / ms / 1e9);
        }
    }
    return 0;
}
