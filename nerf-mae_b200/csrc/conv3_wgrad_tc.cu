// 3x3x3 convolution weight gradient on tcgen05:  dW[co][ci][tap] = sum_voxels dY[v][co] * X[v + tap][ci].
//
// Both operands come from the pre-built bf16 hi/lo image tensors of uimg.cuh (the X image the forward convolution consumed
// and the dY image its dgrad consumes), so staging is pure cp.async.bulk.  The reduction (K) dimension is the voxel
// position, which is the ROW dimension of an image ([8-channel chunk][position][8 x bf16]); that format is exactly the
// canonical no-swizzle MN-major UMMA layout (SBO = chunk stride, LBO = 128 B between 8-position groups).
//
// Per stage = (128-position tile of one (batch, x, z-strip) plane, one dy):
//   A = the dY tile, hi chunks then lo chunks stacked along M: rows 0..47 = dY_hi, 48..95 = dY_lo (96..127 zeros)
//   B = the X rows [p0 + dy*ZP - 1, +130) of the three dx planes stacked along N: 18 chunks = 144 columns (dx, ci)
//   for dz in 0..2 (a dz tap is a +16 B start offset of B), for each 16-position k-step:
//       D[dz] += A x B_hi ;  D[dz] += A x B_lo                       (M=128, N=144, K=16)
//   i.e. all four hi/lo products with two instructions that run at the tensor-pipe floor (72 cycles; measured 75.6 with
//   tools/mma_bench.cu - an N=48 instruction costs 50 cycles because the 4 KB A operand fetch is not amortised).
//   The epilogue adds rows r and r+48.
// The dY image is the "type X" image built for the dgrad (halo columns carry neighbours): the MMA warp zeroes the halo
// rows of the staged tile before issuing, so that every voxel is counted exactly once.
// A CTA owns one (48-channel input group, 48-channel output tile, dy) accumulator set (3 dz x 144 = 432 TMEM columns) over
// a contiguous range of tiles, then flushes it with fp32 atomics into dW (Cout,Cin,3,3,3).
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"
#include "uimg.cuh"

using namespace tc;

#define CG 48
#define KCH 6
#define TILE_K 128   // positions per stage
#define XROWS 130
#define XCHUNKS 18
#define NCOL (XCHUNKS * 8)   // 144

#define Y_CHUNK_BYTES (TILE_K * 16)              // 2048
#define X_CHUNK_BYTES (XROWS * 16)               // 2080
#define Y_BYTES (2 * KCH * Y_CHUNK_BYTES)        // 24576: [hi 6 chunks][lo 6 chunks]
#define X_PART_BYTES (XCHUNKS * X_CHUNK_BYTES)   // 37440: [dx][chunk]
#define Y_PAD_BYTES (4 * Y_CHUNK_BYTES)           // rows 96..127 of the M=128 A operand: kept zero (idle multipliers draw less power)
#define X_OFF (Y_BYTES + Y_PAD_BYTES)
#define LOAD_BYTES (Y_BYTES + 2 * X_PART_BYTES)  // 99456 bytes arrive per stage
#define STAGE_BYTES (X_OFF + 2 * X_PART_BYTES)   // 107648
#define N_STAGES 2

struct WgradTcParams {
    const uint8_t* ximg;
    const uint8_t* yimg;
    long long x_chunk_bytes, x_part_bytes, x_img_bytes, y_chunk_bytes, y_part_bytes, y_img_bytes;
    float* dw;
    int B, Dx, C, N;
    int n_strips, ZP, tpp, H, num_tiles, n_cg, n_nt, n_ident, splits, num_items;
    int dbg;  // NMAE_DBG experiments: 1 no image copies, 2 no MMAs
};

__device__ long long g_conv3_wgrad_cycles[256];   // NMAE_DBG bit 128: cycles the MMA warp of each CTA spent in its main loop

__global__ void __launch_bounds__(256, 1) conv3_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + N_STAGES * STAGE_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto ST_FULL = [&](int s) { return bar0 + 8u * s; };
    auto ST_EMPTY = [&](int s) { return bar0 + 8u * (N_STAGES + s); };
    const uint32_t ACC_FULL = bar0 + 8u * (2 * N_STAGES), ACC_EMPTY = bar0 + 8u * (2 * N_STAGES + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * N_STAGES + 2);

    if (tid == 0) {
        for (int s = 0; s < N_STAGES; s++) {
            mbar_init(ST_FULL(s), 1);
            mbar_init(ST_EMPTY(s), 1);
        }
        mbar_init(ACC_FULL, 1);
        mbar_init(ACC_EMPTY, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int s = 0; s < N_STAGES; s++)
        for (int i = tid; i < Y_PAD_BYTES / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem + (size_t)s * STAGE_BYTES + Y_BYTES)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    // item -> (identity = (cg, nt, dy), tile range); the three dy identities of one range run on neighbouring CTAs at the
    // same time, so the images they share are fetched from DRAM once
    auto item_decode = [&](int item, int& cg, int& nt, int& dyi, int& c_beg, int& c_end) {
        const int sp = item / p.n_ident, ident = item - sp * p.n_ident;
        dyi = ident % 3;
        const int r = ident / 3;
        nt = r % p.n_nt;
        cg = r / p.n_nt;
        c_beg = (int)((long long)p.num_tiles * sp / p.splits);
        c_end = (int)((long long)p.num_tiles * (sp + 1) / p.splits);
    };

    if (warp == 0) {
        // =========================================================== image loader: 12 + 36 bulk copies per stage
        int s = 0, ph = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            for (int ch = c_beg; ch < c_end; ch++) {
                // tile order: x fastest, so that consecutive stages of a CTA share two of their three X planes (L2 hits)
                const int xq = ch % p.Dx;
                const int p0 = ((ch / p.Dx) % p.tpp) * TILE_K;
                const int strip = (ch / (p.Dx * p.tpp)) % p.n_strips, b = ch / (p.Dx * p.tpp * p.n_strips);
                mbar_wait(ST_EMPTY(s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t dst = smem0 + (uint32_t)s * STAGE_BYTES;
                    if (p.dbg & 1) {
                        mbar_arrive(ST_FULL(s));
                    } else {
                        mbar_expect_tx(ST_FULL(s), LOAD_BYTES);
                        const uint8_t* ysrc = p.yimg + ((((long long)(b * (p.Dx + 2) + xq + 1) * p.n_strips + strip) * p.n_nt + nt)) * p.y_img_bytes +
                                              (long long)(p0 + p.H) * 16;
#pragma unroll
                        for (int part = 0; part < 2; part++)
#pragma unroll
                            for (int c = 0; c < KCH; c++)
                                bulk_g2s(dst + (uint32_t)(part * KCH + c) * Y_CHUNK_BYTES, ysrc + part * p.y_part_bytes + c * p.y_chunk_bytes,
                                         Y_CHUNK_BYTES, ST_FULL(s));
                        // X rows [p0 + dy*ZP - 1, +130) in position space = image rows [p0 + dy*ZP, +130)   (H = ZP + 1)
                        const long long xrow = (long long)(p0 + dyi * p.ZP) * 16;
#pragma unroll 1
                        for (int dx = 0; dx < 3; dx++) {
                            const uint8_t* xsrc = p.ximg + ((((long long)(b * (p.Dx + 2) + xq + dx) * p.n_strips + strip) * p.n_cg + cg)) * p.x_img_bytes + xrow;
#pragma unroll
                            for (int part = 0; part < 2; part++)
#pragma unroll
                                for (int c = 0; c < KCH; c++)
                                    bulk_g2s(dst + X_OFF + (uint32_t)part * X_PART_BYTES + (uint32_t)(dx * KCH + c) * X_CHUNK_BYTES,
                                             xsrc + part * p.x_part_bytes + c * p.x_chunk_bytes, X_CHUNK_BYTES, ST_FULL(s));
                        }
                    }
                }
                __syncwarp();
                if (++s == N_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (whole warp converged, one elected lane issues)
        const uint32_t idesc = idesc_bf16(128, (p.dbg & 64) ? 16 : NCOL, 1, 1);   // NMAE_DBG bit 64: N=16 instructions (timing experiment)
        // MN-major operands: SBO = chunk stride (8-channel groups), LBO = 128 B (8-position groups)
        const uint32_t y_hi = desc_hi(Y_CHUNK_BYTES), x_hi = desc_hi(X_CHUNK_BYTES), lbo = (128u >> 4) << 16;
        int s = 0, ph = 0, it = 0;
        const long long t_begin = clock64();
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            mbar_wait(ACC_EMPTY, (it & 1) ^ 1);
            fence_after_sync();
            for (int ch = c_beg; ch < c_end; ch++) {
                const uint32_t first = ch == c_beg ? 0u : 1u;
                const int p0 = ((ch / p.Dx) % p.tpp) * TILE_K;
                mbar_wait(ST_FULL(s), ph);
                // zero the halo rows (zz == 0 or zz == ZP-1) of the dY tile: they duplicate voxels of the neighbouring strips
                if (!(p.dbg & 1)) {
                    uint8_t* ys = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < TILE_K / 32; k++) {
                        const int i = lane + 32 * k;
                        const int zz = (p0 + i) % p.ZP;
                        if (zz == 0 || zz == p.ZP - 1) {
#pragma unroll
                            for (int c = 0; c < 2 * KCH; c++)
                                *reinterpret_cast<uint4*>(ys + (size_t)c * Y_CHUNK_BYTES + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t y16 = (smem0 + (uint32_t)s * STAGE_BYTES) >> 4;
                    const uint32_t xh16 = y16 + (X_OFF >> 4), xl16 = xh16 + (X_PART_BYTES >> 4);
                    if (!(p.dbg & 2)) {
#pragma unroll 1
                        for (int dz = 0; dz < 3; dz++) {
                            const uint32_t d = tmem_base + (uint32_t)(dz * NCOL);
#pragma unroll
                            for (int ks = 0; ks < TILE_K / 16; ks++) {
                                const uint64_t ay = desc_make(y_hi, lbo, y16 + (uint32_t)(16 * ks));
                                const uint32_t xo = (uint32_t)(dz + 16 * ks);
                                mma_bf16(d, ay, desc_make(x_hi, lbo, xh16 + xo), idesc, ks == 0 ? first : 1u);
                                mma_bf16(d, ay, desc_make(x_hi, lbo, xl16 + xo), idesc, 1);
                            }
                        }
                    }
                    mma_commit(ST_EMPTY(s));
                    if (ch == c_end - 1) mma_commit(ACC_FULL);
                }
                __syncwarp();
                if (++s == N_STAGES) { s = 0; ph ^= 1; }
            }
            if (c_end <= c_beg) {
                if (elect_one()) mma_commit(ACC_FULL);
                __syncwarp();
            }
        }
        if ((p.dbg & 128) && lane == 0 && blockIdx.x < 256) g_conv3_wgrad_cycles[blockIdx.x] = clock64() - t_begin;
    } else if (warp >= 4) {
        // =========================================================== epilogue: TMEM -> atomics into dW
        const int q = warp & 3;
        const int row = q * 32 + lane;          // rows 0..47: dY_hi products, 48..95: dY_lo products, 96..127: unused
        int it = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            mbar_wait_warp(ACC_FULL, it & 1);
            fence_after_sync();
            if (q < 3 && c_end > c_beg && !(p.dbg & 2)) {
                const int co = nt * CG + row % CG;
                float* dw_row = p.dw + (long long)co * p.C * 27;
                for (int dz = 0; dz < 3; dz++) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dz * NCOL);
                    for (int j = 0; j < NCOL / 16; j++) {
                        float v[16];
                        tmem_ld16(taddr + j * 16, v);
                        const int dx = j / 3, ci0 = cg * CG + (j - dx * 3) * 16;     // 16 columns never straddle a dx plane
                        const int tap = dx * 9 + dyi * 3 + dz;
                        if (row < 2 * CG) {
#pragma unroll
                            for (int e = 0; e < 16; e++) atomicAdd(dw_row + (ci0 + e) * 27 + tap, v[e]);
                        }
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(ACC_EMPTY);
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

bool k_conv3_wgrad_tc_supported(int C, int N) { return C % CG == 0 && N % CG == 0; }

// ximg: type-X image of the convolution input (C channels); yimg: type-X image of the output gradient (N channels).
// dw (N=Cout, C=Cin, 27) is overwritten.
int k_conv3_wgrad_tc(const void* ximg, const void* yimg, int B, int Dx, int Dy, int Dz, int C, int N, float* dw, cudaStream_t st) {
    NMAE_CHECK_ARG(k_conv3_wgrad_tc_supported(C, N), "conv3_wgrad_tc: unsupported channels C=%d N=%d", C, N);
    const UImgGeom gx = uimg_geom(B, Dx, Dy, Dz, C), gy = uimg_geom(B, Dx, Dy, Dz, N);
    WgradTcParams p;
    memset(&p, 0, sizeof(p));
    p.ximg = reinterpret_cast<const uint8_t*>(ximg);
    p.yimg = reinterpret_cast<const uint8_t*>(yimg);
    p.x_chunk_bytes = gx.chunk_bytes; p.x_part_bytes = gx.part_bytes; p.x_img_bytes = gx.img_bytes;
    p.y_chunk_bytes = gy.chunk_bytes; p.y_part_bytes = gy.part_bytes; p.y_img_bytes = gy.img_bytes;
    p.dw = dw;
    p.B = B; p.Dx = Dx; p.C = C; p.N = N;
    p.n_strips = gx.n_strips; p.ZP = gx.ZP; p.tpp = gx.tpp; p.H = gx.H;
    p.num_tiles = B * Dx * gx.n_strips * gx.tpp;
    p.n_cg = C / CG;
    p.n_nt = N / CG;
    p.n_ident = p.n_cg * p.n_nt * 3;
    p.dbg = nmae_debug_mask();
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // two items per CTA, never a third: rounding the split count UP gave 297 items on 148 CTAs at decoder1, i.e. one CTA with
    // three items and a kernel 1.5x longer than its average CTA (ncu: sm__cycles_elapsed.max 19.7 M vs 12.8 M in the MMA loop)
    p.splits = max(1, min(p.num_tiles, (2 * sms) / p.n_ident));
    if (p.n_ident >= sms) p.splits = 1;
    p.num_items = p.n_ident * p.splits;
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * (size_t)C * N, st));
    const int smem = N_STAGES * STAGE_BYTES + 128;
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(conv3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    conv3_wgrad_tc_kernel<<<min(sms, p.num_items), 256, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    if (p.dbg & 128) {
        long long h[256];
        const int grid = min(sms, p.num_items);
        NMAE_CUDA(cudaStreamSynchronize(st));
        NMAE_CUDA(cudaMemcpyFromSymbol(h, g_conv3_wgrad_cycles, sizeof(h)));
        double avg = 0;
        for (int i = 0; i < grid; i++) avg += (double)h[i];
        fprintf(stderr, "[NMAE_DBG] conv3_wgrad_tc: %.3f Mcycles in the MMA loop (avg over %d CTAs), %d tiles x %d identities\n", avg / grid / 1e6,
                grid, p.num_tiles, p.n_ident);
    }
    return NMAE_OK;
}
