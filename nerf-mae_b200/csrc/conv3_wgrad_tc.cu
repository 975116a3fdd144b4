// 3x3x3 convolution weight gradient on tcgen05:  dW[co][ci][tap] = sum_voxels dY[v][co] * X[v + tap][ci].
//
// The reduction (K) dimension is the voxel position, which is NOT the contiguous dimension of the channels-last
// tensors, so both operands are fed as MN-major UMMA operands.  The shared-memory image format of the forward
// kernel, [8-channel chunk][position][8 x bf16], is exactly the canonical no-swizzle MN-major layout (SBO = chunk
// stride, LBO = 128 B between 8-position groups), and a dz tap is again a +16 B descriptor start offset.
//
// Per stage (128 consecutive positions of one (batch,x) plane, one dy):
//   A = X images of the three dx planes stacked along M: 18 chunks = 144 rows (dx, ci)   [130 positions: dz halo]
//   B = dY image, 48 output channels (N = 48)                                              [128 positions]
//   for dz in 0..2, for 16-position k-step, 3 MMAs (hi*hi, hi*lo, lo*hi) with M=128 on chunks 0..15 and again on
//   chunks 2..17 (second accumulator; only its rows 112..127 = chunks 16,17 are used).
// A CTA owns one (48-channel input group, 48-channel output tile, dy) accumulator set (3 dz x 2 x 48 = 288 TMEM
// columns) over a contiguous range of position tiles, then flushes it with fp32 atomics into dW (Cout,Cin,3,3,3).
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"

using namespace tc;

#define CG 48
#define KCH 6
#define TILE_K 128   // positions per stage
#define NTW 48       // output channels per accumulator
#define N_PROD 512
#define PW (N_PROD / 32)
#define XROWS 130
#define XCHUNKS 18
#define MAXU_W 13   // float4 units per producer thread and stage: (3*130 + 128)*12 = 6216 <= 13*512

struct WgradTcParams {
    const float* x;
    const float* dy;
    float* dw;
    int B, Dx, Dy, Dz, C, N;
    int ZP, P, tpp, num_chunks, n_cg, n_nt, n_ident, splits, num_items;
    int dbg;  // NMAE_DBG experiments: 1 no loads, 2 no MMAs, 16 no smem stores
};

#define X_PART_BYTES (XCHUNKS * XROWS * 16)  // 37440
#define Y_PART_BYTES (KCH * TILE_K * 16)     // 12288
#define STAGE_BYTES (2 * X_PART_BYTES + 2 * Y_PART_BYTES)

__device__ __forceinline__ void store_row_split(const float4* v, uint8_t* hi_base, uint8_t* lo_base, uint32_t chunk_stride, int row) {
#pragma unroll
    for (int c = 0; c < KCH; c++) {
        uint4 h, l;
        split2(v[2 * c].x, v[2 * c].y, h.x, l.x);
        split2(v[2 * c].z, v[2 * c].w, h.y, l.y);
        split2(v[2 * c + 1].x, v[2 * c + 1].y, h.z, l.z);
        split2(v[2 * c + 1].z, v[2 * c + 1].w, h.w, l.w);
        *reinterpret_cast<uint4*>(hi_base + (size_t)c * chunk_stride + (size_t)row * 16) = h;
        *reinterpret_cast<uint4*>(lo_base + (size_t)c * chunk_stride + (size_t)row * 16) = l;
    }
}

__global__ void __launch_bounds__(N_PROD + 160, 1) conv3_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES);
    const uint32_t bar0 = smem_u32(bars);
    auto ST_FULL = [&](int s) { return bar0 + 8u * s; };
    auto ST_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
    const uint32_t ACC_FULL = bar0 + 8u * 4, ACC_EMPTY = bar0 + 8u * 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(ST_FULL(s), PW);
            mbar_init(ST_EMPTY(s), 1);
        }
        mbar_init(ACC_FULL, 1);
        mbar_init(ACC_EMPTY, 4);
        fence_barrier_init();
    }
    if (warp == PW) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    // item -> (identity = (cg, nt, dy), chunk range)
    auto item_decode = [&](int item, int& cg, int& nt, int& dyi, int& c_beg, int& c_end) {
        int ident = item / p.splits, sp = item - ident * p.splits;
        dyi = ident % 3;
        int r = ident / 3;
        nt = r % p.n_nt;
        cg = r / p.n_nt;
        c_beg = (int)((long long)p.num_chunks * sp / p.splits);
        c_end = (int)((long long)p.num_chunks * (sp + 1) / p.splits);
    };

    if (warp < PW) {
        // =========================================================== producers
        int s = 0, ph = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            for (int ch = c_beg; ch < c_end; ch++) {
                const int p0 = (ch % p.tpp) * TILE_K;
                const int xq = (ch / p.tpp) % p.Dx, b = ch / (p.tpp * p.Dx);
                // flat float4 units: [0, 3*130*12) X images (plane, row, j), then 128*12 dY units; all loads are issued
                // before waiting for the stage buffer
                constexpr int X_UNITS = 3 * XROWS * 12, UNITS = X_UNITS + TILE_K * 12;
                const int ybase = p0 / p.ZP, zbase = p0 - ybase * p.ZP;   // one division per stage
                float4 v[MAXU_W];
#pragma unroll
                for (int t = 0; t < MAXU_W; t++) {
                    const int u = tid + t * N_PROD;
                    v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u < UNITS && !(p.dbg & 1)) {
                        int row, j, xx, ld, c0, yy, zz;
                        const float* base;
                        if (u < X_UNITS) {
                            const int dx = u / (XROWS * 12), r = u - dx * (XROWS * 12);
                            row = r / 12; j = r - row * 12;
                            xx = xq + dx - 1;
                            yy = ybase + dyi - 1; zz = zbase - 1 + row;
                            base = p.x; ld = p.C; c0 = cg * CG;
                        } else {
                            const int r = u - X_UNITS;
                            row = r / 12; j = r - row * 12;
                            xx = xq;
                            yy = ybase; zz = zbase + row;
                            base = p.dy; ld = p.N; c0 = nt * NTW;
                        }
                        if (zz < 0) { zz += p.ZP; yy--; }
                        while (zz >= p.ZP) { zz -= p.ZP; yy++; }
                        if (xx >= 0 && xx < p.Dx && yy >= 0 && yy < p.Dy && zz >= 1 && zz <= p.Dz)
                            v[t] = __ldg(reinterpret_cast<const float4*>(
                                             base + ((((long long)(b * p.Dx + xx) * p.Dy + yy) * p.Dz + (zz - 1)) * ld + c0)) + j);
                    }
                }
                mbar_wait_warp(ST_EMPTY(s), ph ^ 1);
                uint8_t* xh = smem + (size_t)s * STAGE_BYTES;
                uint8_t* xl = xh + X_PART_BYTES;
                uint8_t* yh = xl + X_PART_BYTES;
                uint8_t* yl = yh + Y_PART_BYTES;
#pragma unroll
                for (int t = 0; t < MAXU_W; t++) {
                    const int u = tid + t * N_PROD;
                    if (u < UNITS && !(p.dbg & 16)) {
                        uint8_t *dh, *dl;
                        if (u < X_UNITS) {
                            const int dx = u / (XROWS * 12), r = u - dx * (XROWS * 12);
                            const int row = r / 12, j = r - row * 12;
                            const uint32_t off = (uint32_t)(dx * KCH + (j >> 1)) * (XROWS * 16) + (uint32_t)row * 16u + (uint32_t)(j & 1) * 8u;
                            dh = xh + off; dl = xl + off;
                        } else {
                            const int r = u - X_UNITS;
                            const int row = r / 12, j = r - row * 12;
                            const uint32_t off = (uint32_t)(j >> 1) * (TILE_K * 16) + (uint32_t)row * 16u + (uint32_t)(j & 1) * 8u;
                            dh = yh + off; dl = yl + off;
                        }
                        uint2 h, l;
                        split2(v[t].x, v[t].y, h.x, l.x);
                        split2(v[t].z, v[t].w, h.y, l.y);
                        *reinterpret_cast<uint2*>(dh) = h;
                        *reinterpret_cast<uint2*>(dl) = l;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(ST_FULL(s));
                if (++s == 2) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == PW) {
        // =========================================================== MMA issuer (whole warp converged, one elected lane issues)
        {
            const uint32_t idesc = idesc_bf16(128, NTW, 1, 1);
            // MN-major operands: SBO = chunk stride (8-channel groups), LBO = 128 B (8-position groups)
            const uint32_t x_hi = desc_hi(XROWS * 16), y_hi = desc_hi(TILE_K * 16), lbo = (128u >> 4) << 16;
            int s = 0, ph = 0, it = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
                int cg, nt, dyi, c_beg, c_end;
                item_decode(item, cg, nt, dyi, c_beg, c_end);
                mbar_wait(ACC_EMPTY, (it & 1) ^ 1);
                fence_after_sync();
                for (int ch = c_beg; ch < c_end; ch++) {
                    const uint32_t first = ch == c_beg ? 0u : 1u;
                    mbar_wait(ST_FULL(s), ph);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t xh16 = (smem0 + (uint32_t)s * STAGE_BYTES) >> 4, xl16 = xh16 + (X_PART_BYTES >> 4);
                        const uint32_t yh16 = xl16 + (X_PART_BYTES >> 4), yl16 = yh16 + (Y_PART_BYTES >> 4);
                        if (!(p.dbg & 2)) {
#pragma unroll 1
                            for (int dz = 0; dz < 3; dz++) {
#pragma unroll
                                for (int ks = 0; ks < TILE_K / 16; ks++) {
                                    const uint32_t xo = (uint32_t)(dz + 16 * ks), yo = (uint32_t)(16 * ks);
                                    const uint64_t byh = desc_make(y_hi, lbo, yh16 + yo), byl = desc_make(y_hi, lbo, yl16 + yo);
#pragma unroll
                                    for (int half = 0; half < 2; half++) {
                                        const uint32_t co = (uint32_t)half * 2u * XROWS;   // second MMA starts two chunks further
                                        const uint64_t axh = desc_make(x_hi, lbo, xh16 + xo + co), axl = desc_make(x_hi, lbo, xl16 + xo + co);
                                        const uint32_t d = tmem_base + (uint32_t)((dz * 2 + half) * NTW);
                                        mma_bf16(d, axh, byh, idesc, ks == 0 ? first : 1u);
                                        mma_bf16(d, axh, byl, idesc, 1);
                                        mma_bf16(d, axl, byh, idesc, 1);
                                    }
                                }
                            }
                        }
                        mma_commit(ST_EMPTY(s));
                        if (ch == c_end - 1) mma_commit(ACC_FULL);
                    }
                    __syncwarp();
                    if (++s == 2) { s = 0; ph ^= 1; }
                }
                if (c_end <= c_beg) {
                    if (elect_one()) mma_commit(ACC_FULL);
                    __syncwarp();
                }
            }
        }
    } else {
        // =========================================================== epilogue: TMEM -> atomics into dW
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            mbar_wait_warp(ACC_FULL, it & 1);
            fence_after_sync();
            for (int dz = 0; dz < 3; dz++) {
                for (int half = 0; half < 2; half++) {
                    // accumulator `half` holds stacked chunks [2*half, 2*half+16); use rows of chunks 0..15 (half 0) / 16,17 (half 1)
                    const int chunk = row / 8 + 2 * half;
                    const bool use = (c_end > c_beg) && (half == 0 ? true : chunk >= 16);
                    const int dx = chunk / KCH, kc = chunk - dx * KCH;
                    const int ci = cg * CG + kc * 8 + (row & 7);
                    const int tap = dx * 9 + dyi * 3 + dz;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((dz * 2 + half) * NTW);
                    for (int j = 0; j < NTW / 16; j++) {
                        float v[16];
                        tmem_ld16(taddr + j * 16, v);
                        if (use && chunk < XCHUNKS) {
#pragma unroll
                            for (int e = 0; e < 16; e++) {
                                const int co = nt * NTW + j * 16 + e;
                                atomicAdd(p.dw + ((long long)co * p.C + ci) * 27 + tap, v[e]);
                            }
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
    if (warp == PW) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

bool k_conv3_wgrad_tc_supported(int C, int N) { return C % CG == 0 && N % NTW == 0; }

// dw (N=Cout, C=Cin, 27) is overwritten
int k_conv3_wgrad_tc(const float* x, const float* dy, int B, int Dx, int Dy, int Dz, int C, int N, float* dw, cudaStream_t st) {
    NMAE_CHECK_ARG(k_conv3_wgrad_tc_supported(C, N), "conv3_wgrad_tc: unsupported channels C=%d N=%d", C, N);
    WgradTcParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.dy = dy; p.dw = dw;
    p.B = B; p.Dx = Dx; p.Dy = Dy; p.Dz = Dz; p.C = C; p.N = N;
    p.ZP = Dz + 2;
    p.P = Dy * p.ZP;
    p.tpp = cdiv(p.P, TILE_K);
    p.num_chunks = B * Dx * p.tpp;
    p.n_cg = C / CG;
    p.n_nt = N / NTW;
    p.n_ident = p.n_cg * p.n_nt * 3;
    { const char* d = getenv("NMAE_DBG"); p.dbg = d ? atoi(d) : 0; }
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    p.splits = max(1, min(p.num_chunks, (2 * sms + p.n_ident - 1) / p.n_ident));
    if (p.n_ident >= sms) p.splits = 1;
    p.num_items = p.n_ident * p.splits;
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * (size_t)C * N, st));
    const int smem = 2 * STAGE_BYTES + 64;
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(conv3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    conv3_wgrad_tc_kernel<<<min(sms, p.num_items), N_PROD + 160, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
