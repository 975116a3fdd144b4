// Single-pass 3x3x3 convolution weight gradient on tcgen05 with fp16 operands ("fp16" precision mode; conv3_wgrad_tc.cu is the
// three-pass bf16 hi/lo form):  dW[co][ci][tap] = sum_voxels dY[v][co] * X[v + tap][ci].
//
// Both operands are the fp16 "H" images (uimg.cuh) the forward convolution and its dgrad consume; the reduction dimension is
// the voxel position = the ROW dimension of an image, i.e. the canonical no-swizzle MN-major UMMA layout.
//
// Per stage = (128-position tile of one (batch, x, z-strip) plane, one dy):
//   A (M = 128) = the dY tile TWICE along M: rows [0,CG) = dY[p], rows [CG,2CG) = dY[p+1] (a second bulk copy of the same image,
//                 one row further); with CG = 48 the remaining 32 rows are zero
//   B (N = 3*CG) = the X rows of the three dx planes stacked along N
//   for o in {1,2} (row offset of B = dz' + 1), for each 16-position k-step:  D[o] += A x B(o)        (M=128, N=3*CG, K=16)
//   Row block s of D[o] is the tap dz = (o-1) - s:  D[1] holds dz = 0 (s=0) and dz = -1 (s=1), D[2] holds dz = +1 (s=0) and a
//   duplicate of dz = 0 (s=1, ignored): two instructions per k-step produce the three dz taps (an M = 48 operand would leave 5/8 of
//   the tensor pipe's rows idle and need three).
// The dY image is a type X image (halo columns carry neighbours): a dedicated warp zeroes the halo rows of both staged copies (so
// that every voxel is counted once) while the MMA warp is still issuing the previous stage.  dY images are stored scaled by a power of two (uimg_h.cu); the epilogue multiplies by its
// reciprocal.  A CTA owns one (input group, output tile, dy) accumulator pair over a range of tiles and flushes it with fp32 atomics.
#include "kernels.cuh"
#include "tc.cuh"
#include "uimg.cuh"

using namespace tc;

#define WH_TILE_K 128    // positions per stage (64-position stages in a 6-deep ring were slower: 3.7 vs 2.85 ms - the per-stage
#define WH_XROWS 130     // barrier / commit / 30-bulk-copy overhead dominates, not the L2 latency)
#define WH_MAXS 6
#define WH_Y_CHUNK (WH_TILE_K * 16)
#define WH_X_CHUNK (WH_XROWS * 16)

struct WgradHParams {
    const uint8_t* ximg;
    const uint8_t* yimg;
    long long x_chunk_bytes, x_img_bytes, y_chunk_bytes, y_img_bytes;
    const float* inv_scale;
    float* dw;
    int B, Dx, C, N;
    int n_strips, ZP, tpp, H, num_tiles, n_cg, n_nt, n_ident, splits, num_items, n_stages;
};

__host__ __device__ constexpr uint32_t idesc_f16_mn(int M, int N) {   // fp16 operands, fp32 accumulate, A and B MN-major
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CG>
__global__ void __launch_bounds__(256, 1) conv3_wgrad_h_kernel(const __grid_constant__ WgradHParams p) {
    constexpr int KCH = CG / 8, XCH = 3 * KCH, NCOL = 3 * CG;
    constexpr int Y_BYTES = 16 * WH_Y_CHUNK;                 // 16 chunks = 128 rows of the A operand (2*KCH used)
    constexpr int X_BYTES = XCH * WH_X_CHUNK;
    constexpr int STAGE = Y_BYTES + X_BYTES;
    constexpr int LOAD_BYTES = 2 * KCH * WH_Y_CHUNK + X_BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = p.n_stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * STAGE);
    const uint32_t bar0 = smem_u32(bars);
    auto ST_FULL = [&](int s) { return bar0 + 8u * s; };
    auto ST_EMPTY = [&](int s) { return bar0 + 8u * (WH_MAXS + s); };
    auto ST_READY = [&](int s) { return bar0 + 8u * (2 * WH_MAXS + s); };     // halo rows of the stage zeroed: operands final
    const uint32_t ACC_FULL = bar0 + 8u * (3 * WH_MAXS), ACC_EMPTY = bar0 + 8u * (3 * WH_MAXS + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WH_MAXS + 2);

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(ST_FULL(s), 1); mbar_init(ST_EMPTY(s), 1); mbar_init(ST_READY(s), 1); }
        mbar_init(ACC_FULL, 1);
        mbar_init(ACC_EMPTY, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    if (2 * KCH < 16) {          // rows 2*CG..127 of the A operand stay zero
        for (int s = 0; s < NS; s++)
            for (int i = tid; i < (16 - 2 * KCH) * WH_Y_CHUNK / 16; i += blockDim.x)
                reinterpret_cast<uint4*>(smem + (size_t)s * STAGE + 2 * KCH * WH_Y_CHUNK)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

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
        // =========================================================== image loader
        int s = 0, ph = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            for (int ch = c_beg; ch < c_end; ch++) {
                // tile order: x fastest, so that consecutive stages of a CTA share two of their three X planes (L2 hits)
                const int xq = ch % p.Dx;
                const int p0 = ((ch / p.Dx) % p.tpp) * WH_TILE_K;
                const int strip = (ch / (p.Dx * p.tpp)) % p.n_strips, b = ch / (p.Dx * p.tpp * p.n_strips);
                mbar_wait(ST_EMPTY(s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t dst = smem0 + (uint32_t)s * STAGE;
                    mbar_expect_tx(ST_FULL(s), LOAD_BYTES);
                    const uint8_t* ysrc = p.yimg + ((((long long)(b * (p.Dx + 2) + xq + 1) * p.n_strips + strip) * p.n_nt + nt)) * p.y_img_bytes +
                                          (long long)(p0 + p.H) * 16;
#pragma unroll
                    for (int sh = 0; sh < 2; sh++)
#pragma unroll
                        for (int c = 0; c < KCH; c++)
                            bulk_g2s(dst + (uint32_t)(sh * KCH + c) * WH_Y_CHUNK, ysrc + sh * 16 + c * p.y_chunk_bytes, WH_Y_CHUNK, ST_FULL(s));
                    // X rows [p0 + (dy-1)*ZP - 1, +130) in position space = image rows [p0 + dy*ZP, +130)   (H = ZP + 1)
                    const long long xrow = (long long)(p0 + dyi * p.ZP) * 16;
#pragma unroll 1
                    for (int dx = 0; dx < 3; dx++) {
                        const uint8_t* xsrc = p.ximg + ((((long long)(b * (p.Dx + 2) + xq + dx) * p.n_strips + strip) * p.n_cg + cg)) * p.x_img_bytes + xrow;
#pragma unroll
                        for (int c = 0; c < KCH; c++)
                            bulk_g2s(dst + Y_BYTES + (uint32_t)(dx * KCH + c) * WH_X_CHUNK, xsrc + c * p.x_chunk_bytes, WH_X_CHUNK, ST_FULL(s));
                    }
                }
                __syncwarp();
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (whole warp converged, one elected lane issues)
        const uint32_t idesc = idesc_f16_mn(128, NCOL);
        // MN-major operands: SBO = chunk stride (8-channel groups), LBO = 128 B (8-position groups)
        const uint32_t y_hi = desc_hi(WH_Y_CHUNK), x_hi = desc_hi(WH_X_CHUNK), lbo = (128u >> 4) << 16;
        int s = 0, ph = 0, it = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            mbar_wait(ACC_EMPTY, (it & 1) ^ 1);
            fence_after_sync();
            for (int ch = c_beg; ch < c_end; ch++) {
                const uint32_t first = ch == c_beg ? 0u : 1u;
                mbar_wait(ST_READY(s), ph);
                fence_after_sync();
                if (elect_one()) {
                    const uint32_t y16 = (smem0 + (uint32_t)s * STAGE) >> 4;
                    const uint32_t x16 = y16 + (Y_BYTES >> 4);
#pragma unroll 1
                    for (int o = 1; o <= 2; o++) {
                        const uint32_t d = tmem_base + (uint32_t)((o - 1) * 256);
#pragma unroll
                        for (int ks = 0; ks < WH_TILE_K / 16; ks++) {
                            mma_bf16(d, desc_make(y_hi, lbo, y16 + (uint32_t)(16 * ks)), desc_make(x_hi, lbo, x16 + (uint32_t)(o + 16 * ks)), idesc,
                                     ks == 0 ? first : 1u);
                        }
                    }
                    mma_commit(ST_EMPTY(s));
                    if (ch == c_end - 1) mma_commit(ACC_FULL);
                }
                __syncwarp();
                if (++s == NS) { s = 0; ph ^= 1; }
            }
            if (c_end <= c_beg) {
                if (elect_one()) mma_commit(ACC_FULL);
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // =========================================================== halo warp: the dY image is a type X image (halo columns carry
        // the neighbouring strips' voxels); zero the halo rows (zz == 0 or zz == ZP-1) of both staged copies so that every voxel
        // is counted once, while the MMA warp is still issuing the previous stage
        int s = 0, ph = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            for (int ch = c_beg; ch < c_end; ch++) {
                const int p0 = ((ch / p.Dx) % p.tpp) * WH_TILE_K;
                mbar_wait_warp(ST_FULL(s), ph);
                uint8_t* ys = smem + (size_t)s * STAGE;
#pragma unroll
                for (int k = 0; k < WH_TILE_K / 32; k++) {
                    const int i = lane + 32 * k;
                    const int zz = (p0 + i) % p.ZP;
                    const bool h0 = zz == 0 || zz == p.ZP - 1;                 // copy 0: row i is position p0 + i
                    const bool h1 = zz == p.ZP - 1 || zz == p.ZP - 2;          // copy 1: row i is position p0 + i + 1
#pragma unroll
                    for (int c = 0; c < KCH; c++) {
                        if (h0) *reinterpret_cast<uint4*>(ys + (size_t)c * WH_Y_CHUNK + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
                        if (h1) *reinterpret_cast<uint4*>(ys + (size_t)(KCH + c) * WH_Y_CHUNK + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(ST_READY(s));
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue: TMEM -> atomics into dW
        const int q = warp & 3;
        const int row = q * 32 + lane;          // rows [0,CG): shift 0, rows [CG,2CG): shift 1
        const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
        int it = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int cg, nt, dyi, c_beg, c_end;
            item_decode(item, cg, nt, dyi, c_beg, c_end);
            mbar_wait_warp(ACC_FULL, it & 1);
            fence_after_sync();
            if (q * 32 < 2 * CG && c_end > c_beg) {
                const int sft = row >= CG ? 1 : 0;
                const int co = nt * CG + row - sft * CG;
                float* dw_row = p.dw + (long long)co * p.C * 27;
                for (int o = 1; o <= 2; o++) {
                    const int dz = o - sft;                       // tap index 0..2 (dz = -1, 0, +1)
                    const bool use = row < 2 * CG && !(o == 2 && sft == 1);    // (o=2, s=1) duplicates dz = 0
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((o - 1) * 256);
                    for (int j = 0; j < NCOL / 16; j++) {
                        float v[16];
                        tmem_ld16(taddr + j * 16, v);
                        const int dx = (j * 16) / CG, ci0 = cg * CG + (j * 16 - dx * CG);     // 16 columns never straddle a dx plane
                        const int tap = dx * 9 + dyi * 3 + dz;
                        if (use) {
#pragma unroll
                            for (int e = 0; e < 16; e++) atomicAdd(dw_row + (ci0 + e) * 27 + tap, v[e] * inv);
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

bool k_conv3_wgrad_h_supported(int C, int N) {
    const int cg = uimg_h_cg(C);
    return cg != 0 && cg == uimg_h_cg(N);
}

template <int CG>
static int wgrad_h_launch(WgradHParams& p, cudaStream_t st) {
    constexpr int KCH = CG / 8;
    constexpr int STAGE = 16 * WH_Y_CHUNK + 3 * KCH * WH_X_CHUNK;
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // two items per CTA, never a third (see conv3_wgrad_tc.cu)
    p.splits = max(1, min(p.num_tiles, (2 * sms) / p.n_ident));
    if (p.n_ident >= sms) p.splits = 1;
    p.num_items = p.n_ident * p.splits;
    p.n_stages = min(WH_MAXS, (227 * 1024 - 256) / STAGE);
    const int smem = p.n_stages * STAGE + 256;
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(conv3_wgrad_h_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    conv3_wgrad_h_kernel<CG><<<min(sms, p.num_items), 256, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ximg: H image of the convolution input (C channels); yimg: H image of the output gradient (N channels), scaled by 1 / *inv_scale
// (inv_scale NULL: unscaled).  dw (N=Cout, C=Cin, 27) is overwritten.
int k_conv3_wgrad_h(const void* ximg, const void* yimg, const float* inv_scale, int B, int Dx, int Dy, int Dz, int C, int N, float* dw,
                    cudaStream_t st) {
    NMAE_CHECK_ARG(k_conv3_wgrad_h_supported(C, N), "conv3_wgrad_h: unsupported channels C=%d N=%d", C, N);
    const UImgGeom gx = uimg_geom_h(B, Dx, Dy, Dz, C), gy = uimg_geom_h(B, Dx, Dy, Dz, N);
    WgradHParams p;
    memset(&p, 0, sizeof(p));
    p.ximg = reinterpret_cast<const uint8_t*>(ximg);
    p.yimg = reinterpret_cast<const uint8_t*>(yimg);
    p.x_chunk_bytes = gx.chunk_bytes; p.x_img_bytes = gx.img_bytes;
    p.y_chunk_bytes = gy.chunk_bytes; p.y_img_bytes = gy.img_bytes;
    p.inv_scale = inv_scale;
    p.dw = dw;
    p.B = B; p.Dx = Dx; p.C = C; p.N = N;
    p.n_strips = gx.n_strips; p.ZP = gx.ZP; p.H = gx.H;
    p.tpp = (gx.P + WH_TILE_K - 1) / WH_TILE_K;
    p.num_tiles = B * Dx * gx.n_strips * p.tpp;
    p.n_cg = gx.n_cg;
    p.n_nt = gy.n_cg;
    p.n_ident = p.n_cg * p.n_nt * 3;
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * (size_t)C * N, st));
    return gx.cg == 48 ? wgrad_h_launch<48>(p, st) : wgrad_h_launch<64>(p, st);
}
