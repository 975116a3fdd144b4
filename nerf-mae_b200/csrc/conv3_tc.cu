// 3x3x3 convolution (forward and dgrad) as an implicit GEMM on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// GEMM view: rows = output voxels, K = 27 taps x C input channels, N = output channels.
//
// Position space.  Each (batch, x) plane is cut into z-strips of SW <= 40 voxels; inside a strip the (y,z) voxels are
// linearised WITH one halo column on both sides of z:  pos = y*(SW+2) + (z - z0 + 1)  (the halo columns hold the
// neighbouring strips' voxels, or zeros at the volume border).  A CTA tile is 128 consecutive positions, so a (dy,dz)
// tap is the constant row offset dy*(SW+2)+dz.  The loader fetches, per (dx, 48-channel group), one shared-memory "image"
// of 128 + 2*(SW+3) positions (198 rows for SW=32) in the canonical no-swizzle K-major UMMA layout [k-chunk(6)][position][8 x bf16]
// (SBO = 128 B between 8-row groups, LBO = R_img*16 B between k-chunks).  Because rows are 16 B apart, the nine
// (dy,dz) taps of that image are just nine different descriptor START ADDRESSES: no im2col copy, each input
// element is fetched from L2 3x(C/48..) per tile instead of 27x.  Outputs that land on halo positions are
// discarded in the epilogue (2 of every Dz+2 rows).
//
// Precision.  fp32 operands are split into bf16 hi + bf16 lo when the image tensor is built (uimg.cu); each k-step issues
// hi*hi + hi*lo + lo*hi into the fp32 TMEM accumulator (~2^-17 relative, i.e. fp32-class results; the north_star
// tolerance is 1e-3 and a single bf16 pass does not meet it).
//
// Roles (256 threads): warp 0 image loader (cp.async.bulk from the pre-built UMMA-ready image tensor, uimg.cuh), warp 1 MMA
// issuer (+TMEM alloc), warp 2 weight loader (cp.async.bulk of pre-arranged blobs), warps 4-7 epilogue.  Persistent over tiles; TMEM accumulator double
// buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"
#include "uimg.cuh"

using namespace tc;

#define CG 48          // channels per image
#define KCH (CG / 8)   // 16-byte k-chunks per image row
#define TILE_M 128
#define MAX_IMG 4
#define MAX_BST 6

struct ConvTcParams {
    const uint8_t* uimg;   // UMMA-ready bf16 hi/lo image tensor of the input (uimg.cuh)
    long long u_chunk_bytes, u_part_bytes, u_img_bytes;
    float* y;
    const float* bias;
    const __nv_bfloat16* wblob;
    int B, Dx, Dy, Dz, C, N, NT, n_tiles_n;
    int SW, n_strips, ZP, P, tpp, num_m_tiles, H, R_img, n_cg, accumulate;
    int img_part_bytes, b_stage_bytes, b_tap_bytes, tps, n_bst, n_img, tmem_cols;
    int dbg;  // NMAE_DBG bit mask for bottleneck experiments: 1 no image loads, 2 no MMAs, 4 no weight copies, 8 no output stores
};

// weight blobs: [dx][cg][tap9][nt][kc][part(hi,lo)][n][8]  <-  value(n, c, tap) = w[n*s_n + c*s_c + tap']
__global__ void __launch_bounds__(256) conv3_tc_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ blob, int C, int N,
                                                            int NT, long long s_n, long long s_c, int flip) {
    long long total = 27LL * C * N * 2;
    int n_cg = C / CG, ntn = N / NT;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int e = (int)(r % 8); r /= 8;
        int n = (int)(r % NT); r /= NT;
        int part = (int)(r % 2); r /= 2;      // rows [0,NT) of a chunk = hi, [NT,2NT) = lo: one N=2*NT operand
        int kc = (int)(r % KCH); r /= KCH;
        int nt = (int)(r % ntn); r /= ntn;
        int tap9 = (int)(r % 9); r /= 9;
        int cg = (int)(r % n_cg); r /= n_cg;
        int dx = (int)r;
        int tap = dx * 9 + tap9;
        if (flip) tap = 26 - tap;
        float v = w[(long long)(nt * NT + n) * s_n + (long long)(cg * CG + kc * 8 + e) * s_c + tap];
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        blob[i] = part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

__device__ long long g_conv3_tc_cycles[256];   // NMAE_DBG bit 128: cycles the MMA warp of each CTA spent in its main loop

__global__ void __launch_bounds__(256, 1) conv3_tc_kernel(const __grid_constant__ ConvTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- shared memory carve-up
    uint8_t* img = smem;                                                   // [n_img buffers][2 parts][img_part_bytes]
    uint8_t* bst = img + 2 * (size_t)p.n_img * p.img_part_bytes;           // [n_bst][b_stage_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bst + (size_t)p.n_bst * p.b_stage_bytes);
    // barrier indices
    const uint32_t bar0 = smem_u32(bars);
    auto IMG_FULL = [&](int b) { return bar0 + 8u * (0 + b); };
    auto IMG_EMPTY = [&](int b) { return bar0 + 8u * (MAX_IMG + b); };
    auto B_FULL = [&](int s) { return bar0 + 8u * (2 * MAX_IMG + s); };
    auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_IMG + MAX_BST + s); };
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * MAX_IMG + 2 * MAX_BST + a); };
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * MAX_IMG + 2 * MAX_BST + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_IMG + 2 * MAX_BST + 4);

    if (tid == 0) {
        for (int b = 0; b < p.n_img; b++) {
            mbar_init(IMG_FULL(b), 1);             // expect_tx arrive of the image loader
            mbar_init(IMG_EMPTY(b), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(ACC_FULL(b), 1);
            mbar_init(ACC_EMPTY(b), 4);            // one arrive per epilogue warp
        }
        for (int s = 0; s < p.n_bst; s++) {
            mbar_init(B_FULL(s), 1);
            mbar_init(B_EMPTY(s), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int total_work = p.num_m_tiles * p.n_tiles_n;
    const uint32_t img0 = smem_u32(img), bst0 = smem_u32(bst);
    const uint32_t chunk_stride = (uint32_t)p.R_img * 16u;

    if (warp == 0) {
        // =========================================================== image loader: 12 bulk copies per image
        int buf = 0, ph = 0;
        const uint32_t img_bytes = 2u * (uint32_t)p.img_part_bytes, row_bytes = chunk_stride;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const int mt = w / p.n_tiles_n;
            const int p0 = (mt % p.tpp) * TILE_M;
            const int strip = (mt / p.tpp) % p.n_strips;
            const int xq = (mt / (p.tpp * p.n_strips)) % p.Dx, b = mt / (p.tpp * p.n_strips * p.Dx);
            for (int dx = 0; dx < 3; dx++) {
                const int xx = xq + dx - 1;
                if (xx < 0 || xx >= p.Dx) continue;
                for (int cg = 0; cg < p.n_cg; cg++) {
                    mbar_wait(IMG_EMPTY(buf), ph ^ 1);
                    if (elect_one()) {
                        const uint8_t* src = p.uimg + ((((long long)(b * (p.Dx + 2) + xx + 1) * p.n_strips + strip) * p.n_cg + cg)) * p.u_img_bytes +
                                             (long long)p0 * 16;
                        const uint32_t dst = img0 + (uint32_t)buf * img_bytes;
                        if (p.dbg & 1) {
                            mbar_arrive(IMG_FULL(buf));
                        } else {
                            mbar_expect_tx(IMG_FULL(buf), img_bytes);
#pragma unroll
                            for (int part = 0; part < 2; part++)
#pragma unroll
                                for (int c = 0; c < KCH; c++)
                                    bulk_g2s(dst + (uint32_t)(part * KCH + c) * row_bytes, src + part * p.u_part_bytes + c * p.u_chunk_bytes,
                                             row_bytes, IMG_FULL(buf));
                        }
                    }
                    __syncwarp();
                    if (++buf == p.n_img) { buf = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // =========================================================== weight loader (whole warp converged, one lane issues)
        {
            int s = 0, ph = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int mt = w / p.n_tiles_n, nt = w % p.n_tiles_n;
                const int xq = (mt / (p.tpp * p.n_strips)) % p.Dx;
                for (int dx = 0; dx < 3; dx++) {
                    const int xx = xq + dx - 1;
                    if (xx < 0 || xx >= p.Dx) continue;
                    for (int cg = 0; cg < p.n_cg; cg++) {
                        for (int t9 = 0; t9 < 9; t9 += p.tps) {   // tps taps per stage (tps == 3 only when n_tiles_n == 1)
                            mbar_wait(B_EMPTY(s), ph ^ 1);
                            if (elect_one()) {
                                if (p.dbg & 4) {
                                    mbar_arrive(B_FULL(s));
                                } else {
                                    mbar_expect_tx(B_FULL(s), (uint32_t)p.b_stage_bytes);
                                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wblob) +
                                                         ((size_t)(((dx * p.n_cg + cg) * 9 + t9) * p.n_tiles_n + nt)) * p.b_tap_bytes;
                                    bulk_g2s(bst0 + (uint32_t)s * p.b_stage_bytes, src, (uint32_t)p.b_stage_bytes, B_FULL(s));
                                }
                            }
                            __syncwarp();
                            if (++s == p.n_bst) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (whole warp converged, one elected lane issues)
        {
            // NMAE_DBG bit 64: N=16 instructions (garbage results) - separates issue/synchronisation cost from tensor-pipe cost
            const uint32_t idesc = idesc_bf16(TILE_M, (p.dbg & 64) ? 16 : p.NT, 0, 0), idesc2 = idesc_bf16(TILE_M, (p.dbg & 64) ? 16 : 2 * p.NT, 0, 0);
            const uint32_t dhi = desc_hi(128);                              // SBO = 128 B between 8-row groups (A and B)
            const uint32_t a_lbo = (uint32_t)p.R_img << 16, b_lbo = (uint32_t)(2 * p.NT) << 16;   // LBO in 16-byte units, pre-shifted
            const uint32_t b_tap16 = (uint32_t)p.b_tap_bytes >> 4, a_part16 = (uint32_t)p.img_part_bytes >> 4;
            const uint32_t a_k2 = 2u * (uint32_t)p.R_img, b_k4 = 4u * (uint32_t)p.NT;   // k-step strides (16-byte units)
            const int stages_per_img = 9 / p.tps;
            int buf = 0, iph = 0, s = 0, bph = 0, it = 0;
            const long long t_begin = clock64();
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
                const int mt = w / p.n_tiles_n;
                const int xq = (mt / (p.tpp * p.n_strips)) % p.Dx;
                const int acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(ACC_EMPTY(acc), aph ^ 1);
                fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * p.NT);
                uint32_t accum = 0;
                for (int dx = 0; dx < 3; dx++) {
                    const int xx = xq + dx - 1;
                    if (xx < 0 || xx >= p.Dx) continue;
                    for (int cg = 0; cg < p.n_cg; cg++) {
                        mbar_wait(IMG_FULL(buf), iph);
                        fence_after_sync();
                        // low descriptor word of the A operand for tap (dy=0, dz=0), k-step 0, hi part; every other A
                        // descriptor of this image is this word plus a constant (the issuing lane's budget is a handful of
                        // integer instructions per MMA: with N=48/96 an MMA lasts ~55 cycles, ncu showed the lane issue-bound)
                        uint32_t a_tap = a_lbo + ((img0 + (uint32_t)(buf * 2) * p.img_part_bytes) >> 4) + (uint32_t)(p.H - p.ZP - 1);
                        int dz = 0;
                        for (int st = 0; st < stages_per_img; st++) {
                            mbar_wait(B_FULL(s), bph);
                            fence_after_sync();
                            if (elect_one()) {
                                const uint32_t b_st = b_lbo + ((bst0 + (uint32_t)s * p.b_stage_bytes) >> 4);
                                if (!(p.dbg & 2)) {
                                    // per k-step: A_hi x [W_hi | W_lo] (one N = 2*NT instruction, columns [0,2NT)) and
                                    // A_lo x W_hi (N = NT, columns [0,NT)); the epilogue adds the two column blocks.
                                    // A_hi is fetched from shared memory once instead of twice.
                                    if (p.tps == 3) {
#pragma unroll
                                        for (int sub = 0; sub < 3; sub++) {
#pragma unroll
                                            for (int ks = 0; ks < CG / 16; ks++) {
                                                const uint32_t a = a_tap + (uint32_t)sub + (uint32_t)ks * a_k2;
                                                const uint64_t db = desc_pack(b_st + (uint32_t)sub * b_tap16 + (uint32_t)ks * b_k4, dhi);
                                                mma_bf16(d_tmem, desc_pack(a, dhi), db, idesc2, accum);
                                                accum = 1;
                                                mma_bf16(d_tmem, desc_pack(a + a_part16, dhi), db, idesc, 1);
                                            }
                                        }
                                    } else {
#pragma unroll
                                        for (int ks = 0; ks < CG / 16; ks++) {
                                            const uint32_t a = a_tap + (uint32_t)ks * a_k2;
                                            const uint64_t db = desc_pack(b_st + (uint32_t)ks * b_k4, dhi);
                                            mma_bf16(d_tmem, desc_pack(a, dhi), db, idesc2, accum);
                                            accum = 1;
                                            mma_bf16(d_tmem, desc_pack(a + a_part16, dhi), db, idesc, 1);
                                        }
                                    }
                                }
                                mma_commit(B_EMPTY(s));
                                if (st == stages_per_img - 1) mma_commit(IMG_EMPTY(buf));
                            }
                            __syncwarp();
                            // next stage: tps == 3 -> next dy (one row of the plane further); tps == 1 -> next dz, wrapping into the next dy
                            accum = 1;
                            if (p.tps == 3) {
                                a_tap += (uint32_t)p.ZP;
                            } else if (++dz == 3) {
                                dz = 0;
                                a_tap += (uint32_t)(p.ZP - 2);
                            } else {
                                a_tap += 1u;
                            }
                            if (++s == p.n_bst) { s = 0; bph ^= 1; }
                        }
                        if (++buf == p.n_img) { buf = 0; iph ^= 1; }
                    }
                }
                if (elect_one()) mma_commit(ACC_FULL(acc));
                __syncwarp();
            }
            if ((p.dbg & 128) && lane == 0 && blockIdx.x < 256) g_conv3_tc_cycles[blockIdx.x] = clock64() - t_begin;
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue (4 warps, one TMEM lane quarter each)
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
            const int mt = w / p.n_tiles_n, nt = w % p.n_tiles_n;
            const int p0 = (mt % p.tpp) * TILE_M;
            const int strip = (mt / p.tpp) % p.n_strips;
            const int xq = (mt / (p.tpp * p.n_strips)) % p.Dx, b = mt / (p.tpp * p.n_strips * p.Dx);
            const int acc = it & 1, aph = (it >> 1) & 1;
            const int pos = p0 + m;
            bool valid = pos < p.P;
            int yy = 0, z = 0;
            if (valid) {
                yy = pos / p.ZP;
                const int zz = pos - yy * p.ZP;
                z = strip * p.SW + zz - 1;
                valid = zz >= 1 && zz <= p.SW && z < p.Dz;
            }
            float* dst = p.y + ((((long long)(b * p.Dx + xq) * p.Dy + yy) * p.Dz + z) * p.N + nt * p.NT);
            mbar_wait_warp(ACC_FULL(acc), aph);
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.NT);
            for (int j = 0; j < ((p.dbg & 32) ? 0 : p.NT / 16); j++) {
                float v[16], v2[16];
                tmem_ld16(taddr + j * 16, v);
                tmem_ld16(taddr + p.NT + j * 16, v2);       // the A_hi x W_lo column block
#pragma unroll
                for (int e = 0; e < 16; e++) v[e] += v2[e];
                if (valid && !(p.dbg & 8)) {
                    if (p.bias) {
#pragma unroll
                        for (int e = 0; e < 16; e++) v[e] += __ldg(p.bias + nt * p.NT + j * 16 + e);
                    }
                    float4* d4 = reinterpret_cast<float4*>(dst + j * 16);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        float4 o = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        if (p.accumulate) {
                            float4 old = d4[e];
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        d4[e] = o;
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(ACC_EMPTY(acc));
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int pick_nt(int N) {
    for (int nt = 128; nt >= 16; nt -= 16)   // [W_hi | W_lo] is issued as one N = 2*NT <= 256 operand
        if (N % nt == 0) return nt;
    return 0;
}

bool k_conv3_tc_supported(int C, int N) { return C % CG == 0 && N % 16 == 0 && pick_nt(N) >= 16; }

// mode 0: forward (w is (N, C, 27)); mode 1: dgrad (w is (C, N, 27): out channel of the GEMM = w's in-channel, taps flipped).
// `uimg` is the type-X image tensor of the GEMM input (uimg.cuh) with C channels.
int k_conv3_tc(const void* uimg, const float* w, const float* bias, int B, int Dx, int Dy, int Dz, int C, int N, int mode, float* w_ws,
               float* y, int accumulate, cudaStream_t st) {
    NMAE_CHECK_ARG(k_conv3_tc_supported(C, N), "conv3_tc: unsupported channels C=%d N=%d", C, N);
    const UImgGeom g = uimg_geom(B, Dx, Dy, Dz, C);
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.uimg = reinterpret_cast<const uint8_t*>(uimg);
    p.u_chunk_bytes = g.chunk_bytes; p.u_part_bytes = g.part_bytes; p.u_img_bytes = g.img_bytes;
    p.y = y; p.bias = bias; p.wblob = reinterpret_cast<const __nv_bfloat16*>(w_ws);
    p.B = B; p.Dx = Dx; p.Dy = Dy; p.Dz = Dz; p.C = C; p.N = N;
    p.NT = pick_nt(N);
    p.n_tiles_n = N / p.NT;
    p.SW = g.SW; p.n_strips = g.n_strips; p.ZP = g.ZP; p.P = g.P; p.tpp = g.tpp; p.H = g.H; p.R_img = g.R_img; p.n_cg = g.n_cg;
    p.num_m_tiles = B * Dx * p.n_strips * p.tpp;
    p.accumulate = accumulate;
    p.dbg = nmae_debug_mask();
    p.img_part_bytes = KCH * p.R_img * 16;
    p.b_tap_bytes = p.NT * CG * 2 * 2;
    int tm = 4 * p.NT;
    p.tmem_cols = tm <= 32 ? 32 : tm <= 64 ? 64 : tm <= 128 ? 128 : tm <= 256 ? 256 : 512;
    const int bar_bytes = 8 * (2 * MAX_IMG + 2 * MAX_BST + 4) + 16;
    const int max_smem = 227 * 1024;
    // shared-memory budget: 3 image buffers when at least two weight stages still fit, else 2
    p.tps = (p.n_tiles_n == 1) ? 3 : 1;     // three dz taps per weight stage (3x fewer handshakes) when the blobs are contiguous
    p.b_stage_bytes = p.tps * p.b_tap_bytes;
    p.n_img = 3;
    if (2LL * p.n_img * p.img_part_bytes + bar_bytes + 2LL * p.b_stage_bytes > max_smem) p.n_img = 2;
    if (2LL * p.n_img * p.img_part_bytes + bar_bytes + 2LL * p.b_stage_bytes > max_smem) {
        p.tps = 1;
        p.b_stage_bytes = p.b_tap_bytes;
    }
    long long fixed = 2LL * p.n_img * p.img_part_bytes + bar_bytes;
    NMAE_CHECK_ARG(fixed + 2LL * p.b_stage_bytes <= max_smem, "conv3_tc: tile does not fit in shared memory (Dz=%d N=%d)", Dz, N);
    p.n_bst = (int)((max_smem - fixed) / p.b_stage_bytes);
    if (p.n_bst > MAX_BST) p.n_bst = MAX_BST;
    size_t smem = (size_t)fixed + (size_t)p.n_bst * p.b_stage_bytes;

    // weights -> bf16 hi/lo blobs in the UMMA layout
    long long total = 27LL * C * N * 2;
    int gr = (int)min((long long)148 * 8, (total + 255) / 256);
    if (mode == 0)
        conv3_tc_prep_kernel<<<gr, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(w_ws), C, N, p.NT, (long long)C * 27, 27, 0);
    else
        conv3_tc_prep_kernel<<<gr, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(w_ws), C, N, p.NT, 27, (long long)N * 27, 1);
    NMAE_LAUNCH_CHECK();

    static bool attr_set[64] = {false};
    int dev;
    NMAE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(conv3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set[dev] = true;
    }
    int sms = 148;
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int grid = min(sms, p.num_m_tiles * p.n_tiles_n);
    conv3_tc_kernel<<<grid, 256, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    if (p.dbg & 128) {
        long long h[256];
        NMAE_CUDA(cudaStreamSynchronize(st));
        NMAE_CUDA(cudaMemcpyFromSymbol(h, g_conv3_tc_cycles, sizeof(h)));
        double avg = 0;
        for (int i = 0; i < grid; i++) avg += (double)h[i];
        fprintf(stderr, "[NMAE_DBG] conv3_tc: %.3f Mcycles in the MMA loop (avg over %d CTAs), %d tiles\n", avg / grid / 1e6, grid,
                p.num_m_tiles * p.n_tiles_n);
    }
    return NMAE_OK;
}
