// 3D shifted-window attention core (reference swin_mae3d.py:27-197) on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// Work item = (batch, PAIR of windows, head): the 2 x 64 tokens of two windows are the 128 rows of one UMMA tile.
//   S  = Q K^T       one M=128, N=128 (keys of both windows), K=32 product; rows of window w only use the 64 columns
//                    of window w (the off-diagonal blocks are computed and ignored - the tensor pipe has no 64-row mode
//                    that is cheaper, tools/mma_bench.cu)
//   P  = softmax(S + relative-position bias + shift mask), fp32, one thread per row straight out of TMEM
//   O  = P V         P is written back to shared memory as a bf16 hi/lo K-major operand
// fp32 operands are split into bf16 hi + lo and every product is issued as hi*hi + hi*lo + lo*hi (fp32-class accuracy,
// like the convolution and linear kernels).  The cyclic shift, the padding to a multiple of the window and the window
// partition are one index map (slot_map); padding tokens read the qkv bias row and are NOT masked.
//
// Backward (same tiling): S and dP = dO V^T are recomputed on the tensor cores, P = exp(S - lse), dS = P o (dP - D);
// P and dS are written to shared memory as BLOCK-DIAGONAL 128 x 128 operands (the off-diagonal blocks stay zero), so that
//   dV = P^T dO,   dK = dS^T Q,   dQ = dS K
// are three uniform M=128, N=32, K=128 products whose 128 output rows are all useful.  P^T / dS^T cost nothing: the
// K-major image of a matrix is the MN-major image of its transpose.
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"

using namespace tc;

#define WS 4
#define NTOK 64   // tokens per window
#define HD 32     // head dim
#define ROWS 128  // two windows

struct WinGeomTc {
    int H, W, D;     // real token grid
    int PH, PW, PD;  // padded
    int sh, sw, sd;  // effective shift per axis
    int nWh, nWw, nWd;
};

__device__ __forceinline__ void slot_map_tc(const WinGeomTc& g, int win, int slot, int& src, int& region) {
    int wd = win % g.nWd, t = win / g.nWd;
    int ww = t % g.nWw, wh = t / g.nWw;
    int ih = slot >> 4, iw = (slot >> 2) & 3, id = slot & 3;
    int rh = wh * WS + ih, rw = ww * WS + iw, rd = wd * WS + id;
    int h = rh + g.sh; if (h >= g.PH) h -= g.PH;
    int w = rw + g.sw; if (w >= g.PW) w -= g.PW;
    int d = rd + g.sd; if (d >= g.PD) d -= g.PD;
    src = (h < g.H && w < g.W && d < g.D) ? (h * g.W + w) * g.D + d : -1;
    int bh = g.sh ? (rh >= g.PH - g.sh ? 2 : (rh >= g.PH - WS ? 1 : 0)) : 0;
    int bw = g.sw ? (rw >= g.PW - g.sw ? 2 : (rw >= g.PW - WS ? 1 : 0)) : 0;
    int bd = g.sd ? (rd >= g.PD - g.sd ? 2 : (rd >= g.PD - WS ? 1 : 0)) : 0;
    region = (bh * 3 + bw) * 3 + bd;
}

__device__ __forceinline__ int rel_index_tc(int qi, int kj) {
    int dh = (qi >> 4) - (kj >> 4) + 3, dw = ((qi >> 2) & 3) - ((kj >> 2) & 3) + 3, dd = (qi & 3) - (kj & 3) + 3;
    return (dh * 7 + dw) * 7 + dd;
}

static int make_geom_tc(int H, int W, int D, int shift, WinGeomTc& g) {
    g.H = H; g.W = W; g.D = D;
    g.PH = cdiv(H, WS) * WS; g.PW = cdiv(W, WS) * WS; g.PD = cdiv(D, WS) * WS;
    // swin_mae3d.py:68-75: the shift of an axis is dropped when the window covers the padded extent
    g.sh = (WS >= g.PH) ? 0 : shift;
    g.sw = (WS >= g.PW) ? 0 : shift;
    g.sd = (WS >= g.PD) ? 0 : shift;
    g.nWh = g.PH / WS; g.nWw = g.PW / WS; g.nWd = g.PD / WS;
    return g.nWh * g.nWw * g.nWd;
}

int k_wattn_num_windows(int H, int W, int D) {
    WinGeomTc g;
    return make_geom_tc(H, W, D, 0, g);
}

// Operand tiles in shared memory: [part: hi, lo][8-element chunk][128 rows][8 x bf16] (chunk stride 2048 B).
//   as a K-major operand  (rows = M or N, chunks = k):  SBO = 128 B,  LBO = 2048 B
//   as an MN-major operand (chunks = M or N, rows = k): SBO = 2048 B, LBO = 128 B
#define CHUNK_B 2048
#define T32_PART (4 * CHUNK_B)    // 128 x 32 tile, one part: 8 KB
#define T32_BYTES (2 * T32_PART)  // hi + lo: 16 KB

// Staging of NT_ 128 x 32 fp32 tiles (token rows gathered through s_row) into bf16 hi/lo operand tiles.  8 lanes per token
// row: lane j owns float4 j of the row's 32-float head slice.  ALL global loads of a thread are issued before the first
// conversion (the shared-memory stores would otherwise fence them one by one: the kernel is latency-bound there).
struct StageSrc {
    const float* base;     // row-major source
    long long row_stride;  // floats
    int col0;
    float scale;
    int masked;            // 1: rows with s_ok[r] == 0 are staged as zeros
    uint8_t* tile;
};

template <int NT_, int NTHREADS>
__device__ __forceinline__ void stage_tiles(const StageSrc (&src)[NT_], const long long* s_row, const int* s_ok, int tid, float4 (&v)[NT_][ROWS * 8 / NTHREADS]) {
    constexpr int PER = ROWS * 8 / NTHREADS;
#pragma unroll
    for (int t = 0; t < NT_; t++)
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int u = tid + i * NTHREADS, r = u >> 3, j = u & 7;
            v[t][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!src[t].masked || s_ok[r]) v[t][i] = __ldg(reinterpret_cast<const float4*>(src[t].base + s_row[r] * src[t].row_stride + src[t].col0) + j);
        }
#pragma unroll
    for (int t = 0; t < NT_; t++)
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int u = tid + i * NTHREADS, r = u >> 3, j = u & 7;
            const float sc = src[t].scale;
            uint2 h, l;
            split2(v[t][i].x * sc, v[t][i].y * sc, h.x, l.x);
            split2(v[t][i].z * sc, v[t][i].w * sc, h.y, l.y);
            uint8_t* d = src[t].tile + (size_t)(j >> 1) * CHUNK_B + (size_t)r * 16 + (size_t)(j & 1) * 8;
            *reinterpret_cast<uint2*>(d) = h;
            *reinterpret_cast<uint2*>(d + T32_PART) = l;
        }
}

// three-pass product: D (+)= A_hi B_hi + A_hi B_lo + A_lo B_hi over `ksteps` 16-wide k-steps
__device__ __forceinline__ void mma3(uint32_t d_tmem, uint32_t a16, uint32_t a_part16, uint32_t a_hi, uint32_t a_lbo, uint32_t a_step16,
                                     uint32_t b16, uint32_t b_part16, uint32_t b_hi, uint32_t b_lbo, uint32_t b_step16, uint32_t idesc,
                                     int ksteps) {
    for (int ks = 0; ks < ksteps; ks++) {
        const uint64_t ah = desc_make(a_hi, a_lbo, a16 + ks * a_step16), al = desc_make(a_hi, a_lbo, a16 + a_part16 + ks * a_step16);
        const uint64_t bh = desc_make(b_hi, b_lbo, b16 + ks * b_step16), bl = desc_make(b_hi, b_lbo, b16 + b_part16 + ks * b_step16);
        mma_bf16(d_tmem, ah, bh, idesc, ks ? 1u : 0u);
        mma_bf16(d_tmem, ah, bl, idesc, 1);
        mma_bf16(d_tmem, al, bh, idesc, 1);
    }
}

struct WmsaTcParams {
    const float* qkv;
    const float* table;
    float* out;
    float* lse;
    // backward only
    const float* o_saved;
    const float* lse_in;
    const float* dout;
    float* dqkv;
    float* dtable;
    WinGeomTc g;
    int B, nW, nPairs, T, C, nH, num_items;
    long long pad_row;
    float scale;
};

// ------------------------------------------------------------------------------------------------ forward
// shared memory: Q, K, V tiles (16 KB each), P tile [hi, lo][8 key chunks][128 rows] (32 KB), small tables
#define FWD_P_PART (8 * CHUNK_B)
#define FWD_SMEM (3 * T32_BYTES + 2 * FWD_P_PART + 343 * 4 + ROWS * 8 + ROWS * 4 + 64)

__global__ void __launch_bounds__(128, 2) wmsa_tc_fwd_kernel(const __grid_constant__ WmsaTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + T32_BYTES;
    uint8_t* sV = sK + T32_BYTES;
    uint8_t* sP = sV + T32_BYTES;
    float* stab = reinterpret_cast<float*>(sP + 2 * FWD_P_PART);
    long long* s_row = reinterpret_cast<long long*>(stab + 344);    // qkv row of each of the 128 slots
    int* s_reg = reinterpret_cast<int*>(s_row + ROWS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_reg + ROWS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_s = smem_u32(bars), bar_o = bar_s + 8;

    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t q16 = smem_u32(sQ) >> 4, k16 = smem_u32(sK) >> 4, v16 = smem_u32(sV) >> 4, p16 = smem_u32(sP) >> 4;
    const uint32_t kmaj_hi = desc_hi(128), kmaj_lbo = (CHUNK_B >> 4) << 16;      // K-major: SBO 128 B, LBO = chunk stride
    const uint32_t mnmaj_hi = desc_hi(CHUNK_B), mnmaj_lbo = (128u >> 4) << 16;   // MN-major: SBO = chunk stride, LBO 128 B
    const uint32_t idesc_s = idesc_bf16(128, 128, 0, 0), idesc_o = idesc_bf16(128, HD, 0, 1);

    // head of this CTA (fixed: its relative-position-bias table is loaded once); neighbouring CTAs work on the same token
    // rows for different heads at the same time, so the rows are fetched from DRAM once
    const int h = blockIdx.x % p.nH, group = blockIdx.x / p.nH, n_groups = gridDim.x / p.nH;
    for (int i = tid; i < 343; i += 128) stab[i] = p.table[i * p.nH + h];
    const int n_bp = p.B * p.nPairs;
    int it = 0;
    for (int bp = group; bp < n_bp; bp += n_groups, it++) {
        const int wp = bp % p.nPairs, b = bp / p.nPairs;
        const int w = tid >> 6, slot = tid & 63, win = wp * 2 + w;
        // ---- slot map
        {
            int src = -1, region = 0;
            if (win < p.nW) slot_map_tc(p.g, win, slot, src, region);
            s_row[tid] = src >= 0 ? (long long)b * p.T + src : p.pad_row;
            s_reg[tid] = region;
        }
        __syncthreads();
        // ---- stage Q (scaled), K, V
        const long long C3 = 3LL * p.C;
        {
            const StageSrc src[3] = {{p.qkv, C3, h * HD, p.scale, 0, sQ}, {p.qkv, C3, p.C + h * HD, 1.f, 0, sK}, {p.qkv, C3, 2 * p.C + h * HD, 1.f, 0, sV}};
            float4 v[3][8];
            stage_tiles<3, 128>(src, s_row, nullptr, tid, v);
        }
        fence_proxy_async();
        __syncthreads();
        // ---- S = Q K^T  -> TMEM columns [0, 128)
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
                mma3(tmem_base, q16, T32_PART >> 4, kmaj_hi, kmaj_lbo, (2 * CHUNK_B) >> 4, k16, T32_PART >> 4, kmaj_hi, kmaj_lbo,
                     (2 * CHUNK_B) >> 4, idesc_s, HD / 16);
                mma_commit(bar_s);
            }
            __syncwarp();
        }
        mbar_wait_warp(bar_s, it & 1);
        fence_after_sync();
        // ---- softmax of row `tid` over the 64 keys of its own window
        {
            const int qi = slot, myreg = s_reg[tid];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(w * NTOK);
            float s[NTOK];
#pragma unroll
            for (int j = 0; j < NTOK / 16; j++) tmem_ld16(taddr + j * 16, s + j * 16);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NTOK; j++) {
                float v = s[j] + stab[rel_index_tc(qi, j)];
                if (s_reg[w * NTOK + j] != myreg) v += -100.f;
                s[j] = v;
                mx = fmaxf(mx, v);
            }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NTOK; j++) {
                s[j] = expf(s[j] - mx);
                sum += s[j];
            }
            const float inv = 1.f / sum;
            if (win < p.nW) p.lse[(((long long)b * p.nW + win) * p.nH + h) * NTOK + qi] = mx + logf(sum);
            uint8_t* prow = sP + (size_t)tid * 16;
#pragma unroll
            for (int c = 0; c < NTOK / 8; c++) {
                uint4 hi, lo;
                split2(s[8 * c] * inv, s[8 * c + 1] * inv, hi.x, lo.x);
                split2(s[8 * c + 2] * inv, s[8 * c + 3] * inv, hi.y, lo.y);
                split2(s[8 * c + 4] * inv, s[8 * c + 5] * inv, hi.z, lo.z);
                split2(s[8 * c + 6] * inv, s[8 * c + 7] * inv, hi.w, lo.w);
                *reinterpret_cast<uint4*>(prow + (size_t)c * CHUNK_B) = hi;
                *reinterpret_cast<uint4*>(prow + (size_t)c * CHUNK_B + FWD_P_PART) = lo;
            }
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        // ---- O = P V: window 0 rows against V[0:64) -> columns [128,160), window 1 rows against V[64:128) -> [160,192)
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
                for (int ww = 0; ww < 2; ww++)
                    mma3(tmem_base + 128 + ww * HD, p16, FWD_P_PART >> 4, kmaj_hi, kmaj_lbo, (2 * CHUNK_B) >> 4,
                         v16 + ((ww * NTOK * 16) >> 4), T32_PART >> 4, mnmaj_hi, mnmaj_lbo, (16 * 16) >> 4, idesc_o, NTOK / 16);
                mma_commit(bar_o);
            }
            __syncwarp();
        }
        mbar_wait_warp(bar_o, it & 1);
        fence_after_sync();
        {
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(128 + w * HD);
            float o[HD];
            tmem_ld16(taddr, o);
            tmem_ld16(taddr + 16, o + 16);
            if (s_row[tid] != p.pad_row) {
                float4* op = reinterpret_cast<float4*>(p.out + s_row[tid] * p.C + h * HD);
#pragma unroll
                for (int e = 0; e < HD / 4; e++) op[e] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
            }
        }
        fence_before_sync();
        __syncthreads();
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}

static void fill_params(WmsaTcParams& p, int B, int H, int W, int D, int C, int nH, int shift) {
    memset(&p, 0, sizeof(p));
    p.nW = make_geom_tc(H, W, D, shift, p.g);
    p.nPairs = (p.nW + 1) / 2;
    p.B = B; p.T = H * W * D; p.C = C; p.nH = nH;
    p.num_items = B * p.nPairs * nH;
    p.pad_row = (long long)B * p.T;
    p.scale = 1.f / sqrtf((float)HD);
}

int k_wattn_tc_fwd(const float* qkv, const float* table, int B, int H, int W, int D, int C, int nH, int shift, float* out, float* lse,
                   cudaStream_t st) {
    NMAE_CHECK_ARG(C == nH * HD, "window attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
    NMAE_CHECK_ARG(shift >= 0 && shift < WS, "window attention: shift %d out of range", shift);
    WmsaTcParams p;
    fill_params(p, B, H, W, D, C, nH, shift);
    p.qkv = qkv; p.table = table; p.out = out; p.lse = lse;
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(wmsa_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
        attr_set[dev] = true;
    }
    NMAE_CHECK_ARG(nH <= 2 * sms, "window attention: more heads (%d) than resident CTAs", nH);
    const int groups = max(1, min(2 * sms / nH, B * p.nPairs));
    wmsa_tc_fwd_kernel<<<groups * nH, 128, FWD_SMEM, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------ backward
// shared memory: Q (scaled), K, V, dO tiles (16 KB each); P and dS as block-diagonal 128 x 128 operands
// [hi, lo][16 key chunks][128 query rows] (64 KB each); small tables.  One CTA per SM, 8 warps.
#define BWD_PD_PART (16 * CHUNK_B)   // 32 KB
#define BWD_SMEM (4 * T32_BYTES + 4 * BWD_PD_PART + 2 * 344 * 4 + ROWS * 8 + ROWS * 4 * 4 + 64 * 4 + 64)

__global__ void __launch_bounds__(256, 1) wmsa_tc_bwd_kernel(const __grid_constant__ WmsaTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + T32_BYTES;
    uint8_t* sV = sK + T32_BYTES;
    uint8_t* sdO = sV + T32_BYTES;
    uint8_t* sP = sdO + T32_BYTES;
    uint8_t* sdS = sP + 2 * BWD_PD_PART;
    float* stab = reinterpret_cast<float*>(sdS + 2 * BWD_PD_PART);
    float* sdb = stab + 344;
    long long* s_row = reinterpret_cast<long long*>(sdb + 344);
    int* s_reg = reinterpret_cast<int*>(s_row + ROWS);
    int* s_ok = s_reg + ROWS;                                   // 1: real token, 0: padding slot or absent window
    float* s_lse = reinterpret_cast<float*>(s_ok + ROWS);
    float* s_D = s_lse + ROWS;
    float* s_pad = s_D + ROWS;                                // [dK(32) | dV(32)] gradient through this head's padding k/v
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_pad + 64);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar1 = smem_u32(bars), bar2 = bar1 + 8;
    if (tid < 64) s_pad[tid] = 0.f;

    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    // the off-diagonal blocks of P and dS are never written again
    for (int i = tid; i < 4 * BWD_PD_PART / 16; i += 256) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    // head of this CTA (fixed, so that the bias-gradient partial sums can stay in registers across items)
    const int h = blockIdx.x % p.nH, group = blockIdx.x / p.nH, n_groups = gridDim.x / p.nH;
    for (int i = tid; i < 343; i += 256) {
        stab[i] = p.table[i * p.nH + h];
        sdb[i] = 0.f;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t q16 = smem_u32(sQ) >> 4, k16 = smem_u32(sK) >> 4, v16 = smem_u32(sV) >> 4, do16 = smem_u32(sdO) >> 4;
    const uint32_t p16 = smem_u32(sP) >> 4, ds16 = smem_u32(sdS) >> 4;
    const uint32_t kmaj_hi = desc_hi(128), kmaj_lbo = (CHUNK_B >> 4) << 16;
    const uint32_t mnmaj_hi = desc_hi(CHUNK_B), mnmaj_lbo = (128u >> 4) << 16;
    const uint32_t idesc_s = idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_t = idesc_bf16(128, HD, 1, 1);   // A = P^T / dS^T (MN-major), B = dO / Q (MN-major: N = dims, k = rows)
    const uint32_t idesc_q = idesc_bf16(128, HD, 0, 1);   // A = dS (K-major), B = K (MN-major)
    enum { COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 288, COL_DQ = 320 };

    const int row = (warp & 3) * 32 + lane, hf = warp >> 2;   // elementwise phase: thread = (query row, half of its 64 keys)
    const int w_row = row >> 6, qi = row & 63;
    float dsacc[32];
#pragma unroll
    for (int j = 0; j < 32; j++) dsacc[j] = 0.f;

    const int n_bp = p.B * p.nPairs;
    int it = 0;
    for (int bp = group; bp < n_bp; bp += n_groups, it++) {
        const int wp = bp % p.nPairs, b = bp / p.nPairs;
        // ---- slot map, log-sum-exp of each query row
        if (tid < ROWS) {
            const int w = tid >> 6, slot = tid & 63, win = wp * 2 + w;
            int src = -1, region = 0;
            if (win < p.nW) slot_map_tc(p.g, win, slot, src, region);
            s_row[tid] = src >= 0 ? (long long)b * p.T + src : p.pad_row;
            s_reg[tid] = region;
            s_ok[tid] = src >= 0;
            s_lse[tid] = win < p.nW ? p.lse_in[(((long long)b * p.nW + win) * p.nH + h) * NTOK + slot] : 0.f;
        }
        __syncthreads();
        // ---- stage Q (scaled), K, V, dO (zeros for padding slots)
        const long long C3 = 3LL * p.C;
        {
            const StageSrc src[4] = {{p.qkv, C3, h * HD, p.scale, 0, sQ}, {p.qkv, C3, p.C + h * HD, 1.f, 0, sK},
                                     {p.qkv, C3, 2 * p.C + h * HD, 1.f, 0, sV}, {p.dout, (long long)p.C, h * HD, 1.f, 1, sdO}};
            float4 v[4][4];
            // the saved attention output O of the same (row, float4) units: D_r = sum_d dO[r][d] * O[r][d] (== rowsum(dP o P))
            float4 o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int u = tid + i * 256, r = u >> 3, j = u & 7;
                o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s_ok[r]) o[i] = __ldg(reinterpret_cast<const float4*>(p.o_saved + s_row[r] * p.C + h * HD) + j);
            }
            stage_tiles<4, 256>(src, s_row, s_ok, tid, v);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float d = v[3][i].x * o[i].x + v[3][i].y * o[i].y + v[3][i].z * o[i].z + v[3][i].w * o[i].w;
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                d += __shfl_xor_sync(0xffffffffu, d, 2);
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                if ((tid & 7) == 0) s_D[(tid + i * 256) >> 3] = d;
            }
        }
        const bool qvalid = s_ok[row] != 0;
        fence_proxy_async();
        __syncthreads();
        // ---- S = Q K^T, dP = dO V^T
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
                mma3(tmem_base + COL_S, q16, T32_PART >> 4, kmaj_hi, kmaj_lbo, (2 * CHUNK_B) >> 4, k16, T32_PART >> 4, kmaj_hi, kmaj_lbo,
                     (2 * CHUNK_B) >> 4, idesc_s, HD / 16);
                mma3(tmem_base + COL_DP, do16, T32_PART >> 4, kmaj_hi, kmaj_lbo, (2 * CHUNK_B) >> 4, v16, T32_PART >> 4, kmaj_hi, kmaj_lbo,
                     (2 * CHUNK_B) >> 4, idesc_s, HD / 16);
                mma_commit(bar1);
            }
            __syncwarp();
        }
        mbar_wait_warp(bar1, it & 1);
        fence_after_sync();
        // ---- P = exp(S + bias + mask - lse), dS = P o (dP - D): 32 keys per thread
        {
            const int myreg = s_reg[row], kbase = w_row * NTOK + hf * 32;
            const float l = s_lse[row], Dr = s_D[row];
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kbase;
            float s[32], dp[32];
            tmem_ld16(taddr + COL_S, s);
            tmem_ld16(taddr + COL_S + 16, s + 16);
            tmem_ld16(taddr + COL_DP, dp);
            tmem_ld16(taddr + COL_DP + 16, dp + 16);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float v = s[j] + stab[rel_index_tc(qi, hf * 32 + j)];
                if (s_reg[kbase + j] != myreg) v += -100.f;
                const float pj = qvalid ? expf(v - l) : 0.f;
                const float ds = pj * (dp[j] - Dr);
                s[j] = pj;
                dp[j] = ds;
                dsacc[j] += ds;
            }
            uint8_t* prow = sP + (size_t)(kbase >> 3) * CHUNK_B + (size_t)row * 16;
            uint8_t* drow = sdS + (size_t)(kbase >> 3) * CHUNK_B + (size_t)row * 16;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint4 hi, lo;
                split2(s[8 * c], s[8 * c + 1], hi.x, lo.x);
                split2(s[8 * c + 2], s[8 * c + 3], hi.y, lo.y);
                split2(s[8 * c + 4], s[8 * c + 5], hi.z, lo.z);
                split2(s[8 * c + 6], s[8 * c + 7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(prow + (size_t)c * CHUNK_B) = hi;
                *reinterpret_cast<uint4*>(prow + (size_t)c * CHUNK_B + BWD_PD_PART) = lo;
                split2(dp[8 * c], dp[8 * c + 1], hi.x, lo.x);
                split2(dp[8 * c + 2], dp[8 * c + 3], hi.y, lo.y);
                split2(dp[8 * c + 4], dp[8 * c + 5], hi.z, lo.z);
                split2(dp[8 * c + 6], dp[8 * c + 7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(drow + (size_t)c * CHUNK_B) = hi;
                *reinterpret_cast<uint4*>(drow + (size_t)c * CHUNK_B + BWD_PD_PART) = lo;
            }
        }
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        // ---- dV = P^T dO, dK = dS^T Q, dQ = dS K   (K = 128 rows / keys each)
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
                mma3(tmem_base + COL_DV, p16, BWD_PD_PART >> 4, mnmaj_hi, mnmaj_lbo, 16, do16, T32_PART >> 4, mnmaj_hi, mnmaj_lbo, 16, idesc_t,
                     ROWS / 16);
                mma3(tmem_base + COL_DK, ds16, BWD_PD_PART >> 4, mnmaj_hi, mnmaj_lbo, 16, q16, T32_PART >> 4, mnmaj_hi, mnmaj_lbo, 16, idesc_t,
                     ROWS / 16);
                mma3(tmem_base + COL_DQ, ds16, BWD_PD_PART >> 4, kmaj_hi, kmaj_lbo, (2 * CHUNK_B) >> 4, k16, T32_PART >> 4, mnmaj_hi, mnmaj_lbo, 16,
                     idesc_q, ROWS / 16);
                mma_commit(bar2);
            }
            __syncwarp();
        }
        mbar_wait_warp(bar2, it & 1);
        fence_after_sync();
        // ---- epilogue: warps 0-3 write dK, dV of token slot `row`; warps 4-7 write dQ
        {
            const long long grow = s_row[row];
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
            float v[HD];
            if (hf == 1) {
                tmem_ld16(taddr + COL_DQ, v);
                tmem_ld16(taddr + COL_DQ + 16, v + 16);
                if (qvalid) {
                    float4* dst = reinterpret_cast<float4*>(p.dqkv + grow * C3 + h * HD);
#pragma unroll
                    for (int e = 0; e < HD / 4; e++)
                        dst[e] = make_float4(v[4 * e] * p.scale, v[4 * e + 1] * p.scale, v[4 * e + 2] * p.scale, v[4 * e + 3] * p.scale);
                }
            } else {
                // padding slots: their k/v are the qkv bias row, shared by every window - summed per warp, then per CTA in
                // shared memory, one global atomic per CTA and channel at the end (global atomics per slot serialise in L2)
                const bool is_pad = !qvalid && (wp * 2 + w_row < p.nW);
                const bool any_pad = __any_sync(0xffffffffu, is_pad);
#pragma unroll
                for (int which = 0; which < 2; which++) {   // 0: dK, 1: dV
                    tmem_ld16(taddr + (which ? COL_DV : COL_DK), v);
                    tmem_ld16(taddr + (which ? COL_DV : COL_DK) + 16, v + 16);
                    if (qvalid) {
                        float* dst = p.dqkv + grow * C3 + (which ? 2 : 1) * p.C + h * HD;
#pragma unroll
                        for (int e = 0; e < HD / 4; e++)
                            reinterpret_cast<float4*>(dst)[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    }
                    if (any_pad) {
                        float mine = 0.f;
#pragma unroll
                        for (int e = 0; e < HD; e++) {
                            const float t = warp_sum(is_pad ? v[e] : 0.f);
                            if (lane == e) mine = t;
                        }
                        atomicAdd(&s_pad[which * HD + lane], mine);
                    }
                }
            }
        }
        fence_before_sync();
        __syncthreads();
    }

    // ---- relative-position-bias gradient of this head
#pragma unroll
    for (int j = 0; j < 32; j++)
        if (dsacc[j] != 0.f) atomicAdd(&sdb[rel_index_tc(qi, hf * 32 + j)], dsacc[j]);
    __syncthreads();
    for (int i = tid; i < 343; i += 256)
        if (sdb[i] != 0.f) atomicAdd(p.dtable + i * p.nH + h, sdb[i]);
    if (tid < 64 && s_pad[tid] != 0.f) atomicAdd(p.dqkv + p.pad_row * 3 * p.C + (1 + (tid >> 5)) * p.C + h * HD + (tid & 31), s_pad[tid]);

    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// dqkv ((B*T + 1) x 3C; row B*T receives the gradient through padding tokens) is overwritten; dtable must be zeroed by the caller
int k_wattn_tc_bwd(const float* qkv, const float* table, const float* o_saved, const float* dout, const float* lse, int B, int H, int W,
                   int D, int C, int nH, int shift, float* dqkv, float* dtable, cudaStream_t st) {
    NMAE_CHECK_ARG(C == nH * HD, "window attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
    NMAE_CHECK_ARG(shift >= 0 && shift < WS, "window attention: shift %d out of range", shift);
    WmsaTcParams p;
    fill_params(p, B, H, W, D, C, nH, shift);
    p.qkv = qkv; p.table = table; p.o_saved = o_saved; p.dout = dout; p.lse_in = lse; p.dqkv = dqkv; p.dtable = dtable;
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NMAE_CHECK_ARG(nH <= sms, "window attention: more heads (%d) than SMs", nH);
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(wmsa_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
        attr_set[dev] = true;
    }
    NMAE_CUDA(cudaMemsetAsync(dqkv + (long long)B * p.T * 3 * C, 0, sizeof(float) * 3 * C, st));
    const int groups = max(1, min(sms / nH, B * p.nPairs));
    wmsa_tc_bwd_kernel<<<groups * nH, 256, BWD_SMEM, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
