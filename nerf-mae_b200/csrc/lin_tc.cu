// Linear layers on tcgen05: out[M,N] = epilogue(A[M,K] * W^T) for forward / input-gradient GEMMs, and
// dW[N,K] = dY^T X for weight gradients.  Same building blocks as conv3_tc.cu: fp32 operands are split into
// bf16 hi/lo when staged in shared memory (3 MMAs per k-step, fp32-class accuracy), no-swizzle canonical UMMA
// layouts, weights pre-arranged as blobs and fetched with cp.async.bulk, accumulators in TMEM (double buffered),
// persistent warp-specialised CTAs.
//
// Epilogue fusions (forward / dgrad kernel): + bias, GELU with pre-activation saved, residual + per-sample
// stochastic-depth scale, multiply by GELU'(saved pre-activation), accumulate, and the depth-to-space scatter of a
// kernel==stride transposed convolution.
#include <stdlib.h>

#include <type_traits>

#include "kernels.cuh"
#include "tc.cuh"

using namespace tc;

#define TILE_M 128
#define A_CH ((TILE_M + 1) * 16)   // bytes between the 8-channel chunks of the A operand in shared memory (one pad row)
#define N_PROD 256
#define N_EPI_W 8                  // epilogue warps: two per TMEM lane quadrant (= per SM sub-partition), alternating column groups
#define LT_PROD 192                // lin_tc_kernel: producer threads (6 warps) + MMA warp + weight-loader warp + 8 epilogue warps = 512
#define LT_W_MMA (LT_PROD / 32)
#define LT_W_LOAD (LT_PROD / 32 + 1)
#define LT_W_EPI (LT_PROD / 32 + 2)
#define N_THREADS (LT_PROD + 64 + 32 * N_EPI_W)
#define EPI_G 2                    // accumulator chunks (16 columns each) staged per epilogue flush: 128 contiguous bytes per row
#define EPI_F4 (EPI_G * 4)         // float4 per staged row
#define EPI_RPI (32 / EPI_F4)      // rows covered by one flush instruction of a warp
#define EPI_LD (EPI_G * 16 + 4)    // floats per staged row (+16 bytes: conflict-free 16-byte stores of 32 rows)
#define EPI_STAGE_BYTES (N_EPI_W * (32 * EPI_LD * 4 + 32 * 8 + 256 * 4))   // per epilogue warp: 32 staged rows, their 64-bit output offsets (D2S), the tile's bias
#define MAX_ST 6

// ------------------------------------------------------------------------------------------------ blobs
// [kg][nt][part(hi,lo)][kc][n(NT)][8]  <-  value(n, k)
//   mode 0: w[n*s_n + k*s_k]
//   mode 1: transposed conv forward,  n = ijl*cc + co, k = ci       : w[ci][co][ijl]   (w is (Cin, cc, k3))
//   mode 2: transposed conv dgrad,    n = ci,          k = ijl*cc+co : w[ci][co][ijl]
__global__ void __launch_bounds__(256) lin_tc_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ blob, int N, int K,
                                                          int NT, long long s_n, long long s_k, int mode, int cc, int k3, int KG) {
    long long total = (long long)N * K * 2;
    int ntn = N / NT;
    const int KCH = KG / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int e = (int)(r % 8); r /= 8;
        int n = (int)(r % NT); r /= NT;
        int kc = (int)(r % KCH); r /= KCH;
        int part = (int)(r % 2); r /= 2;
        int nt = (int)(r % ntn); r /= ntn;
        int kg = (int)r;
        const int nn = nt * NT + n, kk = kg * KG + kc * 8 + e;
        long long idx;
        if (mode == 0) idx = (long long)nn * s_n + (long long)kk * s_k;
        else if (mode == 1) idx = (long long)kk * cc * k3 + (long long)(nn % cc) * k3 + nn / cc;
        else idx = (long long)nn * cc * k3 + (long long)(kk % cc) * k3 + kk / cc;
        float v = w[idx];
        __nv_bfloat16 hi = __float2bfloat16_rn(v);
        blob[i] = part == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// One launch re-lays MANY weights (mode 0) into their blobs: table rows of 8 int64 {w, blob, N, K, NT, s_n, s_k, KG}; block
// (entry, slice) -> grid (slices, entries).  Used once per optimiser step for every linear of the Swin blocks instead of one
// lin_tc_prep launch per GEMM call.
__global__ void __launch_bounds__(256) lin_tc_prep_batch_kernel(const long long* __restrict__ table) {
    const long long* t = table + (long long)blockIdx.y * 8;
    const float* w = reinterpret_cast<const float*>(t[0]);
    uint4* blob = reinterpret_cast<uint4*>(t[1]);
    const int N = (int)t[2], K = (int)t[3], NT = (int)t[4], KG = (int)t[7];
    const long long s_n = t[5], s_k = t[6];
    const long long total8 = (long long)N * K * 2 / 8;       // 16-byte units: 8 consecutive k of one (n, part)
    const int ntn = N / NT, KCH = KG / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int n = (int)(r % NT); r /= NT;
        const int kc = (int)(r % KCH); r /= KCH;
        const int part = (int)(r % 2); r /= 2;
        const int nt = (int)(r % ntn); r /= ntn;
        const int kg = (int)r;
        const float* src = w + (long long)(nt * NT + n) * s_n + (long long)(kg * KG + kc * 8) * s_k;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __ldg(src + e * s_k);
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            uint32_t hi, lo;
            split2(v[2 * e], v[2 * e + 1], hi, lo);
            o[e] = part == 0 ? hi : lo;
        }
        blob[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

struct LinTcParams {
    const float* a;       // [M, K] row-major, row stride lda
    long long lda;
    const __nv_bfloat16* wblob;
    GEpilogue e;          // out/ldc/bias/aux/resid/row_scale/flags (+ D2S geometry)
    int M, N, K, NT, n_tiles_n, n_kg, num_m_tiles;
    int a_stage_bytes, b_stage_bytes, stage_bytes, n_st, tmem_cols;
    int a_d2s;            // 1: A rows are gathered from the fine volume of a k==s transposed conv (K index = ijl*C + c)
    int gX, gY, gZ, gC, gld, gks;
    int a_patch;          // 1: A rows are the 4x4x4 patches of a (B,4,R,R,R) grid (k = c*64 + i*16 + j*4 + l), gX = R/4, KG = 32
    int dbg;              // NMAE_DBG experiments: 1 no A loads, 2 no MMAs, 8 no output stores / epilogue reads
};

// KG = channels per pipeline stage: 48 (K % 48 == 0: every width of swin_t/s/l) or 32 (K % 32 == 0: swin_b, the patch embed)
template <int KG>
__global__ void __launch_bounds__(N_THREADS, 1) lin_tc_kernel(const __grid_constant__ LinTcParams p) {
    constexpr int KCH = KG / 8;                                      // 8-channel chunks per stage
    constexpr int UNITS = TILE_M * (KG / 4), F4 = (UNITS + LT_PROD - 1) / LT_PROD;   // float4 units per stage / per producer thread
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_st * p.stage_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](int s) { return bar0 + 8u * s; };
    auto B_FULL = [&](int s) { return bar0 + 8u * (MAX_ST + s); };
    auto S_EMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_ST + s); };
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (3 * MAX_ST + a); };
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (3 * MAX_ST + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_ST + 4);
    float* stage_out = reinterpret_cast<float*>(bars + 3 * MAX_ST + 6);     // [N_EPI_W warps][32 rows][EPI_LD floats], row offsets, bias

    if (tid == 0) {
        for (int s = 0; s < p.n_st; s++) {
            mbar_init(A_FULL(s), LT_PROD / 32);
            mbar_init(B_FULL(s), 1);
            mbar_init(S_EMPTY(s), 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(ACC_FULL(a), 1);
            mbar_init(ACC_EMPTY(a), N_EPI_W);
        }
        fence_barrier_init();
    }
    if (warp == LT_W_MMA) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);
    const int total_work = p.num_m_tiles * p.n_tiles_n;

    if (warp < LT_W_MMA) {
        // =========================================================== A producers
        // Unit = one float4 (4 consecutive k of one row); unit u = j*LT_PROD + tid covers row u / (KG/4), float4 u % (KG/4): a warp's
        // load instruction reads 512 contiguous bytes of 2-3 rows (4-6 cache lines; the first version gave each thread half a row:
        // 16+ lines per instruction, and the kernel's L1/TEX pipe - shared with the epilogue's stores - was the bottleneck).
        // Chunks are (128 + 1) rows apart in shared memory, so the 8-byte stores of a warp conflict at most 2-way.
        // Software-pipelined over the flattened (work item, k-group) stage sequence: the global loads of stage g+1 are in
        // flight while stage g is converted and stored.
        constexpr int RF4 = KG / 4;                 // float4 per row and stage
        int urow[F4], uf4[F4];
#pragma unroll
        for (int j = 0; j < F4; j++) {
            const int u = j * LT_PROD + tid;
            urow[j] = u < UNITS ? u / RF4 : TILE_M;      // past the stage (KG = 32: 1024 units over 192 threads): row >= M, never loaded or stored
            uf4[j] = u - (u / RF4) * RF4;
        }
        const int n_my = blockIdx.x < total_work ? (total_work - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int G = n_my * p.n_kg;
        int s = 0, ph = 0;
        int cached_i = -1;
        long long row_base[F4];                     // element offset of each unit's row (mode specific), -1 past M
        const long long Rg = 4LL * p.gX;            // patch mode: grid resolution
        const int Yk = p.gY * p.gks, Zk = p.gZ * p.gks;
        // (Tried: prefetch.global.L2 probes for stage g+3, one per 64 bytes - to get more than the ~1.5 stages the register double buffer
        // keeps in flight.  Slower everywhere: the decoder1 transposed-convolution backward went 3.1 -> 6.3 ms, stage-3 qkv 38 -> 50 us;
        // profiles/r2_lin_epilogue.txt.)
        auto load_stage = [&](int g, float4 (&v)[F4]) {
            const int i = g / p.n_kg, kg = g - i * p.n_kg;
            if (i != cached_i) {
                cached_i = i;
                const int mt = (blockIdx.x + i * gridDim.x) / p.n_tiles_n;
#pragma unroll
                for (int j = 0; j < F4; j++) {
                    const int m = mt * TILE_M + urow[j];
                    if (m >= p.M || urow[j] >= TILE_M || (p.dbg & 1)) {
                        row_base[j] = -1;
                    } else if (p.a_d2s) {
                        const SpIdx sp = decode_sp(p.gX, p.gY, p.gZ, m);
                        row_base[j] = ((((long long)sp.n * (p.gX * p.gks) + sp.x * p.gks) * Yk + sp.y * p.gks) * Zk + sp.z * p.gks) * p.gld;
                    } else if (KG == 32 && p.a_patch) {
                        const SpIdx sp = decode_sp(p.gX, p.gX, p.gX, m);
                        row_base[j] = ((((long long)sp.n * 4) * Rg + 4 * sp.x) * Rg + 4 * sp.y) * Rg + 4 * sp.z;
                    } else {
                        row_base[j] = (long long)m * p.lda;
                    }
                }
            }
            long long kbase = (long long)kg * KG;   // offset of the stage's first k within a row (plain / transposed-conv gather)
            if (p.a_d2s) {                          // k = ijl*C + c: the whole k-group lies in one fine voxel (C % KG == 0)
                const int k0 = kg * KG, ijl = k0 / p.gC, c0 = k0 - ijl * p.gC;
                const int fi = ijl / (p.gks * p.gks), fj = (ijl / p.gks) % p.gks, fl = ijl % p.gks;
                kbase = (((long long)fi * Yk + fj) * Zk + fl) * p.gld + c0;
            }
#pragma unroll
            for (int j = 0; j < F4; j++) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_base[j] >= 0) {
                    long long off = kbase + uf4[j] * 4;
                    if (KG == 32 && p.a_patch) {    // k = c*64 + i*16 + j*4 + l: one float4 = 4 contiguous z voxels
                        const int k0 = kg * KG + uf4[j] * 4;
                        off = (((long long)(k0 >> 6) * Rg + ((k0 >> 4) & 3)) * Rg + ((k0 >> 2) & 3)) * Rg;
                    }
                    v[j] = __ldg(reinterpret_cast<const float4*>(p.a + row_base[j] + off));
                }
            }
        };
        auto store_stage = [&](const float4 (&v)[F4]) {
            mbar_wait_warp(S_EMPTY(s), ph ^ 1);
            uint8_t* hi_base = smem + (size_t)s * p.stage_bytes;
            uint8_t* lo_base = hi_base + KCH * A_CH;
#pragma unroll
            for (int j = 0; j < F4; j++) {
                if (UNITS % LT_PROD != 0 && urow[j] >= TILE_M) continue;
                uint2 h, l;
                split2(v[j].x, v[j].y, h.x, l.x);
                split2(v[j].z, v[j].w, h.y, l.y);
                const size_t off = (size_t)(uf4[j] >> 1) * A_CH + (size_t)urow[j] * 16 + (size_t)(uf4[j] & 1) * 8;
                *reinterpret_cast<uint2*>(hi_base + off) = h;
                *reinterpret_cast<uint2*>(lo_base + off) = l;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(A_FULL(s));
            if (++s == p.n_st) { s = 0; ph ^= 1; }
        };
        float4 v0[F4], v1[F4];
        if (G > 0) load_stage(0, v0);
        for (int g = 0; g < G; g += 2) {
            if (g + 1 < G) load_stage(g + 1, v1);
            store_stage(v0);
            if (g + 1 < G) {
                if (g + 2 < G) load_stage(g + 2, v0);
                store_stage(v1);
            }
        }
    } else if (warp == LT_W_LOAD) {
        // =========================================================== weight loader (converged warp, elected lane issues)
        {
            int s = 0, ph = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int nt = w % p.n_tiles_n;
                for (int kg = 0; kg < p.n_kg; kg++) {
                    mbar_wait(S_EMPTY(s), ph ^ 1);
                    if (elect_one()) {
                        if (p.dbg & 4) {        // experiment: no weight loads
                            mbar_arrive(B_FULL(s));
                        } else {
                            mbar_expect_tx(B_FULL(s), (uint32_t)p.b_stage_bytes);
                            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wblob) + (size_t)(kg * p.n_tiles_n + nt) * p.b_stage_bytes;
                            bulk_g2s(smem0 + (uint32_t)s * p.stage_bytes + p.a_stage_bytes, src, (uint32_t)p.b_stage_bytes, B_FULL(s));
                        }
                    }
                    __syncwarp();
                    if (++s == p.n_st) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == LT_W_MMA) {
        // =========================================================== MMA issuer (converged warp, elected lane issues)
        {
            const uint32_t idesc = idesc_bf16(TILE_M, p.NT, 0, 0);
            const uint32_t dhi = desc_hi(128), a_lbo = (uint32_t)(A_CH >> 4) << 16, b_lbo = (uint32_t)p.NT << 16;
            const uint32_t b_part16 = ((uint32_t)p.NT * KG * 2u) >> 4;
            int s = 0, ph = 0, it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
                const int acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(ACC_EMPTY(acc), aph ^ 1);
                fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.NT);
                for (int kg = 0; kg < p.n_kg; kg++) {
                    mbar_wait(A_FULL(s), ph);
                    mbar_wait(B_FULL(s), ph);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t a_hi16 = (smem0 + (uint32_t)s * p.stage_bytes) >> 4, a_lo16 = a_hi16 + (KCH * (A_CH >> 4));
                        const uint32_t b_hi16 = a_hi16 + ((uint32_t)p.a_stage_bytes >> 4), b_lo16 = b_hi16 + b_part16;
#pragma unroll
                        for (int ks = 0; ks < KG / 16; ks++) {
                            const uint32_t ao = 2u * ks * (A_CH >> 4), bo = 2u * ks * (uint32_t)p.NT;
                            const uint64_t dah = desc_make(dhi, a_lbo, a_hi16 + ao), dal = desc_make(dhi, a_lbo, a_lo16 + ao);
                            const uint64_t dbh = desc_make(dhi, b_lbo, b_hi16 + bo), dbl = desc_make(dhi, b_lbo, b_lo16 + bo);
                            if (p.dbg & 2) continue;
                            mma_bf16(d_tmem, dah, dbh, idesc, (kg | ks) ? 1u : 0u);
                            mma_bf16(d_tmem, dah, dbl, idesc, 1);
                            mma_bf16(d_tmem, dal, dbh, idesc, 1);
                        }
                        mma_commit(S_EMPTY(s));
                        if (kg == p.n_kg - 1) mma_commit(ACC_FULL(acc));
                    }
                    __syncwarp();
                    if (++s == p.n_st) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // =========================================================== epilogue (warps 8..15)
        // Two warps per TMEM lane quadrant: warp pair (q, half) owns rows q*32.. of the tile and the 32-column groups half, half+2, ...
        // (a single warp per SM sub-partition ran its dependent instruction stream at IPC 0.15 - the epilogue, not the tensor pipe
        // or HBM, bounded every large-M layer; profiles/r2_lin_epilogue.txt)
        const GEpilogue& e = p.e;
        const int q = warp & 3, ew = warp - LT_W_EPI, half = ew >> 2;
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
            const int mt = w / p.n_tiles_n, nt = w % p.n_tiles_n;
            const int acc = it & 1, aph = (it >> 1) & 1;
            const int m = mt * TILE_M + q * 32 + lane;
            const bool valid = m < p.M;
            float rs = 1.f;
            if (valid && (e.flags & EPI_RESID) && e.row_scale) rs = e.row_scale[m / e.rows_per_scale];
            SpIdx sp = {0, 0, 0, 0};
            if (valid && (e.flags & EPI_D2S)) sp = decode_sp(e.X, e.Y, e.Z, m);
            // Per-tile set-up done BEFORE waiting for the accumulator (it used to sit on the critical path of every 16-column chunk:
            // a dependent global load of the bias and, for the depth-to-space scatter, five integer divisions per flush - the single
            // epilogue warp of an SM sub-partition has nothing to hide them behind; decoder1 transposed convolution 1.69 -> ... ms):
            // the tile's bias and the rows' output offsets go to shared memory, the per-lane column offsets to registers.
            float* sw = stage_out + ew * (32 * EPI_LD);
            long long* row_off = reinterpret_cast<long long*>(stage_out + N_EPI_W * 32 * EPI_LD) + ew * 32;
            float* sbias = reinterpret_cast<float*>(reinterpret_cast<long long*>(stage_out + N_EPI_W * 32 * EPI_LD) + N_EPI_W * 32) + ew * 256;
            const int m0 = mt * TILE_M + q * 32;
            // depth-to-space scatter (k == s transposed convolution): column n = tap*C + c of row m lands at the row's voxel offset +
            // the tap's offset + c; the 4 columns of a float4 stay inside one voxel (C % 4 == 0), consecutive float4s are contiguous
            // across the channels of a voxel and (ld == C) across the taps along z
            auto d2s_col = [&](int n) -> long long {
                const int ijl = n / e.C, c = n - ijl * e.C, ks = e.ks;
                const int i = ijl / (ks * ks), jj = (ijl / ks) % ks, l = ijl % ks;
                return (((long long)i * (e.Y * ks) + jj) * (long long)(e.Z * ks) + l) * e.ld + c;
            };
            long long d2s_c0 = 0, d2s_c1 = 0, d2s_c2 = 0, d2s_c3 = 0;     // column offsets of this lane's float4 in this warp's column groups
            if (e.flags & EPI_D2S) {   // offset of the row's coarse voxel (tap 0, channel 0) in the fine volume; read back by the flush
                row_off[lane] = ((((long long)sp.n * (e.X * e.ks) + sp.x * e.ks) * (e.Y * e.ks) + sp.y * e.ks) * (long long)(e.Z * e.ks) +
                                 sp.z * e.ks) * e.ld;
                constexpr int GW = EPI_G * 16;       // columns per group; this warp's i-th group is half + 2*i
                const int nb = nt * p.NT + half * GW + (lane & (EPI_F4 - 1)) * 4;
                d2s_c0 = d2s_col(nb);
                if (p.NT > (half + 2) * GW) d2s_c1 = d2s_col(nb + 2 * GW);
                if (p.NT > (half + 4) * GW) d2s_c2 = d2s_col(nb + 4 * GW);
                if (p.NT > (half + 6) * GW) d2s_c3 = d2s_col(nb + 6 * GW);
            }
            if (e.flags & EPI_BIAS) {
                for (int c = lane; c < p.NT; c += 32) {
                    const int n = nt * p.NT + c;
                    sbias[c] = __ldg(e.bias + ((e.flags & EPI_D2S) ? n % e.C : n));     // D2S: n = tap*C + c, the bias is per channel
                }
            }
            __syncwarp();
            mbar_wait_warp(ACC_FULL(acc), aph);
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
            // the TMEM load of chunk j+1 is in flight while chunk j is processed (the load latency was exposed once per chunk)
            auto process = [&](int j, const uint32_t (&r)[16]) {
                if (!valid || (p.dbg & 8)) return;
                float v[16];
#pragma unroll
                for (int t = 0; t < 16; t++) v[t] = __uint_as_float(r[t]);
                const int n0 = nt * p.NT + j * 16;
                long long idx;
                int bias0 = n0;
                if (e.flags & EPI_D2S) {  // n = ijl*C + c : 16 consecutive channels of one fine voxel
                    const int ijl = n0 / e.C, c = n0 - ijl * e.C, ks = e.ks;
                    const int i = ijl / (ks * ks), jj = (ijl / ks) % ks, l = ijl % ks;
                    idx = ((((long long)sp.n * (e.X * ks) + sp.x * ks + i) * (e.Y * ks) + sp.y * ks + jj) * (long long)(e.Z * ks) +
                           sp.z * ks + l) * e.ld + c;
                    bias0 = c;
                } else {
                    idx = (long long)m * e.ldc + n0;
                }
                if (e.flags & EPI_BIAS) {
#pragma unroll
                    for (int t = 0; t < 16; t++) v[t] += __ldg(e.bias + bias0 + t);
                }
                if (e.flags & EPI_GELU) {
                    float4* a4 = reinterpret_cast<float4*>(e.aux + idx);
#pragma unroll
                    for (int t = 0; t < 4; t++) a4[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
#pragma unroll
                    for (int t = 0; t < 16; t++) v[t] = gelu_erf(v[t]);
                }
                if (e.flags & EPI_GELU_GRAD) {
                    const float4* a4 = reinterpret_cast<const float4*>(e.aux + idx);
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        float4 a = a4[t];
                        v[4 * t] *= gelu_erf_grad(a.x); v[4 * t + 1] *= gelu_erf_grad(a.y);
                        v[4 * t + 2] *= gelu_erf_grad(a.z); v[4 * t + 3] *= gelu_erf_grad(a.w);
                    }
                }
                if (e.flags & EPI_RESID) {
                    const float4* r4 = reinterpret_cast<const float4*>(e.resid + idx);
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        float4 r = r4[t];
                        v[4 * t] = r.x + rs * v[4 * t]; v[4 * t + 1] = r.y + rs * v[4 * t + 1];
                        v[4 * t + 2] = r.z + rs * v[4 * t + 2]; v[4 * t + 3] = r.w + rs * v[4 * t + 3];
                    }
                }
                float4* o4 = reinterpret_cast<float4*>(e.out + idx);
                if ((p.dbg & 512) && !(e.flags & EPI_D2S)) {   // experiment: same bytes, each store instruction covers 512 contiguous bytes
                    float4* c4 = reinterpret_cast<float4*>(e.out + (long long)(mt * TILE_M + q * 32) * e.ldc) + (j * 4) * 32 + lane;
                    if (mt * TILE_M + q * 32 + 32 <= p.M) {
#pragma unroll
                        for (int t = 0; t < 4; t++) c4[t * 32] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                    }
                    return;
                }
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    float4 o = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                    if (e.flags & EPI_ACCUM) {
                        float4 old = o4[t];
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    o4[t] = o;
                }
            };
            // Staged epilogue: the tile is staged EPI_G chunks (32 columns) at a time in shared memory - each lane writes its own row - and
            // flushed with every global instruction of the warp covering 4 rows x 128 contiguous bytes instead of 16 B per lane at a row
            // stride (32 L1 wavefronts per instruction).  The flush also does the arithmetic that needs row-major global operands (see
            // finish() below), so those loads are coalesced too.  History and same-box A/Bs: profiles/r2_lin_epilogue.txt.
            auto stage_chunk = [&](int j, const uint32_t (&r)[16]) {
                float v[16];
#pragma unroll
                for (int t = 0; t < 16; t++) v[t] = __uint_as_float(r[t]);
                if (e.flags & EPI_BIAS) {
                    const float4* b4 = reinterpret_cast<const float4*>(sbias + j * 16);
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const float4 bb = b4[t];
                        v[4 * t] += bb.x; v[4 * t + 1] += bb.y; v[4 * t + 2] += bb.z; v[4 * t + 3] += bb.w;
                    }
                }
                float4* d = reinterpret_cast<float4*>(sw + lane * EPI_LD + (j % EPI_G) * 16);
#pragma unroll
                for (int t = 0; t < 4; t++) d[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
            };
            // FL = the epilogue's operand flags (compile-time per instantiation): the arithmetic that needs row-major global operands
            // (GELU pre-activation out, GELU', residual, accumulate) runs HERE, in the coalesced layout - every global load and store
            // of the flush covers 512 contiguous bytes.  All loads of a batch of rows are issued before the first store (the
            // pointers may alias as far as the compiler knows, so it would not hoist them itself).
            auto finish = [&](auto fl_tag, float4 v, const float4& a, const float4& r, const float4& old, float rsc, long long idx) {
                constexpr int FL = decltype(fl_tag)::value;
                if (FL & EPI_GELU) {
                    *reinterpret_cast<float4*>(e.aux + idx) = v;
                    v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
                }
                if (FL & EPI_GELU_GRAD) {
                    v.x *= gelu_erf_grad(a.x); v.y *= gelu_erf_grad(a.y); v.z *= gelu_erf_grad(a.z); v.w *= gelu_erf_grad(a.w);
                }
                if (FL & EPI_RESID) {
                    v.x = fmaf(rsc, v.x, r.x); v.y = fmaf(rsc, v.y, r.y); v.z = fmaf(rsc, v.z, r.z); v.w = fmaf(rsc, v.w, r.w);
                }
                if (FL & EPI_ACCUM) { v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                *reinterpret_cast<float4*>(e.out + idx) = v;
            };
            auto flush_t = [&](auto fl_tag, int j_last) {
                constexpr int FL = decltype(fl_tag)::value;
                constexpr int NIT = 32 / EPI_RPI;                                    // flush instructions per staged group
                constexpr int TB = (FL == (EPI_GELU_GRAD | EPI_ACCUM)) ? 4 : 8;      // rows per batch (registers: up to 3 float4 per row in flight)
                static_assert(NIT % TB == 0, "flush batches");
                const int g0 = j_last - (j_last % EPI_G), gc = j_last % EPI_G + 1;   // first chunk and number of chunks of the group
                const int ncol0 = nt * p.NT + g0 * 16;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
                if (!(p.dbg & 8)) {
                    if (gc == EPI_G) {          // full group: EPI_F4 float4 per row, EPI_RPI rows per instruction
                        const int f = lane & (EPI_F4 - 1), rsub = lane / EPI_F4;
                        const int grp = (g0 / EPI_G) >> 1;          // index among this warp's groups
                        const long long col_off = !(FL & EPI_D2S) ? 0 : grp == 0 ? d2s_c0 : grp == 1 ? d2s_c1 : grp == 2 ? d2s_c2 : d2s_c3;
#pragma unroll
                        for (int t0 = 0; t0 < NIT; t0 += TB) {
                            float4 o[TB], a[TB], r[TB], old[TB];
                            float rsc[TB];
#pragma unroll
                            for (int t = 0; t < TB; t++) {
                                const int mm = m0 + EPI_RPI * (t0 + t) + rsub;
                                const long long idx = (FL & EPI_D2S) ? row_off[EPI_RPI * (t0 + t) + rsub] + col_off : (long long)mm * e.ldc + ncol0 + f * 4;
                                o[t] = *reinterpret_cast<const float4*>(sw + (EPI_RPI * (t0 + t) + rsub) * EPI_LD + f * 4);
                                a[t] = r[t] = old[t] = z4;
                                rsc[t] = 1.f;
                                if (mm < p.M) {
                                    if (FL & EPI_GELU_GRAD) a[t] = *reinterpret_cast<const float4*>(e.aux + idx);
                                    if (FL & EPI_RESID) {
                                        r[t] = *reinterpret_cast<const float4*>(e.resid + idx);
                                        if (e.row_scale) rsc[t] = e.row_scale[mm / e.rows_per_scale];
                                    }
                                    if (FL & EPI_ACCUM) old[t] = *reinterpret_cast<const float4*>(e.out + idx);
                                }
                            }
#pragma unroll
                            for (int t = 0; t < TB; t++) {
                                const int mm = m0 + EPI_RPI * (t0 + t) + rsub;
                                if (mm < p.M)
                                    finish(fl_tag, o[t], a[t], r[t], old[t], rsc[t],
                                           (FL & EPI_D2S) ? row_off[EPI_RPI * (t0 + t) + rsub] + col_off : (long long)mm * e.ldc + ncol0 + f * 4);
                            }
                        }
                    } else {
                        const int f_per_row = gc * 4, total = 32 * f_per_row;
                        for (int u = lane; u < total; u += 32) {
                            const int row = u / f_per_row, f = u - row * f_per_row, mm = m0 + row;
                            if (mm < p.M) {
                                const long long idx = (FL & EPI_D2S) ? row_off[row] + d2s_col(ncol0 + f * 4) : (long long)mm * e.ldc + ncol0 + f * 4;
                                float4 a = z4, r = z4, old = z4;
                                float rsc = 1.f;
                                if (FL & EPI_GELU_GRAD) a = *reinterpret_cast<const float4*>(e.aux + idx);
                                if (FL & EPI_RESID) {
                                    r = *reinterpret_cast<const float4*>(e.resid + idx);
                                    if (e.row_scale) rsc = e.row_scale[mm / e.rows_per_scale];
                                }
                                if (FL & EPI_ACCUM) old = *reinterpret_cast<const float4*>(e.out + idx);
                                finish(fl_tag, *reinterpret_cast<const float4*>(sw + row * EPI_LD + f * 4), a, r, old, rsc, idx);
                            }
                        }
                    }
                }
                __syncwarp();
            };
            // (Tried: flushing bias-only outputs with one cp.async.bulk shared->global per row and group (256 B each) instead of the
            // float4 loop below - fewer instructions, but slower: decoder1 transposed convolution 1.62 -> 1.93 ms, stage-1 qkv 0.184
            // -> 0.216 ms; profiles/r2_lin_epilogue.txt.)
            const int fl_ops = e.flags & (EPI_GELU | EPI_GELU_GRAD | EPI_RESID | EPI_ACCUM | EPI_D2S);
            auto flush = [&](int j_last) {
                switch (fl_ops) {
                    case 0: flush_t(std::integral_constant<int, 0>{}, j_last); break;
                    case EPI_GELU: flush_t(std::integral_constant<int, EPI_GELU>{}, j_last); break;
                    case EPI_GELU_GRAD: flush_t(std::integral_constant<int, EPI_GELU_GRAD>{}, j_last); break;
                    case EPI_RESID: flush_t(std::integral_constant<int, EPI_RESID>{}, j_last); break;
                    case EPI_ACCUM: flush_t(std::integral_constant<int, EPI_ACCUM>{}, j_last); break;
                    case EPI_D2S: flush_t(std::integral_constant<int, EPI_D2S>{}, j_last); break;
                    default: flush_t(std::integral_constant<int, EPI_GELU_GRAD | EPI_ACCUM>{}, j_last); break;
                }
            };
            {
                const int nch = p.NT / 16;
                // staged + coalesced flush for every row-major epilogue and the depth-to-space scatter; atomic (split) outputs keep
                // the row-per-lane form.  dbg 1024: only the plain outputs are staged (the previous behaviour)
                const bool combo_ok = fl_ops == 0 || fl_ops == EPI_GELU || fl_ops == EPI_GELU_GRAD || fl_ops == EPI_RESID ||
                                      fl_ops == EPI_ACCUM || fl_ops == (EPI_GELU_GRAD | EPI_ACCUM) ||
                                      (fl_ops == EPI_D2S && e.C % 16 == 0 && !(p.dbg & 8192));
                const bool staged = !(e.flags & EPI_ATOMIC) && combo_ok && !((p.dbg & 1024) && fl_ops != 0) &&
                                    !((p.dbg & 2048) && (fl_ops & EPI_GELU_GRAD)) && !((p.dbg & 4096) && (fl_ops & (EPI_RESID | EPI_ACCUM)));
                // this warp's groups: half, half + 2, ...; both chunks of a group are loaded from TMEM together, and the next group's loads
                // are issued as soon as the registers are free (before the flush)
                const int ng = (nch + EPI_G - 1) / EPI_G;
                static_assert(EPI_G == 2, "the chunk loop loads a group as two 16-column chunks");
                uint32_t ra[16], rb[16];
                auto issue_group = [&](int g) {
                    if (g >= ng) return;
                    tmem_ld16_issue(taddr + g * EPI_G * 16, ra);
                    if (g * EPI_G + 1 < nch) tmem_ld16_issue(taddr + (g * EPI_G + 1) * 16, rb);
                };
                issue_group(half);
                for (int g = half; g < ng; g += 2) {
                    const int j0 = g * EPI_G;
                    const bool two = j0 + 1 < nch;
                    tmem_ld16_wait(ra);
                    if (two) tmem_ld16_wait(rb);
                    if (p.dbg & 32768) {        // experiment: TMEM loads only
                        issue_group(g + 2);
                    } else if (staged) {
                        stage_chunk(j0, ra);
                        if (two) stage_chunk(j0 + 1, rb);
                        issue_group(g + 2);
                        flush(two ? j0 + 1 : j0);
                    } else {
                        process(j0, ra);
                        if (two) process(j0 + 1, rb);
                        issue_group(g + 2);
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
    if (warp == LT_W_MMA) {
        fence_after_sync();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

static int pick_nt(int N) {
    for (int nt = 256; nt >= 16; nt -= 16)
        if (N % nt == 0) return nt;
    return 0;
}

// N tile of the forward / dgrad GEMM: the largest divisor of N (multiple of 16, <= 256) when there are many M tiles; for the small
// GEMMs of the deep stages (M = 4000 or 500 rows) the tile that fills the SMs best - cost model: waves x (operand staging + NT)
static int pick_nt_for(int M, int N, int sms) {
    if (nmae_debug_mask() & 256) return pick_nt(N);       // experiment: largest tile always
#ifdef NMAE_DBG
    if (const char* f = getenv("NMAE_FORCE_NT")) {        // experiment (debug library only): a fixed N tile where it divides N
        const int nt = atoi(f);
        if (nt >= 16 && nt % 16 == 0 && N % nt == 0) return nt;
    }
#endif
    const int mt = cdiv(M, TILE_M);
    int best = 0;
    long long best_cost = 0;
    for (int nt = 256; nt >= 16; nt -= 16) {
        if (N % nt != 0) continue;
        const long long items = (long long)mt * (N / nt), waves = (items + sms - 1) / sms;
        const long long cost = waves * (64 + nt);
        if (best == 0 || cost < best_cost) { best = nt; best_cost = cost; }
    }
    return best;
}

static int pick_kg(int K) { return K % 48 == 0 ? 48 : (K % 32 == 0 ? 32 : 0); }

bool k_lin_tc_supported(int M, int N, int K, long long lda, long long ldc) {
    return M >= 1 && pick_kg(K) != 0 && N % 16 == 0 && pick_nt(N) >= 16 && lda % 4 == 0 && ldc % 4 == 0;
}

// out = epi( A[M,K] * Wv^T ),  Wv(n,k) = w[n*s_n + k*s_k]   (forward: s_n=K, s_k=1; input gradient: s_n=1, s_k=ldw)
int k_lin_tc(const float* a, long long lda, const float* w, long long s_n, long long s_k, int M, int N, int K, const GEpilogue& e,
             float* w_ws, cudaStream_t st, int prep_mode, const GOperand* a_gather, bool blob_ready) {
    LinTcParams p;
    memset(&p, 0, sizeof(p));
    int cc = 0, k3 = 0;
    int KG = pick_kg(K);
    if (prep_mode == 1) { cc = e.C; k3 = e.ks * e.ks * e.ks; }
    if (a_gather && a_gather->mode == OPM_PATCH) {
        NMAE_CHECK_ARG(a_gather->ks == 4 && K == 256, "lin_tc: the patch gather is specialised for 4x4x4 patches of 4-channel grids");
        p.a_patch = 1;
        p.gX = a_gather->X;
        KG = 32;
    } else if (a_gather) {
        p.a_d2s = 1;
        p.gX = a_gather->X; p.gY = a_gather->Y; p.gZ = a_gather->Z; p.gC = a_gather->C; p.gld = a_gather->ld; p.gks = a_gather->ks;
        if (KG == 48 && p.gC % 48 != 0) KG = K % 32 == 0 ? 32 : 0;
        NMAE_CHECK_ARG(KG != 0 && p.gC % KG == 0, "lin_tc: gathered transposed-conv operand needs channels %% 48 == 0 or %% 32 == 0");
        cc = p.gC; k3 = p.gks * p.gks * p.gks;
    }
    const int KCH = KG / 8;
    p.a = a; p.lda = lda; p.wblob = reinterpret_cast<const __nv_bfloat16*>(w_ws); p.e = e;
    p.dbg = nmae_debug_mask();
    p.M = M; p.N = N; p.K = K;
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    p.NT = pick_nt_for(M, N, sms);
    NMAE_CHECK_ARG(p.NT >= 16 && KG != 0 && K % KG == 0, "lin_tc: unsupported shape N=%d K=%d", N, K);
    if (e.flags & EPI_D2S) NMAE_CHECK_ARG(e.C % 16 == 0, "lin_tc: D2S needs channel count multiple of 16");
    p.n_tiles_n = N / p.NT;
    p.n_kg = K / KG;
    p.num_m_tiles = cdiv(M, TILE_M);
    p.a_stage_bytes = 2 * KCH * A_CH;
    p.b_stage_bytes = p.NT * KG * 2 * 2;
    p.stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    int tm = 2 * p.NT;
    p.tmem_cols = tm <= 32 ? 32 : tm <= 64 ? 64 : tm <= 128 ? 128 : tm <= 256 ? 256 : 512;
    const int bar_bytes = 8 * (3 * MAX_ST + 6) + EPI_STAGE_BYTES;
    const int max_smem = 227 * 1024;
    p.n_st = (max_smem - bar_bytes) / p.stage_bytes;
    if (p.n_st > MAX_ST) p.n_st = MAX_ST;
    NMAE_CHECK_ARG(p.n_st >= 2, "lin_tc: stage does not fit");
    size_t smem = (size_t)p.n_st * p.stage_bytes + bar_bytes;

    long long total = (long long)N * K * 2;
    int g = (int)min((long long)148 * 8, (total + 255) / 256);
    if (!blob_ready) {
        lin_tc_prep_kernel<<<g, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16*>(w_ws), N, K, p.NT, s_n, s_k, prep_mode, cc, k3, KG);
        NMAE_LAUNCH_CHECK();
    }

    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(lin_tc_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        NMAE_CUDA(cudaFuncSetAttribute(lin_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set[dev] = true;
    }
    if (KG == 48) lin_tc_kernel<48><<<min(sms, p.num_m_tiles * p.n_tiles_n), N_THREADS, smem, st>>>(p);
    else lin_tc_kernel<32><<<min(sms, p.num_m_tiles * p.n_tiles_n), N_THREADS, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// tile / k-group the forward / dgrad kernel uses for an (M, N, K) GEMM: the blob layout depends on both
int k_lin_tc_tile(int M, int N, int K, int* nt, int* kg) {
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    *nt = pick_nt_for(M, N, sms);
    *kg = pick_kg(K);
    return NMAE_OK;
}

int k_lin_tc_prep_batch(const long long* table, int n, long long max_elems, cudaStream_t st) {
    if (n <= 0) return NMAE_OK;
    const int gx = (int)max(1LL, min(64LL, (2 * max_elems + 256 * 8 - 1) / (256 * 8)));
    lin_tc_prep_batch_kernel<<<dim3(gx, n), 256, 0, st>>>(table);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ================================================================================================ weight gradient
// dW[n][k] += sum_m dY[m][n] * X[m][k].  Both operands MN-major (reduction over rows).  A CTA owns a
// (128 x-features) x (NT dy-features) accumulator over a contiguous range of 64-row stages, then flushes with atomics.
#define WG_ROWS 64
#define WG_XCH 16   // x-feature chunks per tile (128 features)
#define WG_CSTRIDE ((WG_ROWS + 1) * 16)   // bytes between 8-feature chunks in shared memory (one pad row: conflict-free stores)

struct LinWgParams {
    const float* x;   // [M, K]
    const float* dy;  // [M, N]
    float* dw;        // [N, K]
    long long ldx, ldy;
    int M, N, K, NT, n_tiles_n, n_kb, n_chunks, splits, num_items;
    int x_part_bytes, y_part_bytes, stage_bytes;
    int x_d2s;        // 1: x-side features are gathered from the fine volume of a k==s transposed conv (k = ijl*C + co) and
                      //    dw is the transposed-conv weight (n=ci, co, ijl)
    int gX, gY, gZ, gC, gld, gks;
    int x_patch;      // 1: x rows are the 4x4x4 patches of a (B,4,R,R,R) grid (k = c*64 + i*16 + j*4 + l), gX = R/4
};

__global__ void __launch_bounds__(416, 1) lin_wgrad_tc_kernel(const __grid_constant__ LinWgParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)p.stage_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto ST_FULL = [&](int s) { return bar0 + 8u * s; };
    auto ST_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (4 + a); };
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (6 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(ST_FULL(s), N_PROD / 32);
            mbar_init(ST_EMPTY(s), 1);
            mbar_init(ACC_FULL(s), 1);
            mbar_init(ACC_EMPTY(s), 4);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);
    const int ychunks = p.NT / 8;

    auto item_decode = [&](int item, int& kb, int& nt, int& c_beg, int& c_end) {
        int ident = item / p.splits, sp = item - ident * p.splits;
        nt = ident % p.n_tiles_n;
        kb = ident / p.n_tiles_n;
        c_beg = (int)((long long)p.n_chunks * sp / p.splits);
        c_end = (int)((long long)p.n_chunks * (sp + 1) / p.splits);
    };

    if (warp < 8) {
        // producers: unit = (row, 8-feature chunk) -> one 32-byte load, one 16-byte hi + lo store.  A warp covers 4 rows x 8
        // consecutive chunks, i.e. 256 contiguous bytes per row: a load instruction touches 8 cache lines (the first version had
        // consecutive threads on consecutive ROWS - 32 lines per instruction - and ran at 89 % L1/TEX throughput, 17 % tensor pipe).
        // Chunks are WG_CSTRIDE = (64 + 1) rows apart in shared memory, which makes the 16-byte stores of 8 chunks x 4 rows
        // conflict-free.  All loads of a stage are issued before the stage buffer is waited for and before the first conversion.
        constexpr int YP = 4;                      // dy chunk passes: NT/8 <= 32 chunks, 8 per pass
        int s = 0, ph = 0;
        const int c8 = tid & 7, r32 = tid >> 3;    // chunks c8 + 8*ci, rows r32 + 32*ri
        const int Yk = p.gY * p.gks, Zk = p.gZ * p.gks;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int kb, nt, c_beg, c_end;
            item_decode(item, kb, nt, c_beg, c_end);
            long long feat[2];       // element offset of x-feature chunk c8 + 8*ci within a row, or -1 past K
#pragma unroll
            for (int ci = 0; ci < 2; ci++) {
                const int k = kb * 128 + (c8 + 8 * ci) * 8;
                feat[ci] = -1;
                if (k < p.K) {
                    if (p.x_patch) {    // 8 features = rows j0, j0+1 of 4 contiguous z voxels of patch plane (c, i)
                        const long long R = 4 * p.gX;
                        feat[ci] = (((long long)(k >> 6) * R + ((k >> 4) & 3)) * R + ((k >> 2) & 3)) * R;
                    } else if (p.x_d2s) {      // k = ijl*C + c  ->  fine voxel (i, j, l) of the coarse voxel, channel c
                        const int ijl = k / p.gC, c = k - ijl * p.gC;
                        const int i = ijl / (p.gks * p.gks), j = (ijl / p.gks) % p.gks, l = ijl % p.gks;
                        feat[ci] = (((long long)i * Yk + j) * Zk + l) * p.gld + c;
                    } else {
                        feat[ci] = k;
                    }
                }
            }
            const long long x_second = p.x_patch ? 4LL * p.gX : 4;       // float offset of the second 16 bytes of an x unit
            auto x_row_off = [&](long long m) -> long long {
                if (p.x_d2s) {
                    const SpIdx sp = decode_sp(p.gX, p.gY, p.gZ, (int)m);
                    return ((((long long)sp.n * (p.gX * p.gks) + sp.x * p.gks) * Yk + sp.y * p.gks) * Zk + sp.z * p.gks) * p.gld;
                }
                if (p.x_patch) {
                    const SpIdx sp = decode_sp(p.gX, p.gX, p.gX, (int)m);
                    const long long R = 4 * p.gX;
                    return ((((long long)sp.n * 4) * R + 4 * sp.x) * R + 4 * sp.y) * R + 4 * sp.z;
                }
                return m * p.ldx;
            };
            for (int ch = c_beg; ch < c_end; ch++) {
                float4 vx[2][2][2], vy[YP][2][2];      // [chunk pass][row pass][first / second 16 bytes]
                long long mrow[2];
                bool mvalid[2];
#pragma unroll
                for (int ri = 0; ri < 2; ri++) {
                    const long long m = (long long)ch * WG_ROWS + r32 + 32 * ri;
                    mrow[ri] = m;
                    mvalid[ri] = m < p.M;
                    const long long xrow = mvalid[ri] ? x_row_off(m) : 0;
#pragma unroll
                    for (int ci = 0; ci < 2; ci++) {
                        vx[ci][ri][0] = vx[ci][ri][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (mvalid[ri] && feat[ci] >= 0) {
                            const float* src = p.x + xrow + feat[ci];
                            vx[ci][ri][0] = __ldg(reinterpret_cast<const float4*>(src));
                            vx[ci][ri][1] = __ldg(reinterpret_cast<const float4*>(src + x_second));
                        }
                    }
#pragma unroll
                    for (int ci = 0; ci < YP; ci++) {
                        vy[ci][ri][0] = vy[ci][ri][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        const int chunk = c8 + 8 * ci;
                        if (mvalid[ri] && chunk < ychunks) {
                            const float* src = p.dy + m * p.ldy + nt * p.NT + chunk * 8;
                            vy[ci][ri][0] = __ldg(reinterpret_cast<const float4*>(src));
                            vy[ci][ri][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
                        }
                    }
                }
                mbar_wait_warp(ST_EMPTY(s), ph ^ 1);
                uint8_t* xh = smem + (size_t)s * p.stage_bytes;
                uint8_t* xl = xh + p.x_part_bytes;
                uint8_t* yh = xl + p.x_part_bytes;
                uint8_t* yl = yh + p.y_part_bytes;
                auto put = [&](uint8_t* dh, uint8_t* dl, const float4& a, const float4& b) {
                    uint4 h, l;
                    split2(a.x, a.y, h.x, l.x);
                    split2(a.z, a.w, h.y, l.y);
                    split2(b.x, b.y, h.z, l.z);
                    split2(b.z, b.w, h.w, l.w);
                    *reinterpret_cast<uint4*>(dh) = h;
                    *reinterpret_cast<uint4*>(dl) = l;
                };
#pragma unroll
                for (int ri = 0; ri < 2; ri++) {
                    const size_t roff = (size_t)(r32 + 32 * ri) * 16;
#pragma unroll
                    for (int ci = 0; ci < 2; ci++) {
                        const size_t off = (size_t)(c8 + 8 * ci) * WG_CSTRIDE + roff;
                        put(xh + off, xl + off, vx[ci][ri][0], vx[ci][ri][1]);
                    }
#pragma unroll
                    for (int ci = 0; ci < YP; ci++) {
                        if (c8 + 8 * ci < ychunks) {
                            const size_t off = (size_t)(c8 + 8 * ci) * WG_CSTRIDE + roff;
                            put(yh + off, yl + off, vy[ci][ri][0], vy[ci][ri][1]);
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(ST_FULL(s));
                if (++s == 2) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 8) {
        {
            const uint32_t idesc = idesc_bf16(128, p.NT, 1, 1);
            const uint32_t mhi = desc_hi(WG_CSTRIDE), lbo = (128u >> 4) << 16;
            int s = 0, ph = 0, it = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
                int kb, nt, c_beg, c_end;
                item_decode(item, kb, nt, c_beg, c_end);
                const int acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(ACC_EMPTY(acc), aph ^ 1);
                fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)(acc * 256);
                for (int ch = c_beg; ch < c_end; ch++) {
                    mbar_wait(ST_FULL(s), ph);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t xh16 = (smem0 + (uint32_t)s * p.stage_bytes) >> 4, xl16 = xh16 + ((uint32_t)p.x_part_bytes >> 4);
                        const uint32_t yh16 = xl16 + ((uint32_t)p.x_part_bytes >> 4), yl16 = yh16 + ((uint32_t)p.y_part_bytes >> 4);
#pragma unroll
                        for (int ks = 0; ks < WG_ROWS / 16; ks++) {
                            const uint32_t o = (uint32_t)(16 * ks);
                            const uint64_t axh = desc_make(mhi, lbo, xh16 + o), axl = desc_make(mhi, lbo, xl16 + o);
                            const uint64_t byh = desc_make(mhi, lbo, yh16 + o), byl = desc_make(mhi, lbo, yl16 + o);
                            mma_bf16(d, axh, byh, idesc, (ch == c_beg && ks == 0) ? 0u : 1u);
                            mma_bf16(d, axh, byl, idesc, 1);
                            mma_bf16(d, axl, byh, idesc, 1);
                        }
                        mma_commit(ST_EMPTY(s));
                        if (ch == c_end - 1) mma_commit(ACC_FULL(acc));
                    }
                    __syncwarp();
                    if (++s == 2) { s = 0; ph ^= 1; }
                }
                if (c_end <= c_beg) {
                    if (elect_one()) mma_commit(ACC_FULL(acc));
                    __syncwarp();
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, it++) {
            int kb, nt, c_beg, c_end;
            item_decode(item, kb, nt, c_beg, c_end);
            const int acc = it & 1, aph = (it >> 1) & 1;
            const int k = kb * 128 + row;
            mbar_wait_warp(ACC_FULL(acc), aph);
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
            for (int j = 0; j < p.NT / 16; j++) {
                float v[16];
                tmem_ld16(taddr + j * 16, v);
                if (k < p.K && c_end > c_beg) {
                    if (p.x_d2s) {
                        const int k3 = p.gks * p.gks * p.gks, ijl = k / p.gC, co = k - ijl * p.gC;
#pragma unroll
                        for (int t = 0; t < 16; t++)
                            atomicAdd(p.dw + ((long long)(nt * p.NT + j * 16 + t) * p.gC + co) * k3 + ijl, v[t]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 16; t++) atomicAdd(p.dw + (long long)(nt * p.NT + j * 16 + t) * p.K + k, v[t]);
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
    if (warp == 8) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

bool k_lin_wgrad_tc_supported(int M, int N, int K, long long ldx, long long ldy) {
    return N % 16 == 0 && K % 8 == 0 && pick_nt(N) >= 16 && ldx % 4 == 0 && ldy % 4 == 0 && M >= 1;
}

// dw [N,K] overwritten
int k_lin_wgrad_tc(const float* x, long long ldx, const float* dy, long long ldy, int M, int N, int K, float* dw, cudaStream_t st,
                   const GOperand* x_gather) {
    LinWgParams p;
    memset(&p, 0, sizeof(p));
    if (x_gather && x_gather->mode == OPM_PATCH) {
        NMAE_CHECK_ARG(x_gather->ks == 4 && K == 256, "lin_wgrad_tc: the patch gather is specialised for 4x4x4 patches of 4-channel grids");
        p.x_patch = 1;
        p.gX = x_gather->X;
    } else if (x_gather) {
        p.x_d2s = 1;
        p.gX = x_gather->X; p.gY = x_gather->Y; p.gZ = x_gather->Z; p.gC = x_gather->C; p.gld = x_gather->ld; p.gks = x_gather->ks;
        NMAE_CHECK_ARG(p.gC % 8 == 0 && p.gld % 4 == 0, "lin_wgrad_tc: gathered operand needs channels %% 8 == 0");
    }
    p.x = x; p.dy = dy; p.dw = dw; p.ldx = ldx; p.ldy = ldy;
    p.M = M; p.N = N; p.K = K;
    p.NT = pick_nt(N);
    NMAE_CHECK_ARG(p.NT >= 16, "lin_wgrad_tc: unsupported N=%d", N);
    p.n_tiles_n = N / p.NT;
    p.n_kb = cdiv(K, 128);
    p.n_chunks = cdiv(M, WG_ROWS);
    int ident = p.n_kb * p.n_tiles_n;
    int dev, sms = 148;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    p.splits = max(1, min(p.n_chunks, (2 * sms) / ident));   // rounded down: no CTA gets a third item
    if (ident >= 2 * sms) p.splits = 1;
    p.num_items = ident * p.splits;
    p.x_part_bytes = WG_XCH * WG_CSTRIDE;
    p.y_part_bytes = (p.NT / 8) * WG_CSTRIDE;
    p.stage_bytes = 2 * p.x_part_bytes + 2 * p.y_part_bytes;
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    const int smem = 2 * p.stage_bytes + 128;
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(lin_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    lin_wgrad_tc_kernel<<<min(sms, p.num_items), 416, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
