// Single-pass 3x3x3 convolution (forward and dgrad) on tcgen05 with fp16 operands and fp32 accumulation ("fp16" precision
// mode: 11-bit significands, the class of the TF32 convolutions the reference's GPU path uses by default; the three-pass bf16
// hi/lo kernel of conv3_tc.cu is the fp32-class "bf16x3" mode).
//
// GEMM view.  Rows = 128 consecutive positions of one (batch, x, z-strip) plane (position space of uimg.cuh), K = 9 (dx,dy)
// taps x C input channels, N = 3 dz taps x 48|64 output channels: the dz = -1/0/+1 taps are THREE COLUMN BLOCKS of one
// N = 144|192 instruction, because a 48-column instruction cannot amortise its 4 KB A-operand fetch from shared memory
// (tools/mma_bench.cu: N=48 50 cycles, N=144 76 cycles per MMA).  Column block dz of accumulator row r holds
// sum X[q_r + dy*ZP][c] W[dx,dy,dz][c][co] with q_r the position of row r; the output at row r is
//     out[r] = acc[dz=-1][r-1] + acc[dz=0][r] + acc[dz=+1][r+1]
// which the epilogue forms with one warp shuffle up / down per value (the two rows next to a warp boundary travel through
// shared memory).  Rows 0 and 127 have no complete sum, so tiles advance by 126 positions (UIMGH_STRIDE).
//
// Data movement.  A CTA marches along x through a segment of L tiles of one (batch, strip, position tile) column: tile x needs the
// image planes x-1, x, x+1, so each step fetches ONE new plane (ring of 4 x n_cg plane buffers) - every input element crosses
// L2->SM about once instead of three times.  When the 27 weight blobs of the CTA's output tile fit in shared memory next to the
// ring (C = 48: 124 KB) they are loaded once and stay resident; otherwise they stream through a stage ring.  All staging is
// cp.async.bulk from the pre-built image / blob tensors; the nine (dx,dy) taps of a plane are descriptor start offsets.
//
// Roles (384 threads): warp 0 image loader, warp 1 MMA issuer (+TMEM alloc), warp 2 weight loader, warps 4-7 and 8-11 two
// epilogue groups (TMEM accumulator double buffered, one group per buffer: an epilogue overlaps the MMAs of the next two tiles).
#include "kernels.cuh"
#include "tc.cuh"
#include "uimg.cuh"

using namespace tc;

#define H_MAX_RING 8
#define H_MAX_BST 12
#define H_TILE_M 128

struct ConvHParams {
    const uint8_t* uimg;
    long long u_chunk_bytes, u_img_bytes;
    float* y;
    const float* bias;
    const float* out_scale;      // device scalar multiplied into the result (reciprocal of the gradient image's scale) or NULL
    const uint8_t* wblob;
    int B, Dx, Dy, Dz, C, N, n_cg, n_nt;
    int SW, n_strips, ZP, P, tpp, R_s;
    int L, n_seg, num_items;
    int plane_bytes, n_ring, b_stage_bytes, n_bst, resident, accumulate;
    int dbg;   // nmae_debug_mask(): 1 no image copies, 2 no MMAs, 8 no output stores, 16 no row exchange, 32 no TMEM loads
};

// instruction descriptor, kind::f16 with fp16 operands (format 0) and fp32 accumulation, K-major A and B
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// weight blobs: [dx][cg][dy][nt][kc][n = dz*CG + co][8 x f16]  <-  value(co, c, tap) = w[co*s_n + c*s_c + tap']
template <int CG>
__global__ void __launch_bounds__(256) conv3_h_prep_kernel(const float* __restrict__ w, __half* __restrict__ blob, int C, int N,
                                                           long long s_n, long long s_c, int flip) {
    const long long total = 27LL * C * N;
    const int n_cg = C / CG, n_nt = N / CG, KCH = CG / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % (3 * CG)); r /= 3 * CG;
        const int kc = (int)(r % KCH); r /= KCH;
        const int nt = (int)(r % n_nt); r /= n_nt;
        const int dy = (int)(r % 3); r /= 3;
        const int cg = (int)(r % n_cg); r /= n_cg;
        const int dx = (int)r;
        const int dz = n / CG, co = n - dz * CG;
        int tap = dx * 9 + dy * 3 + dz;
        if (flip) tap = 26 - tap;
        blob[i] = __float2half_rn(w[(long long)(nt * CG + co) * s_n + (long long)(cg * CG + kc * 8 + e) * s_c + tap]);
    }
}

template <int CG>
__global__ void __launch_bounds__(384, 1) conv3_h_kernel(const __grid_constant__ ConvHParams p) {
    constexpr int KCH = CG / 8, N3 = 3 * CG, KSTEPS = CG / 16;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    uint8_t* ring = smem;                                              // [n_ring][plane_bytes]
    uint8_t* bst = ring + (size_t)p.n_ring * p.plane_bytes;            // [n_bst][b_stage_bytes]
    float* xch = reinterpret_cast<float*>(bst + (size_t)p.n_bst * p.b_stage_bytes);   // [group 2][parity 2][warp 4][2][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 2 * 4 * 2 * 16);
    const uint32_t bar0 = smem_u32(bars);
    auto IMG_FULL = [&](int b) { return bar0 + 8u * (0 + b); };
    auto IMG_EMPTY = [&](int b) { return bar0 + 8u * (H_MAX_RING + b); };
    auto B_FULL = [&](int s) { return bar0 + 8u * (2 * H_MAX_RING + s); };
    auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 * H_MAX_RING + H_MAX_BST + s); };
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * H_MAX_RING + 2 * H_MAX_BST + a); };
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * H_MAX_RING + 2 * H_MAX_BST + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * H_MAX_RING + 2 * H_MAX_BST + 4);

    if (tid == 0) {
        for (int b = 0; b < p.n_ring; b++) { mbar_init(IMG_FULL(b), 1); mbar_init(IMG_EMPTY(b), 1); }
        for (int s = 0; s < p.n_bst; s++) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ring0 = smem_u32(ring), bst0 = smem_u32(bst);

    // item -> (nt, segment, position tile, strip, batch)
    auto decode = [&](int item, int& nt, int& x0, int& x1, int& t, int& strip, int& b) {
        int r = item;
        nt = r % p.n_nt; r /= p.n_nt;
        const int seg = r % p.n_seg; r /= p.n_seg;
        t = r % p.tpp; r /= p.tpp;
        strip = r % p.n_strips;
        b = r / p.n_strips;
        x0 = seg * p.L;
        x1 = min(p.Dx, x0 + p.L);
    };

    if (warp == 0) {
        // =========================================================== image loader: one plane = KCH bulk copies
        int slot = 0, ph = 0;
        const uint32_t row_bytes = (uint32_t)p.R_s * 16u;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int nt, x0, x1, t, strip, b;
            decode(item, nt, x0, x1, t, strip, b);
            const int pfirst = max(x0 - 1, 0), plast = min(x1, p.Dx - 1);
            for (int xx = pfirst; xx <= plast; xx++) {
                for (int cg = 0; cg < p.n_cg; cg++) {
                    mbar_wait(IMG_EMPTY(slot), ph ^ 1);
                    if (elect_one()) {
                        const uint8_t* src = p.uimg + ((((long long)(b * (p.Dx + 2) + xx + 1) * p.n_strips + strip) * p.n_cg + cg)) * p.u_img_bytes +
                                             (long long)t * UIMGH_STRIDE * 16;
                        const uint32_t dst = ring0 + (uint32_t)slot * p.plane_bytes;
                        if (p.dbg & 1) {
                            mbar_arrive(IMG_FULL(slot));
                        } else {
                            mbar_expect_tx(IMG_FULL(slot), (uint32_t)p.plane_bytes);
#pragma unroll
                            for (int c = 0; c < KCH; c++)
                                bulk_g2s(dst + (uint32_t)c * row_bytes, src + c * p.u_chunk_bytes, row_bytes, IMG_FULL(slot));
                        }
                    }
                    __syncwarp();
                    if (++slot == p.n_ring) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // =========================================================== weight loader
        if (p.resident) {
            if (blockIdx.x < p.num_items && elect_one()) {
                for (int s = 0; s < 9 * p.n_cg; s++) {
                    mbar_expect_tx(B_FULL(s), (uint32_t)p.b_stage_bytes);
                    bulk_g2s(bst0 + (uint32_t)s * p.b_stage_bytes, p.wblob + (size_t)s * p.b_stage_bytes, (uint32_t)p.b_stage_bytes, B_FULL(s));
                }
            }
            __syncwarp();
        } else {
            int s = 0, ph = 0;
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int nt, x0, x1, t, strip, b;
                decode(item, nt, x0, x1, t, strip, b);
                for (int x = x0; x < x1; x++) {
                    for (int dx = 0; dx < 3; dx++) {
                        const int xx = x + dx - 1;
                        if (xx < 0 || xx >= p.Dx) continue;
                        for (int cg = 0; cg < p.n_cg; cg++) {
                            for (int dy = 0; dy < 3; dy++) {
                                mbar_wait(B_EMPTY(s), ph ^ 1);
                                if (elect_one()) {
                                    mbar_expect_tx(B_FULL(s), (uint32_t)p.b_stage_bytes);
                                    const uint8_t* src = p.wblob + ((size_t)((dx * p.n_cg + cg) * 3 + dy) * p.n_nt + nt) * p.b_stage_bytes;
                                    bulk_g2s(bst0 + (uint32_t)s * p.b_stage_bytes, src, (uint32_t)p.b_stage_bytes, B_FULL(s));
                                }
                                __syncwarp();
                                if (++s == p.n_bst) { s = 0; ph ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (whole warp converged, one elected lane issues)
        const uint32_t idesc = idesc_f16(H_TILE_M, N3);
        const uint32_t dhi = desc_hi(128);                                  // SBO = 128 B between 8-row groups (A and B)
        const uint32_t a_lbo = (uint32_t)p.R_s << 16, b_lbo = (uint32_t)N3 << 16;   // LBO (chunk stride) in 16-byte units, pre-shifted
        const uint32_t a_k2 = 2u * (uint32_t)p.R_s, b_k2 = 2u * (uint32_t)N3;       // one k-step = 2 chunks
        const uint32_t plane16 = (uint32_t)p.plane_bytes >> 4, bst16 = (uint32_t)p.b_stage_bytes >> 4;
        int s = 0, bph = 0, it = 0;
        uint32_t b_waited = 0;          // resident mode: stages whose arrival has been observed
        // ring position (slot, phase) of the first plane the current tile uses; planes occupy consecutive slots in load order,
        // so everything is tracked with adds and compare-subtract wraps (a division here stalls the single issuing lane)
        int win_slot = 0, win_ph = 0;
        auto advance = [&](int& sl, int& ph, int n) {
            sl += n;
            while (sl >= p.n_ring) { sl -= p.n_ring; ph ^= 1; }
        };
        if (p.resident) {
            // ------------------------------------------------------- lean path: weights resident, one channel group.
            // The issuing lane is the throughput limit of this kernel (an N=144 MMA lasts ~76 cycles): all waits of a tile are
            // taken up front and its 27 MMAs are issued from ONE elected region, each descriptor a base word plus a constant.
            for (int bs = 0; bs < 9; bs++) mbar_wait(B_FULL(bs), 0);
            fence_after_sync();
            b_waited = 0x1ffu;
            uint32_t a_off[3 * KSTEPS];                  // (dy, k-step) offset of the A start address inside a plane
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) a_off[dy * KSTEPS + ks] = (uint32_t)(dy * p.ZP) + (uint32_t)ks * a_k2;
            const uint32_t b_base = b_lbo + (bst0 >> 4);
            const uint32_t a_base = a_lbo + (ring0 >> 4);
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int nt, x0, x1, t, strip, b;
                decode(item, nt, x0, x1, t, strip, b);
                const int plast = min(x1, p.Dx - 1);
                int win_plane = max(x0 - 1, 0);
                for (int x = x0; x < x1; x++, it++) {
                    const int acc = it & 1, aph = (it >> 1) & 1;
                    if (x - 1 > win_plane) { advance(win_slot, win_ph, 1); win_plane++; }
                    // ring slots of planes x-1, x, x+1 (consecutive from the window start; plane -1 does not exist)
                    int sl[3], sp[3];
                    {
                        int c_sl = win_slot, c_ph = win_ph;
#pragma unroll
                        for (int dx = 0; dx < 3; dx++) {
                            const int xx = x + dx - 1;
                            sl[dx] = c_sl; sp[dx] = c_ph;
                            if (xx >= 0 && xx < p.Dx) {
                                if (++c_sl == p.n_ring) { c_sl = 0; c_ph ^= 1; }
                            }
                        }
                    }
                    if (x == x0) {
                        if (x - 1 >= 0) mbar_wait(IMG_FULL(sl[0]), sp[0]);
                        mbar_wait(IMG_FULL(sl[1]), sp[1]);
                    }
                    if (x + 1 < p.Dx) mbar_wait(IMG_FULL(sl[2]), sp[2]);
                    mbar_wait(ACC_EMPTY(acc), aph ^ 1);
                    fence_after_sync();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                        uint32_t accum = 0;
                        const bool last = x == x1 - 1;
#pragma unroll
                        for (int dx = 0; dx < 3; dx++) {
                            const int xx = x + dx - 1;
                            if (xx < 0 || xx >= p.Dx) continue;
                            const uint32_t a_plane = a_base + (uint32_t)sl[dx] * plane16;
                            if (!(p.dbg & 2)) {
#pragma unroll
                                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                                    for (int ks = 0; ks < KSTEPS; ks++) {
                                        mma_bf16(d_tmem, desc_pack(a_plane + a_off[dy * KSTEPS + ks], dhi),
                                                 desc_pack(b_base + (uint32_t)((dx * 3 + dy) * (KCH * N3) + ks * 2 * N3), dhi), idesc, accum);
                                        accum = 1;
                                    }
                            }
                            if (dx == 0 || last) mma_commit(IMG_EMPTY(sl[dx]));
                        }
                        mma_commit(ACC_FULL(acc));
                    }
                    __syncwarp();
                }
                advance(win_slot, win_ph, plast - win_plane + 1);
            }
        } else
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int nt, x0, x1, t, strip, b;
            decode(item, nt, x0, x1, t, strip, b);
            const int plast = min(x1, p.Dx - 1);
            int win_plane = max(x0 - 1, 0);               // plane held by ring position (win_slot, win_ph)
            for (int x = x0; x < x1; x++, it++) {
                const int acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait(ACC_EMPTY(acc), aph ^ 1);
                fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                uint32_t accum = 0;
                if (x - 1 > win_plane) { advance(win_slot, win_ph, p.n_cg); win_plane++; }
                int slot = win_slot, iph = win_ph;
                for (int dx = 0; dx < 3; dx++) {
                    const int xx = x + dx - 1;
                    if (xx < 0 || xx >= p.Dx) continue;
                    for (int cg = 0; cg < p.n_cg; cg++) {
                        if (x == x0 || dx == 2) {          // planes x-1 and x of later tiles were observed at the previous tile
                            mbar_wait(IMG_FULL(slot), iph);
                            fence_after_sync();
                        }
                        const uint32_t a_plane = a_lbo + ((ring0 >> 4) + (uint32_t)slot * plane16);
                        const bool release = dx == 0 || x == x1 - 1;
                        for (int dy = 0; dy < 3; dy++) {
                            int bs;
                            if (p.resident) {
                                bs = (dx * p.n_cg + cg) * 3 + dy;
                                if (!(b_waited >> bs & 1u)) {
                                    mbar_wait(B_FULL(bs), 0);
                                    fence_after_sync();
                                    b_waited |= 1u << bs;
                                }
                            } else {
                                bs = s;
                                mbar_wait(B_FULL(bs), bph);
                                fence_after_sync();
                            }
                            if (elect_one()) {
                                const uint32_t a0 = a_plane + (uint32_t)(dy * p.ZP);
                                const uint32_t b0 = b_lbo + ((bst0 >> 4) + (uint32_t)bs * bst16);
                                if (!(p.dbg & 2)) {
#pragma unroll
                                    for (int ks = 0; ks < KSTEPS; ks++) {
                                        mma_bf16(d_tmem, desc_pack(a0 + (uint32_t)ks * a_k2, dhi), desc_pack(b0 + (uint32_t)ks * b_k2, dhi), idesc, accum);
                                        accum = 1;
                                    }
                                }
                                if (!p.resident) mma_commit(B_EMPTY(bs));
                                if (dy == 2 && release) mma_commit(IMG_EMPTY(slot));
                            }
                            __syncwarp();
                            accum = 1;
                            if (!p.resident && ++s == p.n_bst) { s = 0; bph ^= 1; }
                        }
                        if (++slot == p.n_ring) { slot = 0; iph ^= 1; }
                    }
                }
                if (elect_one()) mma_commit(ACC_FULL(acc));
                __syncwarp();
            }
            // next segment starts right behind this segment's last plane
            advance(win_slot, win_ph, (plast - win_plane + 1) * p.n_cg);
        }
        if (p.resident && blockIdx.x < p.num_items) {      // never leave with a weight copy still in flight (Dx == 1 skips taps)
            for (int bs = 0; bs < 9 * p.n_cg; bs++)
                if (!(b_waited >> bs & 1u)) mbar_wait(B_FULL(bs), 0);
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue (4 warps, one TMEM lane quarter each)
        // two groups of four warps: group g drains accumulator g (tiles with it & 1 == g), so each tile's epilogue has two tile
        // times; a warp may only read the TMEM lane quarter warp % 4
        const int q = warp & 3, grp = (warp - 4) >> 2;
        const int m = q * 32 + lane;
        const float oscale = p.out_scale ? __ldg(p.out_scale) : 1.f;
        float* xch_g = xch + grp * (2 * 4 * 2 * 16);
        int it = 0, par = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int nt, x0, x1, t, strip, b;
            decode(item, nt, x0, x1, t, strip, b);
            // row m of the tile is position t*126 - 1 + m; rows 1..126 carry complete sums
            const int pos = t * UIMGH_STRIDE - 1 + m;
            bool valid = m >= 1 && m <= UIMGH_STRIDE && pos < p.P;
            int yy = 0, z = 0;
            if (valid) {
                yy = pos / p.ZP;
                const int zz = pos - yy * p.ZP;
                z = strip * p.SW + zz - 1;
                valid = zz >= 1 && zz <= p.SW && z < p.Dz;
            }
            const float* bias = p.bias ? p.bias + nt * CG : nullptr;
            for (int x = x0; x < x1; x++, it++) {
                const int acc = it & 1, aph = (it >> 1) & 1;
                if (acc != grp) continue;
                float* dst = p.y + ((((long long)(b * p.Dx + x) * p.Dy + yy) * p.Dz + z) * p.N + nt * CG);
                // accumulate (dx += ...): the old values of a 16-channel chunk are loaded one chunk ahead - the first before the accumulator
                // wait - instead of inside the store loop (a dependent DRAM round trip per chunk: the decoder1 conv1 input gradient took
                // 3.64 ms against 2.14 ms for the same convolution without accumulation)
                float4 old_next[4];
#pragma unroll
                for (int e = 0; e < 4; e++) old_next[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.accumulate && valid) {
#pragma unroll
                    for (int e = 0; e < 4; e++) old_next[e] = reinterpret_cast<const float4*>(dst)[e];
                }
                mbar_wait_warp(ACC_FULL(acc), aph);
                fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 1
                for (int j = 0; j < ((p.dbg & 32) ? 0 : CG / 16); j++) {
                    float vm[16], v0[16], vp[16];
                    float4 old_cur[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) old_cur[e] = old_next[e];
                    if (p.accumulate && valid && j + 1 < CG / 16) {
#pragma unroll
                        for (int e = 0; e < 4; e++) old_next[e] = reinterpret_cast<const float4*>(dst + (j + 1) * 16)[e];
                    }
                    {
                        uint32_t ra[16], rb[16], rc[16];
                        tmem_ld16_issue(taddr + j * 16, ra);           // dz = -1 block: wanted by row m+1
                        tmem_ld16_issue(taddr + CG + j * 16, rb);
                        tmem_ld16_issue(taddr + 2 * CG + j * 16, rc);  // dz = +1 block: wanted by row m-1
                        tmem_ld16_wait(ra); tmem_ld16_wait(rb); tmem_ld16_wait(rc);
#pragma unroll
                        for (int e = 0; e < 16; e++) { vm[e] = __uint_as_float(ra[e]); v0[e] = __uint_as_float(rb[e]); vp[e] = __uint_as_float(rc[e]); }
                    }
                    if (p.dbg & 16) {
                        if (valid && !(p.dbg & 8)) {
                            float4* d4 = reinterpret_cast<float4*>(dst + j * 16);
#pragma unroll
                            for (int e = 0; e < 4; e++) d4[e] = make_float4(v0[4 * e] + vm[4 * e], v0[4 * e + 1] + vp[4 * e], v0[4 * e + 2], v0[4 * e + 3]);
                        }
                        continue;
                    }
                    float4* xw = reinterpret_cast<float4*>(xch_g + ((par * 4 + q) * 2) * 16);
                    if (lane == 31) {
#pragma unroll
                        for (int e = 0; e < 4; e++) xw[e] = make_float4(vm[4 * e], vm[4 * e + 1], vm[4 * e + 2], vm[4 * e + 3]);
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int e = 0; e < 4; e++) xw[4 + e] = make_float4(vp[4 * e], vp[4 * e + 1], vp[4 * e + 2], vp[4 * e + 3]);
                    }
                    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                    else asm volatile("bar.sync 2, 128;" ::: "memory");
                    // every lane reads both boundary rows (broadcast loads) and selects: no per-element divergent branches
                    float lo[16], hi[16];
                    {
                        const float4* xlo = reinterpret_cast<const float4*>(xch_g + ((par * 4 + (q > 0 ? q - 1 : 0)) * 2) * 16);       // lane 31 of the warp below: dz=-1 block
                        const float4* xhi = reinterpret_cast<const float4*>(xch_g + ((par * 4 + (q < 3 ? q + 1 : 3)) * 2 + 1) * 16);   // lane 0 of the warp above: dz=+1 block
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float4 a = xlo[e], c = xhi[e];
                            lo[4 * e] = a.x; lo[4 * e + 1] = a.y; lo[4 * e + 2] = a.z; lo[4 * e + 3] = a.w;
                            hi[4 * e] = c.x; hi[4 * e + 1] = c.y; hi[4 * e + 2] = c.z; hi[4 * e + 3] = c.w;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        float um = __shfl_up_sync(0xffffffffu, vm[e], 1);
                        float dp = __shfl_down_sync(0xffffffffu, vp[e], 1);
                        um = lane == 0 ? lo[e] : um;
                        dp = lane == 31 ? hi[e] : dp;
                        v0[e] = (v0[e] + um + dp) * oscale;
                    }
                    par ^= 1;
                    if (valid && !(p.dbg & 8)) {
                        if (bias) {
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + j * 16) + e);
                                v0[4 * e] += bb.x; v0[4 * e + 1] += bb.y; v0[4 * e + 2] += bb.z; v0[4 * e + 3] += bb.w;
                            }
                        }
                        float4* d4 = reinterpret_cast<float4*>(dst + j * 16);
                        if (p.dbg & 64) {   // experiment: same bytes, but each store instruction of a warp covers 512 contiguous bytes (wrong place)
                            d4 = reinterpret_cast<float4*>(p.y + ((long long)(b * p.Dx + x) * p.Dy * p.Dz) * p.N) + (it & 63) * 1536 + q * 384 + j * 128 + lane;
#pragma unroll
                            for (int e = 0; e < 4; e++) d4[e * 32] = make_float4(v0[4 * e], v0[4 * e + 1], v0[4 * e + 2], v0[4 * e + 3]);
                            continue;
                        }
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            float4 o = make_float4(v0[4 * e], v0[4 * e + 1], v0[4 * e + 2], v0[4 * e + 3]);
                            if (p.accumulate) { o.x += old_cur[e].x; o.y += old_cur[e].y; o.z += old_cur[e].z; o.w += old_cur[e].w; }
                            d4[e] = o;
                        }
                    }
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(ACC_EMPTY(acc));
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool k_conv3_h_supported(int C, int N) {
    const int cg = uimg_h_cg(C);
    return cg != 0 && N % cg == 0;
}

long long k_conv3_h_blob_bytes(int C, int N) { return 27LL * C * N * 2; }

template <int CG>
static int conv3_h_launch(ConvHParams& p, const float* w, int mode, void* w_ws, cudaStream_t st) {
    constexpr int KCH = CG / 8;
    const int max_smem = 227 * 1024;
    p.plane_bytes = KCH * p.R_s * 16;
    p.b_stage_bytes = KCH * 3 * CG * 16;
    const int fixed = 2 * 2 * 4 * 2 * 16 * 4 + 8 * (2 * H_MAX_RING + 2 * H_MAX_BST + 4) + 16;
    // marching needs planes x-1, x, x+1 in use plus one prefetched plane per channel group
    int sms = 148, dev;
    NMAE_CUDA(cudaGetDevice(&dev));
    NMAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long columns = (long long)p.B * p.n_strips * p.tpp * p.n_nt;
    p.L = 1;
    p.n_ring = 0;
    if (4 * p.n_cg <= H_MAX_RING && 4LL * p.n_cg * p.plane_bytes + 2LL * p.b_stage_bytes + fixed <= max_smem) {
        // longest segment that still leaves every SM about three items
        p.L = p.Dx;
        while (p.L > 1 && columns * cdiv(p.Dx, p.L) < 3LL * sms) p.L = (p.L + 1) / 2;
        if (p.L > 1) p.n_ring = 4 * p.n_cg;
    }
    if (p.L == 1) {
        p.n_ring = min(H_MAX_RING, max(2, min(3 * p.n_cg, (max_smem - fixed - 2 * p.b_stage_bytes) / p.plane_bytes)));
    }
    NMAE_CHECK_ARG((long long)p.n_ring * p.plane_bytes + 2LL * p.b_stage_bytes + fixed <= max_smem,
                   "conv3_h: tile does not fit in shared memory (Dz=%d C=%d)", p.Dz, p.C);
    p.n_seg = cdiv(p.Dx, p.L);
    p.num_items = (int)(columns * p.n_seg);
    const int room = (max_smem - fixed - p.n_ring * p.plane_bytes) / p.b_stage_bytes;
    p.resident = (p.n_nt == 1 && 9 * p.n_cg <= min(room, H_MAX_BST)) ? 1 : 0;
    p.n_bst = p.resident ? 9 * p.n_cg : min(room, H_MAX_BST);
    const size_t smem = (size_t)fixed + (size_t)p.n_ring * p.plane_bytes + (size_t)p.n_bst * p.b_stage_bytes;

    const long long total = 27LL * p.C * p.N;
    const int gr = (int)min((long long)sms * 8, (total + 255) / 256);
    if (mode == 0)
        conv3_h_prep_kernel<CG><<<gr, 256, 0, st>>>(w, reinterpret_cast<__half*>(w_ws), p.C, p.N, (long long)p.C * 27, 27, 0);
    else
        conv3_h_prep_kernel<CG><<<gr, 256, 0, st>>>(w, reinterpret_cast<__half*>(w_ws), p.C, p.N, 27, (long long)p.N * 27, 1);
    NMAE_LAUNCH_CHECK();
    p.wblob = reinterpret_cast<const uint8_t*>(w_ws);

    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(conv3_h_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set[dev] = true;
    }
    conv3_h_kernel<CG><<<min(sms, p.num_items), 384, smem, st>>>(p);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// mode 0: forward (w is (N, C, 27)); mode 1: dgrad (w is (C, N, 27): the GEMM's output channel is w's input channel, taps flipped).
// `uimg` is the H image (uimg.cuh: uimg_geom_h) of the GEMM input with C channels; w_ws holds k_conv3_h_blob_bytes(C, N) bytes.
int k_conv3_h(const void* uimg, const float* w, const float* bias, const float* out_scale, int B, int Dx, int Dy, int Dz, int C, int N,
              int mode, void* w_ws, float* y, int accumulate, cudaStream_t st) {
    NMAE_CHECK_ARG(k_conv3_h_supported(C, N), "conv3_h: unsupported channels C=%d N=%d", C, N);
    const UImgGeom g = uimg_geom_h(B, Dx, Dy, Dz, C);
    ConvHParams p;
    memset(&p, 0, sizeof(p));
    p.uimg = reinterpret_cast<const uint8_t*>(uimg);
    p.u_chunk_bytes = g.chunk_bytes; p.u_img_bytes = g.img_bytes;
    p.y = y; p.bias = bias; p.out_scale = out_scale;
    p.B = B; p.Dx = Dx; p.Dy = Dy; p.Dz = Dz; p.C = C; p.N = N;
    p.n_cg = g.n_cg; p.n_nt = N / g.cg;
    p.SW = g.SW; p.n_strips = g.n_strips; p.ZP = g.ZP; p.P = g.P; p.tpp = g.tpp; p.R_s = g.R_img;
    p.accumulate = accumulate;
    p.dbg = nmae_debug_mask();
    return g.cg == 48 ? conv3_h_launch<48>(p, w, mode, w_ws, st) : conv3_h_launch<64>(p, w, mode, w_ws, st);
}
