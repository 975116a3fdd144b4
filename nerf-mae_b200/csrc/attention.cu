// 3D shifted-window attention core (reference swin_mae3d.py:27-197), CUDA-core fp32 version.
//
// The cyclic shift, the zero padding to a multiple of the window and the window partition are all
// folded into one index map (no roll / pad / permute copies):
//   window (wh,ww,wd), slot (ih,iw,id)  ->  rolled coordinate r = 4*w + i  (per axis)
//   source coordinate = (r + shift) mod P   (P = padded extent; the roll is modulo the PADDED size)
//   slots whose source lies beyond the real extent are padding tokens: they were zeros before the qkv
//   projection, so their q/k/v equal the qkv bias (row `pad_row` of the qkv buffer) and they are NOT
//   masked out of the softmax.  Padded queries are dropped.
//   region id (shift mask, swin_mae3d.py:126-167): per axis band 0:[0,P-4) 1:[P-4,P-s) 2:[P-s,P).
#include "kernels.cuh"

#define WS 4
#define NT 64   // tokens per window
#define HD 32   // head dim

struct WinGeom {
    int H, W, D;     // real token grid
    int PH, PW, PD;  // padded
    int sh, sw, sd;  // effective shift per axis
    int nWh, nWw, nWd;
};

__device__ __forceinline__ void slot_map(const WinGeom& g, int win, int slot, int& src, int& region) {
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

__device__ __forceinline__ int rel_index(int qi, int kj) {
    int dh = (qi >> 4) - (kj >> 4) + 3, dw = ((qi >> 2) & 3) - ((kj >> 2) & 3) + 3, dd = (qi & 3) - (kj & 3) + 3;
    return (dh * 7 + dw) * 7 + dd;
}

// grid (nW, nH, B), 128 threads: thread t -> query t/2, key half t%2
__global__ void __launch_bounds__(128) wattn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ table, WinGeom g,
                                                        int T, int C, int nH, long long pad_row, float scale,
                                                        float* __restrict__ out, float* __restrict__ lse) {
    __shared__ float sk[NT][HD + 1], sv[NT][HD + 1], stab[343];
    __shared__ int ssrc[NT], sreg[NT];
    const int win = blockIdx.x, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
    const int C3 = 3 * C;
    if (t < NT) slot_map(g, win, t, ssrc[t], sreg[t]);
    for (int i = t; i < 343; i += 128) stab[i] = table[i * nH + h];
    __syncthreads();
    for (int i = t; i < NT * HD; i += 128) {
        int j = i >> 5, d = i & 31;
        long long row = ssrc[j] >= 0 ? (long long)b * T + ssrc[j] : pad_row;
        sk[j][d] = qkv[row * C3 + C + h * HD + d];
        sv[j][d] = qkv[row * C3 + 2 * C + h * HD + d];
    }
    const int qi = t >> 1, kh = t & 1;
    float q[HD];
    {
        long long row = ssrc[qi] >= 0 ? (long long)b * T + ssrc[qi] : pad_row;
        const float* qp = qkv + row * C3 + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d++) q[d] = qp[d] * scale;
    }
    __syncthreads();
    float p[32];
    float mx = -INFINITY;
    const int myreg = sreg[qi];
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
        int j = kh * 32 + jj;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d++) s = fmaf(q[d], sk[j][d], s);
        s += stab[rel_index(qi, j)];
        if (sreg[j] != myreg) s += -100.f;
        p[jj] = s;
        mx = fmaxf(mx, s);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
        p[jj] = expf(p[jj] - mx);
        sum += p[jj];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    float inv = 1.f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; d++) o[d] = 0.f;
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
        float pj = p[jj] * inv;
        int j = kh * 32 + jj;
#pragma unroll
        for (int d = 0; d < HD; d++) o[d] = fmaf(pj, sv[j][d], o[d]);
    }
#pragma unroll
    for (int d = 0; d < HD; d++) o[d] += __shfl_xor_sync(0xffffffffu, o[d], 1);
    if (kh == 0) lse[(((long long)b * gridDim.x + win) * nH + h) * NT + qi] = mx + logf(sum);
    if (ssrc[qi] >= 0) {
        float* op = out + ((long long)b * T + ssrc[qi]) * C + h * HD + kh * 16;
#pragma unroll
        for (int d = 0; d < 16; d++) op[d] = kh ? o[16 + d] : o[d];
    }
}

// grid (nChunk, nH, B): each CTA walks windows win = chunk, chunk+nChunk, ... and keeps the
// relative-position-bias gradient of its head in shared memory until the end.
__global__ void __launch_bounds__(128) wattn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ table,
                                                        const float* __restrict__ o_saved, const float* __restrict__ dout,
                                                        const float* __restrict__ lse, WinGeom g, int nW, int T, int C, int nH,
                                                        long long pad_row, float scale, float* __restrict__ dqkv,
                                                        float* __restrict__ dtable) {
    extern __shared__ float smem[];
    float(*sq)[HD + 1] = reinterpret_cast<float(*)[HD + 1]>(smem);            // scaled q
    float(*sk)[HD + 1] = sq + NT;
    float(*sv)[HD + 1] = sk + NT;
    float(*sdo)[HD + 1] = sv + NT;
    float(*sP)[NT + 1] = reinterpret_cast<float(*)[NT + 1]>(sdo + NT);
    float(*sdS)[NT + 1] = sP + NT;
    float(*sdSsum)[NT + 1] = sdS + NT;
    float* sdb = reinterpret_cast<float*>(sdSsum + NT);  // [343] bias-table gradient of this head
    float* stab = sdb + 343;
    int* ssrc = reinterpret_cast<int*>(stab + 343);
    int* sreg = ssrc + NT;
    const int h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
    const int C3 = 3 * C;
    for (int i = t; i < 343; i += 128) {
        sdb[i] = 0.f;
        stab[i] = table[i * nH + h];
    }
    // relative-position-bias gradient: dS is summed over this CTA's windows in a shared 64x64 matrix (each entry owned by one
    // thread: plain read-modify-write) and binned into the 343 table entries once at the end - per-window shared-memory
    // atomics on 343 hot bins serialise badly
    for (int i = t; i < NT * (NT + 1); i += 128) (&sdSsum[0][0])[i] = 0.f;
    for (int win = blockIdx.x; win < nW; win += gridDim.x) {
        __syncthreads();
        if (t < NT) slot_map(g, win, t, ssrc[t], sreg[t]);
        __syncthreads();
        for (int i = t; i < NT * HD; i += 128) {
            int j = i >> 5, d = i & 31;
            bool valid = ssrc[j] >= 0;
            long long row = valid ? (long long)b * T + ssrc[j] : pad_row;
            sq[j][d] = qkv[row * C3 + h * HD + d] * scale;
            sk[j][d] = qkv[row * C3 + C + h * HD + d];
            sv[j][d] = qkv[row * C3 + 2 * C + h * HD + d];
            sdo[j][d] = valid ? dout[row * C + h * HD + d] : 0.f;
        }
        __syncthreads();
        const int qi = t >> 1, kh = t & 1;
        const bool qvalid = ssrc[qi] >= 0;
        // D_i = sum_d dO*O  (== rowsum(dP*P))
        float Di = 0.f;
        if (qvalid) {
            const float* op = o_saved + ((long long)b * T + ssrc[qi]) * C + h * HD;
#pragma unroll
            for (int d = 0; d < HD; d++) Di = fmaf(sdo[qi][d], op[d], Di);
        }
        const float l = lse[(((long long)b * nW + win) * nH + h) * NT + qi];
        const int myreg = sreg[qi];
        float dq[HD];
#pragma unroll
        for (int d = 0; d < HD; d++) dq[d] = 0.f;
        for (int jj = 0; jj < 32; jj++) {
            int j = kh * 32 + jj;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d++) {
                s = fmaf(sq[qi][d], sk[j][d], s);
                dp = fmaf(sdo[qi][d], sv[j][d], dp);
            }
            int ri = rel_index(qi, j);
            s += stab[ri];
            if (sreg[j] != myreg) s += -100.f;
            float pj = expf(s - l);
            float ds = qvalid ? pj * (dp - Di) : 0.f;
            sP[qi][j] = qvalid ? pj : 0.f;
            sdS[qi][j] = ds;
            sdSsum[qi][j] += ds;
#pragma unroll
            for (int d = 0; d < HD; d++) dq[d] = fmaf(ds, sk[j][d], dq[d]);
        }
#pragma unroll
        for (int d = 0; d < HD; d++) dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], 1);
        if (qvalid) {
            float* p = dqkv + ((long long)b * T + ssrc[qi]) * C3 + h * HD + kh * 16;
#pragma unroll
            for (int d = 0; d < 16; d++) p[d] = (kh ? dq[16 + d] : dq[d]) * scale;
        }
        __syncthreads();
        // phase 2: thread -> key kj = t/2, dims [16*kh, 16*kh+16)
        const int kj = qi, d0 = kh * 16;
        float dk[16], dv[16];
#pragma unroll
        for (int d = 0; d < 16; d++) dk[d] = dv[d] = 0.f;
        for (int i = 0; i < NT; i++) {
            float pij = sP[i][kj], dsij = sdS[i][kj];
#pragma unroll
            for (int d = 0; d < 16; d++) {
                dv[d] = fmaf(pij, sdo[i][d0 + d], dv[d]);
                dk[d] = fmaf(dsij, sq[i][d0 + d], dk[d]);  // sq already carries the 1/sqrt(hd) factor
            }
        }
        if (ssrc[kj] >= 0) {
            float* p = dqkv + ((long long)b * T + ssrc[kj]) * C3 + h * HD + d0;
#pragma unroll
            for (int d = 0; d < 16; d++) {
                p[C + d] = dk[d];
                p[2 * C + d] = dv[d];
            }
        } else {  // padding slot: its k/v are the qkv bias -> accumulate into the shared pad row
            float* p = dqkv + pad_row * C3 + h * HD + d0;
#pragma unroll
            for (int d = 0; d < 16; d++) {
                atomicAdd(p + C + d, dk[d]);
                atomicAdd(p + 2 * C + d, dv[d]);
            }
        }
    }
    {
        const int qi = t >> 1, kh = t & 1;
        for (int jj = 0; jj < 32; jj++) {
            const float v = sdSsum[qi][kh * 32 + jj];
            if (v != 0.f) atomicAdd(&sdb[rel_index(qi, kh * 32 + jj)], v);
        }
    }
    __syncthreads();
    for (int i = t; i < 343; i += 128)
        if (sdb[i] != 0.f) atomicAdd(dtable + i * nH + h, sdb[i]);
}

static int make_geom(int H, int W, int D, int shift, WinGeom& g) {
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
    WinGeom g;
    return make_geom(H, W, D, 0, g);
}

int k_wattn_fwd(const float* qkv, const float* table, int B, int H, int W, int D, int C, int nH, int shift, float* out, float* lse,
                cudaStream_t st) {
    NMAE_CHECK_ARG(C == nH * HD, "window attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
    NMAE_CHECK_ARG(shift >= 0 && shift < WS, "window attention: shift %d out of range", shift);
    WinGeom g;
    int nW = make_geom(H, W, D, shift, g);
    int T = H * W * D;
    NMAE_CHECK_ARG(nH <= 65535 && B <= 65535, "window attention: grid too large");
    wattn_fwd_kernel<<<dim3(nW, nH, B), 128, 0, st>>>(qkv, table, g, T, C, nH, (long long)B * T, 1.f / sqrtf((float)HD), out, lse);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_wattn_bwd(const float* qkv, const float* table, const float* o_saved, const float* dout, const float* lse, int B, int H,
                int W, int D, int C, int nH, int shift, float* dqkv, float* dtable, cudaStream_t st) {
    NMAE_CHECK_ARG(C == nH * HD, "window attention: head_dim must be 32 (C=%d heads=%d)", C, nH);
    WinGeom g;
    int nW = make_geom(H, W, D, shift, g);
    int T = H * W * D;
    size_t smem = sizeof(float) * (4 * NT * (HD + 1) + 3 * NT * (NT + 1) + 2 * 343) + sizeof(int) * 2 * NT;
    static bool attr_set[64] = {false};
    int dev;
    NMAE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(wattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    // the pad row of dqkv accumulates with atomics, the rest is overwritten
    NMAE_CUDA(cudaMemsetAsync(dqkv + (long long)B * T * 3 * C, 0, sizeof(float) * 3 * C, st));
    int nChunk = nW < 48 ? nW : 48;
    wattn_bwd_kernel<<<dim3(nChunk, nH, B), 128, smem, st>>>(qkv, table, o_saved, dout, lse, g, nW, T, C, nH, (long long)B * T,
                                                             1.f / sqrtf((float)HD), dqkv, dtable);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
