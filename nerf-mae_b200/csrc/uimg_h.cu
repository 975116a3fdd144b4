// fp16 single-part ("H") operand images of channels-last volumes: the operands of the single-pass tcgen05 convolution kernels
// (conv3_h.cu, conv3_wgrad_h.cu).  Same position space as uimg.cu (uimg.cuh), one 16-byte row per position and 8-channel chunk.
//
// One CTA = 256 consecutive rows of one image (one (batch, x plane, z-strip, channel group)); one thread per row: it reads its
// voxel's CG channels (contiguous), optionally applies InstanceNorm + LeakyReLU on the fly, converts to fp16 with saturation and
// writes CG/8 16-byte rows (consecutive threads -> consecutive rows of a chunk: coalesced).
#include <cuda_fp16.h>

#include "uimg.cuh"

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source -> upper half
    return r;
}

__device__ __forceinline__ void in_consts_h(const double* st, int V, float eps, float& mu, float& rs) {
    const double m = st[0] / V;
    double var = st[1] / V - m * m;
    if (var < 0) var = 0;
    mu = (float)m;
    rs = (float)(1.0 / sqrt(var + (double)eps));
}

struct RowPos {
    int b, xx, yy, zz, z, cg, r;
    long long image;
    bool in_rows, valid, real;
};

__device__ __forceinline__ RowPos row_decode(const UImgGeom& g) {
    RowPos q;
    const int nrb = (g.R_tot + 255) / 256;
    q.image = blockIdx.x / nrb;
    const int rb = blockIdx.x - (int)q.image * nrb;
    long long t = q.image;
    q.cg = (int)(t % g.n_cg); t /= g.n_cg;
    const int strip = (int)(t % g.n_strips); t /= g.n_strips;
    const int xp = (int)(t % (g.Dx + 2));
    q.b = (int)(t / (g.Dx + 2));
    q.r = rb * 256 + threadIdx.x;
    q.xx = xp - 1;
    const int pos = q.r - g.H;
    q.yy = (pos + 2 * g.ZP) / g.ZP - 2;      // floor division for pos >= -2*ZP
    q.zz = pos - q.yy * g.ZP;
    q.z = strip * g.SW + q.zz - 1;
    q.in_rows = q.r < g.R_tot;
    q.valid = q.in_rows && q.xx >= 0 && q.xx < g.Dx && q.yy >= 0 && q.yy < g.Dy && q.z >= 0 && q.z < g.Dz;
    q.real = q.valid && q.zz >= 1 && q.zz <= g.SW;
    return q;
}

template <int CG>
__global__ void __launch_bounds__(256) uimg_h_build_kernel(const float* __restrict__ x, int ld, int ch_off, UImgGeom g,
                                                           const double* __restrict__ stats, int V, float eps, float slope,
                                                           const float* __restrict__ scale_ptr, uint8_t* __restrict__ out) {
    __shared__ float s_mu[CG], s_rs[CG];
    const float scale = scale_ptr ? __ldg(scale_ptr) : 1.f;
    const RowPos q = row_decode(g);
    if (stats) {
        if (threadIdx.x < CG) in_consts_h(stats + ((long long)q.b * g.C + q.cg * CG + threadIdx.x) * 2, V, eps, s_mu[threadIdx.x], s_rs[threadIdx.x]);
        __syncthreads();
    }
    if (!q.in_rows) return;
    uint8_t* dst = out + q.image * g.img_bytes + (long long)q.r * 16;
    const float4* src = reinterpret_cast<const float4*>(x + ((((long long)q.b * g.Dx + q.xx) * g.Dy + q.yy) * g.Dz + q.z) * ld + ch_off + q.cg * CG);
#pragma unroll
    for (int c = 0; c < CG / 8; c++) {
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (q.valid && ch_off + q.cg * CG + c * 8 + 8 <= ld) {     // channels past the voxel record read as zero
            v0 = __ldg(src + 2 * c);
            v1 = __ldg(src + 2 * c + 1);
            if (stats) {
                float* f0 = reinterpret_cast<float*>(&v0);
                float* f1 = reinterpret_cast<float*>(&v1);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float a = (f0[e] - s_mu[c * 8 + e]) * s_rs[c * 8 + e];
                    f0[e] = a >= 0.f ? a : a * slope;
                    a = (f1[e] - s_mu[c * 8 + 4 + e]) * s_rs[c * 8 + 4 + e];
                    f1[e] = a >= 0.f ? a : a * slope;
                }
            }
        }
        uint4 h;
        h.x = pack_h2(v0.x * scale, v0.y * scale); h.y = pack_h2(v0.z * scale, v0.w * scale);
        h.z = pack_h2(v1.x * scale, v1.y * scale); h.w = pack_h2(v1.z * scale, v1.w * scale);
        *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes) = h;
    }
}

int k_uimg_h_build(const float* x, int ld, int ch_off, const UImgGeom& g, const double* stats, float eps, float slope,
                   const float* scale, void* uimg, cudaStream_t st) {
    NMAE_CHECK_ARG(g.cg != 0 && ld % 4 == 0 && ch_off % 4 == 0, "uimg_h: channels must be a multiple of 48 or 64 (C=%d ld=%d)", g.C, ld);
    NMAE_CHECK_ARG(ch_off + g.C <= ld || (ld - ch_off) % 8 == 0, "uimg_h: a zero-padded image needs (ld - ch_off) %% 8 == 0 (ld=%d)", ld);
    NMAE_CHECK_ARG(stats == nullptr || (ch_off == 0 && ld == g.C), "uimg_h: the fused InstanceNorm needs the whole tensor (ld == C)");
    const long long images = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg;
    const long long ctas = images * ((g.R_tot + 255) / 256);
    NMAE_CHECK_ARG(ctas < (1LL << 31), "uimg_h: volume too large for one launch");
    const int V = g.Dx * g.Dy * g.Dz;
    if (g.cg == 48)
        uimg_h_build_kernel<48><<<(unsigned)ctas, 256, 0, st>>>(x, ld, ch_off, g, stats, V, eps, slope, scale, reinterpret_cast<uint8_t*>(uimg));
    else
        uimg_h_build_kernel<64><<<(unsigned)ctas, 256, 0, st>>>(x, ld, ch_off, g, stats, V, eps, slope, scale, reinterpret_cast<uint8_t*>(uimg));
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of  out = LeakyReLU(IN(x) + R)  writing the gradient wrt x straight into its fp16 operand image (see uimg.cu for the
// bf16 hi/lo form and the closed-form bias gradients).  fp16 has 5 exponent bits, and this gradient is ~1e-7 in magnitude (the loss
// is a mean over 16 M voxels): the image therefore stores  dx * 2^k  with one power-of-two scale per tensor, chosen so that the
// bound  U = 4 * max|g| * max_c(1/std_c)  of |dx| maps to ~2^10 (64x headroom below the fp16 maximum, values down to 2^-34 U stay
// representable); conversions saturate.  Every CTA derives the same scale from the same device scalars; CTA 0 publishes 2^-k for
// the epilogues of the dgrad / weight-gradient kernels.
__global__ void __launch_bounds__(128) in_bwd_bias_h_kernel(const double* __restrict__ stats, const double* __restrict__ stats3,
                                                            const double* __restrict__ sums, int B, int C, int V, float eps,
                                                            float* __restrict__ dbias, float* __restrict__ dbias3) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, a3 = 0.0;
    for (int b = 0; b < B; b++) {
        const double* sm = sums + ((long long)b * C + c) * 3;
        const float m0 = (float)(sm[0] / V), m1 = (float)(sm[1] / V), m2 = (float)(sm[2] / V);
        float mu, rs;
        in_consts_h(stats + ((long long)b * C + c) * 2, V, eps, mu, rs);
        const double sum_xhat = (stats[((long long)b * C + c) * 2] - (double)V * mu) * rs;
        a += (double)rs * ((sm[0] - (double)V * m0) - (double)m1 * sum_xhat);
        if (dbias3) {
            float mu3, rs3;
            in_consts_h(stats3 + ((long long)b * C + c) * 2, V, eps, mu3, rs3);
            const double sum_xhat3 = (stats3[((long long)b * C + c) * 2] - (double)V * mu3) * rs3;
            a3 += (double)rs3 * ((sm[0] - (double)V * m0) - (double)m2 * sum_xhat3);
        }
    }
    if (dbias) dbias[c] = (float)a;
    if (dbias3) dbias3[c] = (float)a3;
}

// One small CTA turns the double statistics / reduction sums into the float constants of the apply kernel - consts[(b*C + c)*8 ..] =
// {mean, 1/std, mean3, 1/std3, S0/V, S1/V, S2/V, -} - and derives the image scale (consts[8*B*C]) and its reciprocal (*inv_scale).
// (The apply kernel used to redo this double-precision arithmetic in each of its ~70 k CTAs: 45 % of its instructions.)
__global__ void __launch_bounds__(256) in_bwd_consts_h_kernel(const double* __restrict__ stats, const double* __restrict__ stats3,
                                                              const double* __restrict__ sums, const float* __restrict__ amax_g, int BC,
                                                              int V, float eps, float* __restrict__ consts, float* __restrict__ inv_scale) {
    __shared__ float s_red[8];
    float rmax = 0.f;
    for (int i = threadIdx.x; i < BC; i += 256) {
        float mu, rs, mu3 = 0.f, rs3 = 1.f;
        in_consts_h(stats + (long long)i * 2, V, eps, mu, rs);
        if (stats3) in_consts_h(stats3 + (long long)i * 2, V, eps, mu3, rs3);
        const double* sm = sums + (long long)i * 3;
        float4* o = reinterpret_cast<float4*>(consts + (long long)i * 8);
        o[0] = make_float4(mu, rs, mu3, rs3);
        o[1] = make_float4((float)(sm[0] / V), (float)(sm[1] / V), stats3 ? (float)(sm[2] / V) : 0.f, 0.f);
        rmax = fmaxf(rmax, rs);
    }
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = rmax;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; i++) rmax = fmaxf(rmax, s_red[i]);
        const float bound = 4.f * __ldg(amax_g) * rmax;
        int k = 0;
        if (bound > 0.f && bound < 3.0e38f) {
            int e;
            frexpf(bound, &e);          // bound = m * 2^e, m in [0.5, 1)
            k = 10 - e;
            k = max(-100, min(100, k));
        }
        consts[(long long)BC * 8] = ldexpf(1.f, k);
        *inv_scale = ldexpf(1.f, -k);
    }
}

template <int CG, bool DP4, bool HAS_OUT, bool HAS_X3>
__global__ void __launch_bounds__(256, 3) in_bwd_apply_image_h_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                      const float* __restrict__ x, const float* __restrict__ x3,
                                                                      const float* __restrict__ consts, UImgGeom g, float slope,
                                                                      uint8_t* __restrict__ img, float* __restrict__ dx3,
                                                                      float* __restrict__ dres, const float4* __restrict__ dp4,
                                                                      const float* __restrict__ w4) {
    __shared__ __align__(16) float s_c[7][CG];     // mu, rs, mu3, rs3, S0/V, S1/V, S2/V of this CTA's channels
    __shared__ __align__(16) float s_w[4][CG];     // DP4: weights of the 1x1x1 output convolution whose input gradient dout is (see norm.cu)
    __shared__ int s_vox[256];      // voxel index | real << 30, or -1 for pad / halo-outside rows
    __shared__ __align__(16) float4 s_dp[256];
    __shared__ __align__(16) uint8_t s_img[(CG / 8) * (256 + 1) * 16];
    const RowPos q = row_decode(g);
    const int vox_q = q.valid ? (int)((((long long)q.b * g.Dx + q.xx) * g.Dy + q.yy) * g.Dz + q.z) : -1;   // < 2^30 (checked by the launcher)
    float4 dpv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (DP4 && vox_q >= 0) dpv = __ldg(dp4 + vox_q);       // in flight together with the constants below
    const float scale = __ldg(consts + (long long)g.B * g.C * 8);
    if (threadIdx.x < CG) {
        const int ch = q.cg * CG + threadIdx.x;
        const float4* c4 = reinterpret_cast<const float4*>(consts + ((long long)q.b * g.C + ch) * 8);
        const float4 ca = __ldg(c4), cb = __ldg(c4 + 1);
        s_c[0][threadIdx.x] = ca.x; s_c[1][threadIdx.x] = ca.y; s_c[2][threadIdx.x] = ca.z; s_c[3][threadIdx.x] = ca.w;
        s_c[4][threadIdx.x] = cb.x; s_c[5][threadIdx.x] = cb.y; s_c[6][threadIdx.x] = cb.z;
        if (DP4) {
#pragma unroll
            for (int k = 0; k < 4; k++) s_w[k][threadIdx.x] = w4[k * g.C + ch];
        }
    }
    __syncthreads();

    // Phase 1 - a warp owns its 32 rows; unit = one float4 (4 channels of one row), unit u = j*32 + lane -> row u / F, float4 u % F:
    // each load / store instruction of the warp covers 512 contiguous bytes (4-5 cache lines).  The first version gave every thread
    // its whole row (lanes 192-256 bytes apart: 32 L1 tag wavefronts per instruction) and ran at the L1 wavefront limit
    // (~16 B/clk/SM = 4.5 TB/s of reads, profiles/r2_in_bwd.txt).  The fp16 results are transposed through shared memory
    // ([chunk][row][16 B], chunks (256+1) rows apart) so that phase 2 writes the image with one 16-byte row per lane as before.
    // Units j and j + P (P = F / gcd(32, F)) of a thread have the same float4 index f, RS = 32*P/F rows apart: the per-channel
    // constants are fetched once per f and the row loop needs no division.
    constexpr int F = CG / 4;
    constexpr int P = (F % 32 == 0) ? F / 32 : (F % 16 == 0) ? F / 16 : (F % 8 == 0) ? F / 8 : F / 4;   // F = 12 -> 3, F = 16 -> 1
    constexpr int NU = F / P, RS = 32 * P / F;       // units per f and their row step (4 units 8 rows apart / 16 units 2 rows apart)
    constexpr int UNR = (HAS_X3 || (HAS_OUT && !DP4)) ? 2 : 4;         // units in flight per thread (85-register budget: three CTAs per SM)
    static_assert(NU % UNR == 0 && F % 4 == 0 && NU * RS == 32, "unit mapping");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s_vox[threadIdx.x] = q.valid ? (vox_q | (q.real ? (1 << 30) : 0)) : -1;
    if (DP4) s_dp[threadIdx.x] = dpv;
    __syncwarp();
    const long long cg_off4 = (long long)q.cg * (CG / 4), C4 = g.C / 4;
#pragma unroll 1
    for (int jj = 0; jj < P; jj++) {
        const int u0 = jj * 32 + lane, rw0 = warp * 32 + u0 / F, f = u0 % F, ch0 = f * 4;
        float k_mu[4], k_rs[4], k_mu3[4], k_rs3[4], k_s0[4], k_s1[4], k_s2[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            k_mu[e] = s_c[0][ch0 + e]; k_rs[e] = s_c[1][ch0 + e]; k_mu3[e] = s_c[2][ch0 + e]; k_rs3[e] = s_c[3][ch0 + e];
            k_s0[e] = s_c[4][ch0 + e]; k_s1[e] = s_c[5][ch0 + e]; k_s2[e] = s_c[6][ch0 + e];
        }
        uint8_t* st_dst = s_img + (size_t)(f >> 1) * ((256 + 1) * 16) + (f & 1) * 8;
#pragma unroll 1
        for (int n0 = 0; n0 < NU; n0 += UNR) {
            float4 d[UNR], xv[UNR], ov[UNR], x3v[UNR];
            int vx[UNR];
#pragma unroll
            for (int t = 0; t < UNR; t++) {
                vx[t] = s_vox[rw0 + (n0 + t) * RS];
                d[t] = xv[t] = ov[t] = x3v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (vx[t] >= 0) {
                    const long long o4 = (long long)(vx[t] & 0x3fffffff) * C4 + cg_off4 + f;
                    if (!DP4) d[t] = __ldg(reinterpret_cast<const float4*>(dout) + o4);
                    xv[t] = __ldg(reinterpret_cast<const float4*>(x) + o4);
                    if (HAS_OUT) ov[t] = __ldg(reinterpret_cast<const float4*>(out) + o4);
                    if (HAS_X3 && (vx[t] >> 30)) x3v[t] = __ldg(reinterpret_cast<const float4*>(x3) + o4);
                }
            }
#pragma unroll
            for (int t = 0; t < UNR; t++) {
                const int rw = rw0 + (n0 + t) * RS;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (vx[t] >= 0) {
                    float dd[4] = {d[t].x, d[t].y, d[t].z, d[t].w};
                    if (DP4) {
                        const float4 dq = s_dp[rw];     // the weights stay in shared memory: 16 more registers would spill
                        const float dqa[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) dd[e] = 0.f;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float4 wk = *reinterpret_cast<const float4*>(&s_w[k][ch0]);
                            dd[0] += wk.x * dqa[k]; dd[1] += wk.y * dqa[k]; dd[2] += wk.z * dqa[k]; dd[3] += wk.w * dqa[k];
                        }
                    }
                    const float xa[4] = {xv[t].x, xv[t].y, xv[t].z, xv[t].w}, oa[4] = {ov[t].x, ov[t].y, ov[t].z, ov[t].w};
                    const float x3a[4] = {x3v[t].x, x3v[t].y, x3v[t].z, x3v[t].w};
                    float gq[4], o3[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float xh = (xa[e] - k_mu[e]) * k_rs[e];
                        gq[e] = dd[e] * ((HAS_OUT ? oa[e] : xh) > 0.f ? 1.f : slope);
                        o[e] = k_rs[e] * (gq[e] - k_s0[e] - xh * k_s1[e]);
                        o3[e] = HAS_X3 ? k_rs3[e] * (gq[e] - k_s0[e] - (x3a[e] - k_mu3[e]) * k_rs3[e] * k_s2[e]) : 0.f;
                    }
                    if (vx[t] >> 30) {      // the strip that owns the voxel writes the fp32 by-products
                        const long long o4 = (long long)(vx[t] & 0x3fffffff) * C4 + cg_off4 + f;
                        if (HAS_X3) reinterpret_cast<float4*>(dx3)[o4] = make_float4(o3[0], o3[1], o3[2], o3[3]);
                        if (dres) reinterpret_cast<float4*>(dres)[o4] = make_float4(gq[0], gq[1], gq[2], gq[3]);
                    }
                }
                uint2 h;
                h.x = pack_h2(o[0] * scale, o[1] * scale);
                h.y = pack_h2(o[2] * scale, o[3] * scale);
                *reinterpret_cast<uint2*>(st_dst + (size_t)rw * 16) = h;
            }
        }
    }
    __syncwarp();
    // Phase 2 - one image row per lane and chunk: 512 contiguous bytes per store instruction
    if (q.in_rows) {
        uint8_t* dst = img + q.image * g.img_bytes + (long long)q.r * 16;
#pragma unroll
        for (int c = 0; c < CG / 8; c++)
            *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes) =
                *reinterpret_cast<const uint4*>(s_img + (size_t)c * ((256 + 1) * 16) + (size_t)threadIdx.x * 16);
    }
}

int k_in_act_bwd_image_h(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                         const double* sums, const float* amax_g, const UImgGeom& g, float eps, float slope, void* dx_image,
                         float* inv_scale, float* dx3, float* dres, float* dbias, float* dbias3, cudaStream_t st, const float* dp4,
                         const float* w4) {
    NMAE_CHECK_ARG(g.cg != 0, "in_lrelu_apply_bwd_image_h: channels must be a multiple of 48 or 64 (C=%d)", g.C);
    const long long images = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg;
    const long long ctas = images * ((g.R_tot + 255) / 256);
    NMAE_CHECK_ARG(ctas < (1LL << 31) && (long long)g.B * g.Dx * g.Dy * g.Dz < (1LL << 30),
                   "in_lrelu_apply_bwd_image_h: volume too large for one launch");
    const int V = g.Dx * g.Dy * g.Dz;
    // the float constants live behind the 3*B*C double sums in the same workspace (nmae_in_lrelu_bwd_sums_ws_bytes)
    float* consts = reinterpret_cast<float*>(const_cast<double*>(sums) + 3LL * g.B * g.C);
    in_bwd_consts_h_kernel<<<1, 256, 0, st>>>(stats, x3 ? stats3 : nullptr, sums, amax_g, g.B * g.C, V, eps, consts, inv_scale);
    NMAE_LAUNCH_CHECK();
#define APPLY_H(CGV, D, O, X3)                                                                                                     \
    in_bwd_apply_image_h_kernel<CGV, D, O, X3><<<(unsigned)ctas, 256, 0, st>>>(dout, out, x, x3, consts, g, slope,                   \
        reinterpret_cast<uint8_t*>(dx_image), dx3, dres, reinterpret_cast<const float4*>(dp4), w4)
#define APPLY_H_CG(CGV)                                                                                                            \
    switch ((dp4 ? 4 : 0) | (out ? 2 : 0) | (x3 ? 1 : 0)) {                                                                        \
        case 0: APPLY_H(CGV, false, false, false); break;                                                                          \
        case 1: APPLY_H(CGV, false, false, true); break;                                                                           \
        case 2: APPLY_H(CGV, false, true, false); break;                                                                           \
        case 3: APPLY_H(CGV, false, true, true); break;                                                                            \
        case 4: APPLY_H(CGV, true, false, false); break;                                                                           \
        case 5: APPLY_H(CGV, true, false, true); break;                                                                            \
        case 6: APPLY_H(CGV, true, true, false); break;                                                                            \
        default: APPLY_H(CGV, true, true, true); break;                                                                            \
    }
    if (g.cg == 48) {
        APPLY_H_CG(48)
    } else {
        APPLY_H_CG(64)
    }
#undef APPLY_H_CG
#undef APPLY_H
    NMAE_LAUNCH_CHECK();
    if (dbias || dbias3) {
        in_bwd_bias_h_kernel<<<(g.C + 127) / 128, 128, 0, st>>>(stats, stats3, sums, g.B, g.C, V, eps, dbias, dbias3);
        NMAE_LAUNCH_CHECK();
    }
    return NMAE_OK;
}
