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

template <int CG>
__global__ void __launch_bounds__(256) in_bwd_apply_image_h_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                   const float* __restrict__ x, const double* __restrict__ stats,
                                                                   const float* __restrict__ x3, const double* __restrict__ stats3,
                                                                   const double* __restrict__ sums, const float* __restrict__ amax_g,
                                                                   UImgGeom g, int V, float eps, float slope, uint8_t* __restrict__ img,
                                                                   float* __restrict__ inv_scale, float* __restrict__ dx3,
                                                                   float* __restrict__ dres, const float4* __restrict__ dp4,
                                                                   const float* __restrict__ w4) {
    __shared__ float s_c[7][CG];     // mu, rs, mu3, rs3, S0/V, S1/V, S2/V of this CTA's channels
    __shared__ __align__(16) float s_w[4][CG];     // dp4 != NULL: weights of the 1x1x1 output convolution whose input gradient dout is (see norm.cu)
    __shared__ float s_red[8];
    const RowPos q = row_decode(g);
    // largest 1/std over every (b, c): identical in all CTAs
    float rmax = 0.f;
    for (int i = threadIdx.x; i < g.B * g.C; i += 256) {
        float mu, rs;
        in_consts_h(stats + (long long)i * 2, V, eps, mu, rs);
        rmax = fmaxf(rmax, rs);
    }
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = rmax;
    if (threadIdx.x < CG) {
        const int ch = q.cg * CG + threadIdx.x;
        in_consts_h(stats + ((long long)q.b * g.C + ch) * 2, V, eps, s_c[0][threadIdx.x], s_c[1][threadIdx.x]);
        s_c[2][threadIdx.x] = 0.f; s_c[3][threadIdx.x] = 1.f;
        if (x3) in_consts_h(stats3 + ((long long)q.b * g.C + ch) * 2, V, eps, s_c[2][threadIdx.x], s_c[3][threadIdx.x]);
        const double* sm = sums + ((long long)q.b * g.C + ch) * 3;
        s_c[4][threadIdx.x] = (float)(sm[0] / V);
        s_c[5][threadIdx.x] = (float)(sm[1] / V);
        s_c[6][threadIdx.x] = x3 ? (float)(sm[2] / V) : 0.f;
        if (dp4) {
#pragma unroll
            for (int k = 0; k < 4; k++) s_w[k][threadIdx.x] = w4[k * g.C + ch];
        }
    }
    __syncthreads();
    rmax = s_red[0];
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
    const float scale = ldexpf(1.f, k);
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = ldexpf(1.f, -k);

    const long long vox = (((long long)q.b * g.Dx + q.xx) * g.Dy + q.yy) * g.Dz + q.z;
    const long long off = vox * g.C + q.cg * CG;
    uint8_t* dst = img + q.image * g.img_bytes + (long long)q.r * 16;
    float4 dpv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dp4 && q.valid) dpv = __ldg(dp4 + vox);
#pragma unroll
    for (int c = 0; c < CG / 8; c++) {
        float o[8], o3[8];
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = o3[e] = 0.f;
        if (q.valid) {
            float d[8], xv[8], ov[8], x3v[8];
            if (dp4) {
                const float dq[4] = {dpv.x, dpv.y, dpv.z, dpv.w};
#pragma unroll
                for (int e = 0; e < 8; e++) d[e] = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 wa = *reinterpret_cast<const float4*>(&s_w[k][c * 8]), wb = *reinterpret_cast<const float4*>(&s_w[k][c * 8 + 4]);
                    d[0] += wa.x * dq[k]; d[1] += wa.y * dq[k]; d[2] += wa.z * dq[k]; d[3] += wa.w * dq[k];
                    d[4] += wb.x * dq[k]; d[5] += wb.y * dq[k]; d[6] += wb.z * dq[k]; d[7] += wb.w * dq[k];
                }
            } else {
                *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(dout + off) + 2 * c);
                *reinterpret_cast<float4*>(d + 4) = __ldg(reinterpret_cast<const float4*>(dout + off) + 2 * c + 1);
            }
            *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x + off) + 2 * c);
            *reinterpret_cast<float4*>(xv + 4) = __ldg(reinterpret_cast<const float4*>(x + off) + 2 * c + 1);
            if (out) {
                *reinterpret_cast<float4*>(ov) = __ldg(reinterpret_cast<const float4*>(out + off) + 2 * c);
                *reinterpret_cast<float4*>(ov + 4) = __ldg(reinterpret_cast<const float4*>(out + off) + 2 * c + 1);
            }
            if (x3 && q.real) {
                *reinterpret_cast<float4*>(x3v) = __ldg(reinterpret_cast<const float4*>(x3 + off) + 2 * c);
                *reinterpret_cast<float4*>(x3v + 4) = __ldg(reinterpret_cast<const float4*>(x3 + off) + 2 * c + 1);
            }
            float gq[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int ch = c * 8 + e;
                const float xh = (xv[e] - s_c[0][ch]) * s_c[1][ch];
                gq[e] = d[e] * ((out ? ov[e] : xh) > 0.f ? 1.f : slope);
                o[e] = s_c[1][ch] * (gq[e] - s_c[4][ch] - xh * s_c[5][ch]);
                if (x3 && q.real) o3[e] = s_c[3][ch] * (gq[e] - s_c[4][ch] - (x3v[e] - s_c[2][ch]) * s_c[3][ch] * s_c[6][ch]);
            }
            if (q.real) {
                if (dx3) {
                    reinterpret_cast<float4*>(dx3 + off)[2 * c] = make_float4(o3[0], o3[1], o3[2], o3[3]);
                    reinterpret_cast<float4*>(dx3 + off)[2 * c + 1] = make_float4(o3[4], o3[5], o3[6], o3[7]);
                }
                if (dres) {
                    reinterpret_cast<float4*>(dres + off)[2 * c] = make_float4(gq[0], gq[1], gq[2], gq[3]);
                    reinterpret_cast<float4*>(dres + off)[2 * c + 1] = make_float4(gq[4], gq[5], gq[6], gq[7]);
                }
            }
        }
        if (q.in_rows) {
            uint4 h;
            h.x = pack_h2(o[0] * scale, o[1] * scale); h.y = pack_h2(o[2] * scale, o[3] * scale);
            h.z = pack_h2(o[4] * scale, o[5] * scale); h.w = pack_h2(o[6] * scale, o[7] * scale);
            *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes) = h;
        }
    }
}

int k_in_act_bwd_image_h(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                         const double* sums, const float* amax_g, const UImgGeom& g, float eps, float slope, void* dx_image,
                         float* inv_scale, float* dx3, float* dres, float* dbias, float* dbias3, cudaStream_t st, const float* dp4,
                         const float* w4) {
    NMAE_CHECK_ARG(g.cg != 0, "in_lrelu_apply_bwd_image_h: channels must be a multiple of 48 or 64 (C=%d)", g.C);
    const long long images = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg;
    const long long ctas = images * ((g.R_tot + 255) / 256);
    NMAE_CHECK_ARG(ctas < (1LL << 31), "in_lrelu_apply_bwd_image_h: volume too large for one launch");
    const int V = g.Dx * g.Dy * g.Dz;
    if (g.cg == 48)
        in_bwd_apply_image_h_kernel<48><<<(unsigned)ctas, 256, 0, st>>>(dout, out, x, stats, x3, stats3, sums, amax_g, g, V, eps, slope,
                                                                       reinterpret_cast<uint8_t*>(dx_image), inv_scale, dx3, dres,
                                                                       reinterpret_cast<const float4*>(dp4), w4);
    else
        in_bwd_apply_image_h_kernel<64><<<(unsigned)ctas, 256, 0, st>>>(dout, out, x, stats, x3, stats3, sums, amax_g, g, V, eps, slope,
                                                                       reinterpret_cast<uint8_t*>(dx_image), inv_scale, dx3, dres,
                                                                       reinterpret_cast<const float4*>(dp4), w4);
    NMAE_LAUNCH_CHECK();
    if (dbias || dbias3) {
        in_bwd_bias_h_kernel<<<(g.C + 127) / 128, 128, 0, st>>>(stats, stats3, sums, g.B, g.C, V, eps, dbias, dbias3);
        NMAE_LAUNCH_CHECK();
    }
    return NMAE_OK;
}
