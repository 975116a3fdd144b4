#include "uimg.cuh"

#include "tc.cuh"

// One CTA = 256 consecutive rows of one image (one (batch, x plane, z-strip, 48-channel group)); one thread per row:
// it reads its voxel's 48 channels (192 contiguous bytes; consecutive rows of a strip are consecutive voxels, so a warp
// reads a contiguous span), optionally applies InstanceNorm + LeakyReLU on the fly (the fused form of
// nmae_in_lrelu_apply_fwd without residual: the fp32 activation is never materialised), and writes 6 x (16 B hi + 16 B lo):
// consecutive threads -> consecutive 16-byte rows of a chunk.
__global__ void __launch_bounds__(256) uimg_build_kernel(const float* __restrict__ x, int ld, int ch_off, UImgGeom g, int type_dy,
                                                         const double* __restrict__ stats, int V, float eps, float slope,
                                                         uint8_t* __restrict__ out) {
    __shared__ float s_mu[UIMG_CG], s_rs[UIMG_CG];
    const int nrb = (g.R_tot + 255) / 256;
    const int image = blockIdx.x / nrb, rb = blockIdx.x - image * nrb;
    int t = image;
    const int cg = t % g.n_cg; t /= g.n_cg;
    const int strip = t % g.n_strips; t /= g.n_strips;
    const int xp = t % (g.Dx + 2);                        // plane index including the two zero pad planes
    const int b = t / (g.Dx + 2);
    if (stats) {
        if (threadIdx.x < UIMG_CG) {
            const double* st = stats + ((long long)b * g.C + cg * UIMG_CG + threadIdx.x) * 2;
            const double m = st[0] / V;
            double var = st[1] / V - m * m;
            if (var < 0) var = 0;
            s_mu[threadIdx.x] = (float)m;
            s_rs[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
        }
        __syncthreads();
    }
    const int r = rb * 256 + threadIdx.x;
    if (r >= g.R_tot) return;
    const int xx = xp - 1;
    const int pos = r - g.H;
    const int yy = (pos + 2 * g.ZP) / g.ZP - 2;      // floor division for pos >= -2*ZP
    const int zz = pos - yy * g.ZP;
    const int z = strip * g.SW + zz - 1;
    bool valid = xx >= 0 && xx < g.Dx && yy >= 0 && yy < g.Dy && z >= 0 && z < g.Dz;
    if (type_dy) valid = valid && zz >= 1 && zz <= g.SW;
    uint8_t* dst = out + (long long)image * g.img_bytes + (long long)r * 16;
    const float4* src = reinterpret_cast<const float4*>(x + ((((long long)b * g.Dx + xx) * g.Dy + yy) * g.Dz + z) * ld + ch_off + cg * UIMG_CG);
#pragma unroll
    for (int c = 0; c < UIMG_KCH; c++) {
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        // channels past the end of the voxel record read as zero: an image with more channels than the tensor (e.g. 256 -> 288,
        // the next multiple of 48) is the zero-padded operand of a convolution whose weights are padded the same way
        if (valid && ch_off + cg * UIMG_CG + c * 8 + 8 <= ld) {
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
        uint4 h, l;
        tc::split2(v0.x, v0.y, h.x, l.x);
        tc::split2(v0.z, v0.w, h.y, l.y);
        tc::split2(v1.x, v1.y, h.z, l.z);
        tc::split2(v1.z, v1.w, h.w, l.w);
        *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes) = h;
        *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes + g.part_bytes) = l;
    }
}

int k_uimg_build(const float* x, int ld, int ch_off, const UImgGeom& g, int type_dy, const double* stats, float eps, float slope,
                 void* uimg, cudaStream_t st) {
    NMAE_CHECK_ARG(g.C % UIMG_CG == 0 && ld % 4 == 0 && ch_off % 4 == 0, "uimg: channels must be a multiple of 48 (C=%d ld=%d)", g.C, ld);
    NMAE_CHECK_ARG(ch_off + g.C <= ld || (ld - ch_off) % 8 == 0, "uimg: a zero-padded image needs (ld - ch_off) %% 8 == 0 (ld=%d)", ld);
    const long long images = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg;
    const long long ctas = images * ((g.R_tot + 255) / 256);
    NMAE_CHECK_ARG(ctas < (1LL << 31), "uimg: volume too large for one launch");
    NMAE_CHECK_ARG(stats == nullptr || (ch_off == 0 && ld == g.C), "uimg: the fused InstanceNorm needs the whole tensor (ld == C)");
    uimg_build_kernel<<<(unsigned)ctas, 256, 0, st>>>(x, ld, ch_off, g, type_dy, stats, g.Dx * g.Dy * g.Dz, eps, slope, reinterpret_cast<uint8_t*>(uimg));
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of  out = LeakyReLU(IN(x) + R)  (norm.cu: in_bwd_apply) that writes the gradient wrt x STRAIGHT INTO ITS OPERAND
// IMAGE (the only consumers of that gradient are the dgrad and weight-gradient convolution kernels, which read images):
//   g = dout * lrelu'(out),  dx = rs * (g - S0/V - xhat * S1/V)
// One thread per image row as in uimg_build_kernel; halo rows recompute the neighbour voxel's gradient (2 of 34 columns).
// Real (non-halo) rows also write the fp32 side outputs: dx3 (gradient of the normalised 1x1x1 residual branch) and dres
// (identity residual).  The column sums of dx / dx3 (bias gradients of the convolutions in front of the InstanceNorms) are not
// accumulated element by element: sum_v dx = rs * ((S0 - V*m0) - m1 * sum_v xhat) is fp32 rounding noise around an exact zero,
// and in_bwd_bias_kernel evaluates that expression from the double-precision sums.
__device__ __forceinline__ void in_consts(const double* st, int V, float eps, float& mu, float& rs) {
    const double m = st[0] / V;
    double var = st[1] / V - m * m;
    if (var < 0) var = 0;
    mu = (float)m;
    rs = (float)(1.0 / sqrt(var + (double)eps));
}

// dbias[c] = sum_b sum_v dx[b][v][c] evaluated in closed form from the per-(b,c) double sums, with the same rounded fp32
// constants the apply kernel uses (m0 = float(S0/V), m1 = float(S1/V), mu, rs): the element-wise sum of dx equals
// rs * ((S0 - V*m0) - m1 * rs * (sum_x - V*mu)) up to the rounding of the individual products.
__global__ void __launch_bounds__(128) in_bwd_bias_kernel(const double* __restrict__ stats, const double* __restrict__ stats3,
                                                          const double* __restrict__ sums, int B, int C, int V, float eps,
                                                          float* __restrict__ dbias, float* __restrict__ dbias3) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, a3 = 0.0;
    for (int b = 0; b < B; b++) {
        const double* sm = sums + ((long long)b * C + c) * 3;
        const float m0 = (float)(sm[0] / V), m1 = (float)(sm[1] / V), m2 = (float)(sm[2] / V);
        float mu, rs;
        in_consts(stats + ((long long)b * C + c) * 2, V, eps, mu, rs);
        const double sum_xhat = (stats[((long long)b * C + c) * 2] - (double)V * mu) * rs;
        a += (double)rs * ((sm[0] - (double)V * m0) - (double)m1 * sum_xhat);
        if (dbias3) {
            float mu3, rs3;
            in_consts(stats3 + ((long long)b * C + c) * 2, V, eps, mu3, rs3);
            const double sum_xhat3 = (stats3[((long long)b * C + c) * 2] - (double)V * mu3) * rs3;
            a3 += (double)rs3 * ((sm[0] - (double)V * m0) - (double)m2 * sum_xhat3);
        }
    }
    if (dbias) dbias[c] = (float)a;
    if (dbias3) dbias3[c] = (float)a3;
}

__global__ void __launch_bounds__(256) in_bwd_apply_image_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                 const float* __restrict__ x, const double* __restrict__ stats,
                                                                 const float* __restrict__ x3, const double* __restrict__ stats3,
                                                                 const double* __restrict__ sums, UImgGeom g, int V, float eps, float slope,
                                                                 uint8_t* __restrict__ img, float* __restrict__ dx3, float* __restrict__ dres) {
    __shared__ float s_c[7][UIMG_CG];     // mu, rs, mu3, rs3, S0/V, S1/V, S2/V of this CTA's 48 channels
    const int nrb = (g.R_tot + 255) / 256;
    const int image = blockIdx.x / nrb, rb = blockIdx.x - image * nrb;
    int t = image;
    const int cg = t % g.n_cg; t /= g.n_cg;
    const int strip = t % g.n_strips; t /= g.n_strips;
    const int xp = t % (g.Dx + 2);
    const int b = t / (g.Dx + 2);
    if (threadIdx.x < UIMG_CG) {
        const int ch = cg * UIMG_CG + threadIdx.x;
        in_consts(stats + ((long long)b * g.C + ch) * 2, V, eps, s_c[0][threadIdx.x], s_c[1][threadIdx.x]);
        s_c[2][threadIdx.x] = 0.f; s_c[3][threadIdx.x] = 1.f;
        if (x3) in_consts(stats3 + ((long long)b * g.C + ch) * 2, V, eps, s_c[2][threadIdx.x], s_c[3][threadIdx.x]);
        const double* sm = sums + ((long long)b * g.C + ch) * 3;
        s_c[4][threadIdx.x] = (float)(sm[0] / V);
        s_c[5][threadIdx.x] = (float)(sm[1] / V);
        s_c[6][threadIdx.x] = x3 ? (float)(sm[2] / V) : 0.f;
    }
    __syncthreads();
    const int r = rb * 256 + threadIdx.x;
    const int xx = xp - 1;
    const int pos = r - g.H;
    const int yy = (pos + 2 * g.ZP) / g.ZP - 2;
    const int zz = pos - yy * g.ZP;
    const int z = strip * g.SW + zz - 1;
    const bool in_rows = r < g.R_tot;
    const bool valid = in_rows && xx >= 0 && xx < g.Dx && yy >= 0 && yy < g.Dy && z >= 0 && z < g.Dz;
    const bool real = valid && zz >= 1 && zz <= g.SW;          // not a halo duplicate: owns the fp32 side outputs
    const long long vox = (((long long)b * g.Dx + xx) * g.Dy + yy) * g.Dz + z;
    const long long off = vox * g.C + cg * UIMG_CG;
    uint8_t* dst = img + (long long)image * g.img_bytes + (long long)r * 16;
#pragma unroll
    for (int c = 0; c < UIMG_KCH; c++) {
        float o[8], o3[8];
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = o3[e] = 0.f;
        if (valid) {
            float d[8], xv[8], ov[8], x3v[8];
            *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(dout + off) + 2 * c);
            *reinterpret_cast<float4*>(d + 4) = __ldg(reinterpret_cast<const float4*>(dout + off) + 2 * c + 1);
            *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x + off) + 2 * c);
            *reinterpret_cast<float4*>(xv + 4) = __ldg(reinterpret_cast<const float4*>(x + off) + 2 * c + 1);
            if (out) {
                *reinterpret_cast<float4*>(ov) = __ldg(reinterpret_cast<const float4*>(out + off) + 2 * c);
                *reinterpret_cast<float4*>(ov + 4) = __ldg(reinterpret_cast<const float4*>(out + off) + 2 * c + 1);
            }
            if (x3 && real) {
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
                if (x3 && real) o3[e] = s_c[3][ch] * (gq[e] - s_c[4][ch] - (x3v[e] - s_c[2][ch]) * s_c[3][ch] * s_c[6][ch]);
            }
            if (real) {
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
        if (in_rows) {
            uint4 h, l;
            tc::split2(o[0], o[1], h.x, l.x);
            tc::split2(o[2], o[3], h.y, l.y);
            tc::split2(o[4], o[5], h.z, l.z);
            tc::split2(o[6], o[7], h.w, l.w);
            *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes) = h;
            *reinterpret_cast<uint4*>(dst + (long long)c * g.chunk_bytes + g.part_bytes) = l;
        }
    }
}

int k_in_act_bwd_image(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                       const double* sums, const UImgGeom& g, float eps, float slope, void* dx_image, float* dx3, float* dres,
                       float* dbias, float* dbias3, cudaStream_t st) {
    NMAE_CHECK_ARG(g.C % UIMG_CG == 0, "in_lrelu_apply_bwd_image: channels must be a multiple of 48 (C=%d)", g.C);
    const long long images = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg;
    const long long ctas = images * ((g.R_tot + 255) / 256);
    NMAE_CHECK_ARG(ctas < (1LL << 31), "in_lrelu_apply_bwd_image: volume too large for one launch");
    const int V = g.Dx * g.Dy * g.Dz;
    in_bwd_apply_image_kernel<<<(unsigned)ctas, 256, 0, st>>>(dout, out, x, stats, x3, stats3, sums, g, V, eps, slope,
                                                             reinterpret_cast<uint8_t*>(dx_image), dx3, dres);
    NMAE_LAUNCH_CHECK();
    if (dbias || dbias3) {
        in_bwd_bias_kernel<<<(g.C + 127) / 128, 128, 0, st>>>(stats, stats3, sums, g.B, g.C, V, eps, dbias, dbias3);
        NMAE_LAUNCH_CHECK();
    }
    return NMAE_OK;
}
