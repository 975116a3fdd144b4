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
