#include "uimg.cuh"

#include "tc.cuh"

// one thread per (image, chunk, row): reads 8 channels (32 B), writes 16 B hi + 16 B lo; consecutive threads -> consecutive rows
__global__ void __launch_bounds__(256) uimg_build_kernel(const float* __restrict__ x, int ld, int ch_off, UImgGeom g, int type_dy,
                                                         uint8_t* __restrict__ out) {
    const long long rows_total = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg * UIMG_KCH * g.R_tot;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < rows_total; u += (long long)gridDim.x * blockDim.x) {
        long long t = u;
        const int r = (int)(t % g.R_tot); t /= g.R_tot;
        const int c = (int)(t % UIMG_KCH); t /= UIMG_KCH;
        const int cg = (int)(t % g.n_cg); t /= g.n_cg;
        const int strip = (int)(t % g.n_strips); t /= g.n_strips;
        const int xp = (int)(t % (g.Dx + 2));      // plane index including the two zero pad planes
        const int b = (int)(t / (g.Dx + 2));
        const int xx = xp - 1;
        const int pos = r - g.H;
        const int yy = (pos + 2 * g.ZP) / g.ZP - 2;      // floor division for pos >= -2*ZP
        const int zz = pos - yy * g.ZP;
        const int z = strip * g.SW + zz - 1;
        bool valid = xx >= 0 && xx < g.Dx && yy >= 0 && yy < g.Dy && z >= 0 && z < g.Dz;
        if (type_dy) valid = valid && zz >= 1 && zz <= g.SW;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (valid) {
            const float4* src = reinterpret_cast<const float4*>(x + ((((long long)b * g.Dx + xx) * g.Dy + yy) * g.Dz + z) * ld + ch_off +
                                                                cg * UIMG_CG + c * 8);
            v0 = __ldg(src);
            v1 = __ldg(src + 1);
        }
        uint4 h, l;
        tc::split2(v0.x, v0.y, h.x, l.x);
        tc::split2(v0.z, v0.w, h.y, l.y);
        tc::split2(v1.x, v1.y, h.z, l.z);
        tc::split2(v1.z, v1.w, h.w, l.w);
        const long long img = (((long long)(b * (g.Dx + 2) + xp) * g.n_strips + strip) * g.n_cg + cg) * g.img_bytes;
        uint8_t* dst = out + img + (long long)c * g.chunk_bytes + (long long)r * 16;
        *reinterpret_cast<uint4*>(dst) = h;
        *reinterpret_cast<uint4*>(dst + g.part_bytes) = l;
    }
}

int k_uimg_build(const float* x, int ld, int ch_off, const UImgGeom& g, int type_dy, void* uimg, cudaStream_t st) {
    NMAE_CHECK_ARG(g.C % UIMG_CG == 0 && ld % 4 == 0 && ch_off % 4 == 0, "uimg: channels must be a multiple of 48 (C=%d ld=%d)", g.C, ld);
    const long long rows_total = (long long)g.B * (g.Dx + 2) * g.n_strips * g.n_cg * UIMG_KCH * g.R_tot;
    int grid = (int)min((long long)148 * 32, (rows_total + 255) / 256);
    uimg_build_kernel<<<grid, 256, 0, st>>>(x, ld, ch_off, g, type_dy, reinterpret_cast<uint8_t*>(uimg));
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
