// "UMMA-ready" activation images.  A channels-last fp32 volume (B,Dx,Dy,Dz,C) is re-laid ONCE into bf16 hi/lo planes in
// exactly the shared-memory layout the tcgen05 convolution kernels consume, so that their operand staging is a handful
// of cp.async.bulk copies (no per-element work, no registers, arbitrarily deep pipelining).
//
// Position space: each (batch, x) plane is cut into z-strips of SW <= 40 voxels; inside a strip
//     pos = y*(SW+2) + (z - strip*SW + 1),   row r = pos + H   (H = SW+3 rows of lead-in so that tile halos never underflow)
// Layout: [b][x+1 in 0..Dx+1][strip][channel group of 48][part: hi, lo][8-channel chunk (6)][R_tot rows][8 x bf16]
// (planes x = -1 and x = Dx are zero pad planes so that the dx = +-1 neighbours of border planes need no special case)
// Type X : halo columns hold the neighbouring strips' voxels (zeros outside the volume)   - convolution inputs
// Type DY: halo columns are zero                                                          - output-side gradients
#pragma once
#include "common.cuh"

#define UIMG_CG 48
#define UIMG_KCH 6
#define UIMG_TILE 128

struct UImgGeom {
    int B, Dx, Dy, Dz, C;
    int SW, n_strips, ZP, P, tpp, H, R_img, R_tot, n_cg;
    long long chunk_bytes, part_bytes, img_bytes, total_bytes;
};

static inline UImgGeom uimg_geom(int B, int Dx, int Dy, int Dz, int C) {
    UImgGeom g;
    g.B = B; g.Dx = Dx; g.Dy = Dy; g.Dz = Dz; g.C = C;
    g.SW = Dz <= 40 ? Dz : 32;
    g.n_strips = (Dz + g.SW - 1) / g.SW;
    g.ZP = g.SW + 2;
    g.P = Dy * g.ZP;
    g.tpp = (g.P + UIMG_TILE - 1) / UIMG_TILE;
    g.H = g.ZP + 1;
    g.R_img = UIMG_TILE + 2 * g.H;
    g.R_tot = g.tpp * UIMG_TILE + 2 * g.H;
    g.n_cg = C / UIMG_CG;
    g.chunk_bytes = (long long)g.R_tot * 16;
    g.part_bytes = UIMG_KCH * g.chunk_bytes;
    g.img_bytes = 2 * g.part_bytes;
    g.total_bytes = (long long)B * (Dx + 2) * g.n_strips * g.n_cg * g.img_bytes;
    return g;
}

// builds the image tensor from channels [ch_off, ch_off + C) of a volume with `ld` floats per voxel
// the apply pass of the InstanceNorm+LeakyReLU backward with the gradient written as a type-X image (uimg.cu)
int k_in_act_bwd_image(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                       const double* sums, const UImgGeom& g, float eps, float slope, void* dx_image, float* dx3, float* dres,
                       float* dbias, float* dbias3, cudaStream_t st);
// stats != NULL: the image of LeakyReLU_slope(InstanceNorm(x)) is built instead (stats = (B,C,2) doubles of nmae_instnorm_stats)
int k_uimg_build(const float* x, int ld, int ch_off, const UImgGeom& g, int type_dy, const double* stats, float eps, float slope,
                 void* uimg, cudaStream_t st);
