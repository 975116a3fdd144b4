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
    int cg, kch, parts;   // channels per group (48 | 64), 16-byte chunks per row, operand parts (2: bf16 hi/lo, 1: fp16)
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
    g.cg = UIMG_CG; g.kch = UIMG_KCH; g.parts = 2;
    return g;
}

// "H" images: ONE fp16 part (single-pass operands: 11-bit significands, the TF32 class) in the same position space, with
// 48- or 64-channel groups (cg == 0: channel count not supported).  Layout [b][x+1][strip][channel group][8-channel chunk][R_tot rows]
// [8 x f16].  The single-pass forward/dgrad kernel (conv3_h.cu) walks a plane in tiles of UIMGH_STRIDE = 126 output positions
// (128 operand rows: the dz = +-1 taps are folded into the GEMM N dimension and re-aligned by one row in the epilogue); the
// weight-gradient kernel walks it in 128-position K tiles.  R_tot covers both.
#define UIMGH_STRIDE 126
static inline int uimg_h_cg(int C) { return C % 48 == 0 ? 48 : (C % 64 == 0 ? 64 : 0); }
static inline UImgGeom uimg_geom_h(int B, int Dx, int Dy, int Dz, int C) {
    UImgGeom g;
    g.B = B; g.Dx = Dx; g.Dy = Dy; g.Dz = Dz; g.C = C;
    g.cg = uimg_h_cg(C);
    g.kch = g.cg / 8; g.parts = 1;
    g.SW = Dz <= 40 ? Dz : 32;
    g.n_strips = (Dz + g.SW - 1) / g.SW;
    g.ZP = g.SW + 2;
    g.P = Dy * g.ZP;
    g.tpp = (g.P + UIMGH_STRIDE - 1) / UIMGH_STRIDE;     // forward/dgrad tiles per (b, x, strip) plane
    g.H = g.ZP + 1;
    g.R_img = UIMG_TILE + 2 * g.ZP;                      // operand rows one tile needs from one plane
    g.R_tot = ((g.P + UIMG_TILE - 1) / UIMG_TILE + 1) * UIMG_TILE + 2 * g.H;
    g.n_cg = g.cg ? C / g.cg : 0;
    g.chunk_bytes = (long long)g.R_tot * 16;
    g.part_bytes = g.kch * g.chunk_bytes;
    g.img_bytes = g.part_bytes;
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

// fp16 "H" images (uimg_h.cu).  stats != NULL fuses LeakyReLU(InstanceNorm(x)) (forward).
// scale != NULL: values are multiplied by the device scalar *scale before the conversion (gradient images).
int k_uimg_h_build(const float* x, int ld, int ch_off, const UImgGeom& g, const double* stats, float eps, float slope,
                   const float* scale, void* uimg, cudaStream_t st);
// InstanceNorm+LeakyReLU backward writing the gradient as an H image scaled by a power of two chosen from amax_g (device float:
// max |dout * lrelu'| over the tensor, from k_in_bwd_sums) and the largest 1/std; the scale's reciprocal is stored to inv_scale.
int k_in_act_bwd_image_h(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                         const double* sums, const float* amax_g, const UImgGeom& g, float eps, float slope, void* dx_image,
                         float* inv_scale, float* dx3, float* dres, float* dbias, float* dbias3, cudaStream_t st,
                         const float* dp4 = nullptr, const float* w4 = nullptr);
