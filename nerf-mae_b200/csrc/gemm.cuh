// Generic implicit GEMM:  C[m,n] = epilogue( sum_k A(m,k) * B(n,k) )
// Operands are addressed through small gather descriptors so that the same kernel serves the
// linear layers, the patch embedding (k=s conv), the 3x3x3 convolutions (fwd / dgrad / wgrad) and
// the k=s transposed convolutions of the decoder without ever materialising an im2col buffer.
#pragma once
#include "common.cuh"

enum { OPM_STRIDED = 0, OPM_CONV3 = 1, OPM_PATCH = 2, OPM_D2S = 3 };

// A gather operand maps (a = spatial index, b = feature index) to an address:
//   STRIDED: p[a*s_a + b*s_b]                                    (a=row, b=k; no swap needed)
//   CONV3  : a=(n,x,y,z) over an NDHWC volume, b=(tap,ci); zero padding 1          [3x3x3 conv im2col]
//   PATCH  : a=(n,tx,ty,tz) tokens, b=(c,i,j,l) inside a ks^3 patch of an NCDHW volume  [k=s conv im2col]
//   D2S    : a=(n,x,y,z) coarse voxel, b=(co,i,j,l) -> fine NDHWC voxel (x*ks+i,..), channel co [k=s convT]
// swap=1 exchanges the roles: the GEMM row index is b and the reduction index is a (used by wgrad).
struct GOperand {
    const float* p;
    int mode, swap;
    long long s_a, s_b;
    int X, Y, Z;  // extent of the spatial index a
    int C;        // CONV3: channels per tap; PATCH: input channels; D2S: output channels
    int ld;       // elements per voxel in the gathered NDHWC tensor (>= C; concat buffers)
    int ks;       // PATCH / D2S kernel (= stride)
};

enum {
    EPI_BIAS = 1,        // + bias[n]
    EPI_GELU = 2,        // aux[m,n] = pre-activation ; out = gelu(pre)
    EPI_RESID = 4,       // out = resid[m,n] + row_scale[m / rows_per_scale] * value
    EPI_GELU_GRAD = 8,   // out = value * gelu'(aux[m,n])
    EPI_ATOMIC = 16,     // atomicAdd into out (split-K)
    EPI_D2S = 32,        // scatter store: m = coarse voxel, n = (co,i,j,l)
    EPI_ACCUM = 64       // out += value
};

struct GEpilogue {
    float* out;
    long long ldc;
    const float* bias;
    float* aux;
    const float* resid;
    const float* row_scale;
    int rows_per_scale;
    int flags;
    int X, Y, Z, C, ld, ks;  // D2S geometry (same meaning as GOperand)
};

struct GemmParams {
    GOperand A, B;
    GEpilogue E;
    int M, N, K;
    int ksplit;  // K elements per grid.z slice
};

struct SpIdx { int n, x, y, z; };

__device__ __forceinline__ SpIdx decode_sp(int X, int Y, int Z, int a) {
    SpIdx s;
    s.z = a % Z; a /= Z;
    s.y = a % Y; a /= Y;
    s.x = a % X; s.n = a / X;
    return s;
}

__device__ __forceinline__ long long d2s_addr(int X, int Y, int Z, int ld, int ks, const SpIdx& s, int b) {
    int k3 = ks * ks * ks;
    int co = b / k3, r = b - co * k3;
    int i = r / (ks * ks), j = (r / ks) % ks, l = r % ks;
    return ((((long long)s.n * (X * ks) + s.x * ks + i) * (Y * ks) + s.y * ks + j) * (long long)(Z * ks) + s.z * ks + l) * ld + co;
}

__device__ __forceinline__ float gather_elem(const GOperand& o, const SpIdx& s, int b) {
    if (o.mode == OPM_CONV3) {
        int tap = b / o.C, ci = b - tap * o.C;
        int xx = s.x + tap / 9 - 1, yy = s.y + (tap / 3) % 3 - 1, zz = s.z + tap % 3 - 1;
        if ((unsigned)xx >= (unsigned)o.X || (unsigned)yy >= (unsigned)o.Y || (unsigned)zz >= (unsigned)o.Z) return 0.f;
        return __ldg(o.p + ((((long long)s.n * o.X + xx) * o.Y + yy) * o.Z + zz) * o.ld + ci);
    } else if (o.mode == OPM_PATCH) {
        int ks = o.ks, k3 = ks * ks * ks;
        int c = b / k3, r = b - c * k3;
        int i = r / (ks * ks), j = (r / ks) % ks, l = r % ks;
        return __ldg(o.p + ((((long long)s.n * o.C + c) * (o.X * ks) + s.x * ks + i) * (o.Y * ks) + s.y * ks + j) * (long long)(o.Z * ks) +
                     s.z * ks + l);
    } else {  // OPM_D2S
        return __ldg(o.p + d2s_addr(o.X, o.Y, o.Z, o.ld, o.ks, s, b));
    }
}

int nmae_gemm_launch(const GemmParams& p, cudaStream_t stream);
