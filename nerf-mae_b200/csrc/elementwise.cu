// HBM-bound helpers: pad-to-cube, small strided gathers (weight re-layout), channel-slice copies
// (skip concat), the masked-voxel MSE loss (reference swin_mae3d.py:1513-1563) and the fused
// clip + AdamW multi-tensor step (reference run_swin_mae3d.py:663-669).
#include <cuda_fp16.h>

#include "kernels.cuh"

// dst (Cc,R,R,R) <- src (Cc,X,Y,Z) zero padded at the high end of every axis (torch_utils.py:56-90)
__global__ void __launch_bounds__(256) pad_grid_kernel(const float* __restrict__ src, int Cc, int X, int Y, int Z,
                                                       float* __restrict__ dst, int R) {
    long long n = (long long)Cc * R * R * R;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int z = (int)(i % R);
        long long t = i / R;
        int y = (int)(t % R); t /= R;
        int x = (int)(t % R);
        int c = (int)(t / R);
        dst[i] = (x < X && y < Y && z < Z) ? src[(((long long)c * X + x) * Y + y) * Z + z] : 0.f;
    }
}

int k_pad_grid(const float* src, int Cc, int X, int Y, int Z, float* dst, int R, cudaStream_t st) {
    long long n = (long long)Cc * R * R * R;
    int g = (int)min((long long)148 * 16, (n + 255) / 256);
    pad_grid_kernel<<<g, 256, 0, st>>>(src, Cc, X, Y, Z, dst, R);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// dst[i][j][k] (contiguous) = src[i*s0 + j*s1 + k*s2]
__global__ void __launch_bounds__(256) gather3_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n0,
                                                      long long n1, long long n2, long long s0, long long s1, long long s2) {
    long long n = n0 * n1 * n2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long k = i % n2, t = i / n2;
        long long j = t % n1, a = t / n1;
        dst[i] = src[a * s0 + j * s1 + k * s2];
    }
}

int k_gather3(float* dst, const float* src, long long n0, long long n1, long long n2, long long s0, long long s1, long long s2,
              cudaStream_t st) {
    long long n = n0 * n1 * n2;
    if (n == 0) return NMAE_OK;
    int g = (int)min((long long)148 * 16, (n + 255) / 256);
    gather3_kernel<<<g, 256, 0, st>>>(dst, src, n0, n1, n2, s0, s1, s2);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

__global__ void __launch_bounds__(256) copy_cols_kernel(float* __restrict__ dst, long long ldd, const float* __restrict__ src,
                                                        long long lds, long long rows, int cols) {
    long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / cols;
        int c = (int)(i - r * cols);
        dst[r * ldd + c] = src[r * lds + c];
    }
}

int k_copy_cols(float* dst, long long ldd, const float* src, long long lds, long long rows, int cols, cudaStream_t st) {
    long long n = rows * cols;
    if (n == 0) return NMAE_OK;
    int g = (int)min((long long)148 * 16, (n + 255) / 256);
    copy_cols_kernel<<<g, 256, 0, st>>>(dst, ldd, src, lds, rows, cols);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// dst[r][c] = src[r][c] * row_scale[r / rows_per_scale]   (stochastic-depth "row" mode, backward side)
__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                         const float* __restrict__ row_scale, int rows_per_scale, long long rows,
                                                         int cols) {
    long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i] * row_scale[(i / cols) / rows_per_scale];
}

int k_scale_rows(float* dst, const float* src, const float* row_scale, int rows_per_scale, long long rows, int cols,
                 cudaStream_t st) {
    long long n = rows * cols;
    if (n == 0) return NMAE_OK;
    int g = (int)min((long long)148 * 16, (n + 255) / 256);
    scale_rows_kernel<<<g, 256, 0, st>>>(dst, src, row_scale, rows_per_scale, rows, cols);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------ loss
// x: (B,4,R,R,R) padded target, pred: (B,R,R,R,4) channels-last, ext: (B,3) un-padded extents (replaces
// the reference's dense 65 MB pad mask), tok_mask: (R/p)^3 token-level block mask shared by the batch.
// sums = {sum_rgb, n_valid, sum_alpha, n_remove}
__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pred,
                                                       const int* __restrict__ ext, const uint8_t* __restrict__ tok_mask, int R,
                                                       int p, double* __restrict__ sums) {
    __shared__ float sh[32];
    const int b = blockIdx.y;
    const long long V = (long long)R * R * R;
    const int n = R / p;
    const int ex = ext[b * 3], ey = ext[b * 3 + 1], ez = ext[b * 3 + 2];
    const float* xb = x + (long long)b * 4 * V;
    const float4* pb = reinterpret_cast<const float4*>(pred) + (long long)b * V;
    float s_rgb = 0.f, s_a = 0.f, n_valid = 0.f, n_rem = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        int z = (int)(i % R);
        long long t = i / R;
        int y = (int)(t % R), xx = (int)(t / R);
        float4 pr = pb[i];
        float t0 = xb[i], t1 = xb[V + i], t2 = xb[2 * V + i], ta = xb[3 * V + i];
        if (ta > 0.01f) {
            float d0 = pr.x - t0, d1 = pr.y - t1, d2 = pr.z - t2;
            s_rgb += d0 * d0 + d1 * d1 + d2 * d2;
            n_valid += 1.f;
        }
        if (xx < ex && y < ey && z < ez && tok_mask[((xx / p) * n + y / p) * n + z / p]) {
            float sg = 1.f / (1.f + expf(-pr.w));
            float d = sg - ta;
            s_a += d * d;
            n_rem += 1.f;
        }
    }
    s_rgb = block_sum(s_rgb, sh);
    n_valid = block_sum(n_valid, sh);
    s_a = block_sum(s_a, sh);
    n_rem = block_sum(n_rem, sh);
    if (threadIdx.x == 0) {
        atomicAdd(sums, (double)s_rgb);
        atomicAdd(sums + 1, (double)n_valid);
        atomicAdd(sums + 2, (double)s_a);
        atomicAdd(sums + 3, (double)n_rem);
    }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, float* __restrict__ out3) {
    float lr = (float)(sums[0] / sums[1]);  // 0/0 -> NaN exactly as the reference
    float la = (float)(sums[2] / sums[3]);
    out3[0] = lr + la;
    out3[1] = lr;
    out3[2] = la;
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pred,
                                                       const int* __restrict__ ext, const uint8_t* __restrict__ tok_mask, int R,
                                                       int p, const double* __restrict__ sums, const float* __restrict__ gout3,
                                                       float* __restrict__ dpred) {
    const int b = blockIdx.y;
    const long long V = (long long)R * R * R;
    const int n = R / p;
    const int ex = ext[b * 3], ey = ext[b * 3 + 1], ez = ext[b * 3 + 2];
    const float* xb = x + (long long)b * 4 * V;
    const float4* pb = reinterpret_cast<const float4*>(pred) + (long long)b * V;
    float4* db = reinterpret_cast<float4*>(dpred) + (long long)b * V;
    const float k_rgb = (float)(2.0 * (double)(gout3[0] + gout3[1]) / sums[1]);
    const float k_a = (float)(2.0 * (double)(gout3[0] + gout3[2]) / sums[3]);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
        int z = (int)(i % R);
        long long t = i / R;
        int y = (int)(t % R), xx = (int)(t / R);
        float4 pr = pb[i];
        float ta = xb[3 * V + i];
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ta > 0.01f) {
            d.x = k_rgb * (pr.x - xb[i]);
            d.y = k_rgb * (pr.y - xb[V + i]);
            d.z = k_rgb * (pr.z - xb[2 * V + i]);
        }
        if (xx < ex && y < ey && z < ez && tok_mask[((xx / p) * n + y / p) * n + z / p]) {
            float sg = 1.f / (1.f + expf(-pr.w));
            d.w = k_a * (sg - ta) * sg * (1.f - sg);
        }
        db[i] = d;
    }
}

int k_loss_fwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p, double* sums,
               float* out3, cudaStream_t st) {
    NMAE_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
    long long V = (long long)R * R * R;
    int gx = (int)min((long long)148 * 4, (V + 255) / 256);
    loss_fwd_kernel<<<dim3(gx, B), 256, 0, st>>>(x, pred, ext, tok_mask, R, p, sums);
    NMAE_LAUNCH_CHECK();
    loss_finalize_kernel<<<1, 1, 0, st>>>(sums, out3);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_loss_bwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p, const double* sums,
               const float* gout3, float* dpred, cudaStream_t st) {
    long long V = (long long)R * R * R;
    int gx = (int)min((long long)148 * 8, (V + 255) / 256);
    loss_bwd_kernel<<<dim3(gx, B), 256, 0, st>>>(x, pred, ext, tok_mask, R, p, sums, gout3, dpred);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------ optimizer
// Multi-tensor chunk table, one row of 6 int64 per chunk: {param, grad, exp_avg, exp_avg_sq, count, dst}.
// The table lives in device memory (the host mirror uploads it); one CTA per chunk.
#define TBL 6

__global__ void __launch_bounds__(256) multi_sumsq_kernel(const long long* __restrict__ table, double* __restrict__ out) {
    __shared__ float sh[32];
    const long long* row = table + (long long)blockIdx.x * TBL;
    const float* g = reinterpret_cast<const float*>(row[1]);
    int n = (int)row[4];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = g[i];
        s += v * v;
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

// dst[i] = grad[i]  (flat-bucket packing for the NCCL all-reduce)
__global__ void __launch_bounds__(256) multi_copy_kernel(const long long* __restrict__ table) {
    const long long* row = table + (long long)blockIdx.x * TBL;
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* d = reinterpret_cast<float*>(row[5]);
    int n = (int)row[4];
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = g[i];
}

// torch.nn.utils.clip_grad_norm_ (coef = clip/(norm+1e-6) clamped to 1) fused with torch.optim.AdamW
__global__ void __launch_bounds__(256) adamw_clip_kernel(const long long* __restrict__ table, const double* __restrict__ norm_sq,
                                                         float clip, float grad_scale, float lr, float b1, float b2, float eps,
                                                         float wd, float bc1, float bc2) {
    const long long* row = table + (long long)blockIdx.x * TBL;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    int n = (int)row[4];
    float coef = grad_scale;
    if (clip > 0.f) {
        float norm = sqrtf((float)(*norm_sq)) * grad_scale;
        coef *= fminf(clip / (norm + 1e-6f), 1.f);
    }
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float gi = g[i] * coef;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] * decay - step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

int k_multi_sumsq(const long long* table, int nchunks, double* out, cudaStream_t st) {
    NMAE_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
    if (nchunks == 0) return NMAE_OK;
    multi_sumsq_kernel<<<nchunks, 256, 0, st>>>(table, out);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_multi_copy(const long long* table, int nchunks, cudaStream_t st) {
    if (nchunks == 0) return NMAE_OK;
    multi_copy_kernel<<<nchunks, 256, 0, st>>>(table);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_adamw_clip(const long long* table, int nchunks, const double* norm_sq, float clip, float grad_scale, float lr, float b1,
                 float b2, float eps, float wd, float bc1, float bc2, cudaStream_t st) {
    if (nchunks == 0) return NMAE_OK;
    adamw_clip_kernel<<<nchunks, 256, 0, st>>>(table, norm_sq, clip, grad_scale, lr, b1, b2, eps, wd, bc1, bc2);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// FPN top-down step (nerf_rpn/model/fpn.py:148-158): fine += nearest-neighbour upsample of coarse to the fine size
// (F.interpolate(mode="nearest", size=...): source index = floor(dst * in / out)), channels-last, float4.
__global__ void __launch_bounds__(256) upsample_nearest_add_kernel(float4* __restrict__ fine, const float4* __restrict__ coarse, int B,
                                                                   int Xf, int Yf, int Zf, int Xc, int Yc, int Zc, int C4) {
    const long long total = (long long)B * Xf * Yf * Zf * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int c = (int)(t % C4); t /= C4;
        const int z = (int)(t % Zf); t /= Zf;
        const int y = (int)(t % Yf); t /= Yf;
        const int x = (int)(t % Xf);
        const int b = (int)(t / Xf);
        const int xs = (int)((long long)x * Xc / Xf), ys = (int)((long long)y * Yc / Yf), zs = (int)((long long)z * Zc / Zf);
        const float4 a = fine[i], u = __ldg(coarse + ((((long long)b * Xc + xs) * Yc + ys) * Zc + zs) * C4 + c);
        fine[i] = make_float4(a.x + u.x, a.y + u.y, a.z + u.z, a.w + u.w);
    }
}

int k_upsample_nearest_add(float* fine, const float* coarse, int B, int Xf, int Yf, int Zf, int Xc, int Yc, int Zc, int C,
                           cudaStream_t st) {
    NMAE_CHECK_ARG(C % 4 == 0, "upsample_nearest_add: channels must be a multiple of 4 (C=%d)", C);
    const long long total = (long long)B * Xf * Yf * Zf * (C / 4);
    if (total == 0) return NMAE_OK;
    const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
    upsample_nearest_add_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(fine), reinterpret_cast<const float4*>(coarse), B, Xf, Yf,
                                                      Zf, Xc, Yc, Zc, C / 4);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// nn.Upsample(size=..., mode="trilinear", align_corners=False) of the legacy decoder (swin_mae3d.py:593-610), channels-last.
// Source coordinate of output index d: max(0, (d + 0.5) * in / out - 0.5); neighbours floor / min(floor + 1, in - 1).
__device__ __forceinline__ void tri_coord(int d, int in, int out, int& i0, int& i1, float& w1) {
    float s = ((float)d + 0.5f) * ((float)in / (float)out) - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 < in - 1 ? i0 + 1 : i0;
    w1 = s - (float)i0;
}

__global__ void __launch_bounds__(256) upsample_trilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int Xi,
                                                                 int Yi, int Zi, int Xo, int Yo, int Zo, int C, int backward) {
    // forward: dst (fine) = interp(src (coarse)); backward: src is the FINE gradient, dst the coarse gradient (atomic scatter)
    const long long total = (long long)B * Xo * Yo * Zo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int c = (int)(t % C); t /= C;
        const int z = (int)(t % Zo); t /= Zo;
        const int y = (int)(t % Yo); t /= Yo;
        const int x = (int)(t % Xo);
        const int b = (int)(t / Xo);
        int x0, x1, y0, y1, z0, z1;
        float wx, wy, wz;
        tri_coord(x, Xi, Xo, x0, x1, wx);
        tri_coord(y, Yi, Yo, y0, y1, wy);
        tri_coord(z, Zi, Zo, z0, z1, wz);
        const long long base = (long long)b * Xi * Yi * Zi;
        float acc = 0.f;
        const float g = backward ? src[i] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int xs = (k & 4) ? x1 : x0, ys = (k & 2) ? y1 : y0, zs = (k & 1) ? z1 : z0;
            const float w = ((k & 4) ? wx : 1.f - wx) * ((k & 2) ? wy : 1.f - wy) * ((k & 1) ? wz : 1.f - wz);
            const long long j = (base + ((long long)xs * Yi + ys) * Zi + zs) * C + c;
            if (backward) atomicAdd(dst + j, w * g);
            else acc += w * __ldg(src + j);
        }
        if (!backward) dst[i] = acc;
    }
}

int k_upsample_trilinear(const float* src, float* dst, int B, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, int C, int backward,
                         cudaStream_t st) {
    const long long total = (long long)B * Xo * Yo * Zo * C;
    if (total == 0) return NMAE_OK;
    if (backward) NMAE_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * Xi * Yi * Zi * C, st));
    const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
    upsample_trilinear_kernel<<<grid, 256, 0, st>>>(src, dst, B, Xi, Yi, Zi, Xo, Yo, Zo, C, backward);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// ------------------------------------------------------------------------------------------------
// Scene ingest: the reference's CPU loader (nerf_rpn/datasets.py:88-104, 172-234) + pad_tensor (torch_utils.py:56-90) in one
// pass over the RAW scene as stored on disk.  src is the `rgbsigma` array (W, L, H, 4), float32 or uint8:
//   value = src[w][l][h][c];  density channel (c == 3): alpha = clip(1 - exp(-exp(sigma) / 100), 0, 1) when normalize != 0
//   (for uint8 files the reference writes alpha back into the uint8 array: truncation, reproduced); uint8 values are / 255;
//   channels first; augmentation = optional 90-degree rotation in the (W, L) plane (transpose + flip of the first axis) followed
//   by optional flips of axes 1 and 2 - applied as an index map; zero padding to R^3 into slot `dst` of the batch.
__device__ __forceinline__ float rh(float x) { return __half2float(__float2half_rn(x)); }

__global__ void __launch_bounds__(256) ingest_scene_kernel(const void* __restrict__ src, int is_u8, int normalize, int W, int L, int H,
                                                           int rot, int flip1, int flip2, float* __restrict__ dst, int R) {
    const int X = rot ? L : W, Y = rot ? W : L;      // extents after the rotation
    const long long n = (long long)R * R * R;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % R);
        long long t = i / R;
        const int y = (int)(t % R), x = (int)(t / R);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (x < X && y < Y && z < H) {
            const int xi = flip1 ? X - 1 - x : x, yi = flip2 ? Y - 1 - y : y;   // undo the flips (applied last)
            const int w = rot ? yi : xi, l = rot ? L - 1 - xi : yi;              // undo flip(transpose(t, 1, 2), [1])
            const long long s = (((long long)w * L + l) * H + z) * 4;
            if (is_u8) {
                const uchar4 u = *reinterpret_cast<const uchar4*>(static_cast<const unsigned char*>(src) + s);
                float a = (float)u.w;
                if (normalize) {
                    // numpy evaluates density_to_alpha on a uint8 array in float16 (every ufunc result rounded to half) and the
                    // reference stores the result back into the uint8 array (truncation): alpha is 1 for sigma >= 7, else 0
                    const float e1 = rh(expf(a)), e2 = rh(-e1 / 100.f), e3 = rh(expf(e2)), e4 = rh(1.f - e3);
                    a = (float)(unsigned char)fminf(fmaxf(e4, 0.f), 1.f);
                }
                v[0] = u.x / 255.f; v[1] = u.y / 255.f; v[2] = u.z / 255.f; v[3] = a / 255.f;
            } else {
                const float4 f = *reinterpret_cast<const float4*>(static_cast<const float*>(src) + s);
                v[0] = f.x; v[1] = f.y; v[2] = f.z;
                v[3] = normalize ? fminf(fmaxf(1.f - expf(-expf(f.w) / 100.f), 0.f), 1.f) : f.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) dst[c * n + i] = v[c];
    }
}

int k_ingest_scene(const void* src, int is_u8, int normalize, int W, int L, int H, int rot, int flip1, int flip2, float* dst, int R,
                   cudaStream_t st) {
    const long long n = (long long)R * R * R;
    const int g = (int)min((long long)148 * 16, (n + 255) / 256);
    ingest_scene_kernel<<<g, 256, 0, st>>>(src, is_u8, normalize, W, L, H, rot, flip1, flip2, dst, R);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}
