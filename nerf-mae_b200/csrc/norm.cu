// LayerNorm (tokens, warp per row), InstanceNorm statistics / apply (decoder volumes, NDHWC) and the
// column reductions that produce bias / affine gradients.  All HBM-bound: one coalesced pass per tensor.
#include "kernels.cuh"

// ------------------------------------------------------------------------------------------------
// Row sources for LayerNorm.  PLAIN: row r is x[r*C .. r*C+C).  MERGE: row r = (n,h2,w2,d2) of the
// 2x2x2 patch-merging gather (reference swin_mae3d.py:390-411): element e = blk*Cin + c with
// blk = dh + 2*dw + 4*dd reads token (2h2+dh, 2w2+dw, 2d2+dd) or 0 beyond the (odd) border.
// ------------------------------------------------------------------------------------------------
struct RowSrc {
    const float* x;
    int C;          // row length
    int merge;      // 0 plain, 1 merge-gather
    int H, W, D, Cin;  // merge: input token grid and channels (C == 8*Cin)
};

__device__ __forceinline__ long long merge_addr(const RowSrc& s, int r, int e) {
    int H2 = (s.H + 1) >> 1, W2 = (s.W + 1) >> 1, D2 = (s.D + 1) >> 1;
    int d2 = r % D2, t = r / D2;
    int w2 = t % W2; t /= W2;
    int h2 = t % H2, n = t / H2;
    int blk = e / s.Cin, c = e - blk * s.Cin;
    int h = 2 * h2 + (blk & 1), w = 2 * w2 + ((blk >> 1) & 1), d = 2 * d2 + (blk >> 2);
    if (h >= s.H || w >= s.W || d >= s.D) return -1;
    return ((((long long)n * s.H + h) * s.W + w) * s.D + d) * s.Cin + c;
}
__device__ __forceinline__ float row_load(const RowSrc& s, int r, int e) {
    if (!s.merge) return s.x[(long long)r * s.C + e];
    long long a = merge_addr(s, r, e);
    return a < 0 ? 0.f : s.x[a];
}

// y = LN(x)*w + b (+ pos[r % pos_rows]) ; rows with mask[r % pos_rows] != 0 are replaced by mask_token
__global__ void __launch_bounds__(256) ln_fwd_kernel(RowSrc s, int rows, const float* __restrict__ w, const float* __restrict__ b,
                                                     float eps, const float* __restrict__ pos, int pos_rows,
                                                     const uint8_t* __restrict__ mask, const float* __restrict__ mask_token,
                                                     float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
    int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int C = s.C;
    float sum = 0.f;
    for (int e = lane; e < C; e += 32) sum += row_load(s, r, e);
    float mu = warp_sum(sum) / C;
    float v = 0.f;
    for (int e = lane; e < C; e += 32) {
        float d = row_load(s, r, e) - mu;
        v += d * d;
    }
    float rs = rsqrtf(warp_sum(v) / C + eps);
    if (lane == 0) {
        mean[r] = mu;
        rstd[r] = rs;
    }
    bool masked = mask && mask[r % pos_rows];
    const float* prow = pos ? pos + (long long)(r % pos_rows) * C : nullptr;
    float* yr = y + (long long)r * C;
    for (int e = lane; e < C; e += 32) {
        float o = (row_load(s, r, e) - mu) * rs * w[e] + b[e];
        if (prow) o += prow[e];
        if (masked) o = mask_token[e];
        yr[e] = o;
    }
}

// dx = rstd * (g - mean(g) - xhat*mean(g*xhat)), g = dy*w ; masked rows get dx = 0.
// MERGE rows scatter dx back to the token grid (each token belongs to exactly one merged row).
__global__ void __launch_bounds__(256) ln_bwd_kernel(RowSrc s, int rows, const float* __restrict__ w, const float* __restrict__ dy,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const uint8_t* __restrict__ mask, int pos_rows, float* dx,
                                                     const float* add_src) {
    int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int C = s.C;
    bool masked = mask && mask[r % pos_rows];
    float mu = mean[r], rs = rstd[r];
    const float* dyr = dy + (long long)r * C;
    float s1 = 0.f, s2 = 0.f;
    if (!masked) {
        for (int e = lane; e < C; e += 32) {
            float g = dyr[e] * w[e];
            float xh = (row_load(s, r, e) - mu) * rs;
            s1 += g;
            s2 += g * xh;
        }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int e = lane; e < C; e += 32) {
        float o = 0.f;
        if (!masked) {
            float g = dyr[e] * w[e];
            float xh = (row_load(s, r, e) - mu) * rs;
            o = rs * (g - s1 - xh * s2);
        }
        long long a = s.merge ? merge_addr(s, r, e) : (long long)r * C + e;
        if (a >= 0) dx[a] = add_src ? add_src[a] + o : o;
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident LayerNorm (C % 4 == 0, C <= 3072): a warp holds RPW rows as float4 (lane l owns float4 l, l+32, ... of each row,
// VPT per row), so every row is read from memory ONCE, all loads of the RPW rows are in flight together, and mean / variance come from
// the registers.  The first kernels (above) walked each row three times with scalar loads - three dependent round trips per 384-byte
// row - and ran at 1.2-1.6 TB/s on the stage-1 tensors (0.12-0.16 ms where 0.03 would do; profiles/r2_layernorm.txt).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 row_load4(const RowSrc& s, int r, int f) {       // float4 f of row r (zero for padded merge tokens)
    if (!s.merge) return __ldg(reinterpret_cast<const float4*>(s.x) + (long long)r * (s.C >> 2) + f);
    const long long a = merge_addr(s, r, f * 4);                                   // Cin % 4 == 0: a float4 stays inside one token
    return a < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(s.x + a));
}

template <int VPT, int RPW>
__global__ void __launch_bounds__(256) ln_fwd_v_kernel(RowSrc s, int rows, const float* __restrict__ w, const float* __restrict__ b,
                                                       float eps, const float* __restrict__ pos, int pos_rows,
                                                       const uint8_t* __restrict__ mask, const float* __restrict__ mask_token,
                                                       float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
    const int lane = threadIdx.x & 31, C = s.C, C4 = C >> 2;
    const int r0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW;
    if (r0 >= rows) return;
    float4 v[RPW][VPT];
#pragma unroll
    for (int rr = 0; rr < RPW; rr++)
#pragma unroll
        for (int i = 0; i < VPT; i++) {
            const int f = lane + 32 * i;
            v[rr][i] = (r0 + rr < rows && f < C4) ? row_load4(s, r0 + rr, f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    float4 w4[VPT], b4[VPT];
#pragma unroll
    for (int i = 0; i < VPT; i++) {
        const int f = lane + 32 * i;
        w4[i] = f < C4 ? __ldg(reinterpret_cast<const float4*>(w) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        b4[i] = f < C4 ? __ldg(reinterpret_cast<const float4*>(b) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int rr = 0; rr < RPW; rr++) {
        const int r = r0 + rr;
        if (r >= rows) break;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; i++) sum += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
        const float mu = warp_sum(sum) / C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; i++) {
            if (lane + 32 * i < C4) {
                const float dx = v[rr][i].x - mu, dy = v[rr][i].y - mu, dz = v[rr][i].z - mu, dw = v[rr][i].w - mu;
                q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
        const float rs = rsqrtf(warp_sum(q) / C + eps);
        if (lane == 0) {
            mean[r] = mu;
            rstd[r] = rs;
        }
        const int pr = r % pos_rows;
        const bool masked = mask && mask[pr];
        float4* yr = reinterpret_cast<float4*>(y) + (long long)r * C4;
#pragma unroll
        for (int i = 0; i < VPT; i++) {
            const int f = lane + 32 * i;
            if (f >= C4) continue;
            float4 o;
            if (masked) {
                o = __ldg(reinterpret_cast<const float4*>(mask_token) + f);
            } else {
                o.x = (v[rr][i].x - mu) * rs * w4[i].x + b4[i].x; o.y = (v[rr][i].y - mu) * rs * w4[i].y + b4[i].y;
                o.z = (v[rr][i].z - mu) * rs * w4[i].z + b4[i].z; o.w = (v[rr][i].w - mu) * rs * w4[i].w + b4[i].w;
                if (pos) {
                    const float4 pp = __ldg(reinterpret_cast<const float4*>(pos) + (long long)pr * C4 + f);
                    o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
                }
            }
            yr[f] = o;
        }
    }
}

// dx = rstd * (g - mean(g) - xhat*mean(g*xhat)), g = dy*w (masked rows: 0; merge rows scatter to the token grid), and - same pass -
// dgamma[c] += sum_r dy*xhat, dbeta[c] += sum_r dy: the warps walk the rows grid-stride with per-lane column accumulators that are
// combined per CTA in shared memory and flushed with one atomic per column (the separate column-reduction kernel re-read x and dy).
template <int VPT, int RPW>
__global__ void __launch_bounds__(256) ln_bwd_v_kernel(RowSrc s, int rows, const float* __restrict__ w, const float* __restrict__ dy,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const uint8_t* __restrict__ mask, int pos_rows, float* dx, const float* add_src,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ float s_acc[];      // [2*C]: dgamma, dbeta partials of this CTA
    const int lane = threadIdx.x & 31, C = s.C, C4 = C >> 2;
    if (dgamma) {
        for (int i = threadIdx.x; i < 2 * C; i += 256) s_acc[i] = 0.f;
        __syncthreads();
    }
    float4 w4[VPT], ag[VPT], ab[VPT];
#pragma unroll
    for (int i = 0; i < VPT; i++) {
        const int f = lane + 32 * i;
        w4[i] = f < C4 ? __ldg(reinterpret_cast<const float4*>(w) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int warps = gridDim.x * 8;
    for (int r0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW; r0 < rows; r0 += warps * RPW) {
        float4 xv[RPW][VPT], dv[RPW][VPT];
        float mu[RPW], rs[RPW];
        bool live[RPW];
#pragma unroll
        for (int rr = 0; rr < RPW; rr++) {
            const int r = r0 + rr;
            live[rr] = r < rows && !(mask && mask[r % pos_rows]);
            mu[rr] = live[rr] ? mean[r] : 0.f;
            rs[rr] = live[rr] ? rstd[r] : 0.f;
#pragma unroll
            for (int i = 0; i < VPT; i++) {
                const int f = lane + 32 * i;
                xv[rr][i] = dv[rr][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[rr] && f < C4) {
                    xv[rr][i] = row_load4(s, r, f);
                    dv[rr][i] = __ldg(reinterpret_cast<const float4*>(dy) + (long long)r * C4 + f);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < RPW; rr++) {
            const int r = r0 + rr;
            if (r >= rows) break;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < VPT; i++) {
                // xhat in place of x, g = dy*w in registers
                float4& x4 = xv[rr][i];
                const float4 d4 = dv[rr][i];
                if (lane + 32 * i < C4) {
                    x4.x = (x4.x - mu[rr]) * rs[rr]; x4.y = (x4.y - mu[rr]) * rs[rr]; x4.z = (x4.z - mu[rr]) * rs[rr]; x4.w = (x4.w - mu[rr]) * rs[rr];
                }
                if (dgamma) {
                    ag[i].x += d4.x * x4.x; ag[i].y += d4.y * x4.y; ag[i].z += d4.z * x4.z; ag[i].w += d4.w * x4.w;
                    ab[i].x += d4.x; ab[i].y += d4.y; ab[i].z += d4.z; ab[i].w += d4.w;
                }
                const float gx = d4.x * w4[i].x, gy = d4.y * w4[i].y, gz = d4.z * w4[i].z, gw = d4.w * w4[i].w;
                s1 += (gx + gy) + (gz + gw);
                s2 += (gx * x4.x + gy * x4.y) + (gz * x4.z + gw * x4.w);
            }
            if (!dx) continue;
            s1 = warp_sum(s1) / C;
            s2 = warp_sum(s2) / C;
#pragma unroll
            for (int i = 0; i < VPT; i++) {
                const int f = lane + 32 * i;
                if (f >= C4) continue;
                const float4 x4 = xv[rr][i], d4 = dv[rr][i];
                float4 o;
                o.x = rs[rr] * (d4.x * w4[i].x - s1 - x4.x * s2); o.y = rs[rr] * (d4.y * w4[i].y - s1 - x4.y * s2);
                o.z = rs[rr] * (d4.z * w4[i].z - s1 - x4.z * s2); o.w = rs[rr] * (d4.w * w4[i].w - s1 - x4.w * s2);
                const long long a = s.merge ? merge_addr(s, r, f * 4) : (long long)r * C + f * 4;
                if (a < 0) continue;
                if (add_src) {
                    const float4 t = *reinterpret_cast<const float4*>(add_src + a);
                    o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                }
                *reinterpret_cast<float4*>(dx + a) = o;
            }
        }
    }
    if (dgamma) {
#pragma unroll
        for (int i = 0; i < VPT; i++) {
            const int f = lane + 32 * i;
            if (f >= C4) continue;
            float* g = s_acc + f * 4;
            float* bb = s_acc + C + f * 4;
            atomicAdd(g, ag[i].x); atomicAdd(g + 1, ag[i].y); atomicAdd(g + 2, ag[i].z); atomicAdd(g + 3, ag[i].w);
            atomicAdd(bb, ab[i].x); atomicAdd(bb + 1, ab[i].y); atomicAdd(bb + 2, ab[i].z); atomicAdd(bb + 3, ab[i].w);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += 256) {
            atomicAdd(dgamma + i, s_acc[i]);
            atomicAdd(dbeta + i, s_acc[C + i]);
        }
    }
}

// dgamma[c] += sum_r dy*xhat ; dbeta[c] += sum_r dy   (masked rows excluded)
__device__ __forceinline__ void colred_finish(float v, float* dst, float (*sh)[33]);
__global__ void __launch_bounds__(256) ln_param_grad_kernel(RowSrc s, int rows, int rows_per_cta, const float* __restrict__ dy,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const uint8_t* __restrict__ mask, int pos_rows,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float g = 0.f, bta = 0.f;
    if (c < s.C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if (mask && mask[r % pos_rows]) continue;
            float d = dy[(long long)r * s.C + c];
            g += d * (row_load(s, r, c) - mean[r]) * rstd[r];
            bta += d;
        }
    }
    colred_finish(g, c < s.C ? dgamma + c : nullptr, sh);
    colred_finish(bta, c < s.C ? dbeta + c : nullptr, sh);
}

// Column reductions use 32x8 thread blocks: threadIdx.x walks 32 consecutive columns (one 128 B line per row),
// threadIdx.y strides over rows; partials are combined through shared memory and one atomic per column and CTA.
__device__ __forceinline__ void colred_finish(float v, float* dst, float (*sh)[33]) {
    sh[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) a += sh[j][threadIdx.x];
        if (dst) atomicAdd(dst, a);
    }
    __syncthreads();
}

// out[c] += sum_r x[r*ld + c] over rows selected by mask (mask given: only rows with mask != 0)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int rows, int C, long long ld, int rows_per_cta,
                                                     const uint8_t* __restrict__ mask, int pos_rows, float* __restrict__ out) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float a = 0.f;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            if (mask && !mask[r % pos_rows]) continue;
            a += x[(long long)r * ld + c];
        }
    }
    colred_finish(a, c < C ? out + c : nullptr, sh);
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm3d (no affine, biased variance, eps; reference unetr_block.py:77) on NDHWC volumes.
// stats[b][c] = {sum, sum of squares} accumulated in double so that var = E[x^2]-E[x]^2 is safe.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void colred_finish_d(float v, double* dst, float (*sh)[33]) {
    sh[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) a += sh[j][threadIdx.x];
        if (dst) atomicAdd(dst, (double)a);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ x, int V, int C, int rows_per_cta,
                                                       double* __restrict__ stats) {
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(V, r0 + rows_per_cta);
    const float* xb = x + (long long)b * V * C;
    float s = 0.f, ss = 0.f;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            float v = xb[(long long)r * C + c];
            s += v;
            ss += v * v;
        }
    }
    double* o = c < C ? stats + ((long long)b * C + c) * 2 : nullptr;
    colred_finish_d(s, o, sh);
    colred_finish_d(ss, o ? o + 1 : nullptr, sh);
}

__device__ __forceinline__ void in_mean_rstd(const double* st, int V, float eps, float& mu, float& rs) {
    double m = st[0] / V;
    double var = st[1] / V - m * m;
    if (var < 0) var = 0;
    mu = (float)m;
    rs = (float)(1.0 / sqrt(var + (double)eps));
}

// out = lrelu( norm(x) + R ),  R = 0 | res | norm(res) (res_stats != null)
__global__ void __launch_bounds__(256) in_act_fwd_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                         const float* __restrict__ res, const double* __restrict__ res_stats,
                                                         int V, int C, float eps, float slope, float* __restrict__ out) {
    extern __shared__ float sm[];  // mu[C], rs[C], mu3[C], rs3[C]
    int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        in_mean_rstd(stats + ((long long)b * C + c) * 2, V, eps, sm[c], sm[C + c]);
        if (res_stats) in_mean_rstd(res_stats + ((long long)b * C + c) * 2, V, eps, sm[2 * C + c], sm[3 * C + c]);
    }
    __syncthreads();
    long long n = (long long)V * C, base = (long long)b * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float v = (x[base + i] - sm[c]) * sm[C + c];
        if (res) {
            float r = res[base + i];
            if (res_stats) r = (r - sm[2 * C + c]) * sm[3 * C + c];
            v += r;
        }
        out[base + i] = v >= 0.f ? v : v * slope;
    }
}

// sums[b][c] = {sum g, sum g*xhat, sum g*xhat3},  g = dout * lrelu'(out)
template <bool DP4>
__global__ void __launch_bounds__(256) in_bwd_sums_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                          const float* __restrict__ x, const double* __restrict__ stats,
                                                          const float* __restrict__ x3, const double* __restrict__ stats3, int V,
                                                          int C, int rows_per_cta, float eps, float slope,
                                                          double* __restrict__ sums, float* __restrict__ amax,
                                                          const float4* __restrict__ dp4, const float* __restrict__ w4) {
    // dp4 != NULL: dout is not materialised - it is the input gradient of a 1x1x1 convolution C -> 4 (the output block,
    // unetr_block.py:96-116): dout[v][c] = sum_k w4[k][c] * dp4[v][k]
    __shared__ float sh[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(V, r0 + rows_per_cta);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, gmax = 0.f;
    if (c < C) {
        float mu, rs, mu3 = 0.f, rs3 = 0.f;
        in_mean_rstd(stats + ((long long)b * C + c) * 2, V, eps, mu, rs);
        if (x3) in_mean_rstd(stats3 + ((long long)b * C + c) * 2, V, eps, mu3, rs3);
        const long long base = (long long)b * V * C;
        float wc[4] = {0.f, 0.f, 0.f, 0.f};
        if (DP4) {
#pragma unroll
            for (int k = 0; k < 4; k++) wc[k] = w4[k * C + c];
        }
#pragma unroll 4
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            long long i = base + (long long)r * C + c;
            const float xh = (x[i] - mu) * rs;
            float d;
            if (DP4) {
                const float4 q = __ldg(dp4 + (long long)b * V + r);
                d = wc[0] * q.x + wc[1] * q.y + wc[2] * q.z + wc[3] * q.w;
            } else {
                d = dout[i];
            }
            // out == NULL: forward without residual, lrelu(xhat) has the sign of xhat
            float g = d * ((out ? out[i] : xh) > 0.f ? 1.f : slope);
            s0 += g;
            s1 += g * xh;
            if (x3) s2 += g * (x3[i] - mu3) * rs3;
            gmax = fmaxf(gmax, fabsf(g));
        }
    }
    if (amax) {     // max |g| over the tensor (fp16 gradient images scale by it); non-negative floats order like their bit patterns
        gmax = warp_max(gmax);
        if (threadIdx.x == 0 && gmax > 0.f && gmax < 3.0e38f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(gmax));
    }
    double* o = c < C ? sums + ((long long)b * C + c) * 3 : nullptr;
    colred_finish_d(s0, o, sh);
    colred_finish_d(s1, o ? o + 1 : nullptr, sh);
    if (x3) colred_finish_d(s2, o ? o + 2 : nullptr, sh);
}

// dx = rs*(g - S0/V - xhat*S1/V) ; dx3 likewise with xhat3/S2 ; dres = g (identity residual) if requested
__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                           const float* __restrict__ x, const double* __restrict__ stats,
                                                           const float* __restrict__ x3, const double* __restrict__ stats3,
                                                           const double* __restrict__ sums, int V, int C, float eps, float slope,
                                                           float* __restrict__ dx, float* __restrict__ dx3, float* __restrict__ dres) {
    extern __shared__ float sm[];  // mu, rs, mu3, rs3, m0, m1, m2  (7*C)
    int b = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        in_mean_rstd(stats + ((long long)b * C + c) * 2, V, eps, sm[c], sm[C + c]);
        if (x3) in_mean_rstd(stats3 + ((long long)b * C + c) * 2, V, eps, sm[2 * C + c], sm[3 * C + c]);
        const double* s = sums + ((long long)b * C + c) * 3;
        sm[4 * C + c] = (float)(s[0] / V);
        sm[5 * C + c] = (float)(s[1] / V);
        sm[6 * C + c] = x3 ? (float)(s[2] / V) : 0.f;
    }
    __syncthreads();
    long long n = (long long)V * C, base = (long long)b * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float xh = (x[base + i] - sm[c]) * sm[C + c];
        float g = dout[base + i] * ((out ? out[base + i] : xh) > 0.f ? 1.f : slope);
        dx[base + i] = sm[C + c] * (g - sm[4 * C + c] - xh * sm[5 * C + c]);
        if (x3) {
            float xh3 = (x3[base + i] - sm[2 * C + c]) * sm[3 * C + c];
            dx3[base + i] = sm[3 * C + c] * (g - sm[4 * C + c] - xh3 * sm[6 * C + c]);
        }
        if (dres) dres[base + i] = g;
    }
}


// ---- float4 variants (C % 4 == 0).  The launcher makes the grid stride a multiple of C/4, so a thread keeps the same
// four channels for its whole loop: per-channel constants live in registers, no per-element index arithmetic.
__device__ __forceinline__ float lrelu_sel(float s, float slope) { return s > 0.f ? 1.f : slope; }

__global__ void __launch_bounds__(256) in_act_fwd_v4_kernel(const float4* __restrict__ x, const double* __restrict__ stats,
                                                            const float4* __restrict__ res, const double* __restrict__ res_stats,
                                                            int V, int C, float eps, float slope, float4* __restrict__ out) {
    const int b = blockIdx.y, C4 = C >> 2;
    const long long n4 = (long long)V * C4, base = (long long)b * n4;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const int c = (int)(i0 % C4) * 4;
    float mu[4], rs[4], mu3[4], rs3[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        in_mean_rstd(stats + ((long long)b * C + c + e) * 2, V, eps, mu[e], rs[e]);
        mu3[e] = 0.f; rs3[e] = 1.f;
        if (res_stats) in_mean_rstd(res_stats + ((long long)b * C + c + e) * 2, V, eps, mu3[e], rs3[e]);
    }
    for (long long i = i0; i < n4; i += stride) {
        const float4 v = __ldcs(x + base + i);
        float o[4] = {(v.x - mu[0]) * rs[0], (v.y - mu[1]) * rs[1], (v.z - mu[2]) * rs[2], (v.w - mu[3]) * rs[3]};
        if (res) {
            const float4 r = __ldcs(res + base + i);
            o[0] += (r.x - mu3[0]) * rs3[0]; o[1] += (r.y - mu3[1]) * rs3[1];
            o[2] += (r.z - mu3[2]) * rs3[2]; o[3] += (r.w - mu3[3]) * rs3[3];
        }
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] = o[e] >= 0.f ? o[e] : o[e] * slope;
        out[base + i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// in_act_fwd fused with the 1x1x1 output convolution C -> 4 that consumes its result (UnetOutBlock, unetr_block.py:96-116):
// out = lrelu(IN(x) + R) is written as before and pred[v][0..3] = W_out . out[v][:] + b_out is formed from the values while they are on
// chip (the separate pass re-read the 3.1 GB `out` volume, thread per row: 0.89 ms at decoder1).  A CTA walks tiles of 256 voxels:
// phase 1 - float4 units, coalesced loads / stores, results also parked in shared memory ([row][C+1]); phase 2 - one thread per voxel
// takes its row from shared memory (conflict-free) against the weights (broadcast).
__global__ void __launch_bounds__(256) in_act_fwd_out_kernel(const float4* __restrict__ x, const double* __restrict__ stats,
                                                             const float4* __restrict__ res, const double* __restrict__ res_stats,
                                                             int V, int C, float eps, float slope, float4* __restrict__ out,
                                                             const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                             float4* __restrict__ pred) {
    extern __shared__ __align__(16) float sm_o[];
    const int b = blockIdx.y, C4 = C >> 2, LD = C + 1;
    float* s_mu = sm_o;                 // [C] mean, [C] rstd, [C] mean3, [C] rstd3
    float4* s_w = reinterpret_cast<float4*>(sm_o + 4 * C);       // [C] {w[0][k], w[1][k], w[2][k], w[3][k]}
    float* tile = sm_o + 8 * C;         // [256][C + 1]
    for (int c = threadIdx.x; c < C; c += 256) {
        in_mean_rstd(stats + ((long long)b * C + c) * 2, V, eps, s_mu[c], s_mu[C + c]);
        s_mu[2 * C + c] = 0.f; s_mu[3 * C + c] = 1.f;
        if (res_stats) in_mean_rstd(res_stats + ((long long)b * C + c) * 2, V, eps, s_mu[2 * C + c], s_mu[3 * C + c]);
        s_w[c] = make_float4(w_out[c], w_out[C + c], w_out[2 * C + c], w_out[3 * C + c]);
    }
    const float4 b4 = b_out ? make_float4(b_out[0], b_out[1], b_out[2], b_out[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const long long base4 = (long long)b * V * C4;
    const int units = 256 * C4;
    for (long long v0 = (long long)blockIdx.x * 256; v0 < V; v0 += (long long)gridDim.x * 256) {
        const int rows = (int)min((long long)256, V - v0);
        for (int u0 = 0; u0 < units; u0 += 4 * 256) {           // four units per thread in flight
            float4 xv[4], rv[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int u = u0 + t * 256 + threadIdx.x;
                xv[t] = rv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < units && u / C4 < rows) {
                    xv[t] = __ldcs(x + base4 + v0 * C4 + u);
                    if (res) rv[t] = __ldcs(res + base4 + v0 * C4 + u);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int u = u0 + t * 256 + threadIdx.x;
                if (u >= units) continue;
                const int row = u / C4, c = (u - row * C4) * 4;
                if (row >= rows) continue;
                const float xa[4] = {xv[t].x, xv[t].y, xv[t].z, xv[t].w}, ra[4] = {rv[t].x, rv[t].y, rv[t].z, rv[t].w};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float a = (xa[e] - s_mu[c + e]) * s_mu[C + c + e];
                    if (res) a += (ra[e] - s_mu[2 * C + c + e]) * s_mu[3 * C + c + e];
                    o[e] = a >= 0.f ? a : a * slope;
                    tile[row * LD + c + e] = o[e];
                }
                out[base4 + v0 * C4 + u] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < rows) {
            const float* tr = tile + threadIdx.x * LD;
            float4 acc = b4;
#pragma unroll 8
            for (int k = 0; k < C; k++) {
                const float v = tr[k];
                const float4 w = s_w[k];
                acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
            }
            pred[(long long)b * V + v0 + threadIdx.x] = acc;
        }
        __syncthreads();
    }
}

// as in_bwd_apply_kernel; additionally accumulates the column sums of dx / dx3 (the bias gradients of the convolutions
// that produced x / x3) when dbias / dbias3 are given.
__global__ void __launch_bounds__(256) in_bwd_apply_v4_kernel(const float4* __restrict__ dout, const float4* __restrict__ out,
                                                              const float4* __restrict__ x, const double* __restrict__ stats,
                                                              const float4* __restrict__ x3, const double* __restrict__ stats3,
                                                              const double* __restrict__ sums, int V, int C, float eps, float slope,
                                                              float4* __restrict__ dx, float4* __restrict__ dx3,
                                                              float4* __restrict__ dres, float* __restrict__ dbias,
                                                              float* __restrict__ dbias3) {
    extern __shared__ float sacc[];   // [2*C] bias-gradient partial sums of this CTA
    const int b = blockIdx.y, C4 = C >> 2;
    const long long n4 = (long long)V * C4, base = (long long)b * n4;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const int c = (int)(i0 % C4) * 4;
    float mu[4], rs[4], mu3[4], rs3[4], m0[4], m1[4], m2[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        in_mean_rstd(stats + ((long long)b * C + c + e) * 2, V, eps, mu[e], rs[e]);
        mu3[e] = 0.f; rs3[e] = 1.f;
        if (x3) in_mean_rstd(stats3 + ((long long)b * C + c + e) * 2, V, eps, mu3[e], rs3[e]);
        const double* s = sums + ((long long)b * C + c + e) * 3;
        m0[e] = (float)(s[0] / V);
        m1[e] = (float)(s[1] / V);
        m2[e] = x3 ? (float)(s[2] / V) : 0.f;
    }
    if (dbias || dbias3) {
        for (int j = threadIdx.x; j < 2 * C; j += blockDim.x) sacc[j] = 0.f;
        __syncthreads();
    }
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a3[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = i0; i < n4; i += stride) {
        const float4 dv = __ldcs(dout + base + i), xv = __ldcs(x + base + i);
        const float d[4] = {dv.x, dv.y, dv.z, dv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
        float xh[4], g[4], o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) xh[e] = (xx[e] - mu[e]) * rs[e];
        if (out) {
            const float4 ov = __ldcs(out + base + i);
            g[0] = d[0] * lrelu_sel(ov.x, slope); g[1] = d[1] * lrelu_sel(ov.y, slope);
            g[2] = d[2] * lrelu_sel(ov.z, slope); g[3] = d[3] * lrelu_sel(ov.w, slope);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) g[e] = d[e] * lrelu_sel(xh[e], slope);
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            o[e] = rs[e] * (g[e] - m0[e] - xh[e] * m1[e]);
            a1[e] += o[e];
        }
        dx[base + i] = make_float4(o[0], o[1], o[2], o[3]);
        if (x3) {
            const float4 x3v = __ldcs(x3 + base + i);
            const float t[4] = {x3v.x, x3v.y, x3v.z, x3v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                o[e] = rs3[e] * (g[e] - m0[e] - (t[e] - mu3[e]) * rs3[e] * m2[e]);
                a3[e] += o[e];
            }
            dx3[base + i] = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (dres) dres[base + i] = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (dbias || dbias3) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (dbias) atomicAdd(&sacc[c + e], a1[e]);
            if (dbias3) atomicAdd(&sacc[C + c + e], a3[e]);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < C; j += blockDim.x) {
            if (dbias) atomicAdd(dbias + j, sacc[j]);
            if (dbias3) atomicAdd(dbias3 + j, sacc[C + j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ float4 column reductions
// InstanceNorm statistics and the reduction pass of its backward with the channel group fixed per thread (grid stride a multiple
// of C/4, like the apply kernels): float4 loads, a batch of independent loads in flight per thread, per-thread partial sums,
// combined through shared-memory atomics and ONE double atomic per channel and CTA.  The (32 x 8)-thread column kernels above
// stay for channel counts that are not a multiple of 4.
// SQ: also accumulate squares (InstanceNorm statistics, double outputs {sum, sumsq}); !SQ: plain column sums into float out[C]
template <bool SQ>
__global__ void __launch_bounds__(256) in_stats_v4_kernel(const float4* __restrict__ x, int V, int C, double* __restrict__ stats,
                                                          float* __restrict__ colsum_out) {
    extern __shared__ float sacc[];   // [2*C]
    const int b = blockIdx.y, C4 = C >> 2;
    const long long n4 = (long long)V * C4, base = (long long)b * n4;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const int c = (int)(i0 % C4) * 4;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    long long i = i0;
    for (; i + 7 * stride < n4; i += 8 * stride) {       // eight independent 16-byte loads in flight per thread
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __ldg(x + base + i + k * stride);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            s[0] += v[k].x; s[1] += v[k].y; s[2] += v[k].z; s[3] += v[k].w;
            if (SQ) { q[0] += v[k].x * v[k].x; q[1] += v[k].y * v[k].y; q[2] += v[k].z * v[k].z; q[3] += v[k].w * v[k].w; }
        }
    }
    for (; i < n4; i += stride) {
        const float4 a = __ldg(x + base + i);
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        if (SQ) { q[0] += a.x * a.x; q[1] += a.y * a.y; q[2] += a.z * a.z; q[3] += a.w * a.w; }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
        atomicAdd(&sacc[c + e], s[e]);
        if (SQ) atomicAdd(&sacc[C + c + e], q[e]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
        if (SQ) {
            atomicAdd(stats + ((long long)b * C + j) * 2, (double)sacc[j]);
            atomicAdd(stats + ((long long)b * C + j) * 2 + 1, (double)sacc[C + j]);
        } else {
            atomicAdd(colsum_out + j, sacc[j]);
        }
    }
}

// WG (with DP4 and HAS_OUT): also accumulate the weight / bias gradient of that 1x1x1 output convolution, dw[k][c] = sum dp4[v][k] *
// out[v][c] and db[k] = sum dp4[v][k] - both operands are already in registers (a separate pass re-read `out`: 1.5 ms at decoder1).
template <bool DP4, bool HAS_OUT, bool HAS_X3, bool WG = false>
__global__ void __launch_bounds__(256) in_bwd_sums_v4_kernel(const float4* __restrict__ dout, const float4* __restrict__ out,
                                                             const float4* __restrict__ x, const double* __restrict__ stats,
                                                             const float4* __restrict__ x3, const double* __restrict__ stats3, int V,
                                                             int C, float eps, float slope, double* __restrict__ sums,
                                                             float* __restrict__ amax, const float4* __restrict__ dp4,
                                                             const float* __restrict__ w4, float* __restrict__ dw_out = nullptr,
                                                             float* __restrict__ db_out = nullptr) {
    extern __shared__ float sacc[];   // [3*C] (+ [4*C + 4] with WG)
    const int b = blockIdx.y, C4 = C >> 2;
    const long long n4 = (long long)V * C4, base = (long long)b * n4;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const int c4 = (int)(i0 % C4), c = c4 * 4;
    for (int i = threadIdx.x; i < (WG ? 7 * C + 4 : 3 * C); i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float mu[4], rs[4], mu3[4], rs3[4], w[4][4];
    float wg[4][4], qs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int e = 0; e < 4; e++) wg[k][e] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        in_mean_rstd(stats + ((long long)b * C + c + e) * 2, V, eps, mu[e], rs[e]);
        mu3[e] = 0.f; rs3[e] = 1.f;
        if (HAS_X3) in_mean_rstd(stats3 + ((long long)b * C + c + e) * 2, V, eps, mu3[e], rs3[e]);
#pragma unroll
        for (int k = 0; k < 4; k++) w[k][e] = DP4 ? w4[k * C + c + e] : 0.f;
    }
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, gmax = 0.f;
    // dv: dout (or, with DP4, the dp4 record of the voxel: dout[v][c] = sum_k w4[k][c] * dp4[v][k], the input gradient of the 1x1x1
    // output convolution, never materialised)
    auto accum = [&](const float4& dv, const float4& ov, const float4& xv, const float4& x3v) {
        const float o[4] = {ov.x, ov.y, ov.z, ov.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w}, y3[4] = {x3v.x, x3v.y, x3v.z, x3v.w};
        float d[4] = {dv.x, dv.y, dv.z, dv.w};
        if (DP4) {
            const float q[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int e = 0; e < 4; e++) d[e] = w[0][e] * q[0] + w[1][e] * q[1] + w[2][e] * q[2] + w[3][e] * q[3];
            if (WG) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    qs[k] += q[k];
#pragma unroll
                    for (int e = 0; e < 4; e++) wg[k][e] += q[k] * o[e];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float xh = (xx[e] - mu[e]) * rs[e];
            // out absent: forward without residual, lrelu(xhat) has the sign of xhat
            const float g = d[e] * ((HAS_OUT ? o[e] : xh) > 0.f ? 1.f : slope);
            s0[e] += g;
            s1[e] += g * xh;
            if (HAS_X3) s2[e] += g * (y3[e] - mu3[e]) * rs3[e];
            gmax = fmaxf(gmax, fabsf(g));
        }
    };
    // the stride is a multiple of C/4: the voxel of element i0 + k*stride is v0 + k*vstep (no division in the loop)
    const long long v0 = i0 / C4, vstep = stride / C4;
    // WG: the loads go through volatile asm so that all six of an iteration are issued before the first use (left to itself the
    // compiler interleaved them with the arithmetic - two loads in flight per thread, 2.7 ms instead of 1.2)
    auto ld4 = [&](const float4* ptr) -> float4 {
        if (!WG) return __ldg(ptr);
        float4 v;
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
        return v;
    };
    auto load_d = [&](long long i, long long vox) -> float4 {
        if (DP4) return ld4(dp4 + (long long)b * V + vox);
        return ld4(dout + base + i);
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    long long i = i0, vox = v0;
    for (; i + stride < n4; i += 2 * stride, vox += 2 * vstep) {      // two independent element groups in flight (up to 8 loads)
        const long long j = i + stride;
        const float4 da = load_d(i, vox), db = load_d(j, vox + vstep);
        const float4 xa = ld4(x + base + i), xb = ld4(x + base + j);
        const float4 oa = HAS_OUT ? ld4(out + base + i) : z4, ob = HAS_OUT ? ld4(out + base + j) : z4;
        const float4 ya = HAS_X3 ? __ldg(x3 + base + i) : z4, yb = HAS_X3 ? __ldg(x3 + base + j) : z4;
        accum(da, oa, xa, ya);
        accum(db, ob, xb, yb);
    }
    for (; i < n4; i += stride, vox += vstep)
        accum(load_d(i, vox), HAS_OUT ? __ldg(out + base + i) : z4, __ldg(x + base + i), HAS_X3 ? __ldg(x3 + base + i) : z4);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        atomicAdd(&sacc[c + e], s0[e]);
        atomicAdd(&sacc[C + c + e], s1[e]);
        if (HAS_X3) atomicAdd(&sacc[2 * C + c + e], s2[e]);
        if (WG) {
#pragma unroll
            for (int k = 0; k < 4; k++) atomicAdd(&sacc[(3 + k) * C + c + e], wg[k][e]);
            if (c4 == 0) atomicAdd(&sacc[7 * C + e], qs[e]);
        }
    }
    if (amax) {     // non-negative floats order like their bit patterns
        gmax = warp_max(gmax);
        if ((threadIdx.x & 31) == 0 && gmax > 0.f && gmax < 3.0e38f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(gmax));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
        double* o = sums + ((long long)b * C + j) * 3;
        atomicAdd(o, (double)sacc[j]);
        atomicAdd(o + 1, (double)sacc[C + j]);
        if (HAS_X3) atomicAdd(o + 2, (double)sacc[2 * C + j]);
    }
    if (WG) {
        for (int j = threadIdx.x; j < 4 * C; j += blockDim.x) atomicAdd(dw_out + j, sacc[3 * C + j]);
        if (threadIdx.x < 4) atomicAdd(db_out + threadIdx.x, sacc[7 * C + threadIdx.x]);
    }
}

// grid.x for the float4 kernels: ~8 CTAs per SM, a multiple of C/4 (so that the stride is), not more than the work
static int v4_grid(long long n4, int C4) {
    long long want = min((long long)148 * 8, (n4 + 255) / 256);
    long long g = ((want + C4 - 1) / C4) * C4;
    return (int)max((long long)C4, g);
}

// ------------------------------------------------------------------------------------------------ launchers
// rows per CTA of a column reduction: enough CTAs for ~8 per SM, but at least 64 rows (8 per thread row)
static int colred_rows_per_cta(int rows, int C, int batch) {
    int col_tiles = cdiv(C, 32);
    int want = max(1, (148 * 8) / max(1, col_tiles * batch));
    return max(64, cdiv(rows, want));
}
static RowSrc make_src(const float* x, int C, const int* merge_dims) {
    RowSrc s;
    s.x = x; s.C = C; s.merge = merge_dims ? 1 : 0;
    s.H = s.W = s.D = s.Cin = 0;
    if (merge_dims) { s.H = merge_dims[0]; s.W = merge_dims[1]; s.D = merge_dims[2]; s.Cin = C / 8; }
    return s;
}

// the float4 kernels need C % 4 == 0 (merge rows: Cin % 4 == 0), at most 24 float4 per lane, and 16-byte aligned operands
static bool ln_vec_ok(const RowSrc& s, const void* p0, const void* p1, const void* p2, const void* p3, const void* p4, const void* p5) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    return s.C % 4 == 0 && s.C <= 3072 && (!s.merge || s.Cin % 4 == 0) && al(p0) && al(p1) && al(p2) && al(p3) && al(p4) && al(p5);
}

int k_layernorm_fwd(const float* x, const int* merge_dims, int rows, int C, const float* w, const float* b, float eps,
                    const float* pos, int pos_rows, const uint8_t* mask, const float* mask_token, float* y, float* mean,
                    float* rstd, cudaStream_t st) {
    if (rows == 0) return NMAE_OK;
    if (pos_rows <= 0) pos_rows = 1;
    const RowSrc s = make_src(x, C, merge_dims);
    if (ln_vec_ok(s, x, w, b, y, pos, mask_token)) {
#define LN_FWD(VPT, RPW)                                                                                                           \
        ln_fwd_v_kernel<VPT, RPW><<<cdiv(rows, 8 * RPW), 256, 0, st>>>(s, rows, w, b, eps, pos, pos_rows, mask, mask_token, y, mean, rstd)
        const int vpt = cdiv(C / 4, 32);
        if (vpt == 1) LN_FWD(1, 4);
        else if (vpt == 2) LN_FWD(2, 2);
        else if (vpt <= 3) LN_FWD(3, 2);
        else if (vpt <= 6) LN_FWD(6, 1);
        else if (vpt <= 12) LN_FWD(12, 1);
        else LN_FWD(24, 1);
#undef LN_FWD
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    ln_fwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(s, rows, w, b, eps, pos, pos_rows, mask, mask_token, y, mean, rstd);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_layernorm_bwd(const float* x, const int* merge_dims, int rows, int C, const float* w, const float* dy, const float* mean,
                    const float* rstd, const uint8_t* mask, int pos_rows, float* dx, const float* add_src, float* dgamma, float* dbeta,
                    cudaStream_t st) {
    if (rows == 0) return NMAE_OK;
    if (pos_rows <= 0) pos_rows = 1;
    RowSrc s = make_src(x, C, merge_dims);
    if (ln_vec_ok(s, x, w, dy, dx, add_src, nullptr) && C <= 1536) {      // C = 3072 (500 merged rows): 5 x 24 float4 per lane would spill
#define LN_BWD(VPT, RPW)                                                                                                           \
        ln_bwd_v_kernel<VPT, RPW><<<min(cdiv(rows, 8 * RPW), 148 * 8), 256, 2 * C * sizeof(float), st>>>(                             \
            s, rows, w, dy, mean, rstd, mask, pos_rows, dx, add_src, dgamma, dbeta)
        const int vpt = cdiv(C / 4, 32);
        if (vpt == 1) LN_BWD(1, 4);
        else if (vpt == 2) LN_BWD(2, 2);
        else if (vpt <= 3) LN_BWD(3, 2);
        else if (vpt <= 6) LN_BWD(6, 1);
        else LN_BWD(12, 1);
#undef LN_BWD
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    if (dx) {
        ln_bwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(s, rows, w, dy, mean, rstd, mask, pos_rows, dx, add_src);
        NMAE_LAUNCH_CHECK();
    }
    if (dgamma) {
        int rpc = colred_rows_per_cta(rows, C, 1);
        dim3 grid(cdiv(C, 32), cdiv(rows, rpc));
        ln_param_grad_kernel<<<grid, dim3(32, 8), 0, st>>>(s, rows, rpc, dy, mean, rstd, mask, pos_rows, dgamma, dbeta);
        NMAE_LAUNCH_CHECK();
    }
    return NMAE_OK;
}

int k_colsum(const float* x, int rows, int C, long long ld, const uint8_t* mask, int pos_rows, float* out, cudaStream_t st) {
    if (rows == 0) return NMAE_OK;
    if (pos_rows <= 0) pos_rows = 1;
    if (!mask && ld == C && C % 4 == 0 && C <= 96 && (long long)rows * C >= (1 << 22) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // large dense narrow tensors (bias gradients of the decoder's transposed convolutions): float4 column sums (out is pre-zeroed);
        // wide tensors keep the column kernel: every CTA of this one ends with C atomics on the same C addresses
        in_stats_v4_kernel<false><<<dim3(v4_grid((long long)rows * C / 4, C / 4), 1), 256, 2 * C * sizeof(float), st>>>(
            reinterpret_cast<const float4*>(x), rows, C, nullptr, out);
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    int rpc = colred_rows_per_cta(rows, C, 1);
    dim3 grid(cdiv(C, 32), cdiv(rows, rpc));
    colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(x, rows, C, ld, rpc, mask, pos_rows, out);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

#define TRY_RET(x) do { int rc_ = (x); if (rc_ != NMAE_OK) return rc_; } while (0)
static int stat_rows_per_cta(int V, int C, int B) { return colred_rows_per_cta(V, C, B); }

int k_in_stats(const float* x, int B, int V, int C, double* stats, cudaStream_t st) {
    NMAE_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * C, st));
    if (C % 4 == 0 && (long long)V * C >= (1 << 16)) {
        in_stats_v4_kernel<true><<<dim3(v4_grid((long long)V * C / 4, C / 4), B), 256, 2 * C * sizeof(float), st>>>(
            reinterpret_cast<const float4*>(x), V, C, stats, nullptr);
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    int rpc = stat_rows_per_cta(V, C, B);
    dim3 grid(cdiv(C, 32), cdiv(V, rpc), B);
    in_stats_kernel<<<grid, dim3(32, 8), 0, st>>>(x, V, C, rpc, stats);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_in_act_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C, float eps,
                 float slope, float* out, cudaStream_t st) {
    long long n = (long long)V * C;
    if (C % 4 == 0) {
        in_act_fwd_v4_kernel<<<dim3(v4_grid(n / 4, C / 4), B), 256, 0, st>>>(
            reinterpret_cast<const float4*>(x), stats, reinterpret_cast<const float4*>(res), res_stats, V, C, eps, slope,
            reinterpret_cast<float4*>(out));
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    int gx = (int)min((long long)148 * 8, (n + 255) / 256);
    in_act_fwd_kernel<<<dim3(gx, B), 256, 4 * C * sizeof(float), st>>>(x, stats, res, res_stats, V, C, eps, slope, out);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// out = lrelu(IN(x) + R) and pred = W_out . out + b_out (W_out (4, C), C % 4 == 0, C <= 128)
int k_in_act_fwd_out(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C, float eps,
                     float slope, float* out, const float* w_out, const float* b_out, float* pred, cudaStream_t st) {
    NMAE_CHECK_ARG(C % 4 == 0 && C >= 4 && C <= 128, "in_lrelu_apply_out_fwd: channels must be a multiple of 4, <= 128 (C=%d)", C);
    const size_t smem = sizeof(float) * (8 * C + 256 * (C + 1));
    static bool attr_set[64] = {false};
    int dev;
    NMAE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        NMAE_CUDA(cudaFuncSetAttribute(in_act_fwd_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (8 * 128 + 256 * 129))));
        attr_set[dev] = true;
    }
    const int gx = (int)min((long long)148 * 4 / max(1, min(B, 4)) * 2, ((long long)V + 255) / 256);
    in_act_fwd_out_kernel<<<dim3(max(1, gx), B), 256, smem, st>>>(reinterpret_cast<const float4*>(x), stats, reinterpret_cast<const float4*>(res),
                                                                 res_stats, V, C, eps, slope, reinterpret_cast<float4*>(out), w_out, b_out,
                                                                 reinterpret_cast<float4*>(pred));
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

// sums[b][c] = {sum g, sum g*xhat, sum g*xhat3}: the reduction pass of the InstanceNorm+LeakyReLU backward
int k_in_bwd_sums(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                  int B, int V, int C, float eps, float slope, double* sums, cudaStream_t st, float* amax, const float* dp4,
                  const float* w4, float* dw_out, float* db_out) {
    NMAE_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * B * C, st));
    if (amax) NMAE_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), st));
    NMAE_CHECK_ARG(dw_out == nullptr || (dp4 && out && !x3 && db_out && C % 4 == 0 && (long long)V * C >= (1 << 16)),
                   "in_bwd_sums: the fused output-convolution weight gradient needs dpred4, out, no shortcut branch, C %% 4 == 0");
    if (dw_out) {
        NMAE_CUDA(cudaMemsetAsync(dw_out, 0, sizeof(float) * 4 * C, st));
        NMAE_CUDA(cudaMemsetAsync(db_out, 0, sizeof(float) * 4, st));
        // two resident CTAs per SM (124 registers) and exactly two waves; every CTA ends with 4*C float atomics on the same few cache
        // lines, so the grid is kept small (the 8-CTAs-per-SM grid of the other variants would issue 8x as many)
        const int C4 = C / 4;
        const int gx = max(C4, (2 * 2 * 148 / max(1, B)) / C4 * C4);
        const dim3 grid(gx, B);
        in_bwd_sums_v4_kernel<true, true, false, true><<<grid, 256, (7 * C + 4) * sizeof(float), st>>>(
            nullptr, reinterpret_cast<const float4*>(out), reinterpret_cast<const float4*>(x), stats, nullptr, nullptr, V, C, eps, slope,
            sums, amax, reinterpret_cast<const float4*>(dp4), w4, dw_out, db_out);
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    if (C % 4 == 0 && (long long)V * C >= (1 << 16)) {
        const dim3 grid(v4_grid((long long)V * C / 4, C / 4), B);
        const size_t sm = 3 * C * sizeof(float);
#define SUMS_V4(D, O, X3)                                                                                                          \
        in_bwd_sums_v4_kernel<D, O, X3><<<grid, 256, sm, st>>>(reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(out), \
            reinterpret_cast<const float4*>(x), stats, reinterpret_cast<const float4*>(x3), stats3, V, C, eps, slope, sums, amax,       \
            reinterpret_cast<const float4*>(dp4), w4)
        const int key = (dp4 ? 4 : 0) | (out ? 2 : 0) | (x3 ? 1 : 0);
        switch (key) {
            case 0: SUMS_V4(false, false, false); break;
            case 1: SUMS_V4(false, false, true); break;
            case 2: SUMS_V4(false, true, false); break;
            case 3: SUMS_V4(false, true, true); break;
            case 4: SUMS_V4(true, false, false); break;
            case 5: SUMS_V4(true, false, true); break;
            case 6: SUMS_V4(true, true, false); break;
            default: SUMS_V4(true, true, true); break;
        }
#undef SUMS_V4
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    int rpc = stat_rows_per_cta(V, C, B);
    dim3 grid(cdiv(C, 32), cdiv(V, rpc), B);
    if (dp4)
        in_bwd_sums_kernel<true><<<grid, dim3(32, 8), 0, st>>>(dout, out, x, stats, x3, stats3, V, C, rpc, eps, slope, sums, amax,
                                                               reinterpret_cast<const float4*>(dp4), w4);
    else
        in_bwd_sums_kernel<false><<<grid, dim3(32, 8), 0, st>>>(dout, out, x, stats, x3, stats3, V, C, rpc, eps, slope, sums, amax,
                                                                nullptr, nullptr);
    NMAE_LAUNCH_CHECK();
    return NMAE_OK;
}

int k_in_act_bwd(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                 int B, int V, int C, float eps, float slope, double* sums, float* dx, float* dx3, float* dres, float* dbias,
                 float* dbias3, cudaStream_t st) {
    TRY_RET(k_in_bwd_sums(dout, out, x, stats, x3, stats3, B, V, C, eps, slope, sums, st));
    if (dbias) NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, st));
    if (dbias3) NMAE_CUDA(cudaMemsetAsync(dbias3, 0, sizeof(float) * C, st));
    long long n = (long long)V * C;
    if (C % 4 == 0) {
        in_bwd_apply_v4_kernel<<<dim3(v4_grid(n / 4, C / 4), B), 256, 2 * C * sizeof(float), st>>>(
            reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(out), reinterpret_cast<const float4*>(x), stats,
            reinterpret_cast<const float4*>(x3), stats3, sums, V, C, eps, slope, reinterpret_cast<float4*>(dx),
            reinterpret_cast<float4*>(dx3), reinterpret_cast<float4*>(dres), dbias, dbias3);
        NMAE_LAUNCH_CHECK();
        return NMAE_OK;
    }
    int gx = (int)min((long long)148 * 8, (n + 255) / 256);
    in_bwd_apply_kernel<<<dim3(gx, B), 256, 7 * C * sizeof(float), st>>>(dout, out, x, stats, x3, stats3, sums, V, C, eps, slope, dx,
                                                                         dx3, dres);
    NMAE_LAUNCH_CHECK();
    // generic channel counts: bias gradients as separate column sums
    if (dbias) TRY_RET(k_colsum(dx, B * V, C, C, nullptr, 1, dbias, st));
    if (dbias3) TRY_RET(k_colsum(dx3, B * V, C, C, nullptr, 1, dbias3, st));
    return NMAE_OK;
}
