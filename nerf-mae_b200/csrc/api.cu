// C ABI of libnmae.so (see include/nmae.h).  Every function only sequences kernel launches on the
// caller's stream; the caller owns all memory.
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/nmae.h"
#include "kernels.cuh"

static thread_local char g_err[512] = "";

void nmae_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

unsigned long long g_nmae_launches = 0;

int nmae_debug_mask(void) {
#ifdef NMAE_DBG
    const char* d = getenv("NMAE_DBG");
    return d ? atoi(d) : 0;
#else
    return 0;
#endif
}

extern "C" const char* nmae_last_error(void) { return g_err; }
extern "C" unsigned long long nmae_launch_count(void) { return g_nmae_launches; }
extern "C" int nmae_version(void) { return 100; }

#define ST(stream) reinterpret_cast<cudaStream_t>(stream)
#define TRY(x)                      \
    do {                            \
        int rc_ = (x);              \
        if (rc_ != NMAE_OK) return rc_; \
    } while (0)

// ------------------------------------------------------------------------------------------------ helpers
static GOperand op_strided(const float* p, long long s_a, long long s_b) {
    GOperand o;
    memset(&o, 0, sizeof(o));
    o.p = p; o.mode = OPM_STRIDED; o.s_a = s_a; o.s_b = s_b;
    return o;
}
static GOperand op_gather(int mode, const float* p, int X, int Y, int Z, int C, int ld, int ks, int swap) {
    GOperand o;
    memset(&o, 0, sizeof(o));
    o.p = p; o.mode = mode; o.swap = swap; o.X = X; o.Y = Y; o.Z = Z; o.C = C; o.ld = ld; o.ks = ks;
    return o;
}
static GEpilogue epi_plain(float* out, long long ldc, int flags = 0) {
    GEpilogue e;
    memset(&e, 0, sizeof(e));
    e.out = out; e.ldc = ldc; e.flags = flags; e.rows_per_scale = 1; e.ks = 1;
    return e;
}
static int pick_ksplit(int M, int N, int K) {
    long long tiles = (long long)cdiv(M, 128) * cdiv(N, 64);
    long long want = (592 + tiles - 1) / tiles;
    long long maxs = (K + 255) / 256;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    long long ks = (K + want - 1) / want;
    ks = (ks + 15) / 16 * 16;
    return (int)ks;
}
static int gemm(const GOperand& A, const GOperand& B, const GEpilogue& E, int M, int N, int K, bool split, cudaStream_t st) {
    GemmParams p;
    p.A = A; p.B = B; p.E = E; p.M = M; p.N = N; p.K = K;
    p.ksplit = split ? pick_ksplit(M, N, K) : K;
    if (split) p.E.flags |= EPI_ATOMIC;
    return nmae_gemm_launch(p, st);
}

extern "C" {

int nmae_pad_grid(const float* grid, int X, int Y, int Z, float* batch, int b, int R, int device, void* stream) {
    // an extent larger than R is cropped at the high end, like the negative pads pad_tensor hands to F.pad (T:76-88)
    NMAE_CHECK_ARG(X > 0 && Y > 0 && Z > 0 && R > 0, "pad_grid: empty extent (%d,%d,%d) / resolution %d", X, Y, Z, R);
    NMAE_SET_DEVICE(device);
    return k_pad_grid(grid, 4, X, Y, Z, batch + (long long)b * 4 * R * R * R, R, ST(stream));
}

int nmae_ingest_scene(const void* rgbsigma, int is_uint8, int normalize_density, int W, int L, int H, int rotate, int flip_axis1,
                      int flip_axis2, float* batch, int b, int R, int device, void* stream) {
    const int X = rotate ? L : W, Y = rotate ? W : L;
    NMAE_CHECK_ARG(W > 0 && L > 0 && H > 0 && R > 0, "ingest_scene: empty extent (%d,%d,%d) / resolution %d", X, Y, H, R);   // oversize: cropped
    NMAE_SET_DEVICE(device);
    return k_ingest_scene(rgbsigma, is_uint8, normalize_density, W, L, H, rotate, flip_axis1, flip_axis2,
                          batch + (long long)b * 4 * R * R * R, R, ST(stream));
}

int nmae_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                         const float* pos, const uint8_t* mask, const float* mask_token, int B, int R, int p, int C,
                         float eps, float* conv, float* mean, float* rstd, float* tokens, float* w_ws, int device, void* stream) {
    NMAE_CHECK_ARG(R % p == 0, "patch_embed: resolution %d not divisible by patch %d", R, p);  // S:1390
    NMAE_SET_DEVICE(device);
    int n = R / p, T = n * n * n, K = 4 * p * p * p;
    GEpilogue e = epi_plain(conv, C, EPI_BIAS);
    e.bias = bias;
    const GOperand patches = op_gather(OPM_PATCH, x, n, n, n, 4, 0, p, 0);
    if (w_ws && p == 4 && k_lin_tc_supported(B * T, C, K, 4, C))     // tcgen05: 4x4x4 patches gathered by the A producers
        TRY(k_lin_tc(x, 0, w, K, 1, B * T, C, K, e, w_ws, ST(stream), 0, &patches));
    else
        TRY(gemm(patches, op_strided(w, K, 1), e, B * T, C, K, false, ST(stream)));
    return k_layernorm_fwd(conv, nullptr, B * T, C, ln_w, ln_b, eps, pos, T, mask, mask_token, tokens, mean, rstd, ST(stream));
}

int nmae_patch_embed_bwd(const float* dtokens, const float* x, const float* w, const float* ln_w, const float* conv,
                         const float* mean, const float* rstd, const uint8_t* mask, int B, int R, int p, int C,
                         float* dconv_ws, float* dw, float* dbias, float* dln_w, float* dln_b, float* dmask_token,
                         int device, void* stream) {
    (void)w;
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    int n = R / p, T = n * n * n, K = 4 * p * p * p;
    NMAE_CUDA(cudaMemsetAsync(dln_w, 0, sizeof(float) * C, st));
    NMAE_CUDA(cudaMemsetAsync(dln_b, 0, sizeof(float) * C, st));
    NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, st));
    NMAE_CUDA(cudaMemsetAsync(dmask_token, 0, sizeof(float) * C, st));
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C * K, st));
    TRY(k_layernorm_bwd(conv, nullptr, B * T, C, ln_w, dtokens, mean, rstd, mask, T, dconv_ws, nullptr, dln_w, dln_b, st));
    if (mask) TRY(k_colsum(dtokens, B * T, C, C, mask, T, dmask_token, st));
    TRY(k_colsum(dconv_ws, B * T, C, C, nullptr, 1, dbias, st));
    if (p == 4 && k_lin_wgrad_tc_supported(B * T, C, K, 4, C)) {     // tcgen05: dW[c][k] = sum_tokens dconv[token][c] * patch[token][k]
        const GOperand patches = op_gather(OPM_PATCH, x, n, n, n, 4, 0, p, 0);
        return k_lin_wgrad_tc(x, 0, dconv_ws, C, B * T, C, K, dw, st, &patches);
    }
    return gemm(op_strided(dconv_ws, 1, C), op_gather(OPM_PATCH, x, n, n, n, 4, 0, p, 1), epi_plain(dw, K), C, K, B * T, true, st);
}

int nmae_layernorm_fwd(const float* x, const float* w, const float* b, int rows, int C, float eps, float* y, float* mean,
                       float* rstd, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_layernorm_fwd(x, nullptr, rows, C, w, b, eps, nullptr, 1, nullptr, nullptr, y, mean, rstd, ST(stream));
}

int nmae_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, int rows, int C,
                       const float* dresid, float* dx, float* dw, float* db, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    if (dw) {
        NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C, st));
        NMAE_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
    }
    return k_layernorm_bwd(x, nullptr, rows, C, w, dy, mean, rstd, nullptr, 1, dx, dresid, dw, db, st);
}

int nmae_linear_fwd(const float* x, const float* w, const float* bias, int M, int N, int K, int flags, float* aux,
                    const float* resid, const float* row_scale, int rows_per_scale, float* out, float* w_ws, int device,
                    void* stream) {
    NMAE_SET_DEVICE(device);
    GEpilogue e = epi_plain(out, N);
    if (bias) { e.flags |= EPI_BIAS; e.bias = bias; }
    if (flags & 1) {
        NMAE_CHECK_ARG(aux != nullptr, "linear_fwd: GELU needs an aux buffer");
        e.flags |= EPI_GELU; e.aux = aux;
    }
    if (flags & 2) {
        NMAE_CHECK_ARG(resid != nullptr, "linear_fwd: residual flag without residual");
        e.flags |= EPI_RESID; e.resid = resid; e.row_scale = row_scale; e.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
    }
    if (k_thin_supported(N, K) && !(flags & 3)) return k_thin_fwd(x, w, bias, M, K, out, ST(stream));
    if (w_ws && k_lin_tc_supported(M, N, K, K, N))
        return k_lin_tc(x, K, w, K, 1, M, N, K, e, w_ws, ST(stream), 0, nullptr, (flags & 256) != 0);
    return gemm(op_strided(x, K, 1), op_strided(w, K, 1), e, M, N, K, false, ST(stream));
}

int nmae_linear_bwd_input(const float* dy, const float* w, int M, int N, int K, int flags, const float* aux, float* dx,
                          float* w_ws, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    GEpilogue e = epi_plain(dx, K);
    if (flags & 1) { e.flags |= EPI_GELU_GRAD; e.aux = const_cast<float*>(aux); }
    if (flags & 4) e.flags |= EPI_ACCUM;
    if (k_thin_supported(N, K) && !(flags & 1)) return k_thin_dgrad(dy, w, M, K, (flags & 4) ? 1 : 0, dx, ST(stream));
    if (w_ws && k_lin_tc_supported(M, K, N, N, K))
        return k_lin_tc(dy, N, w, 1, K, M, K, N, e, w_ws, ST(stream), 0, nullptr, (flags & 256) != 0);
    return gemm(op_strided(dy, N, 1), op_strided(w, 1, K), e, M, K, N, false, ST(stream));
}

int nmae_linear_blob_layout(int M, int N, int K, int* tile_n, int* k_group, int device) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(tile_n != nullptr && k_group != nullptr, "linear_blob_layout: null output");
    if (!k_lin_tc_supported(M, N, K, K, N)) { *tile_n = 0; *k_group = 0; return NMAE_OK; }
    return k_lin_tc_tile(M, N, K, tile_n, k_group);
}

int nmae_linear_prep_batch(const long long* table, int n, long long max_elems, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_lin_tc_prep_batch(table, n, max_elems, ST(stream));
}

int nmae_linear_bwd_weight(const float* dy, const float* x, int M, int N, int K, float* dw, float* db, int device,
                           void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    if (k_thin_supported(N, K)) return k_thin_wgrad(x, dy, M, K, dw, db, st);
    if (db) {
        NMAE_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
        TRY(k_colsum(dy, M, N, N, nullptr, 1, db, st));
    }
    if (k_lin_wgrad_tc_supported(M, N, K, K, N)) return k_lin_wgrad_tc(x, K, dy, N, M, N, K, dw, st);
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    return gemm(op_strided(dy, 1, N), op_strided(x, 1, K), epi_plain(dw, K), N, K, M, true, st);
}

int nmae_window_attention_num_windows(int H, int W, int D) { return k_wattn_num_windows(H, W, D); }

// ---- workspace sizes (bytes) of the caller-provided scratch buffers (the library never allocates)
long long nmae_linear_weight_ws_bytes(int N, int K) { return 4LL * N * K; }
long long nmae_patch_embed_weight_ws_bytes(int C, int p) { return 4LL * C * 4 * p * p * p; }
long long nmae_patch_embed_bwd_ws_bytes(int B, int R, int p, int C) { return 4LL * B * (R / p) * (R / p) * (R / p) * C; }
long long nmae_patch_merge_weight_ws_bytes(int C) { return 4LL * 16 * C * C; }
long long nmae_patch_merge_bwd_ws_bytes(int B, int H, int W, int D, int C) {
    return 4LL * B * ((H + 1) / 2) * ((W + 1) / 2) * ((D + 1) / 2) * 8 * C;
}
long long nmae_convT_weight_ws_bytes(int Cin, int Cout, int k) { return 4LL * Cin * Cout * k * k * k; }
long long nmae_conv3x3x3_weight_ws_bytes(int Cin, int Cout) { return 4LL * 27 * Cin * Cout; }
long long nmae_window_attention_lse_bytes(int B, int H, int W, int D, int num_heads) {
    return 4LL * B * k_wattn_num_windows(H, W, D) * num_heads * 64;
}
long long nmae_instnorm_stats_bytes(int B, int C) { return 8LL * 2 * B * C; }
// 3*B*C double sums, then (fp16-image variant) 8*B*C float constants + the image scale
long long nmae_in_lrelu_bwd_sums_ws_bytes(int B, int C) { return 8LL * 3 * B * C + 4LL * 8 * B * C + 16; }

int nmae_window_attention_fwd(const float* qkv, const float* table, int B, int H, int W, int D, int C, int num_heads,
                              int shift, float* out, float* lse, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_wattn_tc_fwd(qkv, table, B, H, W, D, C, num_heads, shift, out, lse, ST(stream));
}

int nmae_window_attention_bwd(const float* dout, const float* qkv, const float* table, const float* out, const float* lse,
                              int B, int H, int W, int D, int C, int num_heads, int shift, float* dqkv, float* dtable,
                              int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CUDA(cudaMemsetAsync(dtable, 0, sizeof(float) * 343 * num_heads, ST(stream)));
    return k_wattn_tc_bwd(qkv, table, out, dout, lse, B, H, W, D, C, num_heads, shift, dqkv, dtable, ST(stream));
}

int nmae_patch_merge_fwd(const float* x, const float* ln_w, const float* ln_b, const float* red_w, int B, int H, int W,
                         int D, int C, float eps, float* normed, float* mean, float* rstd, float* out, float* w_ws,
                         int device, void* stream) {
    NMAE_SET_DEVICE(device);
    int dims[3] = {H, W, D};
    int rows = B * ((H + 1) / 2) * ((W + 1) / 2) * ((D + 1) / 2);
    TRY(k_layernorm_fwd(x, dims, rows, 8 * C, ln_w, ln_b, eps, nullptr, 1, nullptr, nullptr, normed, mean, rstd, ST(stream)));
    if (w_ws && k_lin_tc_supported(rows, 2 * C, 8 * C, 8 * C, 2 * C))
        return k_lin_tc(normed, 8 * C, red_w, 8 * C, 1, rows, 2 * C, 8 * C, epi_plain(out, 2 * C), w_ws, ST(stream));
    return gemm(op_strided(normed, 8 * C, 1), op_strided(red_w, 8 * C, 1), epi_plain(out, 2 * C), rows, 2 * C, 8 * C, false,
                ST(stream));
}

int nmae_patch_merge_bwd(const float* dout, const float* x, const float* ln_w, const float* red_w, const float* normed,
                         const float* mean, const float* rstd, int B, int H, int W, int D, int C, float* dnormed_ws,
                         float* dx, float* dln_w, float* dln_b, float* dred_w, float* w_ws, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    int dims[3] = {H, W, D};
    int rows = B * ((H + 1) / 2) * ((W + 1) / 2) * ((D + 1) / 2);
    int N = 2 * C, K = 8 * C;
    if (w_ws && k_lin_tc_supported(rows, K, N, N, K))
        TRY(k_lin_tc(dout, N, red_w, 1, K, rows, K, N, epi_plain(dnormed_ws, K), w_ws, st));
    else
        TRY(gemm(op_strided(dout, N, 1), op_strided(red_w, 1, K), epi_plain(dnormed_ws, K), rows, K, N, false, st));
    NMAE_CUDA(cudaMemsetAsync(dln_w, 0, sizeof(float) * K, st));
    NMAE_CUDA(cudaMemsetAsync(dln_b, 0, sizeof(float) * K, st));
    TRY(k_layernorm_bwd(x, dims, rows, K, ln_w, dnormed_ws, mean, rstd, nullptr, 1, dx, nullptr, dln_w, dln_b, st));
    if (k_lin_wgrad_tc_supported(rows, N, K, K, N)) return k_lin_wgrad_tc(normed, K, dout, N, rows, N, K, dred_w, st);
    NMAE_CUDA(cudaMemsetAsync(dred_w, 0, sizeof(float) * (size_t)N * K, st));
    return gemm(op_strided(dout, 1, N), op_strided(normed, 1, K), epi_plain(dred_w, K), N, K, rows, true, st);
}

// tensor-core transposed convolution: Cin is the forward K (groups of 48 or 32), Cout the k-group unit of the gathered backward
static bool convT_tc_ok(int Cin, int Cout, int ld_out) {
    return (Cin % 48 == 0 || Cin % 32 == 0) && (Cout % 48 == 0 || Cout % 32 == 0) && ld_out % 4 == 0;
}

int nmae_convT_k_eq_s_fwd(const float* x, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin, int Cout,
                          int k, float* out, int ld_out, float* w_ws, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(ld_out >= Cout, "convT: ld_out %d < Cout %d", ld_out, Cout);
    int N = Cout * k * k * k;
    GEpilogue e = epi_plain(out, 0, EPI_D2S);
    e.X = X; e.Y = Y; e.Z = Z; e.C = Cout; e.ld = ld_out; e.ks = k;
    if (bias) { e.flags |= EPI_BIAS; e.bias = bias; }
    // tensor-core path: GEMM columns ordered (i,j,l, co) so that 16 consecutive columns are 16 channels of one fine voxel
    if (w_ws && convT_tc_ok(Cin, Cout, ld_out)) return k_lin_tc(x, Cin, w, 0, 0, B * X * Y * Z, N, Cin, e, w_ws, ST(stream), 1);
    return gemm(op_strided(x, Cin, 1), op_strided(w, 1, N), e, B * X * Y * Z, N, Cin, false, ST(stream));
}

int nmae_convT_k_eq_s_bwd(const float* dout, int ld_out, const float* x, const float* w, int B, int X, int Y, int Z, int Cin,
                          int Cout, int k, float* dx, float* dw, float* dbias, float* w_ws, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    int N = Cout * k * k * k, M = B * X * Y * Z;
    if (w_ws && convT_tc_ok(Cin, Cout, ld_out)) {
        GOperand g = op_gather(OPM_D2S, dout, X, Y, Z, Cout, ld_out, k, 0);
        if (dx) TRY(k_lin_tc(dout, 0, w, 0, 0, M, Cin, N, epi_plain(dx, Cin), w_ws, st, 2, &g));
        // dW[ci][co][ijl] = sum_m x[m][ci] * dout_gathered[m][(ijl,co)]
        TRY(k_lin_wgrad_tc(dout, 0, x, Cin, M, Cin, N, dw, st, &g));
        if (dbias) {
            NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, st));
            TRY(k_colsum(dout, M * k * k * k, Cout, ld_out, nullptr, 1, dbias, st));
        }
        return NMAE_OK;
    }
    if (dx) TRY(gemm(op_gather(OPM_D2S, dout, X, Y, Z, Cout, ld_out, k, 0), op_strided(w, N, 1), epi_plain(dx, Cin), M, Cin, N, false, st));
    NMAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cin * N, st));
    TRY(gemm(op_strided(x, 1, Cin), op_gather(OPM_D2S, dout, X, Y, Z, Cout, ld_out, k, 1), epi_plain(dw, N), Cin, N, M, true, st));
    if (dbias) {
        NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, st));
        TRY(k_colsum(dout, M * k * k * k, Cout, ld_out, nullptr, 1, dbias, st));
    }
    return NMAE_OK;
}

long long nmae_conv3_image_bytes(int B, int X, int Y, int Z, int C) {
    if (C % UIMG_CG != 0) return 0;
    return uimg_geom(B, X, Y, Z, C).total_bytes;
}

int nmae_conv3_image_build(const float* x, int ld, int ch_off, int B, int X, int Y, int Z, int C, int type_dy, void* image,
                           int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_uimg_build(x, ld, ch_off, uimg_geom(B, X, Y, Z, C), type_dy, nullptr, 0.f, 0.f, image, ST(stream));
}

int nmae_conv3_image_build_in_lrelu(const float* x, const double* stats, int B, int X, int Y, int Z, int C, float eps, float slope,
                                    void* image, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(stats != nullptr, "conv3_image_build_in_lrelu: statistics required");
    return k_uimg_build(x, C, 0, uimg_geom(B, X, Y, Z, C), 0, stats, eps, slope, image, ST(stream));
}

int nmae_conv3x3x3_fwd(const float* x, const void* x_image, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin,
                       int Cout, float* w_ws, float* out, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    if (x_image && k_conv3_tc_supported(Cin, Cout)) return k_conv3_tc(x_image, w, bias, B, X, Y, Z, Cin, Cout, 0, w_ws, out, 0, st);
    NMAE_CHECK_ARG(x != nullptr, "conv3x3x3_fwd: neither an image nor an fp32 volume given");
    // generic channel counts: CUDA-core implicit GEMM.  w (Cout,Cin,27) -> w_ws [Cout][27][Cin]
    TRY(k_gather3(w_ws, w, Cout, 27, Cin, (long long)Cin * 27, 1, 27, st));
    GEpilogue e = epi_plain(out, Cout);
    if (bias) { e.flags |= EPI_BIAS; e.bias = bias; }
    return gemm(op_gather(OPM_CONV3, x, X, Y, Z, Cin, Cin, 1, 0), op_strided(w_ws, 27LL * Cin, 1), e, B * X * Y * Z, Cout, 27 * Cin,
                false, st);
}

int nmae_conv3x3x3_dgrad(const float* dout, const void* dout_image, const float* w, int B, int X, int Y, int Z, int Cin, int Cout,
                         float* w_ws, float* dx, int accumulate, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    if (dout_image && k_conv3_tc_supported(Cout, Cin))
        return k_conv3_tc(dout_image, w, nullptr, B, X, Y, Z, Cout, Cin, 1, w_ws, dx, accumulate, st);
    NMAE_CHECK_ARG(dout != nullptr, "conv3x3x3_dgrad: neither an image nor an fp32 volume given");
    // w_ws [Cin][27 flipped][Cout] = w[co][ci][26 - tap]
    TRY(k_gather3(w_ws, w + 26, Cin, 27, Cout, 27, -1, (long long)Cin * 27, st));
    return gemm(op_gather(OPM_CONV3, dout, X, Y, Z, Cout, Cout, 1, 0), op_strided(w_ws, 27LL * Cout, 1),
                epi_plain(dx, Cin, accumulate ? EPI_ACCUM : 0), B * X * Y * Z, Cin, 27 * Cout, false, st);
}

int nmae_conv3x3x3_wgrad(const float* dout, const void* dout_image, const float* x, const void* x_image, int B, int X, int Y, int Z,
                         int Cin, int Cout, float* w_ws, float* dw, float* dbias, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    cudaStream_t st = ST(stream);
    int M = B * X * Y * Z;
    NMAE_CHECK_ARG(dout != nullptr || dbias == nullptr, "conv3x3x3_wgrad: the bias gradient needs the fp32 output gradient");
    if (x_image && dout_image && k_conv3_wgrad_tc_supported(Cin, Cout)) {
        TRY(k_conv3_wgrad_tc(x_image, dout_image, B, X, Y, Z, Cin, Cout, dw, st));
        if (dbias) {
            NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, st));
            TRY(k_colsum(dout, M, Cout, Cout, nullptr, 1, dbias, st));
        }
        return NMAE_OK;
    }
    NMAE_CHECK_ARG(x != nullptr && dout != nullptr, "conv3x3x3_wgrad: neither images nor fp32 volumes given");
    NMAE_CUDA(cudaMemsetAsync(w_ws, 0, sizeof(float) * 27 * (size_t)Cin * Cout, st));
    TRY(gemm(op_gather(OPM_CONV3, x, X, Y, Z, Cin, Cin, 1, 1), op_strided(dout, 1, Cout), epi_plain(w_ws, Cout), 27 * Cin, Cout, M,
             true, st));
    // w_ws [(tap,ci)][co] -> dw (Cout,Cin,27)
    TRY(k_gather3(dw, w_ws, Cout, Cin, 27, 1, Cout, (long long)Cin * Cout, st));
    if (dbias) {
        NMAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * Cout, st));
        TRY(k_colsum(dout, M, Cout, Cout, nullptr, 1, dbias, st));
    }
    return NMAE_OK;
}

int nmae_instnorm_stats(const float* x, int B, int V, int C, double* stats, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_in_stats(x, B, V, C, stats, ST(stream));
}

int nmae_in_lrelu_apply_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V,
                            int C, float eps, float slope, float* out, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_in_act_fwd(x, stats, res, res_stats, B, V, C, eps, slope, out, ST(stream));
}

int nmae_in_lrelu_apply_out_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C,
                                float eps, float slope, float* out, const float* w_out, const float* b_out, float* pred, int device,
                                void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(x && stats && out && w_out && pred, "in_lrelu_apply_out_fwd: x, stats, out, w_out and pred are required");
    return k_in_act_fwd_out(x, stats, res, res_stats, B, V, C, eps, slope, out, w_out, b_out, pred, ST(stream));
}

int nmae_in_lrelu_apply_bwd(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                            const double* stats3, int B, int V, int C, float eps, float slope, double* sums_ws, float* dx,
                            float* dx3, float* dres, float* dbias, float* dbias3, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG((x3 == nullptr) == (dx3 == nullptr), "in_lrelu_apply_bwd: x3 and dx3 must be given together");
    NMAE_CHECK_ARG(out != nullptr || (x3 == nullptr && dres == nullptr),
                   "in_lrelu_apply_bwd: out may only be omitted when the forward had no residual");
    NMAE_CHECK_ARG(dbias3 == nullptr || dx3 != nullptr, "in_lrelu_apply_bwd: dbias3 needs dx3");
    return k_in_act_bwd(dout, out, x, stats, x3, stats3, B, V, C, eps, slope, sums_ws, dx, dx3, dres, dbias, dbias3, ST(stream));
}

int nmae_in_lrelu_apply_bwd_image(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                                  const double* stats3, int B, int X, int Y, int Z, int C, float eps, float slope, double* sums_ws,
                                  void* dx_image, float* dx3, float* dres, float* dbias, float* dbias3, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG((x3 == nullptr) == (dx3 == nullptr), "in_lrelu_apply_bwd_image: x3 and dx3 must be given together");
    NMAE_CHECK_ARG(out != nullptr || (x3 == nullptr && dres == nullptr),
                   "in_lrelu_apply_bwd_image: out may only be omitted when the forward had no residual");
    NMAE_CHECK_ARG(dbias3 == nullptr || dx3 != nullptr, "in_lrelu_apply_bwd_image: dbias3 needs dx3");
    NMAE_CHECK_ARG(C % UIMG_CG == 0, "in_lrelu_apply_bwd_image: channels must be a multiple of 48 (C=%d)", C);
    TRY(k_in_bwd_sums(dout, out, x, stats, x3, stats3, B, X * Y * Z, C, eps, slope, sums_ws, ST(stream)));
    return k_in_act_bwd_image(dout, out, x, stats, x3, stats3, sums_ws, uimg_geom(B, X, Y, Z, C), eps, slope, dx_image, dx3, dres, dbias,
                              dbias3, ST(stream));
}

// ---------------------------------------------------------------------------------- single-pass fp16 convolution path
long long nmae_conv3h_image_bytes(int B, int X, int Y, int Z, int C) {
    if (uimg_h_cg(C) == 0) return 0;
    return uimg_geom_h(B, X, Y, Z, C).total_bytes;
}

long long nmae_conv3h_weight_ws_bytes(int Cin, int Cout) { return k_conv3_h_blob_bytes(Cin, Cout); }

int nmae_conv3h_image_build(const float* x, int ld, int ch_off, int B, int X, int Y, int Z, int C, const double* stats, float eps,
                            float slope, const float* scale, void* image, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_uimg_h_build(x, ld, ch_off, uimg_geom_h(B, X, Y, Z, C), stats, eps, slope, scale, image, ST(stream));
}

int nmae_conv3h_fwd(const void* x_image, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin, int Cout, void* w_ws,
                    float* out, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(x_image != nullptr && w_ws != nullptr, "conv3h_fwd: image and weight workspace required");
    return k_conv3_h(x_image, w, bias, nullptr, B, X, Y, Z, Cin, Cout, 0, w_ws, out, 0, ST(stream));
}

int nmae_conv3h_dgrad(const void* dout_image, const float* inv_scale, const float* w, int B, int X, int Y, int Z, int Cin, int Cout,
                      void* w_ws, float* dx, int accumulate, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(dout_image != nullptr && w_ws != nullptr, "conv3h_dgrad: image and weight workspace required");
    return k_conv3_h(dout_image, w, nullptr, inv_scale, B, X, Y, Z, Cout, Cin, 1, w_ws, dx, accumulate, ST(stream));
}

int nmae_conv3h_wgrad(const void* dout_image, const float* inv_scale, const void* x_image, int B, int X, int Y, int Z, int Cin, int Cout,
                      float* dw, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG(dout_image != nullptr && x_image != nullptr, "conv3h_wgrad: both images required");
    return k_conv3_wgrad_h(x_image, dout_image, inv_scale, B, X, Y, Z, Cin, Cout, dw, ST(stream));
}

int nmae_in_lrelu_apply_bwd_image_h(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                                    const double* stats3, int B, int X, int Y, int Z, int C, float eps, float slope, double* sums_ws,
                                    float* amax_ws, void* dx_image, float* inv_scale, float* dx3, float* dres, float* dbias,
                                    float* dbias3, const float* dpred4, const float* w_out, float* dw_out, float* db_out, int device,
                                    void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CHECK_ARG((dw_out == nullptr) == (db_out == nullptr), "in_lrelu_apply_bwd_image_h: dw_out and db_out must be given together");
    NMAE_CHECK_ARG(dw_out == nullptr || (dpred4 != nullptr && x3 == nullptr),
                   "in_lrelu_apply_bwd_image_h: dw_out needs dpred4 / w_out and a block without shortcut convolution");
    NMAE_CHECK_ARG((x3 == nullptr) == (dx3 == nullptr), "in_lrelu_apply_bwd_image_h: x3 and dx3 must be given together");
    NMAE_CHECK_ARG(out != nullptr || (x3 == nullptr && dres == nullptr),
                   "in_lrelu_apply_bwd_image_h: out may only be omitted when the forward had no residual");
    NMAE_CHECK_ARG(dbias3 == nullptr || dx3 != nullptr, "in_lrelu_apply_bwd_image_h: dbias3 needs dx3");
    NMAE_CHECK_ARG(uimg_h_cg(C) != 0, "in_lrelu_apply_bwd_image_h: channels must be a multiple of 48 or 64 (C=%d)", C);
    NMAE_CHECK_ARG(amax_ws != nullptr && inv_scale != nullptr, "in_lrelu_apply_bwd_image_h: amax workspace and inv_scale required");
    NMAE_CHECK_ARG((dpred4 == nullptr) == (w_out == nullptr), "in_lrelu_apply_bwd_image_h: dpred4 and w_out must be given together");
    NMAE_CHECK_ARG(dout != nullptr || dpred4 != nullptr, "in_lrelu_apply_bwd_image_h: neither dout nor (dpred4, w_out) given");
    TRY(k_in_bwd_sums(dout, out, x, stats, x3, stats3, B, X * Y * Z, C, eps, slope, sums_ws, ST(stream), amax_ws, dpred4, w_out,
                      dw_out, db_out));
    return k_in_act_bwd_image_h(dout, out, x, stats, x3, stats3, sums_ws, amax_ws, uimg_geom_h(B, X, Y, Z, C), eps, slope, dx_image,
                                inv_scale, dx3, dres, dbias, dbias3, ST(stream), dpred4, w_out);
}

int nmae_upsample_nearest_add(float* fine, const float* coarse, int B, int Xf, int Yf, int Zf, int Xc, int Yc, int Zc, int C,
                              int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_upsample_nearest_add(fine, coarse, B, Xf, Yf, Zf, Xc, Yc, Zc, C, ST(stream));
}

int nmae_upsample_trilinear_fwd(const float* coarse, float* fine, int B, int Xc, int Yc, int Zc, int Xf, int Yf, int Zf, int C, int device,
                                void* stream) {
    NMAE_SET_DEVICE(device);
    return k_upsample_trilinear(coarse, fine, B, Xc, Yc, Zc, Xf, Yf, Zf, C, 0, ST(stream));
}

int nmae_upsample_trilinear_bwd(const float* dfine, float* dcoarse, int B, int Xc, int Yc, int Zc, int Xf, int Yf, int Zf, int C,
                                int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_upsample_trilinear(dfine, dcoarse, B, Xc, Yc, Zc, Xf, Yf, Zf, C, 1, ST(stream));
}

int nmae_colsum(const float* x, long long rows, int C, long long ld, float* out, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    NMAE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, ST(stream)));
    return k_colsum(x, (int)rows, C, ld, nullptr, 1, out, ST(stream));
}

int nmae_scale_rows(float* dst, const float* src, const float* row_scale, int rows_per_scale, long long rows, int cols,
                    int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_scale_rows(dst, src, row_scale, rows_per_scale, rows, cols, ST(stream));
}

int nmae_copy_cols(float* dst, long long ld_dst, const float* src, long long ld_src, long long rows, int cols, int device,
                   void* stream) {
    NMAE_SET_DEVICE(device);
    return k_copy_cols(dst, ld_dst, src, ld_src, rows, cols, ST(stream));
}

int nmae_mae_loss_fwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p,
                      double* sums_ws, float* out3, int device, void* stream) {
    NMAE_CHECK_ARG(R % p == 0, "mae_loss: resolution %d not divisible by patch %d", R, p);
    NMAE_SET_DEVICE(device);
    return k_loss_fwd(x, pred, ext, tok_mask, B, R, p, sums_ws, out3, ST(stream));
}

int nmae_mae_loss_bwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p,
                      const double* sums_ws, const float* gout3, float* dpred, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_loss_bwd(x, pred, ext, tok_mask, B, R, p, sums_ws, gout3, dpred, ST(stream));
}

int nmae_multi_sumsq(const long long* table, int nchunks, double* norm_sq, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_multi_sumsq(table, nchunks, norm_sq, ST(stream));
}

int nmae_multi_copy(const long long* table, int nchunks, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_multi_copy(table, nchunks, ST(stream));
}

int nmae_adamw_clip_step(const long long* table, int nchunks, const double* norm_sq, float clip, float grad_scale, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                         float bias_correction2, int device, void* stream) {
    NMAE_SET_DEVICE(device);
    return k_adamw_clip(table, nchunks, norm_sq, clip, grad_scale, lr, beta1, beta2, eps, weight_decay, bias_correction1,
                        bias_correction2, ST(stream));
}

}  // extern "C"
