/* nmae.h - C ABI of the B200-native 3D Swin-MAE hot path (libnmae.so).
 *
 * The reference (zubair-irshad/NeRF-MAE @ 721b5ee) has no FFI on this path: every op is a PyTorch
 * eager call inside nerf_mae/model/mae/{swin_mae3d,unetr_block,torch_utils}.py.  Each entry point
 * below names the reference code it replaces (file:line, relative to the reference root;
 * S = nerf_mae/model/mae/swin_mae3d.py, U = nerf_mae/model/mae/unetr_block.py,
 * T = nerf_mae/model/mae/torch_utils.py, R = nerf_mae/run_swin_mae3d.py).
 *
 * Conventions:
 *   - plain pointers and sizes only; all tensors are dense fp32 device buffers unless stated;
 *   - every call takes the CUDA device ordinal and the stream (cudaStream_t as void*) explicitly
 *     and only enqueues work: no host synchronisation, no allocation, no ownership of caller memory;
 *   - workspaces are caller-provided; sizes are stated per function;
 *   - return 0 on success, <0 on error; nmae_last_error() returns a thread-local message;
 *   - token tensors are channels-last (B,H,W,D,C); decoder volumes are channels-last (B,X,Y,Z,C);
 *     the raw grids are (B,4,R,R,R) as in the reference.
 */
#ifndef NMAE_H
#define NMAE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int nmae_version(void);
const char* nmae_last_error(void);
/* number of CUDA kernels this library has launched in this process (not thread-safe: a statistic). */
unsigned long long nmae_launch_count(void);

/* Sizes in bytes of the caller-provided scratch / saved buffers named in the entry points below (the library never allocates). */
long long nmae_linear_weight_ws_bytes(int N, int K);                       /* w_ws of nmae_linear_fwd / _bwd_input */
long long nmae_patch_embed_weight_ws_bytes(int C, int p);                  /* w_ws of nmae_patch_embed_fwd */
long long nmae_patch_embed_bwd_ws_bytes(int B, int R, int p, int C);       /* dconv_ws of nmae_patch_embed_bwd */
long long nmae_patch_merge_weight_ws_bytes(int C);                         /* w_ws of nmae_patch_merge_fwd / _bwd */
long long nmae_patch_merge_bwd_ws_bytes(int B, int H, int W, int D, int C); /* dnormed_ws of nmae_patch_merge_bwd */
long long nmae_convT_weight_ws_bytes(int Cin, int Cout, int k);            /* w_ws of nmae_convT_k_eq_s_fwd / _bwd */
long long nmae_conv3x3x3_weight_ws_bytes(int Cin, int Cout);               /* w_ws of nmae_conv3x3x3_* (>= nmae_conv3h_weight_ws_bytes) */
long long nmae_window_attention_lse_bytes(int B, int H, int W, int D, int num_heads);   /* lse of nmae_window_attention_fwd */
long long nmae_instnorm_stats_bytes(int B, int C);                         /* stats of nmae_instnorm_stats */
long long nmae_in_lrelu_bwd_sums_ws_bytes(int B, int C);                   /* sums_ws of nmae_in_lrelu_apply_bwd* (3*B*C doubles + the float constants of the _image_h variant) */

/* T:56-90 pad_tensor + S:1432-1448 transform: zero-pad one (4,X,Y,Z) grid into slot b of (B,4,R,R,R); an extent larger than R is
 * cropped at the high end (what F.pad does with the negative pads pad_tensor computes). */
int nmae_pad_grid(const float* grid, int X, int Y, int Z, float* batch, int b, int R, int device, void* stream);

/* nerf_rpn/datasets.py:88-104 (scene decoding: density -> alpha, uint8 / 255, channels first) + :172-234 (box-free z-up
 * augmentation: rotate = transpose(1,2) then flip of axis 1, then flips of axes 1 and 2) + T:56-90 (zero padding), in one pass
 * over the RAW `rgbsigma` array (W,L,H,4), float32 or uint8, into slot b of (B,4,R,R,R).  The extents of the result are
 * (rotate ? L : W, rotate ? W : L, H).  The Python RNG draws that decide rotate / flips stay on the host. */
int nmae_ingest_scene(const void* rgbsigma, int is_uint8, int normalize_density, int W, int L, int H, int rotate, int flip_axis1,
                      int flip_axis2, float* batch, int b, int R, int device, void* stream);

/* S:1120-1129,1455-1463 patch_partition (Conv3d k=s=p as implicit GEMM + LayerNorm) + pos_embed add +
 * window_masking_3d's token replacement.  x (B,4,R,R,R); w (C,4*p^3); pos (T,C) with T=(R/p)^3;
 * mask (T) bytes or NULL (1 = replace by mask_token); outputs: conv (B*T,C) saved for backward,
 * mean/rstd (B*T), tokens (B*T,C).  w_ws: C*4*p^3 floats of scratch selects the tcgen05 GEMM (p == 4: the patches are gathered by
 * the operand producers straight from the grid), NULL the CUDA-core kernel. */
int nmae_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                         const float* pos, const uint8_t* mask, const float* mask_token, int B, int R, int p, int C,
                         float eps, float* conv, float* mean, float* rstd, float* tokens, float* w_ws, int device, void* stream);
/* backward of the above; dconv_ws (B*T,C) workspace; dw (C,4*p^3), dbias, dln_w, dln_b, dmask_token (C) are overwritten. */
int nmae_patch_embed_bwd(const float* dtokens, const float* x, const float* w, const float* ln_w, const float* conv,
                         const float* mean, const float* rstd, const uint8_t* mask, int B, int R, int p, int C,
                         float* dconv_ws, float* dw, float* dbias, float* dln_w, float* dln_b, float* dmask_token,
                         int device, void* stream);

/* nn.LayerNorm(C, eps) over rows (S:342,351). */
int nmae_layernorm_fwd(const float* x, const float* w, const float* b, int rows, int C, float eps, float* y, float* mean,
                       float* rstd, int device, void* stream);
/* dx = dresid + LN'(dy) (dresid may be NULL or alias dx: fuses the residual-branch gradient add of S:366-369). */
int nmae_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd, int rows, int C,
                       const float* dresid, float* dx, float* dw, float* db, int device, void* stream);

/* F.linear (S:108,173; torchvision MLP S:352-358): out[M,N] = epi(x[M,K] w[N,K]^T + bias).
 * flags: 1 = GELU (aux[M,N] receives the pre-activation), 2 = residual: out = resid + row_scale[m/rows_per_scale]*value
 * (row_scale NULL = 1; this is the stochastic-depth "row" mode of S:366-369).
 * w_ws: N*K floats of scratch for the tensor-core path (weights re-laid as bf16 hi/lo blobs); NULL selects the
 * CUDA-core kernel.  flags & 256: w_ws already holds the blob (see nmae_linear_prep_batch). */
int nmae_linear_fwd(const float* x, const float* w, const float* bias, int M, int N, int K, int flags, float* aux,
                    const float* resid, const float* row_scale, int rows_per_scale, float* out, float* w_ws, int device,
                    void* stream);
/* dx[M,K] = (dy[M,N] w[N,K]) (* gelu'(aux[M,K]) when flags&1: fuses the GELU backward of the layer below);
 * flags&4: dx += instead of overwrite. */
int nmae_linear_bwd_input(const float* dy, const float* w, int M, int N, int K, int flags, const float* aux, float* dx,
                          float* w_ws, int device, void* stream);
/* Weight blobs outside the GEMM call.  nmae_linear_fwd / _bwd_input re-lay the weight into w_ws on every call unless flags & 256
 * says that w_ws already holds the blob of THIS weight for THIS (M, N, K): a trainer re-lays every linear weight of the model with
 * one nmae_linear_prep_batch launch per optimiser step instead of one small kernel per GEMM.
 * nmae_linear_blob_layout: tile_n / k_group the tensor-core kernel uses for out[M,N] = A[M,K] W^T (0, 0: CUDA-core path, no blob).
 * nmae_linear_prep_batch: table = n rows of 8 int64 on the device {w*, blob*, N, K, tile_n, s_n, s_k, k_group} where the GEMM
 * reads W(n,k) = w[n*s_n + k*s_k] (forward of F.linear: s_n = K, s_k = 1; input gradient, whose "N" is the layer's in_features
 * and "K" its out_features: s_n = 1, s_k = in_features); max_elems = the largest N*K of the table. */
int nmae_linear_blob_layout(int M, int N, int K, int* tile_n, int* k_group, int device);
int nmae_linear_prep_batch(const long long* table, int n, long long max_elems, int device, void* stream);

/* dw[N,K] = dy^T x ; db[N] = colsum(dy) (db may be NULL). Both overwritten. */
int nmae_linear_bwd_weight(const float* dy, const float* x, int M, int N, int K, float* dw, float* db, int device,
                           void* stream);

/* S:27-197 shifted_window_attention core (window 4x4x4, head_dim 32), index-remapped (no roll/pad copies).
 * qkv (B*T+1, 3C): rows 0..B*T-1 are the projected tokens, row B*T must hold the qkv bias (padding tokens);
 * table (343,nH) relative_position_bias_table; out (B*T,C) head-concatenated; lse (B*nW*nH*64) saved. */
int nmae_window_attention_num_windows(int H, int W, int D);
int nmae_window_attention_fwd(const float* qkv, const float* table, int B, int H, int W, int D, int C, int num_heads,
                              int shift, float* out, float* lse, int device, void* stream);
/* dqkv (B*T+1,3C) overwritten (row B*T = gradient reaching the qkv bias through padding tokens);
 * dtable (343,nH) overwritten. */
int nmae_window_attention_bwd(const float* dout, const float* qkv, const float* table, const float* out, const float* lse,
                              int B, int H, int W, int D, int C, int num_heads, int shift, float* dqkv, float* dtable,
                              int device, void* stream);

/* S:372-414 PatchMerging: 2x2x2 gather (zero pad odd dims) + LayerNorm(8C) -> normed (B*T2, 8C) [saved]
 * then reduction Linear(8C->2C, no bias) -> out (B*T2, 2C). */
int nmae_patch_merge_fwd(const float* x, const float* ln_w, const float* ln_b, const float* red_w, int B, int H, int W,
                         int D, int C, float eps, float* normed, float* mean, float* rstd, float* out, float* w_ws,
                         int device, void* stream);
/* dnormed_ws (B*T2,8C) workspace; dx (B,H,W,D,C), dln_w, dln_b (8C), dred_w (2C,8C) overwritten.
 * w_ws (both calls): 16*C*C floats of scratch for the tensor-core path, or NULL. */
int nmae_patch_merge_bwd(const float* dout, const float* x, const float* ln_w, const float* red_w, const float* normed,
                         const float* mean, const float* rstd, int B, int H, int W, int D, int C, float* dnormed_ws,
                         float* dx, float* dln_w, float* dln_b, float* dred_w, float* w_ws, int device, void* stream);

/* U:151-158 ConvTranspose3d with kernel == stride == k: x (B,X,Y,Z,Cin) channels-last, w (Cin,Cout,k,k,k),
 * out written into channels [0,Cout) of a (B,kX,kY,kZ,ld_out) buffer (ld_out > Cout when a skip is concatenated, U:196-198). */
/* w_ws: Cin*Cout*k^3 floats of scratch for the tensor-core path (NULL selects the CUDA-core kernel). */
int nmae_convT_k_eq_s_fwd(const float* x, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin, int Cout,
                          int k, float* out, int ld_out, float* w_ws, int device, void* stream);
int nmae_convT_k_eq_s_bwd(const float* dout, int ld_out, const float* x, const float* w, int B, int X, int Y, int Z, int Cin,
                          int Cout, int k, float* dx, float* dw, float* dbias, float* w_ws, int device, void* stream);

/* U:40-56 3x3x3 Conv3d, padding 1, stride 1, on channels-last volumes; w (Cout,Cin,3,3,3) as in the state dict.
 * w_ws: workspace of 27*Cin*Cout floats (GEMM-ordered weights). */
/* Tensor-core operand images.  The tcgen05 convolution kernels read their activations from a bf16 hi/lo image tensor that
 * is laid out exactly like their shared-memory operand (csrc/uimg.cuh), so staging is pure cp.async.bulk.  Build it once per
 * activation with nmae_conv3_image_build (type_dy = 0: halo columns carry the neighbouring strips' voxels - the form every
 * convolution kernel here takes, forward, dgrad and weight gradient alike; type_dy = 1: halo columns zero, kept for callers
 * that want a position space in which every voxel appears exactly once) and reuse it for every kernel that consumes that
 * activation.  C must be a multiple of 48 (nmae_conv3_image_bytes returns 0 otherwise: use the fp32 path).
 * x: channels [ch_off, ch_off+C) of a channels-last volume with ld floats per voxel; channels past the end of the voxel
 * record (ch_off + c >= ld) read as zero, so C may be the next multiple of 48 above the tensor's channel count (the
 * convolution weights are then zero-padded to C input channels by the caller). */
long long nmae_conv3_image_bytes(int B, int X, int Y, int Z, int C);
int nmae_conv3_image_build(const float* x, int ld, int ch_off, int B, int X, int Y, int Z, int C, int type_dy, void* image,
                           int device, void* stream);
/* Fused U:59-60 + image build: the type 0 image of LeakyReLU_slope(InstanceNorm(x)) straight from the convolution output x
 * and its statistics (nmae_instnorm_stats), without materialising the fp32 activation. */
int nmae_conv3_image_build_in_lrelu(const float* x, const double* stats, int B, int X, int Y, int Z, int C, float eps, float slope,
                                    void* image, int device, void* stream);
/* x_image (type 0 image of x) selects the tcgen05 path; with x_image == NULL the CUDA-core kernel runs on the fp32 volume x. */
int nmae_conv3x3x3_fwd(const float* x, const void* x_image, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin,
                       int Cout, float* w_ws, float* out, int device, void* stream);
/* dx = dgrad; accumulate!=0 adds into dx (identity-residual gradient already stored there). dout_image: type 0 image of dout. */
int nmae_conv3x3x3_dgrad(const float* dout, const void* dout_image, const float* w, int B, int X, int Y, int Z, int Cin, int Cout,
                         float* w_ws, float* dx, int accumulate, int device, void* stream);
/* dw (Cout,Cin,3,3,3), dbias (Cout) overwritten; w_ws as above.  With both x_image and dout_image (type 0 images, i.e. the
 * ones the forward convolution and the dgrad consume) the tcgen05 kernel runs and x may be NULL; otherwise the CUDA-core
 * kernel runs on the fp32 volumes.  dout (fp32) is needed only for dbias (its column sum) and for the CUDA-core kernel. */
int nmae_conv3x3x3_wgrad(const float* dout, const void* dout_image, const float* x, const void* x_image, int B, int X, int Y, int Z,
                         int Cin, int Cout, float* w_ws, float* dw, float* dbias, int device, void* stream);

/* U:77 InstanceNorm3d statistics: stats (B,C,2) doubles {sum, sum of squares} over the V voxels of each (b,c). */
int nmae_instnorm_stats(const float* x, int B, int V, int C, double* stats, int device, void* stream);
/* U:57-71: out = LeakyReLU_slope( IN(x) + R ), R = 0 (res NULL) | res (res_stats NULL) | IN(res). */
int nmae_in_lrelu_apply_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V,
                            int C, float eps, float slope, float* out, int device, void* stream);
/* The same, fused with the 1x1x1 output convolution that consumes the result (UnetOutBlock U:96-116, S:1495): `out` is written as
 * above and pred (B*V, 4) = out . w_out^T + b_out (w_out (4, C), b_out (4) or NULL; C % 4 == 0, C <= 128) is formed in the same pass. */
int nmae_in_lrelu_apply_out_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C,
                                float eps, float slope, float* out, const float* w_out, const float* b_out, float* pred, int device,
                                void* stream);
/* sums_ws: 3*B*C doubles. dx always; dx3 when x3!=NULL (gradient of the normalised residual branch);
 * dres when non-NULL receives the identity-residual gradient.  out (the forward result) may be NULL when the forward had
 * no residual: LeakyReLU(IN(x)) has the sign of IN(x), which is recomputed.  dbias / dbias3 (C floats each, optional)
 * receive the column sums of dx / dx3, i.e. the bias gradients of the convolutions that produced x / x3. */
int nmae_in_lrelu_apply_bwd(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                            const double* stats3, int B, int V, int C, float eps, float slope, double* sums_ws, float* dx,
                            float* dx3, float* dres, float* dbias, float* dbias3, int device, void* stream);

/* nmae_in_lrelu_apply_bwd with the gradient wrt x written as a type 0 operand image (nmae_conv3_image_bytes(B,X,Y,Z,C) bytes)
 * instead of an fp32 volume: its only consumers are nmae_conv3x3x3_dgrad / _wgrad, which read images.  C % 48 == 0. */
int nmae_in_lrelu_apply_bwd_image(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                                  const double* stats3, int B, int X, int Y, int Z, int C, float eps, float slope, double* sums_ws,
                                  void* dx_image, float* dx3, float* dres, float* dbias, float* dbias3, int device, void* stream);

/* ---- Single-pass fp16 convolution path ("fp16" precision mode).  Same operators as the image-based entry points above
 * (U:40-56 Conv3d k3 p1, its input and weight gradients, U:57-71 InstanceNorm+LeakyReLU backward), with ONE fp16 operand image
 * per activation (11-bit significands: the class of the TF32 arithmetic the reference's cuDNN convolutions use on a GPU by
 * default) instead of the bf16 hi/lo pair, fp32 accumulation, and channel groups of 48 or 64 (C % 48 == 0 or C % 64 == 0;
 * nmae_conv3h_image_bytes returns 0 otherwise).  Output-gradient images are stored multiplied by a per-tensor power of two
 * (fp16 has 5 exponent bits); its reciprocal lives in a device float that the dgrad / wgrad calls take as inv_scale (NULL = 1). */
long long nmae_conv3h_image_bytes(int B, int X, int Y, int Z, int C);
/* bytes of w_ws for nmae_conv3h_fwd / _dgrad (fp16 weight blobs in the UMMA layout). */
long long nmae_conv3h_weight_ws_bytes(int Cin, int Cout);
/* image of channels [ch_off, ch_off+C) of a channels-last volume (ld floats per voxel); stats != NULL (requires ch_off == 0,
 * ld == C) builds the image of LeakyReLU_slope(InstanceNorm(x)) from the statistics of nmae_instnorm_stats; scale != NULL
 * multiplies by the device float *scale before the fp16 conversion (gradient images; pass its reciprocal as inv_scale below). */
int nmae_conv3h_image_build(const float* x, int ld, int ch_off, int B, int X, int Y, int Z, int C, const double* stats, float eps,
                            float slope, const float* scale, void* image, int device, void* stream);
int nmae_conv3h_fwd(const void* x_image, const float* w, const float* bias, int B, int X, int Y, int Z, int Cin, int Cout, void* w_ws,
                    float* out, int device, void* stream);
int nmae_conv3h_dgrad(const void* dout_image, const float* inv_scale, const float* w, int B, int X, int Y, int Z, int Cin, int Cout,
                      void* w_ws, float* dx, int accumulate, int device, void* stream);
int nmae_conv3h_wgrad(const void* dout_image, const float* inv_scale, const void* x_image, int B, int X, int Y, int Z, int Cin, int Cout,
                      float* dw, int device, void* stream);
/* nmae_in_lrelu_apply_bwd_image with the gradient written as a scaled fp16 image; amax_ws: one float of scratch;
 * inv_scale: one device float that receives the reciprocal of the image's scale.
 * dpred4 / w_out (both or neither): when the block's output only feeds the 1x1x1 output convolution C -> 4 (U:96-116, S:1495), its
 * input gradient dout[v][c] = sum_k w_out[k][c] * dpred4[v][k] is evaluated on the fly from dpred4 (B*V,4) and w_out (4,C) and
 * `dout` may be NULL: the C-channel gradient volume is never written or read.
 * dw_out (4,C) / db_out (4) (both or neither; need dpred4 and x3 == NULL): also receive the weight / bias gradient of that output
 * convolution, dw_out[k][c] = sum_v dpred4[v][k] * out[v][c], from the same pass over `out`. */
int nmae_in_lrelu_apply_bwd_image_h(const float* dout, const float* out, const float* x, const double* stats, const float* x3,
                                    const double* stats3, int B, int X, int Y, int Z, int C, float eps, float slope, double* sums_ws,
                                    float* amax_ws, void* dx_image, float* inv_scale, float* dx3, float* dres, float* dbias,
                                    float* dbias3, const float* dpred4, const float* w_out, float* dw_out, float* db_out, int device,
                                    void* stream);

/* nerf_rpn/model/fpn.py:148-158 (FPN top-down path): fine (B,Xf,Yf,Zf,C) += nearest-neighbour upsample of coarse
 * (B,Xc,Yc,Zc,C) to the fine size (F.interpolate mode="nearest", size=fine), channels-last, in place. */
int nmae_upsample_nearest_add(float* fine, const float* coarse, int B, int Xf, int Yf, int Zf, int Xc, int Yc, int Zc, int C,
                              int device, void* stream);

/* S:593-610 (legacy SwinTransformer_MAE3D decoder): nn.Upsample(size=(Xf,Yf,Zf), mode="trilinear", align_corners=False) on a
 * channels-last volume, and its backward (dcoarse overwritten). */
int nmae_upsample_trilinear_fwd(const float* coarse, float* fine, int B, int Xc, int Yc, int Zc, int Xf, int Yf, int Zf, int C, int device,
                                void* stream);
int nmae_upsample_trilinear_bwd(const float* dfine, float* dcoarse, int B, int Xc, int Yc, int Zc, int Xf, int Yf, int Zf, int C,
                                int device, void* stream);

/* out[C] = column sums of x (rows x C, row stride ld): bias gradients. */
int nmae_colsum(const float* x, long long rows, int C, long long ld, float* out, int device, void* stream);
/* dst = src * row_scale[row / rows_per_scale]: backward side of torchvision StochasticDepth "row" (S:366-369). */
int nmae_scale_rows(float* dst, const float* src, const float* row_scale, int rows_per_scale, long long rows, int cols,
                    int device, void* stream);

/* copy `cols` channels between channels-last buffers of different channel stride (skip concat, U:196-198). */
int nmae_copy_cols(float* dst, long long ld_dst, const float* src, long long ld_src, long long rows, int cols, int device,
                   void* stream);

/* S:1513-1563 forward_loss without patchify copies.  x (B,4,R,R,R); pred (B,R,R,R,4) channels-last;
 * ext (B,3) int32 un-padded extents; tok_mask ((R/p)^3) bytes; sums_ws 4 doubles (kept for backward);
 * out3 = {loss, loss_rgb, loss_alpha}. */
int nmae_mae_loss_fwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p,
                      double* sums_ws, float* out3, int device, void* stream);
/* gout3: upstream gradients of the three outputs (device). dpred (B,R,R,R,4) overwritten. */
int nmae_mae_loss_bwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p,
                      const double* sums_ws, const float* gout3, float* dpred, int device, void* stream);

/* R:663-669: clip_grad_norm_ + AdamW over a device chunk table (nchunks rows of 6 int64:
 * {param*, grad*, exp_avg*, exp_avg_sq*, count, dst*}).  nmae_multi_sumsq writes sum(grad^2) to norm_sq;
 * nmae_multi_copy packs grad -> dst (flat all-reduce bucket). grad_scale folds the 1/world_size of DDP. */
int nmae_multi_sumsq(const long long* table, int nchunks, double* norm_sq, int device, void* stream);
int nmae_multi_copy(const long long* table, int nchunks, int device, void* stream);
int nmae_adamw_clip_step(const long long* table, int nchunks, const double* norm_sq, float clip, float grad_scale, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                         float bias_correction2, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif
