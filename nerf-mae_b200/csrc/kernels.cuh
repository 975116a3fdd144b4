// Internal kernel launchers shared by the C-ABI layer (api.cu).
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "uimg.cuh"

// norm.cu
int k_layernorm_fwd(const float* x, const int* merge_dims, int rows, int C, const float* w, const float* b, float eps,
                    const float* pos, int pos_rows, const uint8_t* mask, const float* mask_token, float* y, float* mean,
                    float* rstd, cudaStream_t st);
int k_layernorm_bwd(const float* x, const int* merge_dims, int rows, int C, const float* w, const float* dy, const float* mean,
                    const float* rstd, const uint8_t* mask, int pos_rows, float* dx, const float* add_src, float* dgamma, float* dbeta,
                    cudaStream_t st);
int k_colsum(const float* x, int rows, int C, long long ld, const uint8_t* mask, int pos_rows, float* out, cudaStream_t st);
int k_in_stats(const float* x, int B, int V, int C, double* stats, cudaStream_t st);
int k_in_act_fwd(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C, float eps,
                 float slope, float* out, cudaStream_t st);
int k_in_act_fwd_out(const float* x, const double* stats, const float* res, const double* res_stats, int B, int V, int C, float eps,
                     float slope, float* out, const float* w_out, const float* b_out, float* pred, cudaStream_t st);
int k_in_bwd_sums(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                  int B, int V, int C, float eps, float slope, double* sums, cudaStream_t st, float* amax = nullptr,
                  const float* dp4 = nullptr, const float* w4 = nullptr, float* dw_out = nullptr, float* db_out = nullptr);
int k_in_act_bwd(const float* dout, const float* out, const float* x, const double* stats, const float* x3, const double* stats3,
                 int B, int V, int C, float eps, float slope, double* sums, float* dx, float* dx3, float* dres, float* dbias,
                 float* dbias3, cudaStream_t st);

// wmsa_tc.cu (tcgen05 window attention core)
int k_wattn_num_windows(int H, int W, int D);
int k_wattn_tc_fwd(const float* qkv, const float* table, int B, int H, int W, int D, int C, int nH, int shift, float* out, float* lse,
                   cudaStream_t st);
int k_wattn_tc_bwd(const float* qkv, const float* table, const float* o_saved, const float* dout, const float* lse, int B, int H, int W,
                   int D, int C, int nH, int shift, float* dqkv, float* dtable, cudaStream_t st);

// elementwise.cu
int k_pad_grid(const float* src, int Cc, int X, int Y, int Z, float* dst, int R, cudaStream_t st);
int k_gather3(float* dst, const float* src, long long n0, long long n1, long long n2, long long s0, long long s1, long long s2,
              cudaStream_t st);
int k_scale_rows(float* dst, const float* src, const float* row_scale, int rows_per_scale, long long rows, int cols,
                 cudaStream_t st);
int k_copy_cols(float* dst, long long ldd, const float* src, long long lds, long long rows, int cols, cudaStream_t st);
int k_loss_fwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p, double* sums,
               float* out3, cudaStream_t st);
int k_loss_bwd(const float* x, const float* pred, const int* ext, const uint8_t* tok_mask, int B, int R, int p, const double* sums,
               const float* gout3, float* dpred, cudaStream_t st);
int k_ingest_scene(const void* src, int is_u8, int normalize, int W, int L, int H, int rot, int flip1, int flip2, float* dst, int R,
                   cudaStream_t st);
int k_upsample_nearest_add(float* fine, const float* coarse, int B, int Xf, int Yf, int Zf, int Xc, int Yc, int Zc, int C,
                           cudaStream_t st);
int k_upsample_trilinear(const float* src, float* dst, int B, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, int C, int backward,
                         cudaStream_t st);
int k_multi_sumsq(const long long* table, int nchunks, double* out, cudaStream_t st);
int k_multi_copy(const long long* table, int nchunks, cudaStream_t st);
int k_adamw_clip(const long long* table, int nchunks, const double* norm_sq, float clip, float grad_scale, float lr, float b1,
                 float b2, float eps, float wd, float bc1, float bc2, cudaStream_t st);

// conv3_tc.cu (tcgen05 implicit GEMM)
bool k_conv3_tc_supported(int C, int N);
int k_conv3_tc(const void* uimg, const float* w, const float* bias, int B, int Dx, int Dy, int Dz, int C, int N, int mode, float* w_ws,
               float* y, int accumulate, cudaStream_t st);

// conv3_wgrad_tc.cu
bool k_conv3_wgrad_tc_supported(int C, int N);
int k_conv3_wgrad_tc(const void* ximg, const void* yimg, int B, int Dx, int Dy, int Dz, int C, int N, float* dw, cudaStream_t st);

// conv3_h.cu / conv3_wgrad_h.cu (single-pass fp16 operands)
bool k_conv3_h_supported(int C, int N);
long long k_conv3_h_blob_bytes(int C, int N);
int k_conv3_h(const void* uimg, const float* w, const float* bias, const float* out_scale, int B, int Dx, int Dy, int Dz, int C, int N,
              int mode, void* w_ws, float* y, int accumulate, cudaStream_t st);
bool k_conv3_wgrad_h_supported(int C, int N);
int k_conv3_wgrad_h(const void* ximg, const void* yimg, const float* inv_scale, int B, int Dx, int Dy, int Dz, int C, int N, float* dw,
                    cudaStream_t st);

// lin_tc.cu
bool k_lin_tc_supported(int M, int N, int K, long long lda, long long ldc);
int k_lin_tc(const float* a, long long lda, const float* w, long long s_n, long long s_k, int M, int N, int K, const GEpilogue& e,
             float* w_ws, cudaStream_t st, int prep_mode = 0, const GOperand* a_gather = nullptr, bool blob_ready = false);
int k_lin_tc_tile(int M, int N, int K, int* nt, int* kg);
int k_lin_tc_prep_batch(const long long* table, int n, long long max_elems, cudaStream_t st);
bool k_lin_wgrad_tc_supported(int M, int N, int K, long long ldx, long long ldy);
int k_lin_wgrad_tc(const float* x, long long ldx, const float* dy, long long ldy, int M, int N, int K, float* dw, cudaStream_t st,
                   const GOperand* x_gather = nullptr);

// thin_linear.cu (4-wide outputs: the 1x1x1 output convolution)
bool k_thin_supported(int N, int K);
int k_thin_fwd(const float* x, const float* w, const float* bias, long long M, int K, float* y, cudaStream_t st);
int k_thin_dgrad(const float* dy, const float* w, long long M, int K, int accumulate, float* dx, cudaStream_t st);
int k_thin_wgrad(const float* x, const float* dy, long long M, int K, float* dw, float* db, cudaStream_t st);
