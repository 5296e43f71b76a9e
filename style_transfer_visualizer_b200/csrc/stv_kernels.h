// Internal launch functions shared between translation units (the C-ABI in api.cu wraps these).
#pragma once
#include <cuda_runtime.h>

namespace stv {

constexpr int kMaxDevices = 64;  // per-device caches (kernel attributes, SM counts)

// conv_igemm2.cu: persistent, tap-reusing tcgen05 implicit-GEMM convolution ----------------------
// out = gate .* (alpha * conv(x, w) + bias) + add; writes out_pre and / or relu(out) to out_post.
// round_flags: bit 0 = store out_pre rounded to tf32, bit 1 = store out_post rounded to tf32.
// N == 16 writes NCHW 3-channel planes (conv1_1 input gradient).  force_* <= 0 select the rule table.
struct ConvArgs {
  const float* x = nullptr;
  const float* w_packed = nullptr;
  int H = 0, W = 0, C = 0, N = 0, taps = 9;
  const float* bias = nullptr;
  const float* alpha = nullptr;         // device scalar
  const float* mask_src = nullptr;      // fp32 ReLU gate source (x > 0)
  const float* add_src = nullptr;
  float* out_pre = nullptr;
  float* out_post = nullptr;
  int round_flags = 0;
  float* out_nchw3 = nullptr;
  float* out_pool = nullptr;            // fused 2x2 max pool of the post-ReLU output
  unsigned* out_bits = nullptr;         // [H][W][N/32] sign bits of the post-ReLU output
  unsigned* out_code = nullptr;         // [H][W][N/32] pool + ReLU backward routing bits
  const unsigned* mask_bits = nullptr;  // ReLU gate as bits (instead of mask_src)
  const unsigned* unpool_code = nullptr;  // route the result through the 2x2 pool backward
  int H2 = 0, W2 = 0;                   // un-pooled output size (unpool_code)
  // dgrad + Gram backward of the layer whose gradient is being produced, as a second accumulator:
  // out = gate .* conv + style_alpha[0] * style_x * style_s   (style_x: NHWC [H][W][N] features of
  // that layer, style_s: symmetric [N][N] seed matrix)
  const float* style_x = nullptr;
  const float* style_s = nullptr;
  const float* style_alpha = nullptr;
  // row-band sharding: x is a haloed buffer of x_rows rows whose row x_row0 lines up with output
  // row 0 (x_rows = 0: x has exactly H rows and the band edge is zero padding)
  int x_rows = 0, x_row0 = 0;
  int force_n = 0, force_mh = 0, force_tw = 0;
};
// returned (nothing launched) when the tile family chosen for this shape has no second accumulator
constexpr int kConvStyleNotFusable = 3;
int conv_igemm2_launch(const ConvArgs& args, cudaStream_t stream);

// Thread-local tile-policy overrides (tests and sweeps).  pair_mode -1: built-in rule table; 0:
// single-CTA tiles only; 1: CTA pairs (cta_group::2) wherever legal.  a_stages / b_stages / tps > 0
// override the ring depths and weight taps per stage.
void conv_set_tuning(int pair_mode, int a_stages, int b_stages, int tps);
// epilogue store policy of the calling thread: -1 measured rule, 0 direct, 1 coalesced where possible
void conv_set_epilogue(int staged_mode);
void conv_set_split(int mode);
void conv_set_resident(int mode);
void conv_set_pool_smem(int mode);
int conv_plan_override(int H, int W, int C, int N, int backward, int block_n, int mh, int pair,
                       int depth, int tps);

// gram.cu ---------------------------------------------------------------------------------
size_t gram_workspace_bytes(long hw, int C);
int gram_launch(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                const float* target, float clamp_max, float* gram_out, float* s_out,
                float* loss_out, float* raw_out, cudaStream_t stream);
int gram_from_r_launch(const float* r, int C, double n_total, const float* target, float clamp_max,
                       float* gram_out, float* s_out, float* loss_out, float* scratch,
                       cudaStream_t stream);

// conv_direct.cu --------------------------------------------------------------------------
int conv_first_fwd_launch(const float* img_nchw, const float* w, const float* bias, int H, int W,
                          int Cout, float* out_pre, float* out_post, unsigned* out_bits,
                          int round_pre, cudaStream_t stream, int in_rows = 0, int in_row0 = 0);
// conv_first_tc.cu: conv1_1 forward on the tensor cores (TF32, K = 27 padded to 32)
int conv_first_fwd_tc_launch(const float* img_nchw, const float* w, const float* bias, int H, int W,
                             int Cout, int in_rows, int in_row0, float* out_pre, float* out_post,
                             unsigned* out_bits, int round_pre, cudaStream_t stream);
int conv_first_dgrad_launch(const float* dy, const float* w, int H, int W, int Cout,
                            float* dimg_nchw, cudaStream_t stream);
// conv_first_dgrad_tc.cu: conv1_1 input gradient, x taps folded into N (dy patch loaded once)
int conv_first_dgrad_tc_launch(const float* dy, const float* w_rows, int H, int W, int Cout,
                               float* dimg_nchw, cudaStream_t stream);
int conv_ref_launch(const float* x, const float* w_packed, const float* bias, int H, int W, int C,
                    int N, int taps, int relu, float* out, cudaStream_t stream);

// elementwise.cu --------------------------------------------------------------------------
int pack_conv_weights_launch(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin,
                             cudaStream_t stream);
int maxpool2_fwd_launch(const float* x, int H, int W, int C, float* y, cudaStream_t stream);
int maxpool2_bwd_launch(const float* dy, const float* x, int H, int W, int C, int relu_mask,
                        float* dx, cudaStream_t stream);
int relu_fwd_launch(const float* x, long n, float* y, cudaStream_t stream);
int relu_fwd_bits_launch(const float* x, long n, float* y, unsigned* bits, cudaStream_t stream);
int relu_bwd_launch(const float* dy, const float* x, long n, int accumulate, float* dx,
                    cudaStream_t stream);
int add_inplace_launch(float* dst, const float* src, long n, cudaStream_t stream);
int content_fwd_launch(const float* f, const float* t, long n, float* partials, float* loss_out,
                       cudaStream_t stream);
int content_bwd_launch(const float* f, const float* t, long n, const float* grad_w, int accumulate,
                       float* df, cudaStream_t stream);
int adam_step_launch(float* x, const float* g, float* m, float* v, long n, float beta1, float beta2,
                     float eps, float step_size, float bias2_sqrt, cudaStream_t stream);
int adam_step_dev_launch(float* x, const float* g, float* m, float* v, long n, float lr, float beta1,
                         float beta2, float eps, float* state, cudaStream_t stream);
int reduce_scratch_floats();
int dot_launch(const float* a, const float* b, long n, float* partials, float* out,
               cudaStream_t stream);
int absmax_sum_launch(const float* a, long n, float* partials, float* out2, cudaStream_t stream);
int axpy_launch(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
                cudaStream_t stream);
int scale_launch(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
                 cudaStream_t stream);
// lbfgs.cu
size_t lbfgs_workspace_floats(long n, int m);
int lbfgs_step_launch(float* x, const float* g, long n, int m, float* hist_s, float* hist_y,
                      float* prev_g, float* d, float* workspace, float lr, float tol_grad,
                      float tol_change, cudaStream_t stream);
int frame_to_u8_launch(const float* img_nchw, int H, int W, int denormalize, int rounding,
                       unsigned char* out_hwc, cudaStream_t stream);
int image_from_u8_launch(const unsigned char* hwc, int H, int W, int normalize, float* out_nchw,
                         cudaStream_t stream);
int nchw_to_nhwc_launch(const float* src, int C, int H, int W, float* dst, cudaStream_t stream);
int nhwc_to_nchw_launch(const float* src, int C, int H, int W, float* dst, cudaStream_t stream);
int finite_flags_launch(const float* vals, int n, int* flags, cudaStream_t stream);
int step_scores_launch(const float* losses, int n_style, int n_content, float style_w,
                       float content_w, float* scores3, float* loss_ring, int* finite_ring,
                       int capacity, int* counter, cudaStream_t stream);

// halo.cu: row-band sharding, halo rows pushed into the neighbours' buffers over NVLink
int halo_exchange_launch(float* mine, float* up, float* down, int rows, int rows_up, int rows_down,
                         long row_floats, int planes, unsigned* flags_mine, unsigned* flags_up,
                         unsigned* flags_down, unsigned* epoch, unsigned* done, int slot,
                         int wait_ready, cudaStream_t stream);

}  // namespace stv
