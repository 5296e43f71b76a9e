// C-ABI surface (include/stv_b200.h) over the kernel launchers, plus the shared host utilities:
// thread-local error string, cached device properties and the TMA descriptor encoder.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/stv_b200.h"
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int device_sm_count() {
  static int cached[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  STV_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  STV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned",
              base);
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    STV_REQUIRE(gstr[i] % 16 == 0, "TMA stride %llu not a multiple of 16 bytes",
                (unsigned long long)gstr[i]);
  }
  const CUresult r =
      fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank),
         const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         swizzle == kSwizzle128B          ? CU_TENSOR_MAP_SWIZZLE_128B
         : swizzle == kSwizzle128BAtom32B ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  STV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int check_device_impl() {
  int dev = 0;
  STV_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  STV_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  STV_REQUIRE(major == 10,
              "stv_b200 kernels are built for sm_100a only; device %d has compute capability %d.x "
              "(no fallback path)",
              dev, major);
  return 0;
}

}  // namespace stv

using namespace stv;
#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

const char* stv_last_error(void) { return last_error(); }
int stv_abi_version(void) { return STV_ABI_VERSION; }
int stv_device_check(void) { return check_device_impl(); }

int stv_pack_conv_weights(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin,
                          void* stream) {
  return pack_conv_weights_launch(w, w_fwd, w_dgrad, Cout, Cin, S(stream));
}

// fused convolution entry points: thin argument marshalling over conv_igemm2_launch --------------
static ConvArgs fwd_args(const float* x, const float* w_fwd, const float* bias, int H, int W,
                         int Cin, int Cout, float* out_pre, float* out_post, int round_pre) {
  ConvArgs a;
  a.x = x; a.w_packed = w_fwd; a.H = H; a.W = W; a.C = Cin; a.N = Cout; a.taps = 9;
  a.bias = bias; a.out_pre = out_pre; a.out_post = out_post;
  // out_post feeds the next conv's MMA -> stored tf32-rounded; out_pre (read by the losses) exact
  // unless it only feeds the Gram contraction (round_pre)
  a.round_flags = 2 | (round_pre ? 1 : 0);
  return a;
}
static ConvArgs dgrad_args(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                           int accumulate, float* dx) {
  ConvArgs a;
  a.x = dy; a.w_packed = w_dgrad; a.H = H; a.W = W; a.C = Cout; a.N = Cin; a.taps = 9;
  a.add_src = accumulate ? dx : nullptr;
  a.out_pre = dx;
  a.round_flags = 1;  // gradient buffers are the next dgrad's MMA operand
  return a;
}

int stv_conv3x3_first_fwd(const float* img_nchw, const float* w, const float* bias, int H, int W,
                          int Cout, float* out_pre, float* out_post, int round_pre, void* stream) {
  return conv_first_fwd_launch(img_nchw, w, bias, H, W, Cout, out_pre, out_post, nullptr, round_pre,
                               S(stream));
}
int stv_conv3x3_first_fwd_bits(const float* img_nchw, const float* w, const float* bias, int H,
                               int W, int Cout, float* out_pre, float* out_post,
                               unsigned* out_bits, int round_pre, void* stream) {
  return conv_first_fwd_launch(img_nchw, w, bias, H, W, Cout, out_pre, out_post, out_bits,
                               round_pre, S(stream));
}

int stv_conv3x3_fwd(const float* x, const float* w_fwd, const float* bias, int H, int W, int Cin,
                    int Cout, float* out_pre, float* out_post, int round_pre, void* stream) {
  return conv_igemm2_launch(fwd_args(x, w_fwd, bias, H, W, Cin, Cout, out_pre, out_post, round_pre),
                            S(stream));
}
int stv_conv3x3_fwd_bits(const float* x, const float* w_fwd, const float* bias, int H, int W,
                         int Cin, int Cout, float* out_pre, float* out_post, unsigned* out_bits,
                         int round_pre, void* stream) {
  ConvArgs a = fwd_args(x, w_fwd, bias, H, W, Cin, Cout, out_pre, out_post, round_pre);
  a.out_bits = out_bits;
  return conv_igemm2_launch(a, S(stream));
}

int stv_conv3x3_fwd_pool(const float* x, const float* w_fwd, const float* bias, int H, int W,
                         int Cin, int Cout, float* out_pre, float* out_post, float* out_pool,
                         int round_pre, void* stream) {
  STV_REQUIRE(out_post != nullptr && out_pool != nullptr,
              "stv_conv3x3_fwd_pool: out_post and out_pool are required");
  ConvArgs a = fwd_args(x, w_fwd, bias, H, W, Cin, Cout, out_pre, out_post, round_pre);
  a.out_pool = out_pool;
  return conv_igemm2_launch(a, S(stream));
}
int stv_conv3x3_fwd_pool_code(const float* x, const float* w_fwd, const float* bias, int H, int W,
                              int Cin, int Cout, float* out_pre, float* out_post, float* out_pool,
                              unsigned* out_code, int round_pre, void* stream) {
  STV_REQUIRE(out_pool != nullptr && out_code != nullptr,
              "stv_conv3x3_fwd_pool_code: out_pool and out_code are required");
  ConvArgs a = fwd_args(x, w_fwd, bias, H, W, Cin, Cout, out_pre, out_post, round_pre);
  a.out_pool = out_pool;
  a.out_code = out_code;
  return conv_igemm2_launch(a, S(stream));
}

int stv_conv3x3_dgrad(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                      const float* relu_src, int accumulate, float* dx, void* stream) {
  ConvArgs a = dgrad_args(dy, w_dgrad, H, W, Cout, Cin, accumulate, dx);
  a.mask_src = relu_src;
  return conv_igemm2_launch(a, S(stream));
}
int stv_conv3x3_dgrad_bits(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                           const unsigned* relu_bits, int accumulate, float* dx, void* stream) {
  ConvArgs a = dgrad_args(dy, w_dgrad, H, W, Cout, Cin, accumulate, dx);
  a.mask_bits = relu_bits;
  return conv_igemm2_launch(a, S(stream));
}
int stv_conv3x3_dgrad_bits_style(const float* dy, const float* w_dgrad, int H, int W, int Cout,
                                 int Cin, const unsigned* relu_bits, const float* feat,
                                 const float* s, const float* grad_w, float* dx, void* stream) {
  STV_REQUIRE(feat && s && grad_w, "stv_conv3x3_dgrad_bits_style: feat, s and grad_w are required");
  ConvArgs a = dgrad_args(dy, w_dgrad, H, W, Cout, Cin, 0, dx);
  a.mask_bits = relu_bits;
  a.style_x = feat;
  a.style_s = s;
  a.style_alpha = grad_w;
  const int rc = conv_igemm2_launch(a, S(stream));
  if (rc != kConvStyleNotFusable) return rc;
  // tile family without the second accumulator (small feature maps): two launches, same result up
  // to the order of the two additions
  if (int rc2 = stv_style_bwd(feat, s, static_cast<long>(H) * W, Cin, grad_w, 0, dx, stream))
    return rc2;
  return stv_conv3x3_dgrad_bits(dy, w_dgrad, H, W, Cout, Cin, relu_bits, 1, dx, stream);
}
int stv_conv3x3_dgrad_unpool(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                             const unsigned* pool_code, int H2, int W2, float* dx, void* stream) {
  STV_REQUIRE(pool_code != nullptr, "stv_conv3x3_dgrad_unpool: pool_code is required");
  ConvArgs a = dgrad_args(dy, w_dgrad, H, W, Cout, Cin, 0, dx);
  a.unpool_code = pool_code;
  a.H2 = H2;
  a.W2 = W2;
  return conv_igemm2_launch(a, S(stream));
}

int stv_conv3x3_first_dgrad(const float* dy, const float* w, int H, int W, int Cout,
                            float* dimg_nchw, void* stream) {
  return conv_first_dgrad_launch(dy, w, H, W, Cout, dimg_nchw, S(stream));
}
int stv_conv3x3_first_dgrad_tc(const float* dy, const float* w16_dgrad, int H, int W, int Cout,
                               float* dimg_nchw, void* stream) {
  ConvArgs a;
  a.x = dy; a.w_packed = w16_dgrad; a.H = H; a.W = W; a.C = Cout; a.N = 16; a.taps = 9;
  a.out_nchw3 = dimg_nchw;
  return conv_igemm2_launch(a, S(stream));
}

int stv_conv3x3_first_dgrad_rows(const float* dy, const float* w_rows, int H, int W, int Cout,
                                 float* dimg_nchw, void* stream) {
  return conv_first_dgrad_tc_launch(dy, w_rows, H, W, Cout, dimg_nchw, S(stream));
}

int stv_maxpool2_fwd(const float* x, int H, int W, int C, float* y, void* stream) {
  return maxpool2_fwd_launch(x, H, W, C, y, S(stream));
}
int stv_maxpool2_bwd(const float* dy, const float* x, int H, int W, int C, int relu_mask, float* dx,
                     void* stream) {
  return maxpool2_bwd_launch(dy, x, H, W, C, relu_mask, dx, S(stream));
}
int stv_relu_fwd(const float* x, long n, float* y, void* stream) {
  return relu_fwd_launch(x, n, y, S(stream));
}
int stv_relu_fwd_bits(const float* x, long n, float* y, unsigned* bits, void* stream) {
  return relu_fwd_bits_launch(x, n, y, bits, S(stream));
}
int stv_relu_bwd(const float* dy, const float* x, long n, int accumulate, float* dx, void* stream) {
  return relu_bwd_launch(dy, x, n, accumulate, dx, S(stream));
}
int stv_add_inplace(float* dst, const float* src, long n, void* stream) {
  return add_inplace_launch(dst, src, n, S(stream));
}

size_t stv_gram_workspace_bytes(long hw, int C) { return gram_workspace_bytes(hw, C); }
int stv_gram_loss_fwd(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                      const float* target, float clamp_max, float* gram_out, float* s_out,
                      float* loss_out, void* stream) {
  return gram_launch(x, hw, C, workspace, workspace_bytes, target, clamp_max, gram_out, s_out,
                     loss_out, nullptr, S(stream));
}
int stv_gram_partial_r(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                       float* r_out, void* stream) {
  return gram_launch(x, hw, C, workspace, workspace_bytes, nullptr, 0.f, nullptr, nullptr, nullptr,
                     r_out, S(stream));
}
int stv_gram_from_r(const float* r, int C, double n_total, const float* target, float clamp_max,
                    float* gram_out, float* s_out, float* loss_out, float* scratch, void* stream) {
  return gram_from_r_launch(r, C, n_total, target, clamp_max, gram_out, s_out, loss_out, scratch,
                            S(stream));
}
int stv_style_bwd(const float* x, const float* s, long hw, int C, const float* grad_w,
                  int accumulate, float* dy, void* stream) {
  // dY[p, :] = grad_w * X[p, :] * S  -- a 1x1 "conv" over a (1 x hw) image with weight matrix S
  // (symmetric, so its rows serve directly as the K-major B operand).
  STV_REQUIRE(hw <= 0x7fffffffL, "style_bwd: feature map too large");
  // 1x1: view the feature map as a (hw/32 x 32) image when possible so that 2-D patches apply
  int h2 = 1, w2 = static_cast<int>(hw);
  if (hw % 32 == 0) { h2 = static_cast<int>(hw / 32); w2 = 32; }
  else if (hw % 16 == 0) { h2 = static_cast<int>(hw / 16); w2 = 16; }
  else if (hw % 8 == 0) { h2 = static_cast<int>(hw / 8); w2 = 8; }
  ConvArgs a;
  a.x = x; a.w_packed = s; a.H = h2; a.W = w2; a.C = C; a.N = C; a.taps = 1;
  a.alpha = grad_w;
  a.add_src = accumulate ? dy : nullptr;
  a.out_pre = dy;
  a.round_flags = 1;
  return conv_igemm2_launch(a, S(stream));
}

int stv_reduce_scratch_floats(void) { return reduce_scratch_floats(); }
int stv_content_loss_fwd(const float* f, const float* t, long n, float* partials, float* loss_out,
                         void* stream) {
  return content_fwd_launch(f, t, n, partials, loss_out, S(stream));
}
int stv_content_loss_bwd(const float* f, const float* t, long n, const float* grad_w,
                         int accumulate, float* df, void* stream) {
  return content_bwd_launch(f, t, n, grad_w, accumulate, df, S(stream));
}

int stv_adam_step(float* x, const float* g, float* m, float* v, long n, float beta1, float beta2,
                  float eps, float step_size, float bias2_sqrt, void* stream) {
  return adam_step_launch(x, g, m, v, n, beta1, beta2, eps, step_size, bias2_sqrt, S(stream));
}
int stv_adam_step_dev(float* x, const float* g, float* m, float* v, long n, float lr, float beta1,
                      float beta2, float eps, float* state3, void* stream) {
  return adam_step_dev_launch(x, g, m, v, n, lr, beta1, beta2, eps, state3, S(stream));
}
int stv_dot(const float* a, const float* b, long n, float* partials, float* out, void* stream) {
  return dot_launch(a, b, n, partials, out, S(stream));
}
int stv_absmax_sum(const float* a, long n, float* partials, float* out2, void* stream) {
  return absmax_sum_launch(a, n, partials, out2, S(stream));
}
int stv_axpy(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
             void* stream) {
  return axpy_launch(alpha_dev, alpha_host, x, y, n, S(stream));
}
int stv_scale(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
              void* stream) {
  return scale_launch(alpha_dev, alpha_host, x, y, n, S(stream));
}

size_t stv_lbfgs_workspace_floats(long n, int history) { return lbfgs_workspace_floats(n, history); }
int stv_lbfgs_step(float* x, const float* g, long n, int history, float* hist_s, float* hist_y,
                   float* prev_g, float* d, float* workspace, float lr, float tolerance_grad,
                   float tolerance_change, void* stream) {
  return lbfgs_step_launch(x, g, n, history, hist_s, hist_y, prev_g, d, workspace, lr,
                           tolerance_grad, tolerance_change, S(stream));
}

int stv_frame_to_u8(const float* img_nchw, int H, int W, int denormalize, int rounding,
                    unsigned char* out_hwc, void* stream) {
  return frame_to_u8_launch(img_nchw, H, W, denormalize, rounding, out_hwc, S(stream));
}
int stv_image_from_u8(const unsigned char* img_hwc, int H, int W, int normalize, float* out_nchw,
                      void* stream) {
  STV_REQUIRE(H > 0 && W > 0 && img_hwc != nullptr && out_nchw != nullptr,
              "stv_image_from_u8: empty image or null buffer");
  return image_from_u8_launch(img_hwc, H, W, normalize, out_nchw, S(stream));
}
int stv_nchw_to_nhwc(const float* src, int C, int H, int W, float* dst, void* stream) {
  return nchw_to_nhwc_launch(src, C, H, W, dst, S(stream));
}
int stv_nhwc_to_nchw(const float* src, int C, int H, int W, float* dst, void* stream) {
  return nhwc_to_nchw_launch(src, C, H, W, dst, S(stream));
}
int stv_finite_flags(const float* vals, int n, int* flags, void* stream) {
  return finite_flags_launch(vals, n, flags, S(stream));
}
int stv_step_scores(const float* losses, int n_style, int n_content, float style_w, float content_w,
                    float* scores3, float* loss_ring, int* finite_ring, int capacity, int* counter,
                    void* stream) {
  return step_scores_launch(losses, n_style, n_content, style_w, content_w, scores3, loss_ring,
                            finite_ring, capacity, counter, S(stream));
}

int stv_conv3x3_desc(const stv_conv_desc* d, void* stream) {
  STV_REQUIRE(d != nullptr, "stv_conv3x3_desc: null descriptor");
  ConvArgs a;
  a.x = d->x; a.w_packed = d->w_packed; a.H = d->H; a.W = d->W; a.C = d->C; a.N = d->N;
  a.taps = d->taps; a.x_rows = d->x_rows; a.x_row0 = d->x_row0;
  a.bias = d->bias; a.alpha = d->alpha; a.mask_src = d->mask_src; a.add_src = d->add_src;
  a.out_pre = d->out_pre; a.out_post = d->out_post; a.round_flags = d->round_flags;
  a.out_pool = d->out_pool; a.out_bits = d->out_bits; a.out_code = d->out_code;
  a.mask_bits = d->mask_bits; a.unpool_code = d->unpool_code; a.H2 = d->H2; a.W2 = d->W2;
  a.style_x = d->style_x; a.style_s = d->style_s; a.style_alpha = d->style_alpha;
  return conv_igemm2_launch(a, S(stream));
}
int stv_conv3x3_first_fwd_band(const float* img_nchw, const float* w, const float* bias, int H,
                               int W, int Cout, int in_rows, int in_row0, float* out_pre,
                               float* out_post, unsigned* out_bits, int round_pre, void* stream) {
  STV_REQUIRE(in_rows >= H + in_row0 - 1 && in_row0 >= 0,
              "stv_conv3x3_first_fwd_band: %d input rows do not cover %d output rows", in_rows, H);
  return conv_first_fwd_launch(img_nchw, w, bias, H, W, Cout, out_pre, out_post, out_bits,
                               round_pre, S(stream), in_rows, in_row0);
}

int stv_conv3x3_first_fwd_tc(const float* img_nchw, const float* w, const float* bias, int H, int W,
                             int Cout, int in_rows, int in_row0, float* out_pre, float* out_post,
                             unsigned* out_bits, int round_pre, void* stream) {
  return conv_first_fwd_tc_launch(img_nchw, w, bias, H, W, Cout, in_rows, in_row0, out_pre,
                                  out_post, out_bits, round_pre, S(stream));
}

int stv_halo_exchange(float* mine, float* up, float* down, int rows, int rows_up, int rows_down,
                      long row_floats, int planes, unsigned* flags_mine, unsigned* flags_up,
                      unsigned* flags_down, unsigned* epoch, unsigned* done, int slot, int wait_ready,
                      void* stream) {
  return halo_exchange_launch(mine, up, down, rows, rows_up, rows_down, row_floats, planes,
                              flags_mine, flags_up, flags_down, epoch, done, slot, wait_ready,
                              S(stream));
}

int stv_conv_igemm2_ex(const float* x, const float* w_packed, int H, int W, int C, int N, int taps,
                       const float* bias, const float* alpha, const float* mask_src,
                       const float* add_src, float* out_pre, float* out_post, int block_n,
                       int m_halves, int tw, void* stream) {
  ConvArgs a;
  a.x = x; a.w_packed = w_packed; a.H = H; a.W = W; a.C = C; a.N = N; a.taps = taps;
  a.bias = bias; a.alpha = alpha; a.mask_src = mask_src; a.add_src = add_src;
  a.out_pre = out_pre; a.out_post = out_post;
  a.force_n = block_n; a.force_mh = m_halves; a.force_tw = tw;
  return conv_igemm2_launch(a, S(stream));
}
int stv_conv_set_tuning(int pair_mode, int a_stages, int b_stages, int taps_per_stage) {
  STV_REQUIRE(pair_mode >= -1 && pair_mode <= 1, "stv_conv_set_tuning: pair_mode must be -1, 0 or 1");
  STV_REQUIRE(a_stages >= 0 && a_stages <= 8 && b_stages >= 0 && b_stages <= 16,
              "stv_conv_set_tuning: ring depths out of range");
  STV_REQUIRE(taps_per_stage == 0 || taps_per_stage == 1 || taps_per_stage == 3,
              "stv_conv_set_tuning: taps_per_stage must be 0, 1 or 3");
  conv_set_tuning(pair_mode, a_stages, b_stages, taps_per_stage);
  return 0;
}
int stv_conv_set_epilogue(int staged_mode) {
  STV_REQUIRE(staged_mode >= -1 && staged_mode <= 1, "stv_conv_set_epilogue: mode must be -1, 0 or 1");
  conv_set_epilogue(staged_mode);
  return 0;
}
int stv_conv_set_split(int mode) {
  STV_REQUIRE(mode >= -1 && mode <= 1, "stv_conv_set_split: mode must be -1, 0 or 1");
  conv_set_split(mode);
  return 0;
}
int stv_conv_set_resident(int mode) {
  STV_REQUIRE(mode >= -1 && mode <= 0, "stv_conv_set_resident: mode must be -1 or 0");
  conv_set_resident(mode);
  return 0;
}
int stv_conv_set_pool_smem(int mode) {
  STV_REQUIRE(mode >= -1 && mode <= 0, "stv_conv_set_pool_smem: mode must be -1 or 0");
  conv_set_pool_smem(mode);
  return 0;
}
int stv_conv_plan_override(int H, int W, int C, int N, int backward, int block_n, int m_halves,
                           int pair, int depth, int taps_per_stage) {
  if (H > 0) {
    STV_REQUIRE(block_n == 0 || block_n == 64 || block_n == 128 || block_n == 256,
                "stv_conv_plan_override: block_n must be 0, 64, 128 or 256");
    STV_REQUIRE(m_halves >= 0 && m_halves <= 2 && pair >= -1 && pair <= 1 && depth >= 0 &&
                    depth <= 8 && (taps_per_stage == 0 || taps_per_stage == 1 || taps_per_stage == 3),
                "stv_conv_plan_override: parameter out of range");
  }
  STV_REQUIRE(conv_plan_override(H, W, C, N, backward, block_n, m_halves, pair, depth,
                                 taps_per_stage) == 0,
              "stv_conv_plan_override: table full");
  return 0;
}
int stv_conv_ref(const float* x, const float* w_packed, const float* bias, int H, int W, int C,
                 int N, int taps, int relu, float* out, void* stream) {
  return conv_ref_launch(x, w_packed, bias, H, W, C, N, taps, relu, out, S(stream));
}

}  // extern "C"
