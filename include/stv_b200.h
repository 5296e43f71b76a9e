/*
 * stv_b200.h -- C ABI of the B200-native style-transfer optimisation step.
 *
 * The reference (bjg-gh/style_transfer_visualizer) has no FFI: its hot path calls straight into
 * PyTorch from Python.  Each entry point below therefore names the reference call site whose
 * device work it replaces (paths relative to src/style_transfer_visualizer/ in the reference).
 * The Python package binds these with ctypes and exposes them as torch.library custom ops
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; memory is owned by the caller (PyTorch's
 *     caching allocator in the shipped host code); kernels never allocate;
 *   - activations are NHWC fp32 ([H][W][C], C innermost); the image itself and its gradient are
 *     NCHW fp32 exactly as the reference holds them ([1,3,H,W]);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *     so every call is CUDA-graph capturable;
 *   - return value 0 = ok, non-zero = error; stv_last_error() returns a thread-local message;
 *   - the library refuses to run on anything but compute capability 10.x (no fallback path).
 */
#ifndef STV_B200_H_
#define STV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STV_ABI_VERSION 4

const char* stv_last_error(void);
int stv_abi_version(void);
/* 0 when the current device is sm_100-class, else non-zero (with stv_last_error set). */
int stv_device_check(void);

/* ---- weights (one-off; weights are frozen: core_model.py:103-117) ------------------------- */
/* w: torch Conv2d layout [Cout][Cin][3][3].  w_fwd: [9][Cout][Cin]; w_dgrad: [9][Cin][Cout] with
 * the taps flipped (either output may be NULL). */
int stv_pack_conv_weights(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin,
                          void* stream);

/* ---- VGG conv stack: forward (core_model.py:316 `x = block(x)`, torchvision Conv2d+ReLU) --- */
/* conv1_1: image NCHW [3][H][W] -> NHWC [H][W][64].  out_pre = conv + bias, out_post = relu of it;
 * either may be NULL. */
int stv_conv3x3_first_fwd(const float* img_nchw, const float* w /*[64][3][3][3]*/,
                          const float* bias, int H, int W, int Cout, float* out_pre,
                          float* out_post, int round_pre, void* stream);
/* Same, also recording what autograd's ReLU backward needs (nn.ReLU, core_model.py:134-135):
 * out_bits [H][W][Cout/32] words, bit (c % 32) of word c / 32 = (out_post[y][x][c] > 0).  The
 * dgrad of the next layer gates with these bits (stv_conv3x3_dgrad_bits) instead of re-reading the
 * fp32 activation: 1/32 of the bytes. */
int stv_conv3x3_first_fwd_bits(const float* img_nchw, const float* w, const float* bias, int H,
                               int W, int Cout, float* out_pre, float* out_post,
                               unsigned* out_bits, int round_pre, void* stream);
/* conv1_1 forward on the tensor cores: the K = 27 contraction as four tcgen05.mma (TF32 multiply, FP32
 * accumulate -- what cuDNN does for this layer in the reference's CUDA path), im2col rows built in
 * shared memory straight from the NCHW image.  Same outputs as stv_conv3x3_first_fwd_bits; in_rows /
 * in_row0 as in stv_conv3x3_first_fwd_band (0, 0 = plain image).  w: torch layout [64][3][3][3]. */
int stv_conv3x3_first_fwd_tc(const float* img_nchw, const float* w, const float* bias, int H, int W,
                             int Cout, int in_rows, int in_row0, float* out_pre, float* out_post,
                             unsigned* out_bits, int round_pre, void* stream);
/* 3x3 pad-1 conv on the tensor cores (tcgen05, TF32 multiply, FP32 accumulate).
 * x: NHWC [H][W][Cin], w_fwd from stv_pack_conv_weights, Cin % 32 == 0, Cout % 64 == 0.
 * out_post is always stored rounded to TF32 (it is the next conv's MMA operand); round_pre != 0
 * stores out_pre rounded as well -- used when out_pre only feeds the Gram contraction, whose MMA
 * would otherwise truncate it (a systematic shrink of the Gram entries). */
int stv_conv3x3_fwd(const float* x, const float* w_fwd, const float* bias, int H, int W, int Cin,
                    int Cout, float* out_pre, float* out_post, int round_pre, void* stream);
/* Same + out_bits (see stv_conv3x3_first_fwd_bits). */
int stv_conv3x3_fwd_bits(const float* x, const float* w_fwd, const float* bias, int H, int W,
                         int Cin, int Cout, float* out_pre, float* out_post, unsigned* out_bits,
                         int round_pre, void* stream);
/* The same followed by the MaxPool2d(2, 2) that VGG19 applies to conv1_2 / 2_2 / 3_4 / 4_4
 * (core_model.py:134-135): out_pool [H/2][W/2][Cout] = 2x2 / stride-2 max of out_post, floor mode.
 * The pool is computed in the conv epilogue from the accumulator tile (no re-read of out_post);
 * results are identical to stv_conv3x3_fwd + stv_maxpool2_fwd.  out_post is still written: the
 * pool backward needs it. */
int stv_conv3x3_fwd_pool(const float* x, const float* w_fwd, const float* bias, int H, int W,
                         int Cin, int Cout, float* out_pre, float* out_post, float* out_pool,
                         int round_pre, void* stream);
/* Same, also recording what autograd's max_pool2d + ReLU backward need: out_code [H][W][Cout/32]
 * words, bit (c % 32) of word c / 32 = "the pooled gradient of channel c is routed to THIS pixel":
 * the pixel holds the FIRST maximum of its 2x2 window in ATen's scan order and that maximum is > 0
 * (ReLU gate); pixels of a row / column dropped by floor mode get 0.  With the codes the
 * full-resolution activation has no reader left in the backward pass: out_post may be NULL (saves
 * its write), and stv_conv3x3_dgrad_unpool routes the gradient without a pool-backward kernel. */
int stv_conv3x3_fwd_pool_code(const float* x, const float* w_fwd, const float* bias, int H, int W,
                              int Cin, int Cout, float* out_pre, float* out_post, float* out_pool,
                              unsigned* out_code, int round_pre, void* stream);

/* ---- VGG conv stack: input gradient (autograd of the above; optimization.py:313) ---------- */
/* dx = [relu_src > 0] .* conv_transpose(dy) (+ dx when accumulate != 0).
 * dy: NHWC [H][W][Cout]; dx, relu_src: NHWC [H][W][Cin]; relu_src may be NULL (no gating).
 * w_dgrad from stv_pack_conv_weights.  Cout % 32 == 0, Cin % 64 == 0. */
int stv_conv3x3_dgrad(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                      const float* relu_src, int accumulate, float* dx, void* stream);
/* Same with the ReLU gate given as sign bits (out_bits of the forward conv that produced the
 * gated activation; [H][W][Cin/32]); relu_bits may be NULL (no gating). */
int stv_conv3x3_dgrad_bits(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                           const unsigned* relu_bits, int accumulate, float* dx, void* stream);
/* stv_conv3x3_dgrad_bits fused with the Gram backward (stv_style_bwd) of the layer whose gradient it
 * produces -- the gradient at a style-tapped conv output is  relu'(.) .* dgrad(dy) + grad_w * F S
 * (core_model.py:234-264 through autograd):  dx = bits .* conv_transpose(dy) + grad_w[0] * feat * s.
 * feat: that layer's NHWC features [H][W][Cin], s: its seed matrix from stv_gram_loss_fwd.  The 1x1
 * contraction runs as a second TMEM accumulator of the same tiles (64- and 128-channel layers), so
 * the style gradient never makes a round trip through memory; other shapes fall back to the two
 * separate launches inside this call. */
int stv_conv3x3_dgrad_bits_style(const float* dy, const float* w_dgrad, int H, int W, int Cout,
                                 int Cin, const unsigned* relu_bits, const float* feat,
                                 const float* s, const float* grad_w, float* dx, void* stream);
/* dgrad of the conv that FOLLOWS a MaxPool2d(2, 2), fused with the pool's (and the preceding
 * ReLU's) backward: dy [H][W][Cout] at pooled resolution, pool_code from
 * stv_conv3x3_fwd_pool_code ([H2][W2][Cin/32]), dx [H2][W2][Cin] at the resolution before the pool
 * (H = H2/2, W = W2/2, floor).  Every pooled gradient value is written to the recorded argmax
 * position of its window (zero elsewhere, zero everywhere when the gate bit is clear).  A last row /
 * column of dx dropped by floor mode is NOT written: allocate dx zero-filled. */
int stv_conv3x3_dgrad_unpool(const float* dy, const float* w_dgrad, int H, int W, int Cout, int Cin,
                             const unsigned* pool_code, int H2, int W2, float* dx, void* stream);
/* conv1_1 input gradient: dy NHWC [H][W][64] -> dimg NCHW [3][H][W] (this is input_img.grad). */
int stv_conv3x3_first_dgrad(const float* dy, const float* w /*[64][3][3][3]*/, int H, int W,
                            int Cout, float* dimg_nchw, void* stream);

/* Tensor-core variant of the above: w16_dgrad = [9][16][64] (rows 0..2 = the 3 image channels of
 * the flipped/transposed packing, rows 3..15 zero; N is padded to the smallest UMMA width). */
int stv_conv3x3_first_dgrad_tc(const float* dy, const float* w16_dgrad, int H, int W, int Cout,
                               float* dimg_nchw, void* stream);

/* The same gradient with the x taps folded into the GEMM's N (each dy patch is loaded once instead
 * of once per x shift; the three x-shifted partial results are added in the epilogue):
 * w_rows = [3 row taps t][16 rows n = kx * 3 + ci (9 real)][64 co] = w[co][ci][2 - t][kx],
 * tf32-rounded, rows 9..15 zero.  This is the product path. */
int stv_conv3x3_first_dgrad_rows(const float* dy, const float* w_rows, int H, int W, int Cout,
                                 float* dimg_nchw, void* stream);

/* ---- pooling / ReLU (torchvision MaxPool2d(2,2), nn.ReLU; core_model.py:134-135) ----------- */
int stv_maxpool2_fwd(const float* x, int H, int W, int C, float* y, void* stream);
/* dx = route(dy) to the first max of each 2x2 window; relu_mask != 0 additionally gates by x > 0. */
int stv_maxpool2_bwd(const float* dy, const float* x, int H, int W, int C, int relu_mask, float* dx,
                     void* stream);
int stv_relu_fwd(const float* x, long n, float* y, void* stream);
/* y = tf32(relu(x)) plus the sign bits of y (n / 32 words, as stv_conv3x3_fwd_bits records them): the
 * ReLU of a loss-tapped layer whose conv stored only the pre-activation.  n % 32 == 0. */
int stv_relu_fwd_bits(const float* x, long n, float* y, unsigned* bits, void* stream);
int stv_relu_bwd(const float* dy, const float* x, long n, int accumulate, float* dx, void* stream);
int stv_add_inplace(float* dst, const float* src, long n, void* stream);

/* ---- style loss: gram_matrix + mse_loss (core_model.py:29-63, :234-264) ------------------- */
size_t stv_gram_workspace_bytes(long hw, int C);
/* x: NHWC features [hw][C].  Computes R = F F^T on the tensor cores (upper triangle, split-K),
 * G = min(R, clamp_max) / (C*hw).  Optional outputs (NULL to skip):
 *   gram_out [C][C] = G;   loss_out[0] = mean((G - target)^2);
 *   s_out   [C][C] = 4/(C^2 * C*hw) * 1[R <= clamp_max] * (G - target)   (backward seed). */
int stv_gram_loss_fwd(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                      const float* target, float clamp_max, float* gram_out, float* s_out,
                      float* loss_out, void* stream);
/* Row-band sharding (one image split over several GPUs): the contraction of THIS GPU's pixels only,
 * r_out [C][C] = raw F_band F_band^T (no clamp, no 1/N).  The caller all-reduces r_out over the
 * GPUs (clamp is non-linear, so it must see the global sum) and finishes with stv_gram_from_r, where
 * n_total = C * (pixels of the WHOLE feature map) and scratch holds C*C/256 + 1 floats. */
int stv_gram_partial_r(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                       float* r_out, void* stream);
int stv_gram_from_r(const float* r, int C, double n_total, const float* target, float clamp_max,
                    float* gram_out, float* s_out, float* loss_out, float* scratch, void* stream);
/* Backward of the above w.r.t. the features: dy (+)= grad_w[0] * x * S   (one GEMM instead of the
 * two that torch.mm's autograd issues).  grad_w is a DEVICE scalar (upstream dL/dloss). */
int stv_style_bwd(const float* x, const float* s, long hw, int C, const float* grad_w,
                  int accumulate, float* dy, void* stream);

/* ---- content loss: mse_loss(F, T) (core_model.py:266-295) --------------------------------- */
/* partials: scratch of stv_reduce_scratch_floats() floats. */
int stv_reduce_scratch_floats(void);
int stv_content_loss_fwd(const float* f, const float* t, long n, float* partials, float* loss_out,
                         void* stream);
/* df (+)= grad_w[0] * 2 (f - t) / n */
int stv_content_loss_bwd(const float* f, const float* t, long n, const float* grad_w,
                         int accumulate, float* df, void* stream);

/* ---- image update (torch.optim.Adam / LBFGS on the image; optimization.py:175) ------------ */
int stv_adam_step(float* x, const float* g, float* m, float* v, long n, float beta1, float beta2,
                  float eps, float step_size, float bias2_sqrt, void* stream);
/* Same update with the step counter on the device: state[0] = t (incremented here),
 * state[1..2] = derived scalars.  Needs no host values that change per step => graph replayable. */
int stv_adam_step_dev(float* x, const float* g, float* m, float* v, long n, float lr, float beta1,
                      float beta2, float eps, float* state3, void* stream);
int stv_dot(const float* a, const float* b, long n, float* partials, float* out, void* stream);
/* out2[0] = max |a_i|, out2[1] = sum |a_i|; partials: 2 * stv_reduce_scratch_floats() floats. */
int stv_absmax_sum(const float* a, long n, float* partials, float* out2, void* stream);
/* y += alpha * x ; alpha read from alpha_dev[0] when non-NULL, else alpha_host. */
int stv_axpy(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
             void* stream);
/* y = alpha * x */
int stv_scale(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
              void* stream);

/* Whole L-BFGS step (torch.optim.LBFGS semantics for max_iter = 1, line_search_fn = None -- the
 * reference default, core_model.py:344-349) resident on the device: curvature-pair update,
 * two-loop recursion in coefficient space, step-length rule, tolerance tests and x += t*d, with no
 * host synchronisation.  hist_s / hist_y: [(history + 1)][round_up(n, 4)] floats; prev_g, d: [n];
 * workspace: stv_lbfgs_workspace_floats(n, history) floats, ZERO-INITIALISED before the first step
 * (it holds the iteration counters).  State words: workspace[0] = n_iter (int32),
 * [1] = stored pairs, [3] = 1 if x was updated, [4] = 1 if the tolerance_grad early return fired. */
size_t stv_lbfgs_workspace_floats(long n, int history);
int stv_lbfgs_step(float* x, const float* g, long n, int history, float* hist_s, float* hist_y,
                   float* prev_g, float* d, float* workspace, float lr, float tolerance_grad,
                   float tolerance_change, void* stream);

/* ---- timelapse frame readback (image_io.py:129-152, optimization.py:438-452) --------------- */
/* img NCHW [3][H][W] -> out HWC uint8.  denormalize: apply ImageNet std/mean first.
 * rounding 0 = truncate (timelapse frames), 1 = round-half-even (final frame, main.py:203-214). */
int stv_frame_to_u8(const float* img_nchw, int H, int W, int denormalize, int rounding,
                    unsigned char* out_hwc, void* stream);

/* ---- image load (image_io.py:64-84 apply_transforms: ToTensor + optional Normalize) --------- */
/* img HWC uint8 (device) -> out NCHW fp32 [3][H][W]: x / 255, then (x - mean) / std when
 * normalize != 0, in torchvision's operation order (bit-identical).  Lets the host ship the image
 * as bytes: 4x less PCIe traffic than the reference's fp32 upload. */
int stv_image_from_u8(const unsigned char* img_hwc, int H, int W, int normalize, float* out_nchw,
                      void* stream);

/* ---- misc --------------------------------------------------------------------------------- */
int stv_nchw_to_nhwc(const float* src, int C, int H, int W, float* dst, void* stream);
int stv_nhwc_to_nchw(const float* src, int C, int H, int W, float* dst, void* stream);
/* flags[i] |= !isfinite(vals[i]), i < n <= 32 (optimization.py:375-391, checked lazily). */
int stv_finite_flags(const float* vals, int n, int* flags, void* stream);
/* End of one optimisation step, graph-capturable (optimization.py:298-312 weighted total, :375-391
 * finiteness checks, :402-422 loss recording): scores3 = {sum of the n_style style losses, sum of
 * the n_content content losses, style_w * style + content_w * content}.  When counter != NULL the
 * row is also appended at index (*counter % capacity) of loss_ring [capacity][3] and finite_ring
 * [capacity] (bit 0 / 1 / 2 = style / content / total non-finite), and *counter is incremented: the
 * host reads rows at its logging cadence, never per step. */
int stv_step_scores(const float* losses, int n_style, int n_content, float style_w, float content_w,
                    float* scores3, float* loss_ring, int* finite_ring, int capacity, int* counter,
                    void* stream);

/* ---- general form of the tensor-core convolution -------------------------------------------- */
/* Every stv_conv3x3_* entry point above is a fixed choice of the options of ONE kernel family;
 * callers that combine them differently (the row-band sharded engine: haloed inputs, own-row
 * outputs) fill this descriptor.  out = gate .* (alpha * conv(x, w) + bias) + add [+ style term];
 * out_pre = out, out_post = relu(out); see the individual entry points for each field.  Zero /
 * NULL = option off.  Returns 3 (nothing launched) when a fused style backward is requested for a
 * shape whose tile family has no second accumulator. */
typedef struct stv_conv_desc {
  const float* x;          /* NHWC input [x_rows or H][W][C] */
  const float* w_packed;   /* [taps][N][C] from stv_pack_conv_weights */
  int H, W, C, N, taps;    /* output rows / cols, input / output channels, 9 or 1 */
  int x_rows, x_row0;      /* haloed input: rows of x, and the row of x that lines up with output
                              row 0 (0, 0 = x has H rows, zero padding at its edge) */
  const float* bias;
  const float* alpha;      /* device scalar */
  const float* mask_src;   /* fp32 ReLU gate (x > 0) */
  const float* add_src;
  float* out_pre;
  float* out_post;
  int round_flags;         /* bit 0: store out_pre tf32-rounded, bit 1: out_post */
  float* out_pool;
  unsigned* out_bits;
  unsigned* out_code;
  const unsigned* mask_bits;
  const unsigned* unpool_code;
  int H2, W2;
  const float* style_x;
  const float* style_s;
  const float* style_alpha;
} stv_conv_desc;
int stv_conv3x3_desc(const stv_conv_desc* d, void* stream);
/* conv1_1 forward of a haloed image band: img NCHW [3][in_rows][W], output row 0 reads input rows
 * in_row0 - 1 .. in_row0 + 1 (rows outside [0, in_rows) are zero padding). */
int stv_conv3x3_first_fwd_band(const float* img_nchw, const float* w, const float* bias, int H,
                               int W, int Cout, int in_rows, int in_row0, float* out_pre,
                               float* out_post, unsigned* out_bits, int round_pre, void* stream);

/* ---- row-band sharding over several GPUs (BASELINE configs[4]; no reference counterpart) ----- */
/* Halo exchange of one haloed buffer [rows + 2][row_floats] (x planes) through PEER-MAPPED memory:
 * this GPU's kernels store its first / last own row into the lower halo of the rank above (`up`,
 * which has rows_up own rows) and the upper halo of the rank below (`down`), zero its own halo at
 * the image boundary (NULL neighbour), and hand-shake through flag words.  `up` / `down` /
 * `flags_up` / `flags_down` are this process's mappings of the NEIGHBOURS' buffers (symmetric
 * memory); flags_*: [slots][4] words, zero-initialised; epoch / done: local [slots] words,
 * zero-initialised.  Every rank must issue the same sequence of (slot) exchanges.  On completion (in
 * stream order) the two halo rows of `mine` hold the neighbours' rows.  Graph-capturable.
 * wait_ready != 0 adds a first round trip ("my buffer is final") before the push, needed only when
 * some kernel of a rank writes into the halo rows of its own buffer (e.g. a conv run over band +
 * halos); producers that store own rows only pass 0. */
int stv_halo_exchange(float* mine, float* up, float* down, int rows, int rows_up, int rows_down,
                      long row_floats, int planes, unsigned* flags_mine, unsigned* flags_up,
                      unsigned* flags_down, unsigned* epoch, unsigned* done, int slot, int wait_ready,
                      void* stream);

/* ---- test hooks ---------------------------------------------------------------------------- */
/* Explicit tile selection for the tensor-core conv: out = alpha*conv(x,w)+bias, optional relu gate /
 * accumulate, as in stv_conv3x3_fwd / _dgrad.  taps = 9 or 1; block_n in {64,128,256}, m_halves in
 * {1,2} (128 or 256 pixels per CTA), tw in {8,16,32}; 0 = auto. */
int stv_conv_igemm2_ex(const float* x, const float* w_packed, int H, int W, int C, int N, int taps,
                       const float* bias, const float* alpha, const float* mask_src,
                       const float* add_src, float* out_pre, float* out_post, int block_n,
                       int m_halves, int tw, void* stream);
/* Tuning knobs of the tensor-core conv (per calling host thread; tests and sweeps).  pair_mode: 1 = run every
 * layer on CTA pairs (clusters of two CTAs, tcgen05 cta_group::2: M = 256 per instruction, each CTA
 * stages half of every weight tile), 0 = single-CTA tiles only, -1 = the built-in per-shape rule
 * table (default).  a_stages / b_stages (operand ring depths) and taps_per_stage (1 or 3 weight taps
 * per ring stage): 0 = built-in defaults. */
int stv_conv_set_tuning(int pair_mode, int a_stages, int b_stages, int taps_per_stage);
/* Epilogue store policy of the plain / un-pooling epilogues for the calling host thread: -1 = the
 * measured rule (coalesced shared-memory-transposed stores where a value is stored more than once or
 * the epilogue is the bottleneck), 0 = always direct, 1 = coalesced wherever the tile allows. */
int stv_conv_set_epilogue(int staged_mode);
/* Split-K second MMA issuer of one-half tiles (calling host thread): -1 = the built-in rule (128-wide
 * one-half tiles, and 64-wide ones that have their SM to themselves), 0 = never, 1 = wherever the
 * tile family has the variant.  Results with and without differ in fp32 summation order only. */
int stv_conv_set_split(int mode);
/* Weight-stationary mode of the 64 -> 64 layers (the CTA pair keeps all 9 x 2 weight blocks in
 * shared memory, ring stages carry activations only): -1 = built-in rule, 0 = never (A/B runs).
 * Results are bit-identical either way. */
int stv_conv_set_resident(int mode);
/* Fused 2x2 max pool of tiles up to 128 wide through the shared-memory staging tile (one lane holds
 * a whole window: no shuffles): -1 = built-in rule, 0 = never (the shuffle form; A/B runs).  Results
 * are bit-identical either way. */
int stv_conv_set_pool_smem(int mode);
/* Replace the rule table's tile plan for ONE layer shape (calling host thread only; sweeps): output
 * H x W, C -> N channels, backward = 1 for input-gradient launches.  block_n / m_halves / depth /
 * taps_per_stage: 0 = keep the rule's value; pair: -1 = keep.  H <= 0 clears the table. */
int stv_conv_plan_override(int H, int W, int C, int N, int backward, int block_n, int m_halves,
                           int pair, int depth, int taps_per_stage);
/* Naive CUDA-core NHWC conv, same packed weights; on-device cross-check only. */
int stv_conv_ref(const float* x, const float* w_packed, const float* bias, int H, int W, int C,
                 int N, int taps, int relu, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STV_B200_H_ */
