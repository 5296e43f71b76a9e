// HBM-bound kernels of the optimisation step: vectorised, coalesced, warp-reduced.
//   max-pool 2x2 fwd / bwd (+ fused ReLU mask)      torchvision vgg.py MaxPool2d(2,2) via core_model.py:316
//   ReLU fwd / bwd (only for non-default tap sets)  core_model.py:134-135
//   content MSE fwd / bwd                            core_model.py:266-295
//   Adam update                                      torch.optim.Adam injected at optimization.py:104-105
//   dot / axpy / scale / |g| stats for L-BFGS        torch.optim.LBFGS, core_model.py:344-349
//   frame conversion fp32 NCHW -> u8 HWC             image_io.py:129-152 + optimization.py:438-452
//   weight re-packing, layout conversion, finiteness flags (optimization.py:375-391)
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

static inline unsigned grid_for(long work_items, int threads, int max_blocks) {
  long b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<unsigned>(b);
}
static inline int stream_blocks() { return device_sm_count() * 8; }
constexpr int kReduceBlocks = 592;  // 148 SMs x 4; fixed so reductions are reproducible
int reduce_scratch_floats() { return kReduceBlocks; }

// ------------------------------------------------------------------------------------------
// weight packing: torch [Cout][Cin][3][3] -> fwd [tap][Cout][Cin], dgrad [8-tap][Cin][Cout]
// ------------------------------------------------------------------------------------------
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, float* __restrict__ wf,
                                         float* __restrict__ wd, int Cout, int Cin) {
  const long total = static_cast<long>(Cout) * Cin * 9;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % 9);
    const long r = i / 9;
    const int ci = static_cast<int>(r % Cin);
    const int co = static_cast<int>(r / Cin);
    const float v = round_tf32(w[i]);  // MMA operand: round once here instead of truncating
    if (wf) wf[(static_cast<long>(tap) * Cout + co) * Cin + ci] = v;
    if (wd) wd[(static_cast<long>(8 - tap) * Cin + ci) * Cout + co] = v;
  }
}

// ------------------------------------------------------------------------------------------
// max-pool 2x2 stride 2, floor mode, NHWC.  One thread = one window x 4 channels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const float4* __restrict__ x, int H, int W, int C4, float4* __restrict__ y) {
  // let a following tensor-core conv (launched with programmatic stream serialization) run its
  // prologue under this kernel's tail; it still waits for our completion before touching data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int Ho = H >> 1, Wo = W >> 1;
  const long total = static_cast<long>(Ho) * Wo * C4;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long p = i / C4;
    const int ox = static_cast<int>(p % Wo), oy = static_cast<int>(p / Wo);
    const float4* r0 = x + (static_cast<long>(2 * oy) * W + 2 * ox) * C4 + c;
    const float4* r1 = r0 + static_cast<long>(W) * C4;
    const float4 a = __ldg(r0), b = __ldg(r0 + C4), d = __ldg(r1), e = __ldg(r1 + C4);
    float4 o;
    // same scan order and NaN propagation as ATen: (v > max) || isnan(v)
#define STV_MAX4(f)                                                     \
  {                                                                     \
    float mv = a.f;                                                     \
    if (b.f > mv || b.f != b.f) mv = b.f;                               \
    if (d.f > mv || d.f != d.f) mv = d.f;                               \
    if (e.f > mv || e.f != e.f) mv = e.f;                               \
    o.f = mv;                                                           \
  }
    STV_MAX4(x) STV_MAX4(y) STV_MAX4(z) STV_MAX4(w)
#undef STV_MAX4
    y[i] = o;
  }
}

// dx[window] = dy routed to the first maximum of the window (ATen tie rule), optionally gated by
// the ReLU that preceded the pool (x is then the post-ReLU tensor: gate = x > 0).  Rows / columns
// dropped by floor mode receive zero.
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, int H, int W,
                    int C4, int relu_mask, float4* __restrict__ dx) {
  // let a following tensor-core conv (launched with programmatic stream serialization) run its
  // prologue under this kernel's tail; it still waits for our completion before touching data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int Ho = H >> 1, Wo = W >> 1;
  const int Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;
  const long total = static_cast<long>(Hc) * Wc * C4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    const long p = i / C4;
    const int ox = static_cast<int>(p % Wc), oy = static_cast<int>(p / Wc);
    const long base = (static_cast<long>(2 * oy) * W + 2 * ox) * C4 + c;
    const bool has_x1 = (2 * ox + 1 < W), has_y1 = (2 * oy + 1 < H);
    if (ox < Wo && oy < Ho) {
      const long rs = static_cast<long>(W) * C4;
      const float4 a = __ldg(x + base), b = __ldg(x + base + C4), d = __ldg(x + base + rs),
                   e = __ldg(x + base + rs + C4);
      const float4 g = __ldg(dy + (static_cast<long>(oy) * Wo + ox) * C4 + c);
      float4 ga = zero, gb = zero, gd = zero, ge = zero;
#define STV_ROUTE(f)                                                    \
  {                                                                     \
    int k = 0;                                                          \
    float mv = a.f;                                                     \
    if (b.f > mv || b.f != b.f) { mv = b.f; k = 1; }                    \
    if (d.f > mv || d.f != d.f) { mv = d.f; k = 2; }                    \
    if (e.f > mv || e.f != e.f) { mv = e.f; k = 3; }                    \
    const float gv = (relu_mask && !(mv > 0.f)) ? 0.f : g.f;            \
    ga.f = k == 0 ? gv : 0.f;                                           \
    gb.f = k == 1 ? gv : 0.f;                                           \
    gd.f = k == 2 ? gv : 0.f;                                           \
    ge.f = k == 3 ? gv : 0.f;                                           \
  }
      STV_ROUTE(x) STV_ROUTE(y) STV_ROUTE(z) STV_ROUTE(w)
#undef STV_ROUTE
      dx[base] = ga;
      dx[base + C4] = gb;
      dx[base + rs] = gd;
      dx[base + rs + C4] = ge;
    } else {
      dx[base] = zero;
      if (has_x1) dx[base + C4] = zero;
      if (has_y1) {
        dx[base + static_cast<long>(W) * C4] = zero;
        if (has_x1) dx[base + static_cast<long>(W) * C4 + C4] = zero;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// ReLU (only used when a loss taps a ReLU / pool output, i.e. non-default layer sets)
// ------------------------------------------------------------------------------------------
__global__ void relu_fwd_kernel(const float4* __restrict__ x, long n4, float4* __restrict__ y) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = make_float4(relu_nan(v.x), relu_nan(v.y), relu_nan(v.z), relu_nan(v.w));
  }
}
// ReLU forward of a tapped layer's stored pre-activation, for layers whose conv epilogue would
// otherwise write both tensors with scattered stores (256-wide tiles have no shared memory left
// for coalescing them): y = tf32(relu(x)) -- the next conv's MMA operand -- and the sign bits the
// dgrad gates with.  One thread = 4 channels; the 8 threads of a 32-channel word are adjacent lanes.
__global__ void __launch_bounds__(256)
relu_fwd_bits_kernel(const float4* __restrict__ x, long n4, float4* __restrict__ y,
                     unsigned* __restrict__ bits) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long stride = static_cast<long>(gridDim.x) * 256;
  const long n4_up = (n4 + 255) & ~255L;  // whole warps iterate together (shuffles below)
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n4_up; i += stride) {
    const bool live = i < n4;
    float4 v = live ? __ldg(x + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = round_tf32(relu_nan(v.x)); v.y = round_tf32(relu_nan(v.y));
    v.z = round_tf32(relu_nan(v.z)); v.w = round_tf32(relu_nan(v.w));
    if (live) y[i] = v;
    unsigned nib = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) |
                   (v.w > 0.f ? 8u : 0u);
    nib <<= 4 * (threadIdx.x & 7);
    nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
    nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
    nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
    if (live && (threadIdx.x & 7) == 0 && bits != nullptr) bits[i >> 3] = nib;
  }
}
__global__ void relu_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                long n4, int accumulate, float4* dx) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 g = __ldg(dy + i), v = __ldg(x + i);
    float4 o = make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f,
                           v.w > 0.f ? g.w : 0.f);
    if (accumulate) {
      const float4 p = dx[i];
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
    dx[i] = o;  // gradient buffers are dgrad MMA operands: keep them tf32-exact
  }
}
__global__ void add_inplace_kernel(float4* dst, const float4* __restrict__ src, long n4) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 a = dst[i];
    const float4 b = __ldg(src + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
}

// ------------------------------------------------------------------------------------------
// reductions: fixed grid of partials + a single-block finisher => reproducible sums
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float red[8];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256)
finish_sum_kernel(const float* __restrict__ partials, int n, float scale, float* out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];
  const float t = block_sum_256(acc);
  if (threadIdx.x == 0) *out = t * scale;
}

__global__ void __launch_bounds__(256)
sqdiff_partial_kernel(const float* __restrict__ f, const float* __restrict__ t, long n,
                      float* __restrict__ partials) {
  float acc = 0.f;
  const long n4 = n >> 2;
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * 256) {
    const float4 a = __ldg(f4 + i), b = __ldg(t4 + i);
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
    acc += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
      const float d = f[i] - t[i];
      acc += d * d;
    }
  const float s = block_sum_256(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
content_bwd_kernel(const float* __restrict__ f, const float* __restrict__ t, long n,
                   const float* __restrict__ grad_w, float scale, int accumulate, float* df) {
  const float k = (grad_w ? __ldg(grad_w) : 1.f) * scale;
  const long n4 = n >> 2;
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4* d4 = reinterpret_cast<float4*>(df);
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * 256) {
    const float4 a = __ldg(f4 + i), b = __ldg(t4 + i);
    float4 o = make_float4(k * (a.x - b.x), k * (a.y - b.y), k * (a.z - b.z), k * (a.w - b.w));
    if (accumulate) {
      const float4 p = d4[i];
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
    d4[i] = o;
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
      const float o = k * (f[i] - t[i]);
      df[i] = round_tf32(accumulate ? df[i] + o : o);
    }
}

__global__ void __launch_bounds__(256)
dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long n,
                   float* __restrict__ partials) {
  float acc = 0.f;
  const long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * 256) {
    const float4 u = __ldg(a4 + i), v = __ldg(b4 + i);
    acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += a[i] * b[i];
  const float s = block_sum_256(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// partials layout: [0..B) running max |a|, [B..2B) sum |a|
__global__ void __launch_bounds__(256)
absstat_partial_kernel(const float* __restrict__ a, long n, float* __restrict__ partials) {
  float mx = 0.f, sm = 0.f;
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * 256) {
    const float v = fabsf(a[i]);
    mx = (v > mx || v != v) ? v : mx;
    sm += v;
  }
  __shared__ float redm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (other > mx || other != other) ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) redm[threadIdx.x >> 5] = mx;
  const float s = block_sum_256(sm);  // contains the __syncthreads that publishes redm
  if (threadIdx.x == 0) {
    float m = redm[0];
    for (int w = 1; w < 8; ++w) m = (redm[w] > m || redm[w] != redm[w]) ? redm[w] : m;
    partials[blockIdx.x] = m;
    partials[gridDim.x + blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256)
absstat_finish_kernel(const float* __restrict__ partials, int nb, float* out2) {
  float mx = 0.f, sm = 0.f;
  for (int i = threadIdx.x; i < nb; i += 256) {
    const float v = partials[i];
    mx = (v > mx || v != v) ? v : mx;
    sm += partials[nb + i];
  }
  __shared__ float redm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = (other > mx || other != other) ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) redm[threadIdx.x >> 5] = mx;
  const float s = block_sum_256(sm);
  if (threadIdx.x == 0) {
    float m = redm[0];
    for (int w = 1; w < 8; ++w) m = (redm[w] > m || redm[w] != redm[w]) ? redm[w] : m;
    out2[0] = m;
    out2[1] = s;
  }
}

// ------------------------------------------------------------------------------------------
// optimiser vector ops
// ------------------------------------------------------------------------------------------
// Adam, same operation order as torch's single-tensor path (optim/adam.py): lerp on m,
// mul+addcmul on v, denom = sqrt(v)/sqrt(bc2) + eps, x += -step_size * (m / denom).
__device__ __forceinline__ void adam_elem(float& x, float g, float& m, float& v, float w1, float w2,
                                          float beta2, float eps, float step_size,
                                          float bias2_sqrt) {
  m = fmaf(w1, g - m, m);
  v = __fmul_rn(v, beta2);
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(w2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bias2_sqrt), eps);
  x = __fadd_rn(x, __fmul_rn(-step_size, __fdiv_rn(m, denom)));
}
// 16-byte accesses: 4 loads + 3 stores of 16 B per thread and iteration, two iterations in flight
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, long n, float beta1, float beta2, float eps,
                 float step_size, float bias2_sqrt, const float* __restrict__ dev_scalars) {
  if (dev_scalars) {
    step_size = __ldg(dev_scalars);
    bias2_sqrt = __ldg(dev_scalars + 1);
  }
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  const long n4 = n >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const long stride = static_cast<long>(gridDim.x) * 256;
#pragma unroll 2
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += stride) {
    const float4 gi = __ldg(g4 + i);
    float4 xi = x4[i], mi = m4[i], vi = v4[i];
    adam_elem(xi.x, gi.x, mi.x, vi.x, w1, w2, beta2, eps, step_size, bias2_sqrt);
    adam_elem(xi.y, gi.y, mi.y, vi.y, w1, w2, beta2, eps, step_size, bias2_sqrt);
    adam_elem(xi.z, gi.z, mi.z, vi.z, w1, w2, beta2, eps, step_size, bias2_sqrt);
    adam_elem(xi.w, gi.w, mi.w, vi.w, w1, w2, beta2, eps, step_size, bias2_sqrt);
    x4[i] = xi;
    m4[i] = mi;
    v4[i] = vi;
  }
  if (blockIdx.x == 0)
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += 256)
      adam_elem(x[i], g[i], m[i], v[i], w1, w2, beta2, eps, step_size, bias2_sqrt);
}

// state[0] = step count (as float, exact below 2^24); writes {step_size, sqrt(bias_correction2)}.
__global__ void adam_scalars_kernel(float* state, float lr, float beta1, float beta2) {
  const double t = static_cast<double>(state[0]) + 1.0;
  state[0] = static_cast<float>(t);
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), t);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), t);
  state[1] = static_cast<float>(static_cast<double>(lr) / bc1);
  state[2] = static_cast<float>(sqrt(bc2));
}

__global__ void __launch_bounds__(256)
axpy_kernel(const float* __restrict__ alpha_dev, float alpha, const float* __restrict__ x, float* y,
            long n) {
  if (alpha_dev) alpha = __ldg(alpha_dev);
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * 256)
    y[i] = fmaf(alpha, x[i], y[i]);
}
__global__ void __launch_bounds__(256)
scale_kernel(const float* __restrict__ alpha_dev, float alpha, const float* __restrict__ x,
             float* y, long n) {
  if (alpha_dev) alpha = __ldg(alpha_dev);
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * 256)
    y[i] = alpha * x[i];
}

// ------------------------------------------------------------------------------------------
// timelapse frame: denormalise -> nan_to_num(0, 1, 0) -> clamp[0,1] -> *255 -> u8, NCHW -> HWC.
// Separate mul/add roundings (no FMA contraction) so the byte output matches the reference's
// torch/numpy arithmetic exactly.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned frame_px(float v, float mean, float stdv, int denorm,
                                             int rounding) {
  if (denorm) v = __fadd_rn(__fmul_rn(v, stdv), mean);
  if (v != v) v = 0.f;                     // nan -> 0
  else if (isinf(v)) v = v > 0.f ? 1.f : 0.f;  // +inf -> 1, -inf -> 0
  v = fminf(fmaxf(v, 0.f), 1.f);
  const float s = __fmul_rn(v, 255.f);
  return rounding ? static_cast<unsigned>(rintf(s)) : static_cast<unsigned>(s);
}
__global__ void __launch_bounds__(256)
frame_to_u8_kernel(const float* __restrict__ img, long hw, int denorm, int rounding,
                   unsigned char* __restrict__ out) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const long groups = (hw + 3) >> 2;
  for (long gidx = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; gidx < groups;
       gidx += static_cast<long>(gridDim.x) * 256) {
    const long p0 = gidx << 2;
    if (p0 + 3 < hw && (hw & 3) == 0) {
      float4 c[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) c[ch] = __ldg(reinterpret_cast<const float4*>(img + ch * hw + p0));
      unsigned b[12];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        b[0 * 3 + ch] = frame_px(c[ch].x, mean[ch], stdv[ch], denorm, rounding);
        b[1 * 3 + ch] = frame_px(c[ch].y, mean[ch], stdv[ch], denorm, rounding);
        b[2 * 3 + ch] = frame_px(c[ch].z, mean[ch], stdv[ch], denorm, rounding);
        b[3 * 3 + ch] = frame_px(c[ch].w, mean[ch], stdv[ch], denorm, rounding);
      }
      uint3 wv;
      wv.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
      wv.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
      wv.z = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
      unsigned* o = reinterpret_cast<unsigned*>(out + p0 * 3);
      o[0] = wv.x; o[1] = wv.y; o[2] = wv.z;
    } else {
      for (long p = p0; p < hw && p < p0 + 4; ++p)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          out[p * 3 + ch] = static_cast<unsigned char>(
              frame_px(img[ch * hw + p], mean[ch], stdv[ch], denorm, rounding));
    }
  }
}

// ------------------------------------------------------------------------------------------
// image load: HWC uint8 -> NCHW fp32, ToTensor (x / 255) then optional Normalize ((x - mean) / std),
// the same IEEE operations in the same order as torchvision, so the result is bit-identical.
// One thread = one pixel (3 bytes in, 3 planes out).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_from_u8_kernel(const unsigned char* __restrict__ hwc, long hw, int normalize,
                     float* __restrict__ out) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (long p = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; p < hw;
       p += static_cast<long>(gridDim.x) * 256) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = __fdiv_rn(static_cast<float>(hwc[p * 3 + ch]), 255.0f);
      if (normalize) v = __fdiv_rn(__fsub_rn(v, mean[ch]), stdv[ch]);
      out[ch * hw + p] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// layout conversion (API boundary: the reference exposes NCHW tensors) and finiteness flags
// ------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, long hw,
                                    float* __restrict__ dst) {
  const long total = hw * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long p = i / C;
    dst[i] = src[c * hw + p];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int C, long hw,
                                    float* __restrict__ dst) {
  const long total = hw * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = i % hw;
    const int c = static_cast<int>(i / hw);
    dst[i] = src[p * C + c];
  }
}
__global__ void finite_flags_kernel(const float* __restrict__ vals, int n, int* flags) {
  const int i = threadIdx.x;
  if (i < n && !isfinite(vals[i])) atomicOr(flags + i, 1);
}

// End of one optimisation step, inside the captured graph (reference optimization.py:298-312 weighted
// total, :375-391 finiteness checks, :402-422 loss recording): one thread reduces the per-layer
// losses to {style score, content score, total}, appends the row to a device ring indexed by a
// DEVICE step counter and records the three non-finite flags of that row.  The host reads rows at
// its logging cadence only; nothing here needs a per-step host value.
__global__ void step_scores_kernel(const float* __restrict__ losses, int n_style, int n_content,
                                   float style_w, float content_w, float* __restrict__ scores3,
                                   float* __restrict__ loss_ring, int* __restrict__ finite_ring,
                                   int capacity, int* __restrict__ counter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // same left-to-right order as torch.stack(losses).sum() on a handful of elements
  float st = 0.f, ct = 0.f;
  for (int i = 0; i < n_style; ++i) st += losses[i];
  for (int i = 0; i < n_content; ++i) ct += losses[n_style + i];
  const float total = __fadd_rn(__fmul_rn(style_w, st), __fmul_rn(content_w, ct));
  scores3[0] = st;
  scores3[1] = ct;
  scores3[2] = total;
  if (counter != nullptr) {
    const int step = *counter;
    const int row = step % capacity;
    if (loss_ring != nullptr) {
      loss_ring[3 * row + 0] = st;
      loss_ring[3 * row + 1] = ct;
      loss_ring[3 * row + 2] = total;
    }
    if (finite_ring != nullptr)
      finite_ring[row] = (isfinite(st) ? 0 : 1) | (isfinite(ct) ? 0 : 2) | (isfinite(total) ? 0 : 4);
    *counter = step + 1;
  }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#define STV_LAUNCH_CHECK() STV_CHECK_CUDA(cudaGetLastError())

int pack_conv_weights_launch(const float* w, float* w_fwd, float* w_dgrad, int Cout, int Cin,
                             cudaStream_t stream) {
  const long total = static_cast<long>(Cout) * Cin * 9;
  pack_conv_weights_kernel<<<grid_for(total, 256, stream_blocks()), 256, 0, stream>>>(
      w, w_fwd, w_dgrad, Cout, Cin);
  STV_LAUNCH_CHECK();
  return 0;
}

int maxpool2_fwd_launch(const float* x, int H, int W, int C, float* y, cudaStream_t stream) {
  STV_REQUIRE(C % 4 == 0, "maxpool2: C %d must be a multiple of 4", C);
  STV_REQUIRE(H >= 2 && W >= 2, "maxpool2: input %dx%d too small", H, W);
  const long total = static_cast<long>(H / 2) * (W / 2) * (C / 4);
  maxpool2_fwd_kernel<<<grid_for(total, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), H, W, C / 4, reinterpret_cast<float4*>(y));
  STV_LAUNCH_CHECK();
  return 0;
}

int maxpool2_bwd_launch(const float* dy, const float* x, int H, int W, int C, int relu_mask,
                        float* dx, cudaStream_t stream) {
  STV_REQUIRE(C % 4 == 0, "maxpool2_bwd: C %d must be a multiple of 4", C);
  const long total = static_cast<long>((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
  maxpool2_bwd_kernel<<<grid_for(total, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), H, W, C / 4,
      relu_mask, reinterpret_cast<float4*>(dx));
  STV_LAUNCH_CHECK();
  return 0;
}

int relu_fwd_launch(const float* x, long n, float* y, cudaStream_t stream) {
  STV_REQUIRE(n % 4 == 0, "relu: n must be a multiple of 4");
  relu_fwd_kernel<<<grid_for(n / 4, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<float4*>(y));
  STV_LAUNCH_CHECK();
  return 0;
}
int relu_fwd_bits_launch(const float* x, long n, float* y, unsigned* bits, cudaStream_t stream) {
  STV_REQUIRE(n % 32 == 0, "relu_fwd_bits: n must be a multiple of 32 (whole sign-bit words)");
  relu_fwd_bits_kernel<<<grid_for(n / 4, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), n / 4, reinterpret_cast<float4*>(y), bits);
  STV_LAUNCH_CHECK();
  return 0;
}
int relu_bwd_launch(const float* dy, const float* x, long n, int accumulate, float* dx,
                    cudaStream_t stream) {
  STV_REQUIRE(n % 4 == 0, "relu_bwd: n must be a multiple of 4");
  relu_bwd_kernel<<<grid_for(n / 4, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), n / 4, accumulate,
      reinterpret_cast<float4*>(dx));
  STV_LAUNCH_CHECK();
  return 0;
}
int add_inplace_launch(float* dst, const float* src, long n, cudaStream_t stream) {
  STV_REQUIRE(n % 4 == 0, "add_inplace: n must be a multiple of 4");
  add_inplace_kernel<<<grid_for(n / 4, 256, stream_blocks()), 256, 0, stream>>>(
      reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n / 4);
  STV_LAUNCH_CHECK();
  return 0;
}

// partials must hold kReduceBlocks floats (2x for absmax_sum).
int content_fwd_launch(const float* f, const float* t, long n, float* partials, float* loss_out,
                       cudaStream_t stream) {
  sqdiff_partial_kernel<<<kReduceBlocks, 256, 0, stream>>>(f, t, n, partials);
  STV_LAUNCH_CHECK();
  finish_sum_kernel<<<1, 256, 0, stream>>>(partials, kReduceBlocks,
                                           static_cast<float>(1.0 / static_cast<double>(n)),
                                           loss_out);
  STV_LAUNCH_CHECK();
  return 0;
}
int content_bwd_launch(const float* f, const float* t, long n, const float* grad_w, int accumulate,
                       float* df, cudaStream_t stream) {
  content_bwd_kernel<<<grid_for(n / 4 + 1, 256, stream_blocks()), 256, 0, stream>>>(
      f, t, n, grad_w, static_cast<float>(2.0 / static_cast<double>(n)), accumulate, df);
  STV_LAUNCH_CHECK();
  return 0;
}
int dot_launch(const float* a, const float* b, long n, float* partials, float* out,
               cudaStream_t stream) {
  dot_partial_kernel<<<kReduceBlocks, 256, 0, stream>>>(a, b, n, partials);
  STV_LAUNCH_CHECK();
  finish_sum_kernel<<<1, 256, 0, stream>>>(partials, kReduceBlocks, 1.f, out);
  STV_LAUNCH_CHECK();
  return 0;
}
int absmax_sum_launch(const float* a, long n, float* partials, float* out2, cudaStream_t stream) {
  absstat_partial_kernel<<<kReduceBlocks, 256, 0, stream>>>(a, n, partials);
  STV_LAUNCH_CHECK();
  absstat_finish_kernel<<<1, 256, 0, stream>>>(partials, kReduceBlocks, out2);
  STV_LAUNCH_CHECK();
  return 0;
}

int adam_step_launch(float* x, const float* g, float* m, float* v, long n, float beta1, float beta2,
                     float eps, float step_size, float bias2_sqrt, cudaStream_t stream) {
  adam_step_kernel<<<grid_for(n, 256, stream_blocks()), 256, 0, stream>>>(
      x, g, m, v, n, beta1, beta2, eps, step_size, bias2_sqrt, nullptr);
  STV_LAUNCH_CHECK();
  return 0;
}
// Graph-replayable variant: the step counter and the two step-dependent scalars live in
// state[0..2] on the device and advance by one per call.
int adam_step_dev_launch(float* x, const float* g, float* m, float* v, long n, float lr, float beta1,
                         float beta2, float eps, float* state, cudaStream_t stream) {
  adam_scalars_kernel<<<1, 1, 0, stream>>>(state, lr, beta1, beta2);
  STV_LAUNCH_CHECK();
  adam_step_kernel<<<grid_for(n, 256, stream_blocks()), 256, 0, stream>>>(
      x, g, m, v, n, beta1, beta2, eps, 0.f, 1.f, state + 1);
  STV_LAUNCH_CHECK();
  return 0;
}
int axpy_launch(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
                cudaStream_t stream) {
  axpy_kernel<<<grid_for(n, 256, stream_blocks()), 256, 0, stream>>>(alpha_dev, alpha_host, x, y, n);
  STV_LAUNCH_CHECK();
  return 0;
}
int scale_launch(const float* alpha_dev, float alpha_host, const float* x, float* y, long n,
                 cudaStream_t stream) {
  scale_kernel<<<grid_for(n, 256, stream_blocks()), 256, 0, stream>>>(alpha_dev, alpha_host, x, y,
                                                                      n);
  STV_LAUNCH_CHECK();
  return 0;
}

int frame_to_u8_launch(const float* img_nchw, int H, int W, int denormalize, int rounding,
                       unsigned char* out_hwc, cudaStream_t stream) {
  const long hw = static_cast<long>(H) * W;
  frame_to_u8_kernel<<<grid_for((hw + 3) / 4, 256, stream_blocks()), 256, 0, stream>>>(
      img_nchw, hw, denormalize, rounding, out_hwc);
  STV_LAUNCH_CHECK();
  return 0;
}
int image_from_u8_launch(const unsigned char* hwc, int H, int W, int normalize, float* out_nchw,
                         cudaStream_t stream) {
  const long hw = static_cast<long>(H) * W;
  image_from_u8_kernel<<<grid_for(hw, 256, stream_blocks()), 256, 0, stream>>>(hwc, hw, normalize,
                                                                               out_nchw);
  STV_LAUNCH_CHECK();
  return 0;
}
int nchw_to_nhwc_launch(const float* src, int C, int H, int W, float* dst, cudaStream_t stream) {
  const long hw = static_cast<long>(H) * W;
  nchw_to_nhwc_kernel<<<grid_for(hw * C, 256, stream_blocks()), 256, 0, stream>>>(src, C, hw, dst);
  STV_LAUNCH_CHECK();
  return 0;
}
int nhwc_to_nchw_launch(const float* src, int C, int H, int W, float* dst, cudaStream_t stream) {
  const long hw = static_cast<long>(H) * W;
  nhwc_to_nchw_kernel<<<grid_for(hw * C, 256, stream_blocks()), 256, 0, stream>>>(src, C, hw, dst);
  STV_LAUNCH_CHECK();
  return 0;
}
int finite_flags_launch(const float* vals, int n, int* flags, cudaStream_t stream) {
  STV_REQUIRE(n <= 32, "finite_flags: at most 32 values");
  finite_flags_kernel<<<1, 32, 0, stream>>>(vals, n, flags);
  STV_LAUNCH_CHECK();
  return 0;
}
int step_scores_launch(const float* losses, int n_style, int n_content, float style_w,
                       float content_w, float* scores3, float* loss_ring, int* finite_ring,
                       int capacity, int* counter, cudaStream_t stream) {
  STV_REQUIRE(losses && scores3 && n_style >= 0 && n_content >= 0, "step_scores: bad arguments");
  STV_REQUIRE(counter == nullptr || capacity > 0, "step_scores: ring capacity must be positive");
  step_scores_kernel<<<1, 32, 0, stream>>>(losses, n_style, n_content, style_w, content_w, scores3,
                                           loss_ring, finite_ring, capacity, counter);
  STV_LAUNCH_CHECK();
  return 0;
}

}  // namespace stv
