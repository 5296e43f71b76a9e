// Device-resident L-BFGS step (torch.optim.LBFGS semantics for max_iter = 1, no line search: the
// reference's default optimiser, core_model.py:344-349 / optimization.py:212-217).
//
// torch's implementation runs the two-loop recursion as ~2m dependent dot/axpy pairs, each followed
// by a host sync (`.item()`-style scalar reads), and streams the 2m history vectors ~5 times per
// step.  Here the recursion is done in COEFFICIENT space: the direction is
//     d = cg * g + sum_j cy_j * y_j + sum_j cs_j * s_j
// and cg, cy, cs follow from the small Gram matrices SY_ij = s_i.y_j, YY_ij = y_i.y_j and the
// vectors s_i.g, y_i.g (the same algebra as the two loops, reordered).  One step is three kernels
// and two passes over the history, with every branch of the algorithm (tolerance_grad early
// return, curvature test y.s > 1e-10, first-iteration step scaling, descent test g.d > -tol)
// decided on the device -- no host synchronisation, CUDA-graph capturable:
//   1. lbfgs_dots_kernel   : y_new = g - g_prev, s_new = t*d written to the spare history slot;
//                            per-block partial sums of {s_p.g, y_p.g, s_p.y_new, y_p.y_new,
//                            y_p.s_new} for every stored pair p and of the new pair's own dots
//   2. lbfgs_solve_kernel  : one block; fixed-order reduction of the partials (reproducible),
//                            Gram update, acceptance test, the two recursions on m x m data (fp64)
//   3. lbfgs_apply_kernel  : d, x += t*d, g_prev = g in one pass over the history
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

constexpr int kLbfgsChunk = 3072;   // elements per block in the dots pass (3 x 12 KB in smem)
constexpr int kLbfgsThreads = 256;
constexpr int kDotsPerSlot = 5;
constexpr int kExtraDots = 8;       // ys, yy, s_new.g, y_new.g, g.g, (3 spare)

// integer state: [0] n_iter  [1] count  [2] head  [3] do_update  [4] skip_all  [5] accepted
// float state  : [0] t  [1] H_diag  [2] coef_g  [3] gtd  [4] max|g|  [5] sum|g|
struct LbfgsArgs {
  float* x;
  const float* g;
  long n;
  long stride;           // history row stride in floats (n rounded up to a multiple of 4)
  int m;                 // history size; m + 1 physical slots
  float* hist_s;         // [(m + 1)][n]
  float* hist_y;         // [(m + 1)][n]
  float* prev_g;         // [n]
  float* d;              // [n]
  int* state_i;
  float* state_f;
  double* gram;          // SY [(m+1)^2] then YY [(m+1)^2]
  double* sg;            // [(m + 1)] s_p.g   (and yg right after: [(m + 1)])
  float* coef;           // cy [(m + 1)] then cs [(m + 1)]
  float* partials;       // [blocks][kDotsPerSlot * (m + 1) + kExtraDots]
  int blocks;
  float lr, tol_grad, tol_change;
};

__device__ __forceinline__ float block_reduce_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  return t;
}

__global__ void __launch_bounds__(kLbfgsThreads) lbfgs_dots_kernel(const LbfgsArgs a) {
  __shared__ float sh_g[kLbfgsChunk];
  __shared__ float sh_y[kLbfgsChunk];
  __shared__ float sh_s[kLbfgsChunk];
  __shared__ float red[8];
  const int n_iter = a.state_i[0], count = a.state_i[1], head = a.state_i[2];
  const int slots = a.m + 1;
  const int spare = (head + count) % slots;
  const float t = a.state_f[0];
  const long base = static_cast<long>(blockIdx.x) * kLbfgsChunk;
  const int len = static_cast<int>(min(static_cast<long>(kLbfgsChunk), a.n - base));
  const bool have_prev = n_iter >= 1;  // a previous direction / gradient exists
  float* out = a.partials + static_cast<size_t>(blockIdx.x) * (kDotsPerSlot * slots + kExtraDots);

  float ys = 0.f, yy = 0.f, sng = 0.f, yng = 0.f, gg = 0.f;
  for (int i = threadIdx.x; i < len; i += kLbfgsThreads) {
    const float gi = a.g[base + i];
    float yn = 0.f, sn = 0.f;
    if (have_prev) {
      yn = gi - a.prev_g[base + i];
      sn = t * a.d[base + i];
      a.hist_y[static_cast<size_t>(spare) * a.stride + base + i] = yn;
      a.hist_s[static_cast<size_t>(spare) * a.stride + base + i] = sn;
    }
    sh_g[i] = gi; sh_y[i] = yn; sh_s[i] = sn;
    ys += yn * sn; yy += yn * yn; sng += sn * gi; yng += yn * gi; gg += gi * gi;
  }
  __syncthreads();
  float r;
  r = block_reduce_256(ys, red);  if (threadIdx.x == 0) out[kDotsPerSlot * slots + 0] = r;
  r = block_reduce_256(yy, red);  if (threadIdx.x == 0) out[kDotsPerSlot * slots + 1] = r;
  r = block_reduce_256(sng, red); if (threadIdx.x == 0) out[kDotsPerSlot * slots + 2] = r;
  r = block_reduce_256(yng, red); if (threadIdx.x == 0) out[kDotsPerSlot * slots + 3] = r;
  r = block_reduce_256(gg, red);  if (threadIdx.x == 0) out[kDotsPerSlot * slots + 4] = r;

  // stored pairs: one warp per pair, so the slot loop needs no block-wide barrier
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < count; j += kLbfgsThreads / 32) {
    const int p = (head + j) % slots;
    const float* sp = a.hist_s + static_cast<size_t>(p) * a.stride + base;
    const float* yp = a.hist_y + static_cast<size_t>(p) * a.stride + base;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
    for (int i = lane; i < len; i += 32) {
      const float s = sp[i], y = yp[i];
      d0 += s * sh_g[i];   // s_p . g
      d1 += y * sh_g[i];   // y_p . g
      d2 += s * sh_y[i];   // s_p . y_new
      d3 += y * sh_y[i];   // y_p . y_new
      d4 += y * sh_s[i];   // y_p . s_new
    }
    d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3); d4 = warp_sum(d4);
    if (lane == 0) {
      out[kDotsPerSlot * p + 0] = d0; out[kDotsPerSlot * p + 1] = d1;
      out[kDotsPerSlot * p + 2] = d2; out[kDotsPerSlot * p + 3] = d3;
      out[kDotsPerSlot * p + 4] = d4;
    }
  }
}

// dynamic smem: double red[cols] + double alpha[m+1] + double beta[m+1]
__global__ void __launch_bounds__(kLbfgsThreads) lbfgs_solve_kernel(const LbfgsArgs a) {
  extern __shared__ double sm[];
  const int slots = a.m + 1;
  const int cols = kDotsPerSlot * slots + kExtraDots;
  double* dots = sm;                 // [cols]
  double* alpha = sm + cols;         // [slots]
  double* beta = alpha + slots;      // [slots]
  const int n_iter0 = a.state_i[0];
  int count = a.state_i[1], head = a.state_i[2];

  // fixed-order reduction over blocks: thread c owns column c
  for (int c = threadIdx.x; c < cols; c += kLbfgsThreads) {
    const bool slot_col = c < kDotsPerSlot * slots;
    bool valid = true;
    if (slot_col) {
      const int p = c / kDotsPerSlot;
      const int rel = (p - head + slots) % slots;
      valid = rel < count;
    }
    double acc = 0.0;
    if (valid)
      for (int b = 0; b < a.blocks; ++b) acc += a.partials[static_cast<size_t>(b) * cols + c];
    dots[c] = acc;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;  // warp 0 runs the (small, sequential) algebra, lane-parallel sums
  const int lane = threadIdx.x;

  const float gmax = a.state_f[4], gl1 = a.state_f[5];
  if (gmax <= a.tol_grad) {  // torch: opt_cond -> return before touching any state (a NaN gradient
                             // is NOT <= tol: like torch.optim.LBFGS the step then proceeds)
    if (lane == 0) {
      a.state_i[4] = 1;
      a.state_i[3] = 0;
    }
    return;
  }
  const int n_iter = n_iter0 + 1;
  double* SY = a.gram;
  double* YY = a.gram + static_cast<size_t>(slots) * slots;
  double* sg = a.sg;
  double* yg = a.sg + slots;
  float* cy = a.coef;
  float* cs = a.coef + slots;
  for (int p = lane; p < slots; p += 32) { cy[p] = 0.f; cs[p] = 0.f; }
  __syncwarp();
  const double gg = dots[kDotsPerSlot * slots + 4];
  double H = a.state_f[1];
  double coef_g;
  int accepted = 0;
  if (n_iter == 1) {
    count = 0; head = 0; H = 1.0;
    coef_g = -1.0;
  } else {
    const double ys = dots[kDotsPerSlot * slots + 0];
    const double yy = dots[kDotsPerSlot * slots + 1];
    for (int j = lane; j < count; j += 32) {
      const int p = (head + j) % slots;
      sg[p] = dots[kDotsPerSlot * p + 0];
      yg[p] = dots[kDotsPerSlot * p + 1];
    }
    if (ys > 1e-10) {
      accepted = 1;
      const int nw = (head + count) % slots;
      for (int j = lane; j < count; j += 32) {
        const int p = (head + j) % slots;
        SY[static_cast<size_t>(p) * slots + nw] = dots[kDotsPerSlot * p + 2];   // s_p . y_new
        YY[static_cast<size_t>(p) * slots + nw] = dots[kDotsPerSlot * p + 3];   // y_p . y_new
        YY[static_cast<size_t>(nw) * slots + p] = dots[kDotsPerSlot * p + 3];
        SY[static_cast<size_t>(nw) * slots + p] = dots[kDotsPerSlot * p + 4];   // s_new . y_p
      }
      if (lane == 0) {
        SY[static_cast<size_t>(nw) * slots + nw] = ys;
        YY[static_cast<size_t>(nw) * slots + nw] = yy;
        sg[nw] = dots[kDotsPerSlot * slots + 2];
        yg[nw] = dots[kDotsPerSlot * slots + 3];
      }
      if (count == a.m) head = (head + 1) % slots;  // drop the oldest pair
      else ++count;
      H = ys / yy;
    }
    __syncwarp();
    // first loop (newest -> oldest): alpha_i = rho_i * s_i . q,  q = -g - sum_{j>i} alpha_j y_j
    for (int i = count - 1; i >= 0; --i) {
      const int pi = (head + i) % slots;
      double part = 0.0;
      for (int j = i + 1 + lane; j < count; j += 32)
        part += alpha[j] * SY[static_cast<size_t>(pi) * slots + (head + j) % slots];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) alpha[i] = (-sg[pi] - part) / SY[static_cast<size_t>(pi) * slots + pi];
      __syncwarp();
    }
    // second loop (oldest -> newest): beta_i = rho_i * y_i . r,  r = H q + sum_{j<i} (a_j-b_j) s_j
    for (int i = 0; i < count; ++i) {
      const int pi = (head + i) % slots;
      double pq = 0.0, pr = 0.0;
      for (int j = lane; j < count; j += 32)
        pq += alpha[j] * YY[static_cast<size_t>(pi) * slots + (head + j) % slots];
      for (int j = lane; j < i; j += 32)
        pr += (alpha[j] - beta[j]) * SY[static_cast<size_t>((head + j) % slots) * slots + pi];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        pq += __shfl_xor_sync(0xffffffffu, pq, o);
        pr += __shfl_xor_sync(0xffffffffu, pr, o);
      }
      if (lane == 0)
        beta[i] = (H * (-yg[pi] - pq) + pr) / SY[static_cast<size_t>(pi) * slots + pi];
      __syncwarp();
    }
    coef_g = -H;
    for (int j = lane; j < count; j += 32) {
      const int p = (head + j) % slots;
      cy[p] = static_cast<float>(-H * alpha[j]);
      cs[p] = static_cast<float>(alpha[j] - beta[j]);
    }
    __syncwarp();
  }
  double gtd = 0.0;
  for (int j = lane; j < count; j += 32) {
    const int p = (head + j) % slots;
    gtd += static_cast<double>(cy[p]) * yg[p] + static_cast<double>(cs[p]) * sg[p];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gtd += __shfl_xor_sync(0xffffffffu, gtd, o);
  gtd += coef_g * gg;
  if (lane != 0) return;
  a.state_i[4] = 0;
  a.state_i[0] = n_iter;
  const float t = (n_iter == 1) ? fminf(1.f, 1.f / gl1) * a.lr : a.lr;
  a.state_i[1] = count;
  a.state_i[2] = head;
  a.state_i[3] = (gtd > -static_cast<double>(a.tol_change)) ? 0 : 1;
  a.state_i[5] = accepted;
  a.state_f[0] = t;
  a.state_f[1] = static_cast<float>(H);
  a.state_f[2] = static_cast<float>(coef_g);
  a.state_f[3] = static_cast<float>(gtd);
}

__global__ void __launch_bounds__(kLbfgsThreads) lbfgs_apply_kernel(const LbfgsArgs a) {
  if (a.state_i[4]) return;  // tolerance_grad early return: nothing changes
  const int count = a.state_i[1], head = a.state_i[2], do_update = a.state_i[3];
  const int slots = a.m + 1;
  const float t = a.state_f[0], cg = a.state_f[2];
  const float* cy = a.coef;
  const float* cs = a.coef + slots;
  const long n4 = a.n >> 2;
  for (long i = static_cast<long>(blockIdx.x) * kLbfgsThreads + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * kLbfgsThreads) {
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 acc = make_float4(cg * g.x, cg * g.y, cg * g.z, cg * g.w);
    for (int j = 0; j < count; ++j) {
      const int p = (head + j) % slots;
      const float4 y = reinterpret_cast<const float4*>(a.hist_y + static_cast<size_t>(p) * a.stride)[i];
      const float4 s = reinterpret_cast<const float4*>(a.hist_s + static_cast<size_t>(p) * a.stride)[i];
      const float wy = cy[p], wsv = cs[p];
      acc.x += wy * y.x + wsv * s.x; acc.y += wy * y.y + wsv * s.y;
      acc.z += wy * y.z + wsv * s.z; acc.w += wy * y.w + wsv * s.w;
    }
    reinterpret_cast<float4*>(a.d)[i] = acc;
    reinterpret_cast<float4*>(a.prev_g)[i] = g;
    if (do_update) {
      float4 xv = reinterpret_cast<float4*>(a.x)[i];
      xv.x += t * acc.x; xv.y += t * acc.y; xv.z += t * acc.z; xv.w += t * acc.w;
      reinterpret_cast<float4*>(a.x)[i] = xv;
    }
  }
  if (blockIdx.x == 0) {
    for (long i = (n4 << 2) + threadIdx.x; i < a.n; i += kLbfgsThreads) {
      const float g = a.g[i];
      float acc = cg * g;
      for (int j = 0; j < count; ++j) {
        const int p = (head + j) % slots;
        acc += cy[p] * a.hist_y[static_cast<size_t>(p) * a.stride + i] +
               cs[p] * a.hist_s[static_cast<size_t>(p) * a.stride + i];
      }
      a.d[i] = acc;
      a.prev_g[i] = g;
      if (do_update) a.x[i] += t * acc;
    }
  }
}

static int lbfgs_blocks(long n) { return static_cast<int>((n + kLbfgsChunk - 1) / kLbfgsChunk); }

// Scratch layout (floats): see lbfgs_workspace_floats.
size_t lbfgs_workspace_floats(long n, int m) {
  const size_t slots = static_cast<size_t>(m) + 1;
  const size_t cols = kDotsPerSlot * slots + kExtraDots;
  size_t f = 0;
  f += 16;                                   // state_i (as 32-bit words)
  f += 16;                                   // state_f
  f += 2 * (2 * slots * slots);              // gram (doubles)
  f += 2 * (2 * slots);                      // sg, yg (doubles)
  f += 2 * slots;                            // coef
  f += static_cast<size_t>(lbfgs_blocks(n)) * cols;  // partials
  f += 2 * 592 + 64;                         // |g| statistics scratch (kReduceBlocks x 2)
  return f + 64;
}

int lbfgs_step_launch(float* x, const float* g, long n, int m, float* hist_s, float* hist_y,
                      float* prev_g, float* d, float* workspace, float lr, float tol_grad,
                      float tol_change, cudaStream_t stream) {
  STV_REQUIRE(m >= 1 && m <= 256, "lbfgs: history size %d out of range", m);
  STV_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "lbfgs: workspace not 8-byte aligned");
  const size_t slots = static_cast<size_t>(m) + 1;
  const size_t cols = kDotsPerSlot * slots + kExtraDots;
  LbfgsArgs a;
  a.x = x; a.g = g; a.n = n; a.stride = (n + 3) & ~3L; a.m = m; a.hist_s = hist_s; a.hist_y = hist_y; a.prev_g = prev_g;
  a.d = d; a.lr = lr; a.tol_grad = tol_grad; a.tol_change = tol_change;
  float* w = workspace;
  a.state_i = reinterpret_cast<int*>(w); w += 16;
  a.state_f = w; w += 16;
  a.gram = reinterpret_cast<double*>(w); w += 2 * (2 * slots * slots);
  a.sg = reinterpret_cast<double*>(w); w += 2 * (2 * slots);
  a.coef = w; w += 2 * slots;
  a.blocks = lbfgs_blocks(n);
  a.partials = w; w += static_cast<size_t>(a.blocks) * cols;
  float* stat_scratch = w;

  // max|g|, sum|g| -> state_f[4..5]
  if (int rc = absmax_sum_launch(g, n, stat_scratch, a.state_f + 4, stream)) return rc;
  lbfgs_dots_kernel<<<a.blocks, kLbfgsThreads, 0, stream>>>(a);
  STV_CHECK_CUDA(cudaGetLastError());
  const size_t smem = (cols + 2 * slots) * sizeof(double);
  lbfgs_solve_kernel<<<1, kLbfgsThreads, smem, stream>>>(a);
  STV_CHECK_CUDA(cudaGetLastError());
  int apply_blocks = device_sm_count() * 8;
  const long want = (n / 4 + kLbfgsThreads - 1) / kLbfgsThreads;
  if (want < apply_blocks) apply_blocks = static_cast<int>(want < 1 ? 1 : want);
  lbfgs_apply_kernel<<<apply_blocks, kLbfgsThreads, 0, stream>>>(a);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stv
