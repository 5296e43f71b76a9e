// Gram matrix R = F * F^T as a symmetric, upper-triangle, split-K tensor-core contraction
// (tcgen05, TF32 in / FP32 accumulate in TMEM), with the clamp, 1/N normalisation, (G - A),
// the MSE style loss and the backward seed matrix S fused into the finalize pass.
//
// Reference semantics (core_model.py:29-63, :234-264):
//   F = x.reshape(b*c, h*w);  R = mm(F, F^T).clamp(max=5e5);  G = R / (b*c*h*w)
//   loss = mse_loss(G, A) = mean((G - A)^2)
// and, through autograd, dF = (4 / (C^2 N)) * (1[R <= clamp] .* (G - A)) * F  =: S * F.
//
// Activations are NHWC, i.e. X[pixel][channel] = F^T, so both MMA operands are "MN-major"
// (channel contiguous, pixel = K strided): the smem tile written by one TMA box
// {32 channels, 32 pixels} (TMA swizzle 128B with 32-byte atoms) is consumed directly through an
// MN-major SWIZZLE_128B_BASE32B descriptor -- the only swizzled layout valid for MN-major TF32.
// Work split: (upper-triangle 128x128 block pair) x (pixel range).  Each CTA accumulates its
// block over its pixel range in TMEM and writes one fp32 partial tile; gram_finalize_kernel sums
// the partials in a fixed order (deterministic), mirrors, and applies the fused epilogue.
// C == 64 is handled by viewing two consecutive pixels as one 128-channel row: the two diagonal
// 64x64 blocks of that 128x128 product are the even- and odd-pixel halves of R.
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

constexpr int kGramThreads = 192;
#ifndef STV_GRAM_PIX
#define STV_GRAM_PIX 64
#endif
// pixels (K) per pipeline stage: 64 = 8 MMAs per tcgen05.commit (with 32 the ~350-cycle commit stall
// of the issuing thread outweighed the 4 x 66 cycles of MMAs it followed: 4.2 TB/s, 36 % tensor pipe)
constexpr int kGramPix = STV_GRAM_PIX;
constexpr int kGramChunkBytes = kGramPix * 128;    // one {32ch x 32pix} box
constexpr int kGramOperandBytes = 4 * kGramChunkBytes;
constexpr int kGramStageBytes = 2 * kGramOperandBytes;
constexpr int kGramStages = kGramPix >= 64 ? 3 : 4;  // 3 x 64 KB or 4 x 32 KB of operands in flight
constexpr int kGramBarOffset = kGramStages * kGramStageBytes;
constexpr int kGramSmemTotal = kGramBarOffset + (2 * kGramStages + 1) * 8 + 16 + 1024;

struct GramParams {
  int pairs;             // number of upper-triangle block pairs
  int nb;                // 128-channel blocks per side
  int stages_total;      // ceil(rows / kGramPix)
  int stages_per_split;  // pipeline stages handled by one CTA
  float* partials;       // [splits][pairs][128*128]
};

__device__ __forceinline__ void pair_to_blocks(int pair, int nb, int* bi, int* bj) {
  int i = 0, rem = pair;
  while (rem >= nb - i) {
    rem -= nb - i;
    ++i;
  }
  *bi = i;
  *bj = i + rem;
}

__global__ void __launch_bounds__(kGramThreads)
gram_partial_kernel(const __grid_constant__ CUtensorMap tmap_x, const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kGramBarOffset;
  const uint32_t tmem_full_bar = bar_base + 16 * kGramStages;
  volatile uint32_t* tmem_ptr_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kGramBarOffset + 16 * kGramStages + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x;
  const int split = blockIdx.y;
  int bi, bj;
  pair_to_blocks(pair, p.nb, &bi, &bj);
  const bool diag = (bi == bj);
  const int st_begin = split * p.stages_per_split;
  const int st_end = min(st_begin + p.stages_per_split, p.stages_total);
  const int n_st = st_end - st_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < kGramStages; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (kGramStages + s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int st = st_begin; st < st_end; ++st) {
        mbar_wait(bar_base + 8 * (kGramStages + s), phase ^ 1);
        const uint32_t full = bar_base + 8 * s;
        const uint32_t a_dst = smem_base + s * kGramStageBytes;
        mbar_expect_tx(full, diag ? kGramOperandBytes : kGramStageBytes);
        const int row0 = st * kGramPix;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_2d(a_dst + c * kGramChunkBytes, &tmap_x, full, bi * 128 + c * 32, row0);
        if (!diag) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_2d(a_dst + kGramOperandBytes + c * kGramChunkBytes, &tmap_x, full,
                        bj * 128 + c * 32, row0);
        }
        if (++s == kGramStages) {
          s = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 128, 1, 1);
      int s = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_st; ++it) {
        mbar_wait(bar_base + 8 * s, phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kGramStageBytes;
        const uint32_t b_addr = diag ? a_addr : a_addr + kGramOperandBytes;
#pragma unroll
        for (int k = 0; k < kGramPix / 8; ++k) {
          // MN-major TF32 => SWIZZLE_128B_BASE32B: 32-channel groups are kGramChunkBytes apart
          // (LBO), 4-pixel K groups 512 B apart (SBO); one MMA consumes 8 pixels = 1024 B.
          const uint64_t adesc = make_smem_desc(a_addr + k * 1024, kGramChunkBytes, 512, 1);
          const uint64_t bdesc = make_smem_desc(b_addr + k * 1024, kGramChunkBytes, 512, 1);
          umma_tf32(tmem_d, adesc, bdesc, idesc, (it | k) != 0);
        }
        umma_commit(bar_base + 8 * (kGramStages + s));
        if (++s == kGramStages) {
          s = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* dst = p.partials + (static_cast<size_t>(split) * p.pairs + pair) * (128 * 128) +
                 static_cast<size_t>(m) * 128;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cb = 0; cb < 128; cb += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + cb, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v;
        v.x = __uint_as_float(r[4 * j + 0]);
        v.y = __uint_as_float(r[4 * j + 1]);
        v.z = __uint_as_float(r[4 * j + 2]);
        v.w = __uint_as_float(r[4 * j + 3]);
        *reinterpret_cast<float4*>(dst + cb + 4 * j) = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 128);
  }
}

struct GramFinalizeParams {
  const float* partials;
  int splits, pairs, nb;
  int C;                 // true channel count
  int merged;            // 1 when C == 64 (pixel pairs viewed as 128 channels)
  const float* x_tail;   // last pixel's 64 channels when hw is odd and merged, else null
  float n_total;         // C * hw (exact in fp32 for every supported size)
  float s_scale;         // 4 / (C^2 * C * hw)
  float clamp_max;
  const float* target;   // [C][C] or null
  float* gram_out;       // [C][C] or null
  float* s_out;          // [C][C] or null
  float* loss_partials;  // [gridDim.x]
  float* raw_out;        // [C][C] raw (un-clamped, un-normalised) R; when set nothing else is written
};

__global__ void __launch_bounds__(256) gram_finalize_kernel(const GramFinalizeParams p) {
  const int e = blockIdx.x * 256 + threadIdx.x;  // element within the concatenated pair tiles
  const int pair = e >> 14;
  const int m = (e >> 7) & 127;
  const int n = e & 127;
  float contrib = 0.f;
  bool active = pair < p.pairs;
  int bi = 0, bj = 0;
  if (active) pair_to_blocks(pair, p.nb, &bi, &bj);
  if (active && p.merged) active = (m < 64 && n < 64);
  if (active) {
    const size_t tile_stride = static_cast<size_t>(p.pairs) * 16384;
    const float* src = p.partials + static_cast<size_t>(pair) * 16384 + m * 128 + n;
    // four independent partial sums keep several loads in flight (the order is still fixed, so the
    // result is bit-reproducible run to run)
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
    const size_t off2 = p.merged ? 64 * 128 + 64 : 0;
    int sp = 0;
    for (; sp + 4 <= p.splits; sp += 4) {
      const float* q = src + sp * tile_stride;
      float a0 = q[0], a1 = q[tile_stride], a2 = q[2 * tile_stride], a3 = q[3 * tile_stride];
      if (p.merged) {
        a0 += q[off2]; a1 += q[tile_stride + off2];
        a2 += q[2 * tile_stride + off2]; a3 += q[3 * tile_stride + off2];
      }
      r0 += a0; r1 += a1; r2 += a2; r3 += a3;
    }
    for (; sp < p.splits; ++sp) {
      const float* q = src + sp * tile_stride;
      r0 += p.merged ? q[0] + q[off2] : q[0];
    }
    float r = (r0 + r1) + (r2 + r3);
    if (p.merged && p.x_tail) r += p.x_tail[m] * p.x_tail[n];
    const int i = bi * 128 + m;
    const int j = bj * 128 + n;
    const bool off_diag = (bi != bj);
    if (p.raw_out) {  // row-band sharding: R partials are summed across GPUs before the clamp
      p.raw_out[static_cast<size_t>(i) * p.C + j] = r;
      if (off_diag) p.raw_out[static_cast<size_t>(j) * p.C + i] = r;
    }
    const float g = fminf(r, p.clamp_max) / p.n_total;
    if (p.gram_out) {
      p.gram_out[static_cast<size_t>(i) * p.C + j] = g;
      if (off_diag) p.gram_out[static_cast<size_t>(j) * p.C + i] = g;
    }
    if (p.target) {
      const float d = g - p.target[static_cast<size_t>(i) * p.C + j];
      contrib = off_diag ? 2.f * d * d : d * d;
      if (p.s_out) {
        const float sv = (r <= p.clamp_max) ? round_tf32(d * p.s_scale) : 0.f;  // MMA operand
        p.s_out[static_cast<size_t>(i) * p.C + j] = sv;
        if (off_diag) p.s_out[static_cast<size_t>(j) * p.C + i] = sv;
      }
    }
  }
  __shared__ float red[8];
  float v = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    p.loss_partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
gram_loss_reduce_kernel(const float* partials, int n, float scale, float* loss_out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    *loss_out = t * scale;
  }
}

struct GramPlan {
  int Cv, nb, pairs, stages_total, stages_per_split, splits;
  long rows;
  int merged;
};

static GramPlan plan_gram(long hw, int C) {
  GramPlan g;
  g.merged = (C == 64);
  g.Cv = g.merged ? 128 : C;
  g.rows = g.merged ? hw / 2 : hw;
  g.nb = g.Cv / 128;
  g.pairs = g.nb * (g.nb + 1) / 2;
  g.stages_total = static_cast<int>((g.rows + kGramPix - 1) / kGramPix);
  if (g.stages_total < 1) g.stages_total = 1;
  int want = device_sm_count() / g.pairs;
  if (want < 1) want = 1;
  if (want > g.stages_total) want = g.stages_total;
  g.stages_per_split = (g.stages_total + want - 1) / want;
  g.splits = (g.stages_total + g.stages_per_split - 1) / g.stages_per_split;
  return g;
}

size_t gram_workspace_bytes(long hw, int C) {
  const GramPlan g = plan_gram(hw, C);
  const size_t tiles = static_cast<size_t>(g.splits) * g.pairs * 16384 * sizeof(float);
  const size_t loss = static_cast<size_t>(g.pairs) * 64 * sizeof(float);
  return tiles + loss + 256;
}

// Epilogue of the Gram pipeline applied to an already reduced raw R (C x C, both triangles):
// G = min(R, clamp)/N, loss partial sums and the backward seed S.  Used after the cross-GPU
// all-reduce of R in the row-band sharded path.
struct GramFromRParams {
  const float* r;
  int C;
  float n_total, s_scale, clamp_max;
  const float* target;
  float* gram_out;
  float* s_out;
  float* loss_partials;
};

__global__ void __launch_bounds__(256) gram_from_r_kernel(const GramFromRParams p) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  float contrib = 0.f;
  if (e < p.C * p.C) {
    const float r = p.r[e];
    const float g = fminf(r, p.clamp_max) / p.n_total;
    if (p.gram_out) p.gram_out[e] = g;
    if (p.target) {
      const float d = g - p.target[e];
      contrib = d * d;
      if (p.s_out) p.s_out[e] = (r <= p.clamp_max) ? round_tf32(d * p.s_scale) : 0.f;
    }
  }
  __shared__ float red[8];
  float v = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    p.loss_partials[blockIdx.x] = t;
  }
}

int gram_from_r_launch(const float* r, int C, double n_total, const float* target, float clamp_max,
                       float* gram_out, float* s_out, float* loss_out, float* scratch,
                       cudaStream_t stream) {
  STV_REQUIRE(!loss_out || target, "gram_from_r: loss requested without a target");
  GramFromRParams p;
  p.r = r; p.C = C;
  p.n_total = static_cast<float>(n_total);
  p.s_scale = static_cast<float>(4.0 / (static_cast<double>(C) * C * n_total));
  p.clamp_max = clamp_max; p.target = target; p.gram_out = gram_out; p.s_out = s_out;
  p.loss_partials = scratch;
  const int blocks = (C * C + 255) / 256;
  gram_from_r_kernel<<<blocks, 256, 0, stream>>>(p);
  STV_CHECK_CUDA(cudaGetLastError());
  if (loss_out) {
    gram_loss_reduce_kernel<<<1, 256, 0, stream>>>(
        scratch, blocks, static_cast<float>(1.0 / (static_cast<double>(C) * C)), loss_out);
    STV_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int gram_launch(const float* x, long hw, int C, float* workspace, size_t workspace_bytes,
                const float* target, float clamp_max, float* gram_out, float* s_out,
                float* loss_out, float* raw_out, cudaStream_t stream) {
  STV_REQUIRE(C == 64 || (C % 128 == 0 && C <= 1024), "gram: unsupported channel count %d", C);
  STV_REQUIRE(hw >= 1, "gram: empty feature map");
  STV_REQUIRE(workspace_bytes >= gram_workspace_bytes(hw, C), "gram: workspace too small");
  STV_REQUIRE(!loss_out || target, "gram: loss requested without a target");
  const GramPlan g = plan_gram(hw, C);

  static bool attr_set[kMaxDevices] = {};  // the shared-memory opt-in is per device
  int dev = 0;
  STV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  if (!attr_set[dev]) {
    STV_CHECK_CUDA(cudaFuncSetAttribute(gram_partial_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kGramSmemTotal));
    attr_set[dev] = true;
  }
  float* partials = workspace;
  float* loss_partials = workspace + static_cast<size_t>(g.splits) * g.pairs * 16384;

  if (g.rows >= 1) {
    CUtensorMap tx;
    const uint64_t dims[2] = {(uint64_t)g.Cv, (uint64_t)g.rows};
    const uint64_t strides[1] = {(uint64_t)g.Cv * 4};
    const uint32_t box[2] = {32, (uint32_t)kGramPix};
    if (int rc = encode_tmap_f32(&tx, x, 2, dims, strides, box, kSwizzle128BAtom32B)) return rc;
    GramParams p;
    p.pairs = g.pairs; p.nb = g.nb; p.stages_total = g.stages_total;
    p.stages_per_split = g.stages_per_split; p.partials = partials;
    gram_partial_kernel<<<dim3(g.pairs, g.splits), kGramThreads, kGramSmemTotal, stream>>>(tx, p);
    STV_CHECK_CUDA(cudaGetLastError());
  } else {
    // hw == 1 with C == 64: everything is in the tail pixel.
    STV_CHECK_CUDA(cudaMemsetAsync(partials, 0, static_cast<size_t>(g.splits) * g.pairs * 16384 *
                                                    sizeof(float), stream));
  }

  GramFinalizeParams f;
  f.partials = partials; f.splits = g.splits; f.pairs = g.pairs; f.nb = g.nb; f.C = C;
  f.merged = g.merged;
  f.x_tail = (g.merged && (hw & 1)) ? x + (hw - 1) * 64 : nullptr;
  const double n_total = static_cast<double>(C) * static_cast<double>(hw);
  f.n_total = static_cast<float>(n_total);
  f.s_scale = static_cast<float>(4.0 / (static_cast<double>(C) * C * n_total));
  f.clamp_max = clamp_max; f.target = target; f.gram_out = gram_out; f.s_out = s_out;
  f.loss_partials = loss_partials;
  f.raw_out = raw_out;
  const int fin_blocks = g.pairs * 64;
  gram_finalize_kernel<<<fin_blocks, 256, 0, stream>>>(f);
  STV_CHECK_CUDA(cudaGetLastError());
  if (loss_out) {
    gram_loss_reduce_kernel<<<1, 256, 0, stream>>>(
        loss_partials, fin_blocks, static_cast<float>(1.0 / (static_cast<double>(C) * C)),
        loss_out);
    STV_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace stv
