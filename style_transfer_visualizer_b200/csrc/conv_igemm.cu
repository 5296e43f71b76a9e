// Implicit-GEMM 3x3 (pad 1) / 1x1 convolution on the sm_100a tensor cores (tcgen05, TF32 in,
// FP32 accumulate in TMEM), NHWC activations.
//
// Replaces, on the reference's hot path, the nn.Conv2d forward calls issued by
// StyleContentModel.forward (core_model.py:316 -> torchvision vgg.py conv layers), their
// autograd dgrad (optimization.py:313 loss.backward(); weights are frozen, core_model.py:115-116,
// so no wgrad), and the torch.mm backward of gram_matrix (core_model.py:60) which is a 1x1 conv
// with the symmetric matrix S as the weight.
//
// GEMM view:  D[pixel, n] = sum_{tap, c} X[pixel + off(tap), c] * Wp[tap][n][c]
//   M tile  = 128 pixels arranged as a (th x tw) spatial patch  (TMEM lane = pixel)
//   N tile  = BLOCK_N output channels                           (TMEM column = channel)
//   K loop  = taps x (C/32); each step is a 32-channel slab = one 128-byte swizzled smem row.
// The A operand of one K step is ONE 3-D TMA box {32 ch, tw, th} at the tap-shifted coordinate;
// TMA out-of-bounds zero fill implements the conv padding, so there is no im2col buffer.
// The same kernel computes dgrad: the host passes weights re-packed as [flipped tap][Cin][Cout].
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias / scale / ReLU / ReLU-mask / add ->
// global).  Two CTAs can be resident per SM for BLOCK_N<=128 so that one CTA's epilogue overlaps
// the other's main loop.
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

struct ConvIgemmParams {
  int H, W;        // spatial size of input == output
  int C;           // input channels (GEMM K per tap), multiple of 32
  int N;           // output channels total
  int taps;        // 9 (3x3, pad 1) or 1 (1x1)
  int th, tw;      // spatial patch, th*tw == 128
  int tiles_x;     // ceil(W / tw)
  const float* bias;      // [N] or null
  const float* alpha;     // device scalar multiplier or null
  const float* mask_src;  // NHWC [H,W,N]: keep acc where mask_src > 0, else 0 (ReLU backward)
  const float* add_src;   // NHWC [H,W,N]: added after masking (gradient accumulation)
  float* out_pre;         // result before ReLU (or the only result); may be null
  float* out_post;        // max(result, 0); may be null
  int round_pre;          // store out_pre rounded to tf32 (it feeds another MMA as an operand)
  int round_post;         // store out_post rounded to tf32
};

constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kSlabBytes = 128;  // 32 fp32 channels
constexpr int kABytes = kTileM * kSlabBytes;

template <int BLOCK_N, int STAGES>
struct ConvSmem {
  static constexpr int kBBytes = BLOCK_N * kSlabBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;  // + align slack
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads)
conv_igemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x,
                       const __grid_constant__ CUtensorMap tmap_w, const ConvIgemmParams p) {
  using L = ConvSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barriers: full[s] at +8s, empty[s] at +8(STAGES+s), tmem_full at +16*STAGES, tmem ptr after
  const uint32_t tmem_full_bar = bar_base + 16 * STAGES;
  volatile uint32_t* tmem_ptr_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 16 * STAGES + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int ty0 = (tile / p.tiles_x) * p.th;
  const int tx0 = (tile % p.tiles_x) * p.tw;
  const int n0 = blockIdx.y * BLOCK_N;
  const int kc = p.C >> 5;
  const int k_iters = p.taps * kc;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_ptr_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int it = 0; it < k_iters; ++it) {
        const int tap = it / kc;
        const int c0 = (it - tap * kc) << 5;
        int dy = 0, dx = 0;
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        mbar_wait(bar_base + 8 * (STAGES + s), phase ^ 1);
        const uint32_t full = bar_base + 8 * s;
        const uint32_t a_dst = smem_base + s * L::kStageBytes;
        mbar_expect_tx(full, L::kStageBytes);
        tma_load_3d(a_dst, &tmap_x, full, c0, tx0 + dx, ty0 + dy);
        tma_load_2d(a_dst + kABytes, &tmap_w, full, c0, tap * p.N + n0);
        if (++s == STAGES) {
          s = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ---------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kTileM, BLOCK_N, 0, 0);
      int s = 0;
      uint32_t phase = 0;
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait(bar_base + 8 * s, phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStageBytes;
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_tf32(tmem_d, adesc, bdesc, idesc, (it | k) != 0);
        }
        umma_commit(bar_base + 8 * (STAGES + s));  // frees the smem stage when MMAs retire
        if (++s == STAGES) {
          s = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ------------------------------ epilogue ----------------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    const int py = ty0 + m / p.tw;
    const int px = tx0 + m % p.tw;
    const bool valid = (py < p.H) && (px < p.W);
    const size_t row_off = (static_cast<size_t>(py) * p.W + px) * p.N + n0;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cb = 0; cb < BLOCK_N; cb += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + cb, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v;
          v.x = __uint_as_float(r[4 * j + 0]) * alpha;
          v.y = __uint_as_float(r[4 * j + 1]) * alpha;
          v.z = __uint_as_float(r[4 * j + 2]) * alpha;
          v.w = __uint_as_float(r[4 * j + 3]) * alpha;
          const int col = cb + 4 * j;
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          if (p.mask_src) {
            const float4 mk = __ldg(reinterpret_cast<const float4*>(p.mask_src + row_off + col));
            v.x = mk.x > 0.f ? v.x : 0.f;
            v.y = mk.y > 0.f ? v.y : 0.f;
            v.z = mk.z > 0.f ? v.z : 0.f;
            v.w = mk.w > 0.f ? v.w : 0.f;
          }
          if (p.add_src) {
            const float4 a = *reinterpret_cast<const float4*>(p.add_src + row_off + col);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
          }
          if (p.out_pre) {
            float4 o = v;
            if (p.round_pre) {
              o.x = round_tf32(o.x); o.y = round_tf32(o.y);
              o.z = round_tf32(o.z); o.w = round_tf32(o.w);
            }
            *reinterpret_cast<float4*>(p.out_pre + row_off + col) = o;
          }
          if (p.out_post) {
            float4 o;
            o.x = fmaxf(v.x, 0.f); o.y = fmaxf(v.y, 0.f);
            o.z = fmaxf(v.z, 0.f); o.w = fmaxf(v.w, 0.f);
            if (p.round_post) {
              o.x = round_tf32(o.x); o.y = round_tf32(o.y);
              o.z = round_tf32(o.z); o.w = round_tf32(o.w);
            }
            *reinterpret_cast<float4*>(p.out_post + row_off + col) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BLOCK_N);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvIgemmParams& p,
                       int tiles, cudaStream_t stream) {
  using L = ConvSmem<BLOCK_N, STAGES>;
  auto kern = conv_igemm_tf32_kernel<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    STV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        L::kTotal));
    attr_set = true;
  }
  dim3 grid(tiles, p.N / BLOCK_N);
  kern<<<grid, kThreads, L::kTotal, stream>>>(tx, tw, p);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Pick the (th, tw) patch with the least padded area; ties prefer wider rows.
static void pick_patch(int H, int W, int* th, int* tw) {
  const int cand[5][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}, {1, 128}};
  long best = -1;
  for (int i = 0; i < 5; ++i) {
    const int a = cand[i][0], b = cand[i][1];
    const long area = static_cast<long>((H + a - 1) / a) * a * ((W + b - 1) / b) * b;
    if (best < 0 || area < best) {
      best = area;
      *th = a;
      *tw = b;
    }
  }
}

int conv_igemm_launch(const float* x, const float* w_packed, int H, int W, int C, int N, int taps,
                      const float* bias, const float* alpha, const float* mask_src,
                      const float* add_src, float* out_pre, float* out_post, int round_flags,
                      int block_n, int th, int tw, cudaStream_t stream) {
  STV_REQUIRE(C % 32 == 0 && C >= 32, "conv_igemm: input channels %d must be a multiple of 32", C);
  STV_REQUIRE(N % 64 == 0, "conv_igemm: output channels %d must be a multiple of 64", N);
  STV_REQUIRE(taps == 9 || taps == 1, "conv_igemm: taps must be 9 or 1 (got %d)", taps);
  STV_REQUIRE(H > 0 && W > 0, "conv_igemm: empty image");
  STV_REQUIRE(out_pre || out_post, "conv_igemm: no output buffer");
  if (th <= 0 || tw <= 0) pick_patch(H, W, &th, &tw);
  STV_REQUIRE(th * tw == kTileM && tw <= 256 && th <= 256, "conv_igemm: bad patch %dx%d", th, tw);

  const int tiles_x = (W + tw - 1) / tw;
  const int tiles_y = (H + th - 1) / th;
  const int tiles = tiles_x * tiles_y;
  if (block_n <= 0) {
    // Largest N tile that still gives every SM at least ~2 CTAs of work.
    const int sms = device_sm_count();
    block_n = 256;
    while (block_n > 64 && (N % block_n != 0 || (long)tiles * (N / block_n) < 2L * sms))
      block_n >>= 1;
    if (N % block_n != 0) block_n = 64;
  }
  STV_REQUIRE(N % block_n == 0, "conv_igemm: N %d not divisible by tile %d", N, block_n);

  CUtensorMap tx, twm;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)W, (uint64_t)H};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)W * C * 4};
    const uint32_t box[3] = {32, (uint32_t)tw, (uint32_t)th};
    if (int rc = encode_tmap_f32(&tx, x, 3, dims, strides, box, kSwizzle128B)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)taps * N};
    const uint64_t strides[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {32, (uint32_t)block_n};
    if (int rc = encode_tmap_f32(&twm, w_packed, 2, dims, strides, box, kSwizzle128B)) return rc;
  }
  ConvIgemmParams p;
  p.H = H; p.W = W; p.C = C; p.N = N; p.taps = taps; p.th = th; p.tw = tw; p.tiles_x = tiles_x;
  p.bias = bias; p.alpha = alpha; p.mask_src = mask_src; p.add_src = add_src;
  p.out_pre = out_pre; p.out_post = out_post;
  p.round_pre = round_flags & 1; p.round_post = (round_flags >> 1) & 1;

  switch (block_n) {
    case 64:  return launch_conv<64, 4>(tx, twm, p, tiles, stream);
    case 128: return launch_conv<128, 3>(tx, twm, p, tiles, stream);
    case 256: return launch_conv<256, 4>(tx, twm, p, tiles, stream);
    default:
      set_error("conv_igemm: unsupported N tile %d", block_n);
      return 2;
  }
}

}  // namespace stv
