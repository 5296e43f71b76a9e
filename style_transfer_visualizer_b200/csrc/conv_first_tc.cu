// conv1_1 forward on the tensor cores: 3 -> 64 channels, K = 3*3*3 = 27 padded to 32.
//
// Reference: the first Conv2d of torchvision's VGG19 reached through `x = block(x)`
// (core_model.py:316); on CUDA the reference runs it through cuDNN in TF32 like every other conv
// (torch.backends.cudnn.allow_tf32 defaults to True).  The CUDA-core kernel in conv_direct.cu does
// the 3456 FLOP per pixel in exact fp32 and is bound by its FMA / shared-memory pipes (3.0-3.4 TB/s
// of output at 1080p); here the K = 27 contraction is four tcgen05.mma instructions per 128 pixels
// and the layer is left with its real cost: writing 2 x 256 bytes per pixel.
//
// One tile = 128 consecutive pixels of an image row x 64 channels:
//   gather   every thread builds ONE im2col row (27 image values, tf32-rounded, k = ci*9+ky*3+kx)
//            straight from the NCHW image and stores it as the K-major, 128-byte-swizzled A operand
//            (no TMA: the operand does not exist in memory);
//   MMA      one thread issues 4 x (M=128, N=64, K=8) into TMEM and commits to an mbarrier;
//   epilogue warp w drains TMEM lanes 32w..32w+31 (= its own 32 pixels): bias, tf32 rounding, ReLU,
//            sign bits, and coalesced stores of `pre` and `post` through a shared-memory transpose.
// The three phases of a CTA are sequential; four to five co-resident CTAs per SM (42 KB of shared
// memory, 64 TMEM columns each) overlap each other's phases.
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

constexpr int kFtcThreads = 128;
constexpr int kFtcA = 128 * 128;        // A tile: 128 rows x 32 floats
constexpr int kFtcB = 64 * 128;         // B tile: 64 rows x 32 floats
constexpr int kFtcStage = 4 * 4096;     // one 32x32 transpose tile per warp
constexpr int kFtcSmem = kFtcA + kFtcB + kFtcStage + 64 + 256 + 1024;

struct FirstTcParams {
  const float* img;   // NCHW [3][in_rows][W]
  const float* w;     // torch layout [64][3][3][3]
  const float* bias;
  int H, W, in_rows, in_row0;
  int tiles_x, tiles_total;
  int round_pre;
  float* out_pre;
  float* out_post;
  uint32_t* out_bits;  // [H][W][2]
};

__global__ void __launch_bounds__(kFtcThreads, 4) conv_first_tc_kernel(const FirstTcParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base, b_smem = base + kFtcA, stage = b_smem + kFtcB;
  const uint32_t bar = stage + kFtcStage;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + kFtcA + kFtcB + kFtcStage + 8);
  float* sbias = reinterpret_cast<float*>(gen + kFtcA + kFtcB + kFtcStage + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // B operand: row n = output channel, 32 floats of k (27 real), tf32-rounded, swizzled like A
  {
    const int n = tid >> 1, c0 = (tid & 1) * 4;  // two threads per row, four 16-byte chunks each
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * c + e;
        v[e] = k < 27 ? round_tf32(__ldg(p.w + n * 27 + k)) : 0.f;
      }
      const uint32_t a = b_smem + n * 128 + ((c ^ (n & 7)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[0]), "f"(v[1]),
                   "f"(v[2]), "f"(v[3])
                   : "memory");
    }
  }
  if (tid < 64) sbias[tid] = p.bias ? __ldg(p.bias + tid) : 0.f;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 64);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
  constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t a_lo = ((a_smem & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = ((b_smem & 0x3FFFFu) >> 4) | (1u << 16);
  const long plane = static_cast<long>(p.in_rows) * p.W;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
    const int y = tile / p.tiles_x;
    const int x0 = (tile - y * p.tiles_x) * 128;
    // ---- gather: this thread's pixel (y, x0 + tid) -> 27 neighbours, k = ci*9 + ky*3 + kx
    {
      const int x = x0 + tid;
      float v[32];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1 + p.in_row0;
          const bool rok = yy >= 0 && yy < p.in_rows;
          const float* row = p.img + ci * plane + static_cast<long>(yy) * p.W;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            v[ci * 9 + ky * 3 + kx] =
                (rok && xx >= 0 && xx < p.W) ? round_tf32(__ldg(row + xx)) : 0.f;
          }
        }
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t a = a_smem + tid * 128 + ((c ^ (tid & 7)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * c]),
                     "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                     : "memory");
      }
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    // ---- MMA: D[128 px][64 ch] = A[128][32] * B[64][32]^T
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k);
        const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
        umma_tf32(tmem, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: lane = pixel x0 + 32 * warp + lane
    {
      const int x = x0 + 32 * warp + lane;
      const bool valid = x < p.W;
      const size_t pix = static_cast<size_t>(y) * p.W + x;
      const uint32_t stg = stage + warp * 4096;
      auto row_ptr = [&](float* out, int cb, int rr) -> float* {
        const int xr = x0 + 32 * warp + rr;
        if (xr >= p.W) return nullptr;
        return out + (static_cast<size_t>(y) * p.W + xr) * 64 + cb;
      };
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(32 * warp) << 16) + cb, r);
        tmem_ld_wait();
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = __uint_as_float(r[k]) + sbias[cb + k];
          r[k] = __float_as_uint(v);
          bits |= (round_tf32(relu_nan(v)) > 0.f ? 1u : 0u) << k;
        }
        if (p.out_bits != nullptr && valid) p.out_bits[pix * 2 + (cb >> 5)] = bits;
        if (p.out_pre != nullptr) {
          const bool rnd = p.round_pre != 0;
          staged_store_32x32(
              stg, lane, r,
              [&](uint32_t b, int) {
                return rnd ? __float_as_uint(round_tf32(__uint_as_float(b))) : b;
              },
              [&](int rr) { return row_ptr(p.out_pre, cb, rr); });
        }
        if (p.out_post != nullptr) {
          staged_store_32x32(
              stg, lane, r,
              [&](uint32_t b, int) {
                return __float_as_uint(round_tf32(relu_nan(__uint_as_float(b))));
              },
              [&](int rr) { return row_ptr(p.out_post, cb, rr); });
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // A tile and the accumulator are free for the next tile
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int conv_first_fwd_tc_launch(const float* img_nchw, const float* w, const float* bias, int H, int W,
                             int Cout, int in_rows, int in_row0, float* out_pre, float* out_post,
                             unsigned* out_bits, int round_pre, cudaStream_t stream) {
  STV_REQUIRE(Cout == 64, "conv_first_fwd_tc: Cout must be 64 (got %d)", Cout);
  STV_REQUIRE(out_pre || out_post, "conv_first_fwd_tc: no output buffer");
  STV_REQUIRE(H > 0 && W > 0, "conv_first_fwd_tc: empty image");
  if (in_rows <= 0) in_rows = H;
  static bool attr_set[kMaxDevices] = {};
  int dev = 0;
  STV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  if (!attr_set[dev]) {
    STV_CHECK_CUDA(cudaFuncSetAttribute(conv_first_tc_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kFtcSmem));
    attr_set[dev] = true;
  }
  FirstTcParams p;
  p.img = img_nchw; p.w = w; p.bias = bias; p.H = H; p.W = W; p.in_rows = in_rows;
  p.in_row0 = in_row0;
  p.tiles_x = (W + 127) / 128;
  p.tiles_total = p.tiles_x * H;
  p.round_pre = round_pre;
  p.out_pre = out_pre; p.out_post = out_post; p.out_bits = out_bits;
  const int max_ctas = device_sm_count() * 4;
  const int grid = p.tiles_total < max_ctas ? p.tiles_total : max_ctas;
  conv_first_tc_kernel<<<grid, kFtcThreads, kFtcSmem, stream>>>(p);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stv
