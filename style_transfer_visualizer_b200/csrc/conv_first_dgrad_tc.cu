// conv1_1 input gradient on the tensor cores with the x taps folded into N.
//
// Reference: autograd of the first Conv2d of VGG19 (`loss.backward()`, optimization.py:316-317 ->
// input_img.grad): dimg[y][x][ci] = sum_{co,ky,kx} dy[y+1-ky][x+1-kx][co] * w[co][ci][ky][kx].
//
// The generic implicit-GEMM kernel (conv_igemm2.cu, N = 16 variant) treats the 9 taps alike: per
// 128-pixel tile it loads the dy patch three times (once per x shift) for MMAs that are only 16
// columns wide, and is bound by those loads (1.9 GB of L2 -> SM traffic per 1080p launch against the
// ~12 TB/s the L2 slices deliver, profiles/r2_ncu_conv_gram_metrics_1080p_v2.csv: 205-220 us).  Here
// the x shift moves from the A operand to the OUTPUT side:
//   U[y][x'][kx][ci] = sum_{ky,co} dy[y+1-ky][x'][co] * w[co][ci][ky][kx]      (N = 3*3 = 9 -> 16)
//   dimg[y][x][ci]   = U[y][x+1][0][ci] + U[y][x][1][ci] + U[y][x-1][2][ci]
// so a tile loads its dy patch ONCE (two 32-channel halves), issues 3 (row taps) x 4 MMAs per half,
// and the epilogue adds the three x-shifted partial results of neighbouring pixels with two warp
// shuffles per channel.  Tiles overlap by one pixel column on each side (16 columns in, 14 out), so
// every output pixel finds both neighbours inside its own warp: one kernel, no atomics, a fixed
// summation order.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lanes 32 * (warp % 4)).
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

constexpr int kFdThreads = 192;
constexpr int kFdStages = 4;
constexpr int kFdTw = 16, kFdTh = 8;          // M = 128 pixels: 8 rows x 16 columns
constexpr int kFdOutCols = kFdTw - 2;         // columns 1..14 of a tile are outputs
constexpr int kFdABytes = (kFdTh + 2) * kFdTw * 128;  // 10 patch rows x 16 px x 32 channels
constexpr int kFdBBytes = 16 * 128;           // one (row tap, channel half) weight block
constexpr int kFdSmem = kFdStages * kFdABytes + 6 * kFdBBytes + 8 * (2 * kFdStages + 5) + 64 + 1024;

struct FirstDgradParams {
  int H, W, tiles_x, tiles_total;
  float* dimg;  // NCHW [3][H][W]
};

__global__ void __launch_bounds__(kFdThreads, 2)
conv_first_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                           const __grid_constant__ CUtensorMap tmap_w, const FirstDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base, b_base = a_base + kFdStages * kFdABytes;
  const uint32_t bar_base = b_base + 6 * kFdBBytes;
  const uint32_t full = bar_base, empty = full + 8 * kFdStages;
  const uint32_t acc_full = empty + 8 * kFdStages, acc_empty = acc_full + 16;
  const uint32_t w_full = acc_empty + 16;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gen + (bar_base - base) + 8 * (2 * kFdStages + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < kFdStages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + 8 * s, 1); mbar_init(acc_empty + 8 * s, 4); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlaps the previous kernel's tail; dy is read only after it has finished
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // weights: 3 row taps x 2 channel halves of [16][32], once per CTA
      mbar_expect_tx(w_full, 6 * kFdBBytes);
      for (int t = 0; t < 3; ++t)
        for (int c = 0; c < 2; ++c)
          tma_load_2d(b_base + (t * 2 + c) * kFdBBytes, &tmap_w, w_full, c << 5, t * 16);
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int xo = tx * kFdOutCols - 1, yo = ty * kFdTh - 1;  // patch origin (may be -1: zero fill)
        for (int c = 0; c < 2; ++c) {
          mbar_wait(empty + 8 * s, ph ^ 1);
          mbar_expect_tx(full + 8 * s, kFdABytes);
          tma_load_3d(a_base + s * kFdABytes, &tmap_dy, full + 8 * s, c << 5, xo, yo);
          if (++s == kFdStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 16, 0, 0);
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t lbo_lo = 1u << 16;
      mbar_wait(w_full, 0);
      int s = 0, acc = 0;
      uint32_t ph = 0, accph = 0;
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        mbar_wait(acc_empty + 8 * acc, accph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 16;
        for (int c = 0; c < 2; ++c) {
          mbar_wait(full + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_lo = (((a_base + s * kFdABytes) & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
          for (int tdy = 0; tdy < 3; ++tdy) {
            // output row r reads patch row r + tdy: a whole-row offset keeps the swizzle atoms aligned
            const uint32_t av = a_lo + tdy * ((kFdTw * 128) >> 4);
            const uint32_t bv = (((b_base + (tdy * 2 + c) * kFdBBytes) & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (av + 2 * k);
              const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (bv + 2 * k);
              umma_tf32(d, adesc, bdesc, idesc, (c | tdy | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty + 8 * s);
          if (++s == kFdStages) { s = 0; ph ^= 1; }
        }
        umma_commit(acc_full + 8 * acc);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = 2 * q + (lane >> 4), col = lane & 15;
    const size_t hw = static_cast<size_t>(p.H) * p.W;
    int acc = 0;
    uint32_t accph = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int y = ty * kFdTh + row, x = tx * kFdOutCols - 1 + col;
      mbar_wait(acc_full + 8 * acc, accph);
      tc_fence_after();
      uint32_t r[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 16;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
            "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
            "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
      tmem_ld_wait();
      // the accumulator is in registers: hand it back before the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * acc);
      if (++acc == 2) { acc = 0; accph ^= 1; }
      const bool st = col >= 1 && col <= kFdOutCols && y < p.H && x < p.W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        // kx = 0 comes from the pixel to the right, kx = 2 from the pixel to the left
        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(r[ci]), 1);
        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(r[6 + ci]), 1);
        const float v = (right + __uint_as_float(r[3 + ci])) + left;
        if (st) p.dimg[ci * hw + static_cast<size_t>(y) * p.W + x] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

int conv_first_dgrad_tc_launch(const float* dy, const float* w_rows, int H, int W, int Cout,
                               float* dimg_nchw, cudaStream_t stream) {
  STV_REQUIRE(Cout == 64, "conv_first_dgrad_tc: Cout must be 64 (got %d)", Cout);
  STV_REQUIRE(H > 0 && W > 0, "conv_first_dgrad_tc: empty image");
  CUtensorMap tdy, tw;
  {
    const uint64_t dims[3] = {64, (uint64_t)W, (uint64_t)H};
    const uint64_t strides[2] = {64 * 4, (uint64_t)W * 64 * 4};
    const uint32_t box[3] = {32, kFdTw, kFdTh + 2};
    if (int rc = encode_tmap_f32(&tdy, dy, 3, dims, strides, box, kSwizzle128B)) return rc;
  }
  {
    const uint64_t dims[2] = {64, 48};
    const uint64_t strides[1] = {64 * 4};
    const uint32_t box[2] = {32, 16};
    if (int rc = encode_tmap_f32(&tw, w_rows, 2, dims, strides, box, kSwizzle128B)) return rc;
  }
  static bool attr_set[kMaxDevices] = {};
  int dev = 0;
  STV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  if (!attr_set[dev]) {
    STV_CHECK_CUDA(cudaFuncSetAttribute(conv_first_dgrad_tc_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kFdSmem));
    attr_set[dev] = true;
  }
  FirstDgradParams p;
  p.H = H; p.W = W;
  p.tiles_x = (W + kFdOutCols - 1) / kFdOutCols;
  p.tiles_total = p.tiles_x * ((H + kFdTh - 1) / kFdTh);
  p.dimg = dimg_nchw;
  const int max_ctas = 2 * device_sm_count();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.tiles_total < max_ctas ? p.tiles_total : max_ctas);
  cfg.blockDim = dim3(kFdThreads);
  cfg.dynamicSmemBytes = kFdSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  STV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_first_dgrad_tc_kernel, tdy, tw, p));
  return 0;
}

}  // namespace stv
