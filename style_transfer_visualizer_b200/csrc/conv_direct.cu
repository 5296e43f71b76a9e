// CUDA-core convolutions for the 3-channel image boundary of the VGG stack, plus a naive generic
// NHWC convolution used only as an on-device cross-check in the tests.
//
//   conv_first_fwd   : conv1_1 forward, 3 -> 64 channels.  Reads the optimised image in the
//                      reference's own NCHW layout (core_model.py:316, first block) and writes
//                      NHWC activations for the tensor-core layers.  K = 27, so the layer is
//                      HBM-write bound (64 fp32 per pixel out, 3 in); fp32 FMA on CUDA cores.
//   conv_first_dgrad : input gradient of conv1_1, 64 -> 3 channels, written as NCHW so that it
//                      is directly `input_img.grad` (optimization.py:313).  N = 3: no tensor-core
//                      shape fits; HBM-read bound.
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

constexpr int kFirstCout = 64;

// One thread = kFirstPx consecutive pixels of a row x 16 output channels (the four channel quads
// q = 4 j + cq, cq = lane % 4, so that the four cq lanes of a pixel write 64 contiguous bytes per
// store instruction), all 16 channels in ONE pass over the 27 taps:
//   * packed fp32 FMAs (fma.rn.f32x2, two channels per instruction): each half is an ordinary
//     fma.rn, so the result is bit-identical to a scalar FMA chain in the same tap order;
//   * the image values of one (channel, row) are loaded, broadcast into (v, v) pairs once and used
//     by 3 taps x 4 pixels x 8 channel pairs;
//   * persistent CTAs: the weights are staged in shared memory once per CTA, not once per 32 pixel
//     groups.
// Measured at 1080p (profiles/r2_ncu_conv_first.txt):
//   8 pixels x 4 channels per pass, scalar FMAs, one CTA per 256 pixels: 219 M warp instructions
//     (51 % FFMA), 59 % issue utilisation, 337 us (3.06 TB/s of stores);
//   this version with the image loads in front of their FMAs: 144 M instructions, 301 us; the
//     schedulers waited for the loads (4.7 warps per issue in long-scoreboard stalls) -> all 54
//     values are loaded first: 272 us (4.06 TB/s).  What is left is the LSU data pipe (64 % busy: a
//     128-bit shared load costs four wavefronts however many lanes share an address);
//   16 pixels x one channel quad per thread (a quarter of the weight LDS, full-line stores, sign
//     bits by redux.sync): 455 us; 8 pixels x 8 channels (half the weight LDS, 236 registers, two
//     CTAs per SM): 388 us -- both rejected: occupancy matters more here than LDS wavefronts.
constexpr int kFirstPx = 4;
constexpr int kFirstThreads = 128;
constexpr int kFirstGroupsPerCta = kFirstThreads / 4;
__global__ void __launch_bounds__(kFirstThreads, 3)
conv_first_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                      const float* __restrict__ bias, int H, int W, int groups_per_row,
                      int round_pre, float* __restrict__ out_pre, float* __restrict__ out_post,
                      uint2* __restrict__ out_bits, int in_rows, int in_row0) {
  // let a following tensor-core conv (launched with programmatic stream serialization) run its
  // prologue under this kernel's tail; it still waits for our completion before touching data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float4 ws[27][kFirstCout / 4];  // ws[k][c/4] = w[c..c+3][k], k = ci*9 + ky*3 + kx
  __shared__ float4 bs[kFirstCout / 4];
  for (int i = threadIdx.x; i < 27 * kFirstCout; i += kFirstThreads) {
    const int c = i / 27, k = i % 27;  // torch layout [Cout][3][3][3]
    reinterpret_cast<float*>(&ws[k][c >> 2])[c & 3] = w[i];
  }
  if (threadIdx.x < kFirstCout)
    reinterpret_cast<float*>(bs)[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();

  // the image may be a haloed band: in_rows rows per plane, row in_row0 = output row 0
  const long hw = static_cast<long>(in_rows) * W;
  const int cq = threadIdx.x & 3;
  const long total_groups = static_cast<long>(groups_per_row) * H;
  // whole warps stay in the loop together (the sign-bit shuffles are warp-wide)
  for (long base = static_cast<long>(blockIdx.x) * kFirstGroupsPerCta; base < total_groups;
       base += static_cast<long>(gridDim.x) * kFirstGroupsPerCta) {
    const long grp_raw = base + (threadIdx.x >> 2);
    const bool active = grp_raw < total_groups;  // groups past the end compute but store nothing
    const long grp = active ? grp_raw : 0;
    const int y = static_cast<int>(grp / groups_per_row);
    const int x0 = static_cast<int>(grp % groups_per_row) * kFirstPx;

    float2 acc[kFirstPx][8];  // [pixel][channel pair]: pairs 2 j, 2 j + 1 = quad 4 j + cq
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b = bs[4 * j + cq];
#pragma unroll
      for (int px = 0; px < kFirstPx; ++px) {
        acc[px][2 * j] = make_float2(b.x, b.y);
        acc[px][2 * j + 1] = make_float2(b.z, b.w);
      }
    }
    const bool interior = x0 >= 1 && x0 + kFirstPx < W;
    // all 54 image values of the 3 x 3 x (kFirstPx + 2) neighbourhood first: the loads are in
    // flight together instead of one (channel, row) at a time in front of its FMAs
    float raw[9][kFirstPx + 2];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int ci = r / 3, ky = r - 3 * ci;
      const int yy = y + ky - 1 + in_row0;
      const bool rok = (yy >= 0) && (yy < in_rows);
      const float* row = img + ci * hw + static_cast<long>(rok ? yy : 0) * W;
      if (interior) {
#pragma unroll
        for (int j = 0; j < kFirstPx + 2; ++j) raw[r][j] = rok ? __ldg(row + x0 + j - 1) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < kFirstPx + 2; ++j) {
          const int xx = x0 + j - 1;
          raw[r][j] = (rok && xx >= 0 && xx < W) ? __ldg(row + xx) : 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      float2 vv[kFirstPx + 2];  // (v, v) of x0-1 .. x0+kFirstPx
#pragma unroll
      for (int j = 0; j < kFirstPx + 2; ++j) vv[j] = make_float2(raw[r][j], raw[r][j]);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int k = r * 3 + kx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wk = ws[k][4 * j + cq];
          const float2 w01 = make_float2(wk.x, wk.y), w23 = make_float2(wk.z, wk.w);
#pragma unroll
          for (int px = 0; px < kFirstPx; ++px) {
            acc[px][2 * j] = __ffma2_rn(vv[px + kx], w01, acc[px][2 * j]);
            acc[px][2 * j + 1] = __ffma2_rn(vv[px + kx], w23, acc[px][2 * j + 1]);
          }
        }
      }
    }

    const long pix0 = static_cast<long>(y) * W + x0;
    // sign bits of the post-ReLU values (the ReLU gate of conv1_2's dgrad): word 0 = channels
    // 0..31, word 1 = channels 32..63 of each of this thread's pixels
#pragma unroll
    for (int px = 0; px < kFirstPx; ++px) {
      const bool st = active && x0 + px < W;
      uint32_t sb0 = 0u, sb1 = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = make_float4(acc[px][2 * j].x, acc[px][2 * j].y, acc[px][2 * j + 1].x,
                                     acc[px][2 * j + 1].y);
        const long o = (pix0 + px) * (kFirstCout / 4) + 4 * j + cq;
        if (out_pre && st) {
          float4 r = a;
          if (round_pre) {  // pre only feeds the Gram MMA: store it tf32-rounded
            r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w);
          }
          reinterpret_cast<float4*>(out_pre)[o] = r;
        }
        // post feeds conv1_2's MMA: store it tf32-rounded (see round_tf32)
        float4 r;
        r.x = round_tf32(relu_nan(a.x)); r.y = round_tf32(relu_nan(a.y));
        r.z = round_tf32(relu_nan(a.z)); r.w = round_tf32(relu_nan(a.w));
        if (out_post && st) reinterpret_cast<float4*>(out_post)[o] = r;
        const uint32_t nib = (r.x > 0.f ? 1u : 0u) | (r.y > 0.f ? 2u : 0u) |
                             (r.z > 0.f ? 4u : 0u) | (r.w > 0.f ? 8u : 0u);
        // channels 16 j + 4 cq .. + 3
        if (j < 2) sb0 |= nib << (16 * j + 4 * cq);
        else sb1 |= nib << (16 * (j - 2) + 4 * cq);
      }
      if (out_bits) {  // warp-uniform
        // OR over the four channel-quad lanes of the pixel group (adjacent lanes: cq = lane % 4)
        sb0 |= __shfl_xor_sync(0xffffffffu, sb0, 1);
        sb0 |= __shfl_xor_sync(0xffffffffu, sb0, 2);
        sb1 |= __shfl_xor_sync(0xffffffffu, sb1, 1);
        sb1 |= __shfl_xor_sync(0xffffffffu, sb1, 2);
        if (px == cq && st) out_bits[pix0 + px] = make_uint2(sb0, sb1);
      }
    }
  }
}

// One thread per image pixel; dY (NHWC, 64 ch) neighbours come through L1 (each element is reused
// by the 9 surrounding pixels of the same CTA row segment).
__global__ void __launch_bounds__(128)
conv_first_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int H, int W,
                        float* __restrict__ dimg) {
  // wt[tap][co] = {w[co][0][ky][kx], w[co][1][ky][kx], w[co][2][ky][kx], 0}
  __shared__ float4 wt[9][kFirstCout];
  for (int i = threadIdx.x; i < 9 * kFirstCout; i += 128) {
    const int tap = i / kFirstCout, co = i % kFirstCout;
    float4 v;
    v.x = w[(co * 3 + 0) * 9 + tap];
    v.y = w[(co * 3 + 1) * 9 + tap];
    v.z = w[(co * 3 + 2) * 9 + tap];
    v.w = 0.f;
    wt[tap][co] = v;
  }
  __syncthreads();
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y - (ky - 1);
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x - (kx - 1);
      if (xx < 0 || xx >= W) continue;
      const float4* src =
          reinterpret_cast<const float4*>(dy + (static_cast<long>(yy) * W + xx) * kFirstCout);
      const float4* wk = wt[ky * 3 + kx];
#pragma unroll 4
      for (int c4 = 0; c4 < kFirstCout / 4; ++c4) {
        const float4 g = __ldg(src + c4);
        const float4 w0 = wk[c4 * 4 + 0], w1 = wk[c4 * 4 + 1], w2 = wk[c4 * 4 + 2],
                     w3 = wk[c4 * 4 + 3];
        a0 = fmaf(g.x, w0.x, a0); a1 = fmaf(g.x, w0.y, a1); a2 = fmaf(g.x, w0.z, a2);
        a0 = fmaf(g.y, w1.x, a0); a1 = fmaf(g.y, w1.y, a1); a2 = fmaf(g.y, w1.z, a2);
        a0 = fmaf(g.z, w2.x, a0); a1 = fmaf(g.z, w2.y, a1); a2 = fmaf(g.z, w2.z, a2);
        a0 = fmaf(g.w, w3.x, a0); a1 = fmaf(g.w, w3.y, a1); a2 = fmaf(g.w, w3.z, a2);
      }
    }
  }
  const long hw = static_cast<long>(H) * W;
  const long o = static_cast<long>(y) * W + x;
  dimg[o] = a0;
  dimg[hw + o] = a1;
  dimg[2 * hw + o] = a2;
}

// Naive NHWC conv with packed weights [tap][N][C]; one thread per (pixel, n).  Test-only.
__global__ void conv_ref_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                const float* __restrict__ bias, int H, int W, int C, int N,
                                int taps, int relu, float* __restrict__ out) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = static_cast<long>(H) * W * N;
  if (idx >= total) return;
  const int n = static_cast<int>(idx % N);
  const long pix = idx / N;
  const int py = static_cast<int>(pix / W), px = static_cast<int>(pix % W);
  float acc = bias ? bias[n] : 0.f;
  for (int tap = 0; tap < taps; ++tap) {
    const int dy = taps == 9 ? tap / 3 - 1 : 0, dx = taps == 9 ? tap % 3 - 1 : 0;
    const int yy = py + dy, xx = px + dx;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float* xr = x + (static_cast<long>(yy) * W + xx) * C;
    const float* wr = wp + (static_cast<long>(tap) * N + n) * C;
    for (int c = 0; c < C; ++c) acc = fmaf(xr[c], wr[c], acc);
  }
  out[idx] = relu ? fmaxf(acc, 0.f) : acc;
}

int conv_first_fwd_launch(const float* img_nchw, const float* w, const float* bias, int H, int W,
                          int Cout, float* out_pre, float* out_post, unsigned* out_bits,
                          int round_pre, cudaStream_t stream, int in_rows, int in_row0) {
  STV_REQUIRE(Cout == kFirstCout, "conv_first_fwd: Cout must be %d (got %d)", kFirstCout, Cout);
  STV_REQUIRE(out_pre || out_post, "conv_first_fwd: no output buffer");
  STV_REQUIRE((reinterpret_cast<uintptr_t>(out_bits) & 7) == 0, "conv_first_fwd: out_bits alignment");
  const int groups_per_row = (W + kFirstPx - 1) / kFirstPx;
  const long groups = static_cast<long>(groups_per_row) * H;
  const long want = (groups + kFirstGroupsPerCta - 1) / kFirstGroupsPerCta;
  const long resident = 3L * device_sm_count();  // persistent CTAs: weights staged once each
  const unsigned blocks = static_cast<unsigned>(want < 2 * resident ? want : 2 * resident);
  conv_first_fwd_kernel<<<blocks, kFirstThreads, 0, stream>>>(img_nchw, w, bias, H, W, groups_per_row,
                                                    round_pre, out_pre, out_post,
                                                    reinterpret_cast<uint2*>(out_bits),
                                                    in_rows > 0 ? in_rows : H, in_row0);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_first_dgrad_launch(const float* dy, const float* w, int H, int W, int Cout,
                            float* dimg_nchw, cudaStream_t stream) {
  STV_REQUIRE(Cout == kFirstCout, "conv_first_dgrad: Cout must be %d (got %d)", kFirstCout, Cout);
  dim3 grid((W + 127) / 128, H);
  conv_first_dgrad_kernel<<<grid, 128, 0, stream>>>(dy, w, H, W, dimg_nchw);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int conv_ref_launch(const float* x, const float* w_packed, const float* bias, int H, int W, int C,
                    int N, int taps, int relu, float* out, cudaStream_t stream) {
  const long total = static_cast<long>(H) * W * N;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  conv_ref_kernel<<<blocks, 256, 0, stream>>>(x, w_packed, bias, H, W, C, N, taps, relu, out);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stv
