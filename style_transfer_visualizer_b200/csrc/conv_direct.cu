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

// One thread = kFirstPx consecutive pixels of a row x 16 output channels (quads cq, cq+4, cq+8,
// cq+12 with cq = lane % 4).  Each weight quad read from shared memory (one LDS.128) feeds
// 4 * kFirstPx FMAs -- with 8 pixels the kernel is FMA- rather than LDS-bound (with 4 pixels the two
// pipes were balanced at ~50 % each and the layer ran at 3.0 TB/s of stores) -- and the four cq
// lanes of a pixel write 64 contiguous bytes (full 32-byte sectors) per store instruction.
constexpr int kFirstPx = 8;
__global__ void __launch_bounds__(128)
conv_first_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                      const float* __restrict__ bias, int H, int W, int groups_per_row,
                      int round_pre, float* __restrict__ out_pre, float* __restrict__ out_post,
                      uint2* __restrict__ out_bits, int in_rows, int in_row0) {
  // let a following tensor-core conv (launched with programmatic stream serialization) run its
  // prologue under this kernel's tail; it still waits for our completion before touching data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float4 ws[27][kFirstCout / 4];  // ws[k][c/4] = w[c..c+3][k], k = ci*9 + ky*3 + kx
  __shared__ float4 bs[kFirstCout / 4];
  for (int i = threadIdx.x; i < 27 * kFirstCout; i += 128) {
    const int c = i / 27, k = i % 27;  // torch layout [Cout][3][3][3]
    reinterpret_cast<float*>(&ws[k][c >> 2])[c & 3] = w[i];
  }
  if (threadIdx.x < kFirstCout)
    reinterpret_cast<float*>(bs)[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();

  // the image may be a haloed band: in_rows rows per plane, row in_row0 = output row 0
  const long hw = static_cast<long>(in_rows) * W;
  const int cq = threadIdx.x & 3;
  const long grp_raw = static_cast<long>(blockIdx.x) * 32 + (threadIdx.x >> 2);
  // groups past the end stay alive (the sign-bit shuffles below are warp-wide) but store nothing
  const bool active = grp_raw < static_cast<long>(groups_per_row) * H;
  const long grp = active ? grp_raw : 0;
  const int y = static_cast<int>(grp / groups_per_row);
  const int x0 = static_cast<int>(grp % groups_per_row) * kFirstPx;

  float in[3][3][kFirstPx + 2];  // [channel][row][x0-1 .. x0+kFirstPx]
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1 + in_row0;
      const bool rok = (yy >= 0) && (yy < in_rows);
      const float* row = img + ci * hw + static_cast<long>(yy) * W;
#pragma unroll
      for (int j = 0; j < kFirstPx + 2; ++j) {
        const int xx = x0 + j - 1;
        in[ci][ky][j] = (rok && xx >= 0 && xx < W) ? __ldg(row + xx) : 0.f;
      }
    }

  const long pix0 = static_cast<long>(y) * W + x0;
  // sign bits of the post-ReLU values (the ReLU gate of conv1_2's dgrad): word 0 = channels 0..31,
  // word 1 = channels 32..63 of each of this thread's pixels
  uint32_t sb0[kFirstPx], sb1[kFirstPx];
#pragma unroll
  for (int px = 0; px < kFirstPx; ++px) { sb0[px] = 0u; sb1[px] = 0u; }
#pragma unroll 1
  for (int i = 0; i < kFirstCout / 16; ++i) {
    const int c4 = i * 4 + cq;
    float4 acc[kFirstPx];
#pragma unroll
    for (int px = 0; px < kFirstPx; ++px) acc[px] = bs[c4];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 wk = ws[ci * 9 + ky * 3 + kx][c4];
#pragma unroll
          for (int px = 0; px < kFirstPx; ++px) {
            const float v = in[ci][ky][px + kx];
            acc[px].x = fmaf(v, wk.x, acc[px].x);
            acc[px].y = fmaf(v, wk.y, acc[px].y);
            acc[px].z = fmaf(v, wk.z, acc[px].z);
            acc[px].w = fmaf(v, wk.w, acc[px].w);
          }
        }
#pragma unroll
    for (int px = 0; px < kFirstPx; ++px) {
      if (active && x0 + px < W) {
        const long o = (pix0 + px) * (kFirstCout / 4) + c4;
        if (out_pre) {
          float4 r = acc[px];
          if (round_pre) {  // pre only feeds the Gram MMA: store it tf32-rounded
            r.x = round_tf32(r.x); r.y = round_tf32(r.y); r.z = round_tf32(r.z); r.w = round_tf32(r.w);
          }
          reinterpret_cast<float4*>(out_pre)[o] = r;
        }
        if (out_post || out_bits) {
          float4 r;
          // post feeds conv1_2's MMA: store it tf32-rounded (see round_tf32)
          r.x = round_tf32(relu_nan(acc[px].x)); r.y = round_tf32(relu_nan(acc[px].y));
          r.z = round_tf32(relu_nan(acc[px].z)); r.w = round_tf32(relu_nan(acc[px].w));
          if (out_post) reinterpret_cast<float4*>(out_post)[o] = r;
          const uint32_t nib = (r.x > 0.f ? 1u : 0u) | (r.y > 0.f ? 2u : 0u) |
                               (r.z > 0.f ? 4u : 0u) | (r.w > 0.f ? 8u : 0u);
          const uint32_t sh = nib << (16 * (i & 1) + 4 * cq);  // channels 16 i + 4 cq .. + 3
          if (i < 2) sb0[px] |= sh;
          else sb1[px] |= sh;
        }
      }
    }
  }
  if (out_bits) {  // warp-uniform
#pragma unroll
    for (int px = 0; px < kFirstPx; ++px) {
      // OR over the four channel-quad lanes of the pixel group (adjacent lanes: cq = lane % 4)
      sb0[px] |= __shfl_xor_sync(0xffffffffu, sb0[px], 1);
      sb0[px] |= __shfl_xor_sync(0xffffffffu, sb0[px], 2);
      sb1[px] |= __shfl_xor_sync(0xffffffffu, sb1[px], 1);
      sb1[px] |= __shfl_xor_sync(0xffffffffu, sb1[px], 2);
      if ((px & 3) == cq && active && x0 + px < W)
        out_bits[pix0 + px] = make_uint2(sb0[px], sb1[px]);
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
  const unsigned blocks = static_cast<unsigned>((groups + 31) / 32);
  conv_first_fwd_kernel<<<blocks, 128, 0, stream>>>(img_nchw, w, bias, H, W, groups_per_row,
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
