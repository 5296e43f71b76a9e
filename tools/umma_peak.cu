// Micro-benchmark: the tcgen05 kind::tf32 issue ceiling of one B200 (no loads: every MMA re-reads
// the same shared-memory operand tiles).  Gives the practical upper bound for the conv kernels,
// next to the "bf16 peak / 2" figure the roofline is quoted against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I style_transfer_visualizer_b200/csrc \
//        tools/umma_peak.cu -o tools/bin/umma_peak && tools/bin/umma_peak
#include <cstdio>
#include <cstdlib>

#include "stv_common.cuh"

namespace stv {
void set_error(const char*, ...) {}
}
using namespace stv;

template <int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int commit_every) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 16384;   // A: 128 x 32 floats, B: N x 32 floats
  const uint32_t bar = b_addr + 32768;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar + 32 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: zeros are fine for timing? use small finite values instead (avoid data-dependent gating)
  float* sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) sm[i] = 1.0f + (i % 7) * 0.125f;
  const uint32_t bar2 = bar + 8;   // dummy barrier that absorbs the in-loop commits
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (PAIR) { tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(slot)), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(const_cast<uint32_t*>(slot)), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (warp == 1 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_tf32(PAIR ? 256 : 128, N, 0, 0);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k);
        const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
        // alternate between two accumulators like the double-buffered conv
        if (PAIR) umma_tf32_pair(tmem + (it & 1) * N, ad, bd, idesc, it > 1);
        else umma_tf32(tmem + (it & 1) * N, ad, bd, idesc, it > 1);
      }
      if (commit_every > 0 && (it + 1) % commit_every == 0) {
        if (PAIR) umma_commit_pair(bar2); else umma_commit(bar2);
      }
    }
    if (PAIR) umma_commit_pair(bar); else umma_commit(bar);
  }
  if (warp == 1 && lane == 0) mbar_wait(bar, 0);
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

// Two issuing threads (lane 0 of warps 1 and 2) alternate groups of `group` MMAs, each group followed
// by that thread's own tcgen05.commit: does the ~350-cycle commit stall of one thread hide behind the
// other thread's issue?  `ordered` != 0 adds a shared-memory ticket so that group g+1 is issued only
// after group g has been issued (what an accumulating K loop would need).
template <int N>
__global__ void __launch_bounds__(128, 1) two_issuer_kernel(int groups, int group, int ordered,
                                                            int issuers) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, b_addr = base + 16384;
  const uint32_t bar = b_addr + 32768;          // final completion barrier (count = issuers)
  const uint32_t bar2 = bar + 8;                // sink for the in-loop commits
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (bar + 32 - smem_u32(smem_raw)));
  volatile int* ticket = reinterpret_cast<volatile int*>(smem_raw + (bar + 48 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) sm[i] = 1.0f + (i % 7) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(bar, issuers); mbar_init(bar2, 1); *ticket = 0; fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) { tmem_alloc(smem_u32(const_cast<uint32_t*>(slot)), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int me = warp - 1;   // issuer index 0 / 1
  if (lane == 0 && me >= 0 && me < issuers) {
    constexpr uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
    for (int g = me; g < groups; g += issuers) {
      if (ordered) {
        unsigned spins = 0;
        while (*ticket != g) {
          if (++spins > 400000000u) { printf("ticket watchdog\n"); __trap(); }
        }
        tc_fence_after();
      }
      for (int m = 0; m < group; ++m) {
        const int k = m & 3;
        const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k);
        const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
        umma_tf32(tmem, ad, bd, idesc, (g | m) != 0);
      }
      if (ordered) {
        tc_fence_before();
        __threadfence_block();
        *ticket = g + 1;
      }
      umma_commit(bar2);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
static void run_two(int ctas, int groups, int group, int ordered, int issuers) {
  auto kern = two_issuer_kernel<N>;
  const int smem = 16384 + 32768 + 1024 + 128;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    kern<<<ctas, 128, smem>>>(groups, group, ordered, issuers);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("two_issuer failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double mmas = static_cast<double>(groups) * group;
  printf("N=%3d issuers=%d ordered=%d group=%2d MMAs + commit: %.1f cycles/MMA @1.9GHz (%.1f TF/s)\n", N,
         issuers, ordered, group, best * 1e-3 * 1.9e9 / mmas,
         2.0 * 128 * N * 8 * mmas * ctas / best / 1e9);
}

template <int N, bool PAIR>
static void run(int ctas, int iters, int commit_every = 0) {
  auto kern = peak_kernel<N, PAIR>;
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = PAIR ? 1 : 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, iters, commit_every);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaEventSynchronize(e1);
    if (err != cudaSuccess || e2 != cudaSuccess) { printf("launch failed: %s %s\n", cudaGetErrorString(err), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flop = 2.0 * 128 * N * 32 * static_cast<double>(iters) * ctas;
  printf("N=%3d %s commit/%d ctas=%d iters=%d: %.3f ms  %.1f TF/s  (%.1f cycles/MMA @1.9GHz)\n", N,
         PAIR ? "pair  " : "single", 4 * commit_every, ctas, iters, best, flop / best / 1e9,
         best * 1e-3 * 1.9e9 / (4.0 * iters));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  run<256, false>(sms, iters);
  run<256, false>(sms, iters, 1);
  run<256, false>(sms, iters, 3);
  run<256, false>(sms, iters, 9);
  run<128, false>(sms, iters, 1);
  run<128, false>(sms, iters, 3);
  run<64, false>(sms, iters, 3);
  run<256, true>(sms & ~1, iters, 1);
  run<128, false>(sms, iters);
  run<64, false>(sms, iters);
  run<256, true>(sms & ~1, iters);
  run<128, true>(sms & ~1, iters);
  run<64, true>(sms & ~1, iters);
  for (int group : {12, 24}) {
    run_two<64>(sms, 4000, group, 0, 1);
    run_two<64>(sms, 4000, group, 0, 2);
    run_two<64>(sms, 4000, group, 1, 2);
  }
  run_two<128>(sms, 4000, 12, 0, 1);
  run_two<128>(sms, 4000, 12, 0, 2);
  run_two<128>(sms, 4000, 12, 1, 2);
  return 0;
}
