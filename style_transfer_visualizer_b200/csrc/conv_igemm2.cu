// Tensor-core 3x3 / 1x1 convolution, version 2: persistent, tap-reusing implicit GEMM.
//
// Same math and epilogue options as conv_igemm.cu (which it supersedes on the hot path); the
// difference is how operands reach shared memory.  Version 1 is bound by L2->SM bandwidth: every
// K step (one tap x 32 channels) re-loads a 128-pixel A tile and a BLOCK_N weight tile, i.e.
// (128 + BLOCK_N) * 128 bytes per 2*128*BLOCK_N*32 FLOP, which is 11-14 TB/s at the rates measured.
// Here:
//   * one A stage is a TALL patch {32 ch, tw, TH + 2 rows} loaded once per (channel slab, dx) and
//     used by the three dy taps through row-shifted UMMA descriptors.  With tw a multiple of 8
//     every 8-row core-matrix group is one 1024-byte-aligned run of 8 pixels, so the shifted views
//     keep the canonical K-major SWIZZLE_128B layout (SBO = 1024) -- A traffic drops 3x;
//   * a CTA owns M = 128 or 256 pixels (one or two 128-row accumulators in TMEM) that share every
//     weight stage;
//   * ONE operand ring: stage s = {A patch, the three dy weight taps of that (slab, dx)} behind one
//     full / empty mbarrier pair, so the single MMA-issuing thread pays one tcgen05.commit (a ~350
//     cycle stall of that thread on this part) per 12-24 MMAs -- two-half tiles have one issuing
//     thread per half, so one thread's stall hides behind the other's MMAs; 3-4 stages in flight cover the
//     commit -> producer -> TMA -> consumer chain (~1.7 us) that re-arms a stage.  (A legacy mode
//     with separate A / B rings and one tap per weight stage remains for forced configurations.)
//   * the kernel is persistent (one CTA per SM, static tile schedule, n-tile-major so concurrent
//     CTAs share weights in L2); the ring keeps streaming across tile boundaries and, when TMEM
//     allows (2 * M_HALVES * BLOCK_N <= 512 columns), the accumulator is double-buffered so the
//     epilogue of tile i overlaps the main loop of tile i+1;
//   * PAIR variant (cluster of 2 CTAs, tcgen05 cta_group::2): the two CTAs of a pair own two
//     different pixel patches but the SAME output-channel tile; each stages only HALF of every
//     weight tap (BLOCK_N/2 rows) and the leader's M = 256 MMAs read both halves.  Halving the
//     weight bytes is what lets a three-tap stage and a 3-deep ring fit for 256-wide tiles;
//   * epilogue options: bias / alpha, ReLU + tf32 rounding, dual pre / post outputs, fused 2x2 max
//     pool (forward); ReLU gate + accumulate through a shared-memory transpose (dgrad, Gram backward).
// Also hosts the N = 16 variant used for conv1_1's input gradient (64 -> 3 channels, written as
// NCHW planes): the 3 real output channels are padded to the smallest legal UMMA N.
#include <stdlib.h>

#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

struct Conv2Params {
  int H, W, C, N;
  int taps;            // 9 or 1
  int tw, th;          // CTA patch: th * tw == 128 * M_HALVES
  int tw_shift;        // log2(tw)
  int tiles_x, tiles_m, tiles_total;  // PAIR: tiles_m counts PAIRS of patches (tiles_total too)
  int a_stage_bytes;   // (th + ndy - 1) * tw * 128, multiple of 1024
  int a_stages, b_stages;  // ring depths
  int tps;                 // weight taps per B stage (1 or 3): narrow N tiles batch the three dy taps
                           // of one dx into a stage so the MMA thread synchronises 3x less often
  int uni;                 // unified ring: one weight stage per activation stage (tps == 3, or a 1x1
                           // conv), so stage s = {A tile, its weight taps} behind ONE full / empty
                           // barrier pair.  tcgen05.commit stalls the issuing thread for ~350
                           // cycles (profiles/r1_conv_bottleneck.log): one commit per 12 MMAs, not two
  const float* bias;
  const float* alpha;
  const float* mask_src;
  const float* add_src;
  float* out_pre;
  float* out_post;
  int round_pre, round_post;
  float* out_pool;     // optional fused 2x2 / stride-2 max pool of the post-ReLU output
                       // ([H/2][W/2][N], floor mode); plain forward epilogue, tw in {8, 16} only
  float* out_nchw3;    // N == 16 variant: [3][H][W] planes
  // ReLU / pool bookkeeping for the backward pass, so that it never re-reads an fp32 activation
  // just to test a sign or find an argmax:
  uint32_t* out_bits;          // forward: [H][W][N/32] words, bit c%32 = (out_post[.., c] > 0)
  uint32_t* out_code;          // forward + fused pool: [H][W][N/32] words, bit c%32 = "the pool + ReLU
                               // backward routes the pooled gradient of channel c to THIS pixel"
                               // (it holds the first maximum of its 2x2 window in ATen's scan
                               // order, and that maximum is > 0)
  const uint32_t* mask_bits;   // dgrad: ReLU gate from out_bits of the gated layer (instead of mask_src)
  const uint32_t* unpool_code; // dgrad: the output is the gradient of a POOLED map; route it through
                               // the 2x2 max-pool (+ ReLU gate) straight into out_pre [H2][W2][N]
  int H2, W2;
  // STYLE variant (dgrad of the conv that follows a style-tapped layer): the Gram backward
  // dF = gl * F * S of the tapped layer is a SECOND accumulator of the same tile, fed by style_kc
  // extra ring stages {F patch, S rows}; out = gate .* conv + gl * F S.  No separate 1x1 launch, no
  // round trip of the style gradient through memory.
  int x_row0;                  // input row that lines up with output row 0 (1 when the input buffer
                               // carries a halo row above the band: row-band sharding)
  int pool_smem;               // fused pool through the staging tile: one lane sees a whole 2x2
                               // window (no shuffles); needs the staging tiles
  int staged;                  // plain / un-pooling epilogue stores go through the shared-memory
                               // transpose (coalesced 128-byte lines); needs the staging tiles
  // Weight-stationary mode (64 -> 64 layers): all 9 x C/32 weight blocks of the CTA's N tile (and the
  // style slabs) are loaded ONCE into shared memory before the first tile, ring stages carry the
  // activation patch only.  These layers are bound by L2 -> SM traffic (~12 TB/s over all slices),
  // and more than half of it was the same 147 KB of weights re-read for every 128-pixel tile.
  int w_resident;
  int w_res_bytes;             // bytes of the resident region (replaces the weight ring)
  int style_kc;                // N / 32 slabs of the 1x1 contraction (0 = off)
  int style_a_bytes;           // th * tw * 128: an F patch has no halo rows
  const float* style_alpha;    // gl, device scalar
#ifdef STV_EXPERIMENTS
  int debug;           // bottleneck experiments (results are then garbage): 1 skip weight loads,
                       // 2 skip activation loads, 4 skip stores, 8 skip MMAs
#endif
};

#ifdef STV_EXPERIMENTS
#define STV_DBG(p, bit) ((p).debug & (bit))
#else
#define STV_DBG(p, bit) 0
#endif

// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2.. = epilogue.  A warp may only read the TMEM lanes
// 32*(warp%4)..+31, so epilogue warps come in groups of four; 128- and 256-wide tiles use two
// groups, each draining half of the columns (the epilogue of a CTA's LAST tile is not overlapped
// with anything, and on small feature maps every CTA has just one tile).
#ifndef STV_EPI_WIDE
#define STV_EPI_WIDE 1
#endif
// 64-wide two-half tiles as well: with four warps the epilogue of a 256 x 64 tile (pool shuffles,
// four-pixel un-pooling stores) took longer than its 144 MMAs and set the pace of conv1_2 forward
// and conv2_1 dgrad at 1080p (profiles/r2_ncu_conv_gram_metrics_1080p_v1.csv: 22 % tensor active).
__host__ __device__ constexpr int conv2_epi_warps(int block_n, int mh = 2) {
  return (STV_EPI_WIDE && (block_n >= 128 || (block_n == 64 && mh == 2))) ? 8 : 4;
}
// Two-half tiles get one MMA-issuing thread PER HALF (its own accumulator, so no ordering between
// them): tcgen05.commit stalls its thread for ~500 cycles, which starves the tensor pipe when a
// stage is only 12-24 short (N <= 128) MMAs; the other issuer's MMAs fill that hole
// (profiles/r1_umma_peak.log: N=64 72 -> 53 cycles per MMA, N=128 87 -> 66).
//
// SPLIT gives ONE-half tiles the same second issuer: the two threads take alternate ring stages
// (split-K) into two accumulators of the tile, which the epilogue adds.  It is used where a CTA has
// its SM to itself -- 128-wide one-half tiles (shared memory allows one CTA per SM) and the lone
// 64-wide CTAs of small feature maps; 64-wide tiles with two co-resident CTAs already have two
// issuing threads per SM.
__host__ __device__ constexpr int conv2_issuers(int mh, bool split = false) {
  return (mh == 2 || split) ? 2 : 1;
}
__host__ __device__ constexpr int conv2_threads(int block_n, int mh, bool split = false) {
  return 32 * (1 + conv2_issuers(mh, split) + conv2_epi_warps(block_n, mh));
}

template <int BLOCK_N, int MH, bool PAIR, bool STYLE = false, bool SPLIT = false>
struct Conv2Cfg {
  static_assert(!SPLIT || (MH == 1 && !PAIR && !STYLE && BLOCK_N >= 64), "split-K: one-half tiles");
  // accumulators per 128-pixel half: (conv, style) or the two split-K partial sums
  static constexpr int kAccW = (STYLE || SPLIT) ? 2 : 1;
  static constexpr int kTileCols = MH * BLOCK_N * kAccW;
  static constexpr int kAcc = (2 * kTileCols <= 512) ? 2 : 1;
  static constexpr int kTmemColsRaw = kAcc * kTileCols;
  static constexpr int kTmemCols = kTmemColsRaw < 32 ? 32 : kTmemColsRaw;
  static constexpr int kBRows = PAIR ? BLOCK_N / 2 : BLOCK_N;  // weight rows staged by ONE CTA
  static constexpr int kBBytes = kBRows * 128;
};

// 64- and 16-wide one-half tiles run as two co-resident CTAs per SM: cap their registers accordingly
template <int BLOCK_N, int MH, int TPS, bool PAIR, bool STYLE = false, bool SPLIT = false>
__global__ void __launch_bounds__(conv2_threads(BLOCK_N, MH, SPLIT),
                                  (BLOCK_N <= 64 && MH == 1 && !SPLIT) ? 2 : 1)
conv_igemm2_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x,
                        const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_f,
                        const __grid_constant__ CUtensorMap tmap_s, const Conv2Params p) {
  using Cfg = Conv2Cfg<BLOCK_N, MH, PAIR, STYLE, SPLIT>;
  constexpr int kTileCols = Cfg::kTileCols;
  constexpr int EW = conv2_epi_warps(BLOCK_N, MH);
  constexpr int NI = conv2_issuers(MH, SPLIT);
  constexpr int kStageUsers = SPLIT ? 1 : NI;  // issuers that consume (and commit) one ring stage
  constexpr int kThreads2 = conv2_threads(BLOCK_N, MH, SPLIT);
  constexpr int kFirstEpiWarp = 1 + NI;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // tile schedule: a "worker" is a CTA, or a CTA pair
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
  constexpr int kAcc = Cfg::kAcc;
  const int AS = p.a_stages, BS = p.b_stages;
  // (+2, not +1, for w_full: what follows the barriers is read as float4 and must stay 16-byte aligned)
  const int num_bars = 2 * AS + 2 * BS + 2 * kAcc + 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + AS * p.a_stage_bytes;
  const uint32_t bar_base =
      b_base + (p.w_resident ? static_cast<uint32_t>(p.w_res_bytes) : BS * TPS * Cfg::kBBytes);
  const uint32_t a_full = bar_base, a_empty = a_full + 8 * AS;
  const uint32_t b_full = a_empty + 8 * AS, b_empty = b_full + 8 * BS;
  const uint32_t acc_full = b_empty + 8 * BS, acc_empty = acc_full + 8 * kAcc;
  const uint32_t w_full = acc_empty + 8 * kAcc;  // resident weights have landed
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (bar_base - smem_base) + 8 * num_bars);
  // bias for all N output channels, staged once (epilogue reads it as smem broadcasts)
  float* sbias = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 8 * num_bars + 16);
  // per-epilogue-warp 32x32 fp32 transpose tile (4 KB each), 128-byte aligned
  const uint32_t stage_base =
      (bar_base + 8 * num_bars + 16 + static_cast<uint32_t>(p.N) * 4 + 127u) & ~127u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc = p.C >> 5;
  const int ndx = p.taps == 9 ? 3 : 1;
  const int ndy = ndx;
  const int row_bytes = p.tw * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if constexpr (STYLE) {
      tma_prefetch_desc(&tmap_f);
      tma_prefetch_desc(&tmap_s);
    }
    // PAIR: the leader's "full" barriers collect one expect_tx arrival from each CTA's producer,
    // its "accumulator drained" barriers one arrival from each epilogue warp of both CTAs
    const uint32_t np = PAIR ? 2u : 1u;
    // every issuer commits each stage / accumulator it has used
    for (int s = 0; s < AS; ++s) {
      mbar_init(a_full + 8 * s, np);
      mbar_init(a_empty + 8 * s, kStageUsers);
    }
    for (int s = 0; s < BS; ++s) {
      mbar_init(b_full + 8 * s, np);
      mbar_init(b_empty + 8 * s, kStageUsers);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(acc_full + 8 * s, NI);
      mbar_init(acc_empty + 8 * s, EW * np);
    }
    mbar_init(w_full, np);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  if (warp >= kFirstEpiWarp && BLOCK_N != 16) {
    for (int i = threadIdx.x - 32 * kFirstEpiWarp; i < p.N; i += kThreads2 - 32 * kFirstEpiWarp)
      sbias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers must exist before remote arrives
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.w_resident && warp == 0 && lane == 0) {
    // weights are never written by a kernel of the step: fetch them under the previous kernel's tail.
    // Block order = the ring's: ((c * 3 + dx) * TPS + dy) so that the issuer addresses a resident
    // "slot" c * 3 + dx exactly like a ring stage; style slabs follow.
    const uint32_t w_full_c = PAIR ? mapa_u32(w_full, 0) : w_full;
    const int n0 = static_cast<int>(rank) * Cfg::kBRows;
    const uint32_t bytes = static_cast<uint32_t>((kc * ndx * TPS + p.style_kc) * Cfg::kBBytes);
    if constexpr (PAIR) mbar_expect_tx_cluster(w_full_c, bytes);
    else mbar_expect_tx(w_full, bytes);
    for (int c = 0; c < kc; ++c)
      for (int dxi = 0; dxi < ndx; ++dxi)
#pragma unroll
        for (int u = 0; u < TPS; ++u) {
          const uint32_t dst = b_base + ((c * ndx + dxi) * TPS + u) * Cfg::kBBytes;
          if constexpr (PAIR) tma_load_2d_pair(dst, &tmap_w, w_full_c, c << 5, (u * ndx + dxi) * p.N + n0);
          else tma_load_2d(dst, &tmap_w, w_full, c << 5, (u * ndx + dxi) * p.N + n0);
        }
    if constexpr (STYLE) {
      for (int c = 0; c < p.style_kc; ++c) {
        const uint32_t dst = b_base + (kc * ndx * TPS + c) * Cfg::kBBytes;
        if constexpr (PAIR) tma_load_2d_pair(dst, &tmap_s, w_full_c, c << 5, n0);
        else tma_load_2d(dst, &tmap_s, w_full, c << 5, n0);
      }
    }
  }
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch, bias staging -- none of it touches activations) may overlap the tail of the previous
  // kernel in the stream; from here on we read / write tensors it may still be using.  (Issuing
  // the first ring of weight stages before this wait was tried and measured neutral-to-negative,
  // profiles/r1_conv_ab_epilogue_prefetch.log.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------ TMA producer --------------------------------------------
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      // PAIR: completion bytes and the expect_tx arrival go to the LEADER's barrier
      const uint32_t a_full_c = PAIR ? mapa_u32(a_full, 0) : a_full;
      const uint32_t b_full_c = PAIR ? mapa_u32(b_full, 0) : b_full;
      for (int t = worker; t < p.tiles_total; t += workers) {
        const int nt = t / p.tiles_m;
        const int mt = PAIR ? 2 * (t - nt * p.tiles_m) + static_cast<int>(rank) : t - nt * p.tiles_m;
        // (an odd patch count leaves the last pair's second CTA a patch below the image: its
        // loads are all out-of-bounds zero fill and its epilogue stores nothing)
        const int ty0 = (mt / p.tiles_x) * p.th, tx0 = (mt % p.tiles_x) * p.tw;
        const int n0 = nt * BLOCK_N + static_cast<int>(rank) * Cfg::kBRows;
        for (int c = 0; c < kc; ++c) {
          for (int dxi = 0; dxi < ndx; ++dxi) {
            mbar_wait(a_empty + 8 * as, aph ^ 1);
            const int cur = as;
            const uint32_t a_bytes = STV_DBG(p, 2) ? 0u : static_cast<uint32_t>(p.a_stage_bytes);
            const uint32_t b_bytes = (STV_DBG(p, 1) || p.w_resident)
                                         ? 0u : static_cast<uint32_t>(TPS * Cfg::kBBytes);
            const uint32_t tx_a = a_bytes + (p.uni ? b_bytes : 0u);
            if (tx_a == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(a_full_c + 8 * cur);
              else mbar_arrive(a_full + 8 * cur);
            } else if constexpr (PAIR) {
              mbar_expect_tx_cluster(a_full_c + 8 * cur, tx_a);
            } else {
              mbar_expect_tx(a_full + 8 * cur, tx_a);
            }
            if (a_bytes) {
              if constexpr (PAIR)
                tma_load_3d_pair(a_base + cur * p.a_stage_bytes, &tmap_x, a_full_c + 8 * cur, c << 5,
                                 tx0 + dxi - (ndx >> 1), ty0 - (ndy >> 1) + p.x_row0);
              else
                tma_load_3d(a_base + cur * p.a_stage_bytes, &tmap_x, a_full + 8 * cur, c << 5,
                            tx0 + dxi - (ndx >> 1), ty0 - (ndy >> 1) + p.x_row0);
            }
            if (++as == AS) { as = 0; aph ^= 1; }
            for (int dyi = 0; dyi < ndy; dyi += TPS) {
              uint32_t slot, bar_l, bar_c;   // weight slot and the barrier its bytes are credited to
              if (p.uni) {
                slot = cur; bar_l = a_full + 8 * cur; bar_c = a_full_c + 8 * cur;
              } else {
                mbar_wait(b_empty + 8 * bs, bph ^ 1);
                slot = bs; bar_l = b_full + 8 * bs; bar_c = b_full_c + 8 * bs;
                if (++bs == BS) { bs = 0; bph ^= 1; }
                if (b_bytes == 0) {
                  if constexpr (PAIR) mbar_arrive_cluster(bar_c);
                  else mbar_arrive(bar_l);
                } else if constexpr (PAIR) {
                  mbar_expect_tx_cluster(bar_c, b_bytes);
                } else {
                  mbar_expect_tx(bar_l, b_bytes);
                }
              }
              if (b_bytes == 0) continue;
#pragma unroll
              for (int u = 0; u < TPS; ++u) {
                const int tap = (dyi + u) * ndx + dxi;
                if constexpr (PAIR)
                  tma_load_2d_pair(b_base + (slot * TPS + u) * Cfg::kBBytes, &tmap_w, bar_c, c << 5,
                                   tap * p.N + n0);
                else
                  tma_load_2d(b_base + (slot * TPS + u) * Cfg::kBBytes, &tmap_w, bar_l, c << 5,
                              tap * p.N + n0);
              }
            }
          }
        }
        if constexpr (STYLE) {
          // Gram backward of the tapped layer as extra K steps of the same tile: stage =
          // {F patch (no halo), 32-channel slab of S rows n0..}; unified ring only
          const uint32_t tx_s =
              static_cast<uint32_t>(p.style_a_bytes + (p.w_resident ? 0 : Cfg::kBBytes));
          for (int c = 0; c < p.style_kc; ++c) {
            mbar_wait(a_empty + 8 * as, aph ^ 1);
            const int cur = as;
            if constexpr (PAIR) {
              mbar_expect_tx_cluster(a_full_c + 8 * cur, tx_s);
              tma_load_3d_pair(a_base + cur * p.a_stage_bytes, &tmap_f, a_full_c + 8 * cur, c << 5,
                               tx0, ty0);
              if (!p.w_resident)
                tma_load_2d_pair(b_base + cur * TPS * Cfg::kBBytes, &tmap_s, a_full_c + 8 * cur,
                                 c << 5, n0);
            } else {
              mbar_expect_tx(a_full + 8 * cur, tx_s);
              tma_load_3d(a_base + cur * p.a_stage_bytes, &tmap_f, a_full + 8 * cur, c << 5, tx0,
                          ty0);
              if (!p.w_resident)
                tma_load_2d(b_base + cur * TPS * Cfg::kBBytes, &tmap_s, a_full + 8 * cur, c << 5, n0);
            }
            if (++as == AS) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ------------------------------ MMA issuer(s) --------------------------------------------
    // issuer `me` owns the 128-pixel halves hf0 .. hf1-1 of every tile (all of them when NI == 1)
    const int me = warp - 1;
    const int hf0 = (NI == 2 && !SPLIT) ? me : 0, hf1 = (NI == 2 && !SPLIT) ? me + 1 : MH;
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(PAIR ? 256 : 128, BLOCK_N, 0, 0);
      // Descriptors: the upper word (SBO = 1024, version, SWIZZLE_128B) is constant; the lower word
      // is (address >> 4) | LBO, so every operand view is one 32-bit add away from the stage base.
      // The issue loop is a single thread's dependent instruction stream: for narrow N (32-64
      // tensor cycles per MMA) its length, not the tensor pipe, sets the pace.
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t lbo_lo = 1u << 16;
      int as = 0, bs = 0, acc = 0;
      int gs = 0;  // ring stages walked so far (SPLIT: stage gs belongs to issuer gs & 1)
      uint32_t aph = 0, bph = 0, accph = 0;
      const uint32_t row16 = static_cast<uint32_t>(row_bytes) >> 4;
      const uint32_t half16 = (128u >> p.tw_shift) * row16;  // rows per 128-pixel half
      if (p.w_resident) {
        mbar_wait(w_full, 0);
        tc_fence_after();
      }
      for (int t = worker; t < p.tiles_total; t += workers) {
        // epilogue (of both CTAs) has drained this accumulator
        if constexpr (PAIR) mbar_wait_cluster(acc_empty + 8 * acc, accph ^ 1);
        else mbar_wait(acc_empty + 8 * acc, accph ^ 1);
        tc_fence_after();
        // SPLIT: this issuer's own partial-sum accumulator (columns me * BLOCK_N ..)
        const uint32_t d0 = tmem_base + acc * kTileCols + (SPLIT ? me * BLOCK_N : 0);
        uint32_t accum = 0;
        for (int c = 0; c < kc; ++c) {
          for (int dxi = 0; dxi < ndx; ++dxi, ++gs) {
            // SPLIT: alternate ring stages, counted over the whole tile sequence; the ring depth is
            // even (host), so a ring slot -- and its pair of barriers -- always belongs to the same
            // issuer and no waiter can fall a barrier phase behind the other issuer's progress
            if (SPLIT && (gs & 1) != me) {
              // the other issuer's stage: step over it (and over its weight stages)
              if (++as == AS) { as = 0; aph ^= 1; }
              if (!p.uni) {
                for (int dyi = 0; dyi < ndy; dyi += TPS)
                  if (++bs == BS) { bs = 0; bph ^= 1; }
              }
              continue;
            }
            mbar_wait(a_full + 8 * as, aph);
            const uint32_t a_lo =
                (((a_base + as * p.a_stage_bytes) & 0x3FFFFu) >> 4) | lbo_lo;
            for (int dyi = 0; dyi < ndy; dyi += TPS) {
              if (!p.uni) mbar_wait(b_full + 8 * bs, bph);
              tc_fence_after();
              const uint32_t b_slot = p.w_resident ? c * ndx + dxi : (p.uni ? as : bs);
              const uint32_t b_lo =
                  (((b_base + b_slot * (TPS * Cfg::kBBytes)) & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
              for (int u = 0; u < TPS; ++u) {
#pragma unroll
                for (int hf = hf0; hf < hf1; ++hf) {
                  const uint32_t av = a_lo + (dyi + u) * row16 + hf * half16;
                  const uint32_t bv = b_lo + u * (Cfg::kBBytes >> 4);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (STV_DBG(p, 8)) continue;
                    const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (av + 2 * k);
                    const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (bv + 2 * k);
                    if constexpr (PAIR)
                      umma_tf32_pair(d0 + hf * BLOCK_N, adesc, bdesc, idesc,
                                     (u == 0 && k == 0) ? accum : 1u);
                    else
                      umma_tf32(d0 + hf * BLOCK_N, adesc, bdesc, idesc,
                                (u == 0 && k == 0) ? accum : 1u);
                  }
                }
                if (u == 0) accum = 1;
              }
              if (!p.uni) {
                if constexpr (PAIR) umma_commit_pair(b_empty + 8 * bs);
                else umma_commit(b_empty + 8 * bs);
                if (++bs == BS) { bs = 0; bph ^= 1; }
              }
            }
            if constexpr (PAIR) umma_commit_pair(a_empty + 8 * as);
            else umma_commit(a_empty + 8 * as);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
        }
        if constexpr (STYLE) {
          // second accumulator of the tile (columns MH * BLOCK_N ..): F patch x S slab, K = N
          for (int c = 0; c < p.style_kc; ++c) {
            mbar_wait(a_full + 8 * as, aph);
            tc_fence_after();
            const uint32_t a_lo = (((a_base + as * p.a_stage_bytes) & 0x3FFFFu) >> 4) | lbo_lo;
            const uint32_t s_addr = p.w_resident ? b_base + (kc * ndx * TPS + c) * Cfg::kBBytes
                                                 : b_base + as * (TPS * Cfg::kBBytes);
            const uint32_t b_lo = ((s_addr & 0x3FFFFu) >> 4) | lbo_lo;
#pragma unroll
            for (int hf = hf0; hf < hf1; ++hf) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adesc =
                    (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + hf * half16 + 2 * k);
                const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k);
                if constexpr (PAIR)
                  umma_tf32_pair(d0 + MH * BLOCK_N + hf * BLOCK_N, adesc, bdesc, idesc,
                                 (c | k) != 0 ? 1u : 0u);
                else
                  umma_tf32(d0 + MH * BLOCK_N + hf * BLOCK_N, adesc, bdesc, idesc,
                            (c | k) != 0 ? 1u : 0u);
              }
            }
            if constexpr (PAIR) umma_commit_pair(a_empty + 8 * as);
            else umma_commit(a_empty + 8 * as);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
        }
        if constexpr (PAIR) umma_commit_pair(acc_full + 8 * acc);
        else umma_commit(acc_full + 8 * acc);
        if (++acc == kAcc) { acc = 0; accph ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue -------------------------------------------------
    const int q = warp & 3;
    constexpr int kCols = BLOCK_N / (EW / 4);      // columns drained by this warp
    const int cb0 = ((warp - kFirstEpiWarp) >> 2) * kCols;
    int acc = 0;
    uint32_t accph = 0;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const uint32_t acc_empty_c = PAIR ? mapa_u32(acc_empty, 0) : acc_empty;
    // SPLIT: a tile with a single ring stage (C = 32, one tap) has only one partial sum, and which
    // issuer produced it alternates from tile to tile
    const int tile_stages = kc * ndx;
    int gs_epi = 0;  // ring stages before the current tile
    // one 32-column chunk of the tile's accumulator (the sum of the two split-K partials)
    auto ld_acc = [&](uint32_t taddr, uint32_t (&r)[32]) {
      if constexpr (SPLIT) {
        if (tile_stages < 2) {
          tmem_ld_32x32(taddr + (gs_epi & 1) * BLOCK_N, r);
        } else {
          tmem_ld_32x32(taddr, r);
          uint32_t s2[32];
          tmem_ld_32x32(taddr + BLOCK_N, s2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k)
            r[k] = __float_as_uint(__uint_as_float(r[k]) + __uint_as_float(s2[k]));
          return;
        }
      } else {
        tmem_ld_32x32(taddr, r);
      }
      tmem_ld_wait();
    };
    for (int t = worker; t < p.tiles_total; t += workers) {
      const int nt = t / p.tiles_m;
      const int mt = PAIR ? 2 * (t - nt * p.tiles_m) + static_cast<int>(rank) : t - nt * p.tiles_m;
      const int ty0 = (mt / p.tiles_x) * p.th, tx0 = (mt % p.tiles_x) * p.tw;
      const int n0 = nt * BLOCK_N;
      mbar_wait(acc_full + 8 * acc, accph);
      tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < MH; ++hf) {
        const int m = hf * 128 + q * 32 + lane;
        const int py = ty0 + (m >> p.tw_shift);
        const int px = tx0 + (m & (p.tw - 1));
        const bool valid = (py < p.H) && (px < p.W) && !STV_DBG(p, 4);
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              acc * kTileCols + hf * BLOCK_N;
        if constexpr (BLOCK_N == 16) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(trow)
              : "memory");
          tmem_ld_wait();
          if (valid) {
            const size_t hw = static_cast<size_t>(p.H) * p.W;
            const size_t o = static_cast<size_t>(py) * p.W + px;
            p.out_nchw3[o] = __uint_as_float(r[0]);
            p.out_nchw3[hw + o] = __uint_as_float(r[1]);
            p.out_nchw3[2 * hw + o] = __uint_as_float(r[2]);
          }
        } else {
          // Coalesced stores (p.staged): a warp's 32 pixels x 32 channels go through a 4 KB
          // shared-memory tile so that every st.global.v4 covers four complete 128-byte lines (8
          // lanes per pixel) instead of 32 different lines with 16 bytes each.  The scattered form
          // costs ~4 LSU cycles per 16 bytes: with it a dual-output forward layer took 2.5x and an
          // un-pooling dgrad 4x the time of the same MMAs (profiles/r2_epilogue_ab_v1.log).
          // (Staging the LAST tile of a CTA in the then idle operand ring -- so that the one-tile
          // CTAs of the small 512x512 feature maps store coalesced too -- was measured and
          // rejected: those launches are not store-bound, 845 -> 816 steps/s,
          // profiles/r2_epilogue_ab_512.log.)
          constexpr bool kCanStage = BLOCK_N <= 128;  // wider tiles: no shared memory to spare
          const bool staged = kCanStage && p.staged != 0;
          const uint32_t stg = stage_base + (warp - kFirstEpiWarp) * 4096;
          const int row_m0 = hf * 128 + q * 32;  // first pixel (tile-local) of this warp's 32 rows
          if (p.unpool_code != nullptr) {
            // dgrad whose output pixel is a POOLED pixel: the 2x2 max-pool backward (+ the ReLU gate
            // in front of the pool) is applied here, so the pooled gradient never goes to memory
            // and no pool-backward kernel re-reads the fp32 activation for its argmax.  Each lane
            // owns one pooled pixel; the four pixels of its window receive the value or zero
            // according to the routing bits the forward pass recorded.
            const size_t o00 = (static_cast<size_t>(2 * py) * p.W2 + 2 * px) * p.N + n0;
            const size_t wrow = static_cast<size_t>(p.W2) * p.N;
#pragma unroll 1
            for (int cb = cb0; cb < cb0 + kCols; cb += 32) {
              uint32_t r[32];
              ld_acc(trow + cb, r);
              uint32_t rw[4] = {0u, 0u, 0u, 0u};  // route bits of the four window pixels
              if (valid) {
                const size_t wq = (static_cast<size_t>(2 * py) * p.W2 + 2 * px) * (p.N >> 5) +
                                  ((n0 + cb) >> 5);
                const size_t wrow_w = static_cast<size_t>(p.W2) * (p.N >> 5);
                rw[0] = __ldg(p.unpool_code + wq);
                rw[1] = __ldg(p.unpool_code + wq + (p.N >> 5));
                rw[2] = __ldg(p.unpool_code + wq + wrow_w);
                rw[3] = __ldg(p.unpool_code + wq + wrow_w + (p.N >> 5));
              }
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                float v = __uint_as_float(r[k]) * alpha;
                if (p.round_pre) v = round_tf32(v);
                r[k] = __float_as_uint(v);
              }
              if (staged) {
#pragma unroll
                for (int wq4 = 0; wq4 < 4; ++wq4) {
                  const uint32_t bitsq = rw[wq4];
                  staged_store_32x32(
                      stg, lane, r,
                      [&](uint32_t bits_v, int k) { return ((bitsq >> k) & 1u) ? bits_v : 0u; },
                      [&](int rr) -> float* {
                        const int mm = row_m0 + rr;
                        const int yy = ty0 + (mm >> p.tw_shift), xx = tx0 + (mm & (p.tw - 1));
                        if (yy >= p.H || xx >= p.W) return nullptr;
                        return p.out_pre + (static_cast<size_t>(2 * yy + (wq4 >> 1)) * p.W2 +
                                            2 * xx + (wq4 & 1)) * p.N + n0 + cb;
                      });
                }
              } else if (valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                  for (int wq4 = 0; wq4 < 4; ++wq4) {
                    const uint32_t b4 = rw[wq4] >> (4 * j);
                    float4 o;
                    o.x = (b4 & 1u) ? __uint_as_float(r[4 * j + 0]) : 0.f;
                    o.y = (b4 & 2u) ? __uint_as_float(r[4 * j + 1]) : 0.f;
                    o.z = (b4 & 4u) ? __uint_as_float(r[4 * j + 2]) : 0.f;
                    o.w = (b4 & 8u) ? __uint_as_float(r[4 * j + 3]) : 0.f;
                    const size_t off = o00 + (wq4 >> 1) * wrow + (wq4 & 1) * p.N + cb + 4 * j;
                    *reinterpret_cast<float4*>(p.out_pre + off) = o;
                  }
                }
              }
            }
            continue;
          }
          if (p.mask_src == nullptr && p.add_src == nullptr) {
            // Plain epilogue (forward; dgrad without accumulation): each lane owns one pixel row.
            const size_t pix = static_cast<size_t>(py) * p.W + px;
            const size_t row_off = pix * p.N + n0;
            // Fused max pool: the 2x2 window of a pooled pixel lives in four lanes of this warp
            // (x neighbour = lane ^ 1, y neighbour = lane ^ tw; tile origins are even), so two
            // shuffles per value replace the separate pool kernel's re-read of `post`.  Max commutes
            // with the monotonic tf32 rounding, so the result equals pooling the stored tensor.
            const bool pool = p.out_pool != nullptr;
            const bool want_code = pool && p.out_code != nullptr;
            // the post-ReLU value is needed for the pool and the sign bits even when it is not stored
            // (a pooled layer's full-resolution activation has no reader once the backward pass
            // works from the pool codes)
            const bool want_post = p.out_post != nullptr || pool || p.out_bits != nullptr;
            const bool odd_x = (px & 1) != 0, odd_y = (py & 1) != 0;
            const int Ho = p.H >> 1, Wo = p.W >> 1;
            const bool in_window = (py >> 1) < Ho && (px >> 1) < Wo;
            const bool pool_writer = pool && !(py & 1) && !(px & 1) && in_window && !STV_DBG(p, 4);
            const size_t pool_off =
                (static_cast<size_t>(py >> 1) * Wo + (px >> 1)) * p.N + n0;
            const bool direct = !staged;  // scattered 16-byte stores straight from the registers
            // Pool through the staging tile (p.pool_smem): the warp's 32 pixels x 32 channels of
            // post-ReLU values are transposed through shared memory so that ONE lane holds the four
            // pixels of a 2x2 window for 8 channels: window k = lane % 8 (the warp's 32 pixels are 8
            // windows), channels 8 * (lane / 8) .. + 7.  Maxima and the first-maximum routing rule
            // become plain register compares (~3 instructions per value instead of two shuffles and
            // ~14 compare / select instructions: the shuffle form made conv1_2 forward spend 480 us
            // on 219 us of MMAs, profiles/r2_ncu_conv1_2_fwd_summary.txt), the pooled pixel is
            // written as 4 lanes x 32 contiguous bytes, and the routing bits travel back to the
            // pixel-owning lanes as bytes through the same tile.
            const bool pool_smem = kCanStage && pool && p.pool_smem != 0;
            const int wk = lane & 7, wcg = lane >> 3;
            const int wq = wk >> (p.tw_shift - 1), wr = wk & ((p.tw >> 1) - 1);
            const int wp00 = ((wq * 2) << p.tw_shift) + wr * 2;  // warp-local pixel of the window's corner
            const int wm = row_m0 + wp00;
            const int wyy = ty0 + (wm >> p.tw_shift), wxx = tx0 + (wm & (p.tw - 1));
            const bool w_ok = (wyy >> 1) < Ho && (wxx >> 1) < Wo && !STV_DBG(p, 4);
            const size_t w_off =
                (static_cast<size_t>(wyy >> 1) * Wo + (wxx >> 1)) * p.N + n0 + 8 * wcg;
            auto row_ptr = [&](float* base, int cb, int rr) -> float* {
              const int mm = row_m0 + rr;
              const int yy = ty0 + (mm >> p.tw_shift), xx = tx0 + (mm & (p.tw - 1));
              if (yy >= p.H || xx >= p.W || STV_DBG(p, 4)) return nullptr;
              return base + (static_cast<size_t>(yy) * p.W + xx) * p.N + n0 + cb;
            };
#pragma unroll 1
            for (int cb = cb0; cb < cb0 + kCols; cb += 32) {
              uint32_t r[32];
              ld_acc(trow + cb, r);
              uint32_t gate = 0xFFFFFFFFu;  // dgrad: ReLU gate of this pixel's 32 channels
              if (p.mask_bits != nullptr && valid)
                gate = __ldg(p.mask_bits + pix * (p.N >> 5) + ((n0 + cb) >> 5));
              uint32_t r2[STYLE ? 32 : 1];  // style accumulator chunk: gl * (F S)
              float gl = 0.f;
              if constexpr (STYLE) {
                tmem_ld_32x32(trow + MH * BLOCK_N + cb, r2);
                tmem_ld_wait();
                gl = __ldg(p.style_alpha);
              }
              uint32_t bits = 0;            // forward: sign bits of the post-ReLU values
              uint32_t route = 0;           // forward + pool: pool / ReLU backward routing bits
              if (valid || pool || staged) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = *reinterpret_cast<const float4*>(sbias + n0 + cb + 4 * j);
                  float4 v;
                  v.x = fmaf(__uint_as_float(r[4 * j + 0]), alpha, b.x);
                  v.y = fmaf(__uint_as_float(r[4 * j + 1]), alpha, b.y);
                  v.z = fmaf(__uint_as_float(r[4 * j + 2]), alpha, b.z);
                  v.w = fmaf(__uint_as_float(r[4 * j + 3]), alpha, b.w);
                  if (p.mask_bits != nullptr) {  // (uniform: forward launches skip the selects)
                    const uint32_t g4 = gate >> (4 * j);
                    v.x = (g4 & 1u) ? v.x : 0.f;
                    v.y = (g4 & 2u) ? v.y : 0.f;
                    v.z = (g4 & 4u) ? v.z : 0.f;
                    v.w = (g4 & 8u) ? v.w : 0.f;
                  }
                  if constexpr (STYLE) {  // the loss gradient of the tapped layer is not gated
                    v.x = fmaf(__uint_as_float(r2[4 * j + 0]), gl, v.x);
                    v.y = fmaf(__uint_as_float(r2[4 * j + 1]), gl, v.y);
                    v.z = fmaf(__uint_as_float(r2[4 * j + 2]), gl, v.z);
                    v.w = fmaf(__uint_as_float(r2[4 * j + 3]), gl, v.w);
                  }
                  // keep the (un-rounded) value for the staged stores below
                  r[4 * j + 0] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
                  r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
                  const int col = cb + 4 * j;
                  if (direct && p.out_pre && valid) {
                    float4 o = v;
                    if (p.round_pre) {
                      o.x = round_tf32(o.x); o.y = round_tf32(o.y);
                      o.z = round_tf32(o.z); o.w = round_tf32(o.w);
                    }
                    *reinterpret_cast<float4*>(p.out_pre + row_off + col) = o;
                  }
                  if (want_post) {
                    float4 o;
                    o.x = relu_nan(v.x); o.y = relu_nan(v.y);
                    o.z = relu_nan(v.z); o.w = relu_nan(v.w);
                    if (p.round_post) {
                      o.x = round_tf32(o.x); o.y = round_tf32(o.y);
                      o.z = round_tf32(o.z); o.w = round_tf32(o.w);
                    }
                    if (direct && p.out_post != nullptr && valid)
                      *reinterpret_cast<float4*>(p.out_post + row_off + col) = o;
                    bits |= ((o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) |
                             (o.w > 0.f ? 8u : 0u)) << (4 * j);
                    if (pool_smem) {
                      const uint32_t sa = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
                      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "f"(o.x),
                                   "f"(o.y), "f"(o.z), "f"(o.w)
                                   : "memory");
                    } else if (pool) {  // warp-uniform branch: every lane takes part in the shuffles
                      // 2x2 window = lanes {l, l^1 (x neighbour), l^tw, l^tw^1 (next / previous
                      // row)}: two shuffles per value give every lane the window maximum.  A lane
                      // routes the pooled gradient iff it holds the FIRST maximum in ATen's scan
                      // order (a, b, then the next row) and that maximum is positive (ReLU gate).
                      const float ov[4] = {o.x, o.y, o.z, o.w};
                      float mx[4];
                      uint32_t rb = 0;
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float nb = __shfl_xor_sync(0xffffffffu, ov[e], 1);
                        const float m0 = fmax_nan(nb, ov[e]);  // row maximum
                        const float mo = __shfl_xor_sync(0xffffffffu, m0, p.tw);
                        const float m = fmax_nan(mo, m0);
                        mx[e] = m;
                        const bool first = (ov[e] == m) && (m > 0.f) && !(odd_x && nb == m) &&
                                           !(odd_y && mo == m);
                        rb |= (first ? 1u : 0u) << e;
                      }
                      if (pool_writer)
                        *reinterpret_cast<float4*>(p.out_pool + pool_off + col) =
                            make_float4(mx[0], mx[1], mx[2], mx[3]);
                      route |= rb << (4 * j);
                    }
                  }
                }
              }
              if (pool_smem) {
                __syncwarp();
                float w4[4][8];  // [window pixel a, b, c, d in ATen's scan order][channel]
#pragma unroll
                for (int wp = 0; wp < 4; ++wp) {
                  const int prow = wp00 + (wp >> 1) * p.tw + (wp & 1);
#pragma unroll
                  for (int hc = 0; hc < 2; ++hc) {
                    const uint32_t sa = stg + prow * 128 + (((2 * wcg + hc) ^ (prow & 7)) << 4);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(w4[wp][4 * hc]), "=f"(w4[wp][4 * hc + 1]),
                                   "=f"(w4[wp][4 * hc + 2]), "=f"(w4[wp][4 * hc + 3])
                                 : "r"(sa)
                                 : "memory");
                  }
                }
                float mx[8];
                uint32_t rbyte[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float m = fmax_nan(fmax_nan(w4[0][e], w4[1][e]), fmax_nan(w4[2][e], w4[3][e]));
                  mx[e] = m;
                  // the FIRST maximum in scan order routes, and only a positive one (ReLU gate)
                  const bool pos = m > 0.f;
                  const bool ea = w4[0][e] == m, eb = w4[1][e] == m, ec = w4[2][e] == m,
                             ed = w4[3][e] == m;
                  rbyte[0] |= ((pos && ea) ? 1u : 0u) << e;
                  rbyte[1] |= ((pos && eb && !ea) ? 1u : 0u) << e;
                  rbyte[2] |= ((pos && ec && !ea && !eb) ? 1u : 0u) << e;
                  rbyte[3] |= ((pos && ed && !ea && !eb && !ec) ? 1u : 0u) << e;
                }
                if (w_ok) {
                  float4* dst = reinterpret_cast<float4*>(p.out_pool + w_off + cb);
                  dst[0] = make_float4(mx[0], mx[1], mx[2], mx[3]);
                  dst[1] = make_float4(mx[4], mx[5], mx[6], mx[7]);
                }
                if (want_code) {
                  __syncwarp();  // every lane has its window in registers: the tile can be reused
#pragma unroll
                  for (int wp = 0; wp < 4; ++wp) {
                    const int prow = wp00 + (wp >> 1) * p.tw + (wp & 1);
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(stg + prow * 4 + wcg), "r"(rbyte[wp])
                                 : "memory");
                  }
                  __syncwarp();
                  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(route) : "r"(stg + lane * 4) : "memory");
                }
                __syncwarp();  // the staged stores below (and the next chunk) rewrite the tile
              }
              if (p.out_bits != nullptr && valid)
                p.out_bits[pix * (p.N >> 5) + ((n0 + cb) >> 5)] = bits;
              if (want_code && valid)  // pixels of a row / column dropped by floor mode route nothing
                p.out_code[pix * (p.N >> 5) + ((n0 + cb) >> 5)] = in_window ? route : 0u;
              if (staged) {
                if (p.out_pre != nullptr) {
                  const bool rnd = p.round_pre != 0;
                  staged_store_32x32(
                      stg, lane, r,
                      [&](uint32_t bits_v, int) {
                        return rnd ? __float_as_uint(round_tf32(__uint_as_float(bits_v))) : bits_v;
                      },
                      [&](int rr) { return row_ptr(p.out_pre, cb, rr); });
                }
                if (p.out_post != nullptr) {
                  const bool rnd = p.round_post != 0;
                  staged_store_32x32(
                      stg, lane, r,
                      [&](uint32_t bits_v, int) {
                        const float o = relu_nan(__uint_as_float(bits_v));
                        return __float_as_uint(rnd ? round_tf32(o) : o);
                      },
                      [&](int rr) { return row_ptr(p.out_post, cb, rr); });
                }
              }
            }
            continue;
          }
          // TMEM gives each lane one pixel row (32 consecutive channels).  Writing that straight to
          // global memory would touch 32 different 128-byte lines per instruction, 16 bytes each;
          // the tile is transposed through shared memory instead so that 8 lanes cover one
          // pixel's 128 bytes: every global load/store instruction moves 4 complete lines.
          const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll 1
          for (int cb = cb0; cb < cb0 + kCols; cb += 32) {
            uint32_t r[32];
            ld_acc(trow + cb, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t a = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[4 * j]),
                           "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                           : "memory");
            }
            __syncwarp();
            const int col = cb + 4 * chunk;
            const float4 b = *reinterpret_cast<const float4*>(sbias + n0 + col);
#pragma unroll 1
            for (int h4 = 0; h4 < 8; h4 += 4) {  // two batches of 4 rows bound register use
              uint32_t offs[4];
              bool ok[4];
              float4 mk[4], ad[4];
              uint32_t gb[4];  // ReLU gate bits of this thread's four channels
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int mm = hf * 128 + q * 32 + (h4 + it) * 4 + sub;
                const int yy = ty0 + (mm >> p.tw_shift), xx = tx0 + (mm & (p.tw - 1));
                ok[it] = (yy < p.H) && (xx < p.W);
                const uint32_t pixi = static_cast<uint32_t>(yy) * p.W + xx;
                offs[it] = pixi * p.N + n0 + col;
                gb[it] = 0xFu;
                if (p.mask_bits != nullptr && ok[it])  // one word per pixel and 32-channel chunk:
                  gb[it] = __ldg(p.mask_bits + pixi * (p.N >> 5) + ((n0 + cb) >> 5)) >> (4 * chunk);
              }
              if (p.mask_src) {
#pragma unroll
                for (int it = 0; it < 4; ++it)
                  if (ok[it])
                    mk[it] = __ldg(reinterpret_cast<const float4*>(p.mask_src + offs[it]));
              }
              if (p.add_src) {
#pragma unroll
                for (int it = 0; it < 4; ++it)
                  if (ok[it]) ad[it] = *reinterpret_cast<const float4*>(p.add_src + offs[it]);
              }
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int rr = (h4 + it) * 4 + sub;
                const uint32_t a = stg + rr * 128 + ((chunk ^ (rr & 7)) << 4);
                float4 acc4;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(acc4.x), "=f"(acc4.y), "=f"(acc4.z), "=f"(acc4.w)
                             : "r"(a)
                             : "memory");
                if (!ok[it]) continue;
                float4 v;
                v.x = fmaf(acc4.x, alpha, b.x);
                v.y = fmaf(acc4.y, alpha, b.y);
                v.z = fmaf(acc4.z, alpha, b.z);
                v.w = fmaf(acc4.w, alpha, b.w);
                if (p.mask_src) {
                  v.x = mk[it].x > 0.f ? v.x : 0.f;
                  v.y = mk[it].y > 0.f ? v.y : 0.f;
                  v.z = mk[it].z > 0.f ? v.z : 0.f;
                  v.w = mk[it].w > 0.f ? v.w : 0.f;
                }
                v.x = (gb[it] & 1u) ? v.x : 0.f;
                v.y = (gb[it] & 2u) ? v.y : 0.f;
                v.z = (gb[it] & 4u) ? v.z : 0.f;
                v.w = (gb[it] & 8u) ? v.w : 0.f;
                if (p.add_src) {
                  v.x += ad[it].x; v.y += ad[it].y; v.z += ad[it].z; v.w += ad[it].w;
                }
                if (p.out_pre) {
                  float4 o = v;
                  if (p.round_pre) {
                    o.x = round_tf32(o.x); o.y = round_tf32(o.y);
                    o.z = round_tf32(o.z); o.w = round_tf32(o.w);
                  }
                  *reinterpret_cast<float4*>(p.out_pre + offs[it]) = o;
                }
                if (p.out_post) {
                  float4 o;
                  o.x = relu_nan(v.x); o.y = relu_nan(v.y);
                  o.z = relu_nan(v.z); o.w = relu_nan(v.w);
                  if (p.round_post) {
                    o.x = round_tf32(o.x); o.y = round_tf32(o.y);
                    o.z = round_tf32(o.z); o.w = round_tf32(o.w);
                  }
                  *reinterpret_cast<float4*>(p.out_post + offs[it]) = o;
                }
              }
            }
            __syncwarp();  // the staging tile is rewritten by the next chunk
          }
        }
      }
      // release the accumulator to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(acc_empty_c + 8 * acc);
        else mbar_arrive(acc_empty + 8 * acc);
      }
      if (++acc == kAcc) { acc = 0; accph ^= 1; }
      gs_epi += tile_stages;
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may exit while the other can still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Tile-policy overrides for sweeps and tests (stv_conv_set_tuning).  Thread-local: a launch only
// ever reads the calling thread's copy, so the C ABI stays re-entrant across host threads.
struct ConvTuning {
  int pair_mode = -1;                      // -1: rule table, 0: never, 1: whenever the shape allows
  int a_stages = 0, b_stages = 0, tps = 0;  // ring depth / taps-per-stage overrides (0 = defaults)
  int staged = -1;                          // epilogue stores: -1 rule, 0 direct, 1 coalesced
  int split = -1;                           // split-K second issuer: -1 rule, 0 never, 1 where legal
  int resident = -1;                        // weight-stationary 64 -> 64 layers: -1 rule, 0 never
  int pool_smem = -1;                       // fused pool through the staging tile: -1 rule, 0 never
};
static thread_local ConvTuning g_tuning;

// Per-shape plan overrides (stv_conv_plan_override): the sweep tool replaces the rule table's
// choice for ONE layer of a running step and times the whole step.  Thread-local like g_tuning.
struct ConvPlan { int H, W, C, N, backward, block_n, mh, pair, depth, tps; };
constexpr int kMaxPlans = 64;
static thread_local ConvPlan g_plans[kMaxPlans];
static thread_local int g_num_plans = 0;
int conv_plan_override(int H, int W, int C, int N, int backward, int block_n, int mh, int pair,
                       int depth, int tps) {
  if (H <= 0) { g_num_plans = 0; return 0; }
  for (int i = 0; i < g_num_plans; ++i) {
    ConvPlan& q = g_plans[i];
    if (q.H == H && q.W == W && q.C == C && q.N == N && q.backward == backward) {
      q = ConvPlan{H, W, C, N, backward, block_n, mh, pair, depth, tps};
      return 0;
    }
  }
  if (g_num_plans == kMaxPlans) return 1;
  g_plans[g_num_plans++] = ConvPlan{H, W, C, N, backward, block_n, mh, pair, depth, tps};
  return 0;
}
static bool is_backward(const ConvArgs& a) {
  return a.mask_bits || a.mask_src || a.unpool_code || a.style_x || a.out_nchw3;
}
static const ConvPlan* find_plan(const ConvArgs& a) {
  const int bwd = is_backward(a) ? 1 : 0;
  for (int i = 0; i < g_num_plans; ++i) {
    const ConvPlan& q = g_plans[i];
    if (q.H == a.H && q.W == a.W && q.C == a.C && q.N == a.N && q.backward == bwd) return &q;
  }
  return nullptr;
}
void conv_set_epilogue(int staged_mode) { g_tuning.staged = staged_mode; }
void conv_set_split(int mode) { g_tuning.split = mode; }
void conv_set_resident(int mode) { g_tuning.resident = mode; }
void conv_set_pool_smem(int mode) { g_tuning.pool_smem = mode; }
void conv_set_tuning(int pair_mode, int a_stages, int b_stages, int tps) {
  g_tuning.pair_mode = pair_mode;
  g_tuning.a_stages = a_stages;
  g_tuning.b_stages = b_stages;
  g_tuning.tps = tps;
}

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

// b_rows: weight rows one CTA stages per tap (the N tile, or half of it for a CTA pair)
// staging: the transposing epilogue (ReLU gate / accumulate) needs one 4 KB tile per epilogue warp
// res_bytes > 0: weight-stationary mode, the resident region replaces the weight ring
static int conv2_smem_bytes(int a_stage_bytes, int as, int bs, int tps, int b_rows,
                            int n_total, int block_n, int mh, int staging, int res_bytes = 0) {
  return as * a_stage_bytes + (res_bytes > 0 ? res_bytes : bs * tps * b_rows * 128) +
         8 * (2 * as + 2 * bs + 6) + 32 + 1024 + n_total * 4 + 128 +
         (staging ? conv2_epi_warps(block_n, mh) * 4096 : 0);
}

template <int BLOCK_N, int MH, int TPS, bool PAIR, bool STYLE = false, bool SPLIT = false>
static int launch2(const CUtensorMap& tx, const CUtensorMap& tw, const Conv2Params& p, int grid,
                   cudaStream_t stream, const CUtensorMap* tf = nullptr,
                   const CUtensorMap* ts = nullptr) {
  auto kern = conv_igemm2_tf32_kernel<BLOCK_N, MH, TPS, PAIR, STYLE, SPLIT>;
  const int smem = conv2_smem_bytes(p.a_stage_bytes, p.a_stages, p.b_stages, p.tps,
                                    Conv2Cfg<BLOCK_N, MH, PAIR>::kBRows, p.N, BLOCK_N, MH,
                                    p.staged || p.pool_smem || p.mask_src != nullptr ||
                                        p.add_src != nullptr,
                                    p.w_resident ? p.w_res_bytes : 0);
  STV_REQUIRE(smem <= 227 * 1024, "conv_igemm2: %d bytes of shared memory exceed the SM", smem);
  const int dev = current_device();
  // per kernel instantiation AND per device: the opt-in is a property of the (function, context)
  static bool attr_set[kMaxDevices] = {};
  if (!attr_set[dev]) {
    STV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024));
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(conv2_threads(BLOCK_N, MH, SPLIT));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  // programmatic dependent launch: the prologue overlaps the previous kernel's tail
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (PAIR) {
    // clusters of 2 cannot always use every SM (a GPC with an odd SM count strands one): size the
    // persistent grid by what the device can co-schedule
    static int max_clusters[kMaxDevices] = {};
    static int max_clusters_smem[kMaxDevices] = {};
    if (max_clusters[dev] <= 0 || max_clusters_smem[dev] != smem) {
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3(2 * device_sm_count());
      int n = 0;
      STV_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
      STV_REQUIRE(n > 0, "conv_igemm2: no CTA pair of %d bytes fits the device", smem);
      max_clusters[dev] = n;
      max_clusters_smem[dev] = smem;
    }
    if (grid > 2 * max_clusters[dev]) cfg.gridDim = dim3(2 * max_clusters[dev]);
  }
  // kernels without the STYLE accumulator ignore the two extra descriptors
  STV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tw, tf ? *tf : tx, ts ? *ts : tw, p));
  return 0;
}

// Tile selection (measured on B200, profiles/r1_selftest_perf_v2.log): N tile as wide as Cout
// allows; 256-wide tiles run best with one 128-pixel half per CTA (double-buffered accumulator),
// narrower tiles with two halves sharing each weight stage.  When that leaves SMs idle the tile is
// narrowed.  tw is the patch width with the least padded area (ties: wider rows = smaller halo).
struct TileChoice { int mh, tw, th, block_n, pair; };

// number of CTA-sized work items (a CTA pair counts as two)
static long count_tiles(int H, int W, int N, int mh, int tw, int bn) {
  const int th = 128 * mh / tw;
  return static_cast<long>((H + th - 1) / th) * ((W + tw - 1) / tw) * (N / bn);
}

// want_pool: the launch asks for the fused pool, whose 2x2 window must sit inside one epilogue
// warp: tw = 32 is excluded
static int pick_tw(int H, int W, int mh, bool want_pool) {
  // least padded pixels (= MMA work); the dy halo only inflates A traffic, so it is a tie-breaker
  const int tw_opts[3] = {16, 32, 8};
  int best = 16;
  double best_cost = -1.0;
  for (int i = 0; i < 3; ++i) {
    const int tw = tw_opts[i], th = 128 * mh / tw;
    if (want_pool && tw == 32) continue;
    const double area = static_cast<double>((H + th - 1) / th) * th * (((W + tw - 1) / tw) * tw);
    const double cost = area * (1.0 + 0.15 * 2.0 / th);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = tw;
    }
  }
  return best;
}

static TileChoice choose_tiles(const ConvArgs& a, bool heavy_epilogue, bool want_pool) {
  const int H = a.H, W = a.W, C = a.C, N = a.N;
  const int sms = device_sm_count();
  int bn, mh;
  if (N == 16) { bn = 16; mh = 1; }
  else if (N % 256 == 0) { bn = 256; mh = (C >= 512 && !heavy_epilogue) ? 2 : 1; }
  else if (N % 128 == 0) { bn = 128; mh = 2; }
  else { bn = 64; mh = heavy_epilogue ? 1 : 2; }  // fwd: 24 MMAs per commit beat two co-resident CTAs
  if (a.force_n <= 0 && a.force_mh <= 0 && N != 16) {
    // keep the machine busy on small feature maps: narrow the tile until ~every SM has one
    const int order[5][2] = {{256, 2}, {256, 1}, {128, 2}, {128, 1}, {64, 1}};
    int i = 0;
    while (i < 5 && !(order[i][0] == bn && order[i][1] == mh)) ++i;
    while (i < 4) {
      const int tw = pick_tw(H, W, mh, want_pool);
      if (count_tiles(H, W, N, mh, tw, bn) >= (3L * sms) / 4) break;
      ++i;
      if (N % order[i][0] != 0) continue;
      bn = order[i][0];
      mh = order[i][1];
    }
  }
  // 64 -> 64 layers with several tiles per CTA run weight-stationary (Conv2Params::w_resident) on
  // CTA pairs: each CTA keeps its half of the weights (74 KB) next to a 3-4 deep ring of 256-pixel
  // activation patches
  // (input-gradient launches only: the forward 64 -> 64 layer is bound by its fused-pool epilogue,
  // 481 us on a weight-stationary CTA pair vs 475 us as it was at 1080p,
  // profiles/r2_l2_traffic_1080p_step_v3.csv vs r2_ncu_conv_gram_metrics_1080p_v3.csv)
  const bool stationary = g_tuning.resident != 0 && N == 64 && C == 64 && a.taps == 9 &&
                          is_backward(a) &&
                          count_tiles(H, W, N, 2, pick_tw(H, W, 2, want_pool), 64) >= 2L * sms;
  if (stationary) { bn = 64; mh = 2; }
  const ConvPlan* plan = find_plan(a);
  if (plan && plan->block_n > 0 && N % plan->block_n == 0) bn = plan->block_n;
  if (plan && plan->mh > 0) mh = plan->mh;
  if (a.force_n > 0) bn = a.force_n;
  if (a.force_mh > 0) mh = a.force_mh;
  // CTA pairs (profiles/r1_pair_sweep.log, r1_ring_sweep.log): with half-size weight stages a
  // 256-wide tile affords three taps per stage (12 MMAs per barrier round trip) AND 3-deep rings;
  // a stage is only re-armed a full commit -> producer -> TMA -> consumer chain (~1.7 us) after its
  // MMAs retire, so ring depth x MMA time per stage must cover that chain.  +10-15 % on 256-wide
  // tiles, +5 % on 128-wide two-half tiles; neutral or worse elsewhere, which stays single-CTA.
  int pair = 0;
  if (N != 16) {
    if (g_tuning.pair_mode >= 0) pair = g_tuning.pair_mode;
    else pair = (a.taps == 9 && (bn == 256 || (bn == 128 && mh == 2) || (stationary && bn == 64 && mh == 2))) ? 1 : 0;
    if (plan && plan->pair >= 0) pair = plan->pair;
    if (pair && bn == 256 && a.force_mh <= 0 && !(plan && plan->mh > 0)) mh = 1;
  }
  const int tw2 = a.force_tw > 0 ? a.force_tw : pick_tw(H, W, mh, want_pool);
  return TileChoice{mh, tw2, 128 * mh / tw2, bn, pair};
}

int conv_igemm2_launch(const ConvArgs& a, cudaStream_t stream) {
  const int H = a.H, W = a.W, C = a.C, N = a.N, taps = a.taps;
  STV_REQUIRE(C % 32 == 0 && C >= 32, "conv_igemm2: input channels %d must be a multiple of 32", C);
  STV_REQUIRE(N % 64 == 0 || N == 16, "conv_igemm2: output channels %d must be 16 or a multiple of 64", N);
  STV_REQUIRE(taps == 9 || taps == 1, "conv_igemm2: taps must be 9 or 1 (got %d)", taps);
  STV_REQUIRE(H > 0 && W > 0, "conv_igemm2: empty image");
  STV_REQUIRE((N == 16) == (a.out_nchw3 != nullptr), "conv_igemm2: N == 16 <=> NCHW3 output");
  STV_REQUIRE(a.out_pre || a.out_post || a.out_nchw3 || a.out_pool, "conv_igemm2: no output buffer");
  STV_REQUIRE(static_cast<double>(H) * W * N < 4.0e9, "conv_igemm2: tensor exceeds 32-bit indexing");
  STV_REQUIRE(a.out_pool == nullptr || N != 16, "conv_igemm2: no pool on the N = 16 variant");
  STV_REQUIRE(a.out_code == nullptr || a.out_pool != nullptr,
              "conv_igemm2: pool codes are produced by the fused pool only");
  STV_REQUIRE(a.out_bits == nullptr || N != 16, "conv_igemm2: no sign bits on the N = 16 variant");
  STV_REQUIRE(!(a.mask_bits && a.mask_src), "conv_igemm2: give the ReLU gate as bits OR as fp32");
  STV_REQUIRE(a.x_row0 >= 0 && (a.x_rows == 0 || a.x_rows >= H + a.x_row0),
              "conv_igemm2: input rows %d / first row %d do not cover %d output rows", a.x_rows,
              a.x_row0, H);
  if (a.unpool_code != nullptr) {
    STV_REQUIRE(a.out_pre != nullptr && a.out_post == nullptr && !a.mask_src && !a.add_src &&
                    !a.mask_bits && !a.out_pool && !a.bias && N != 16,
                "conv_igemm2: the un-pooling epilogue writes out_pre only");
    STV_REQUIRE(a.H2 / 2 == H && a.W2 / 2 == W,
                "conv_igemm2: %dx%d is not the 2x2 floor-pooled size of %dx%d", H, W, a.H2, a.W2);
    STV_REQUIRE(static_cast<double>(a.H2) * a.W2 * N < 4.0e9, "conv_igemm2: un-pooled tensor too large");
  }

  const bool style = a.style_x != nullptr;
  if (style) {
    STV_REQUIRE(a.style_s != nullptr && a.style_alpha != nullptr && taps == 9 && N != 16 &&
                    a.out_pre != nullptr && a.out_post == nullptr && !a.mask_src && !a.add_src &&
                    !a.unpool_code && !a.out_pool && !a.bias && !a.alpha,
                "conv_igemm2: the fused style backward is a dgrad option (gate bits only)");
  }

  // big global reads in the epilogue (fp32 gate or accumulate source) go through the transposing
  // path; it also selects one-half tiles for the 64-wide layers (two co-resident CTAs)
  const bool staging_needed = a.mask_src != nullptr || a.add_src != nullptr;
  const bool heavy = staging_needed || a.mask_bits != nullptr;
  const bool pool_ok = a.out_pool != nullptr && !staging_needed && a.force_tw != 32;
  const TileChoice tc = choose_tiles(a, heavy, pool_ok);
  Conv2Params p;
  p.H = H; p.W = W; p.C = C; p.N = N; p.taps = taps;
  p.tw = tc.tw; p.th = tc.th;
  p.tw_shift = tc.tw == 8 ? 3 : (tc.tw == 16 ? 4 : 5);
  p.tiles_x = (W + tc.tw - 1) / tc.tw;
  p.tiles_m = p.tiles_x * ((H + tc.th - 1) / tc.th);
  if (tc.pair) p.tiles_m = (p.tiles_m + 1) / 2;  // pairs of patches
  p.tiles_total = p.tiles_m * (N / tc.block_n);
  const int halo = taps == 9 ? 2 : 0;
  p.a_stage_bytes = (tc.th + halo) * tc.tw * 128;
  p.bias = a.bias; p.alpha = a.alpha; p.mask_src = a.mask_src; p.add_src = a.add_src;
  p.out_pre = a.out_pre; p.out_post = a.out_post;
  p.round_pre = a.round_flags & 1; p.round_post = (a.round_flags >> 1) & 1;
  p.out_nchw3 = a.out_nchw3;
  p.out_pool = (pool_ok && tc.tw != 32) ? a.out_pool : nullptr;
  p.out_bits = a.out_bits;
  p.out_code = p.out_pool ? a.out_code : nullptr;
  p.mask_bits = a.mask_bits;
  p.unpool_code = a.unpool_code;
  p.H2 = a.H2; p.W2 = a.W2;
  p.x_row0 = a.x_row0;
  p.style_kc = style ? N / 32 : 0;
  p.style_a_bytes = tc.th * tc.tw * 128;
  p.style_alpha = a.style_alpha;
  // the second accumulator exists for the two tile families the tapped 64- and 128-wide layers
  // use; any other choice (small feature maps narrow the tile) reports "not fusable" to the caller
  if (style && !((tc.block_n == 64 && (tc.mh == 2 || !tc.pair)) ||
                 (tc.block_n == 128 && tc.mh == 2 && tc.pair)))
    return kConvStyleNotFusable;
  STV_REQUIRE(a.out_code == nullptr || p.out_code != nullptr,
              "conv_igemm2: pool codes requested but the pool cannot be fused for this shape");
#ifdef STV_EXPERIMENTS
  {
    const char* e = getenv("STV_CONV_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
#endif

  CUtensorMap tx, twm;
  {
    // x_rows > H: the input buffer has halo rows around the band (the conv then reads real
    // neighbour rows instead of the zero fill at the band edge)
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)W, (uint64_t)(a.x_rows > 0 ? a.x_rows : H)};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)W * C * 4};
    const uint32_t box[3] = {32, (uint32_t)tc.tw, (uint32_t)(tc.th + halo)};
    if (int rc = encode_tmap_f32(&tx, a.x, 3, dims, strides, box, kSwizzle128B)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)taps * N};
    const uint64_t strides[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {32, (uint32_t)(tc.pair ? tc.block_n / 2 : tc.block_n)};
    if (int rc = encode_tmap_f32(&twm, a.w_packed, 2, dims, strides, box, kSwizzle128B)) return rc;
  }
  CUtensorMap tf, ts;
  if (style) {
    const uint64_t dims[3] = {(uint64_t)N, (uint64_t)W, (uint64_t)H};
    const uint64_t strides[2] = {(uint64_t)N * 4, (uint64_t)W * N * 4};
    const uint32_t box[3] = {32, (uint32_t)tc.tw, (uint32_t)tc.th};
    if (int rc = encode_tmap_f32(&tf, a.style_x, 3, dims, strides, box, kSwizzle128B)) return rc;
    const uint64_t sdims[2] = {(uint64_t)N, (uint64_t)N};
    const uint64_t sstrides[1] = {(uint64_t)N * 4};
    const uint32_t sbox[2] = {32, (uint32_t)(tc.pair ? tc.block_n / 2 : tc.block_n)};
    if (int rc = encode_tmap_f32(&ts, a.style_s, 2, sdims, sstrides, sbox, kSwizzle128B)) return rc;
  }
  const int sms = device_sm_count();
  // ring depths: defaults keep >= ~2000 MMA cycles of weight stages in flight (TMA latency under
  // load) within the shared-memory budget; stv_conv_set_tuning overrides them for sweeps
  const ConvPlan* plan = find_plan(a);
  const int env_as = (plan && plan->depth > 0) ? plan->depth : g_tuning.a_stages;
  const int env_bs = (plan && plan->depth > 0) ? plan->depth : g_tuning.b_stages;
  const int env_tps = (plan && plan->tps > 0) ? plan->tps : g_tuning.tps;
  p.tps = (taps == 9 && (tc.block_n <= 128 || tc.pair)) ? 3 : 1;
  if (env_tps > 0 && taps == 9 && (tc.block_n <= 128 || tc.pair)) p.tps = env_tps;
  const int b_rows = tc.pair ? tc.block_n / 2 : tc.block_n;
  p.uni = (taps == 1 || p.tps == 3) ? 1 : 0;
  // Coalesced epilogue stores need one 4 KB staging tile per epilogue warp.  The transposing
  // (gate / accumulate) epilogue always has them; the plain and un-pooling epilogues of tiles up to
  // 128 wide take them when that does not cost a ring stage (256-wide tiles are tensor-bound and
  // their shared memory is full: they keep the direct stores).
  // weight-stationary mode: the N tile's 9 * C/32 weight blocks (+ style slabs) stay in shared memory
  int res_bytes = 0;
  if (g_tuning.resident != 0 && taps == 9 && p.uni && p.tps == 3 && N == tc.block_n && N == 64 &&
      C == 64 && tc.mh == 2 && tc.pair)
    res_bytes = (9 * (C / 32) + (style ? N / 32 : 0)) * b_rows * 128;
  // 64-wide one-half tiles run two CTAs per SM on a 2-deep ring each -- unless there are not even
  // enough tiles for one CTA per SM (small feature maps): then a lone CTA needs the deep ring to
  // cover the ~1.7 us re-arm chain (32x32x512 layer at depth 2: 0.9 us per 12-MMA stage).
  const bool two_per_sm = tc.block_n == 64 && tc.mh == 1 && p.tiles_total > sms;
  bool staging = staging_needed;
  for (int attempt = 0; attempt < 2; ++attempt) {
  staging = staging_needed;
  p.staged = 0;
  // Measured per launch type (profiles/r2_epilogue_ab_v2.log): the transpose pays where a value is
  // stored more than once or the epilogue is the bottleneck -- un-pooling dgrads (4 stores per value:
  // 2.1-3.3x faster), dual-output forward layers (1.4-1.6x), single-output dgrads (+3-8 %) -- and
  // costs 15-35 % on single-output forward layers, which therefore keep the direct stores.
  const bool staged_pays = a.unpool_code != nullptr || (a.out_pre != nullptr && a.out_post != nullptr) ||
                           (a.out_pre != nullptr && a.out_post == nullptr && a.out_pool == nullptr);
  const bool want_staged = g_tuning.staged < 0 ? staged_pays : g_tuning.staged == 1;
  // the fused pool wants the tiles too (window maxima / routing bits without shuffles)
  const bool want_pool_tiles = p.out_pool != nullptr && g_tuning.pool_smem != 0;
  p.pool_smem = 0;
  if (!staging_needed && (want_staged || want_pool_tiles) && N != 16 && tc.block_n <= 128) {
    auto depth_for = [&](bool stg) {
      int depth = p.uni ? (two_per_sm ? 2 : 4) : 4;
      if (p.uni && env_as > 0) depth = env_as;
      const int a_st = p.uni ? 0 : (tc.block_n >= 256 ? 2 : 3);
      while (depth > 2 && conv2_smem_bytes(p.a_stage_bytes, p.uni ? depth : a_st, depth, p.tps,
                                           b_rows, N, tc.block_n, tc.mh, stg, res_bytes) > 227 * 1024)
        --depth;
      return depth;
    };
    // (the explicit policy override may trade a ring stage for the staging tiles: experiments)
    if (p.uni && (depth_for(true) == depth_for(false) || g_tuning.staged == 1) &&
        conv2_smem_bytes(p.a_stage_bytes, depth_for(true), depth_for(true), p.tps, b_rows, N,
                         tc.block_n, tc.mh, true, res_bytes) <= 227 * 1024) {
      staging = true;
      p.staged = want_staged ? 1 : 0;
      p.pool_smem = want_pool_tiles ? 1 : 0;
    }
  }
  if (p.uni) {
    // one ring of {A tile, its weight taps}: as deep as shared memory allows, up to 4; 64-wide and
    // 16-wide tiles stay at 2 so that two CTAs share an SM
    int depth = two_per_sm ? 2 : 4;
    if (env_as > 0) depth = env_as;
    while (depth > 2 && conv2_smem_bytes(p.a_stage_bytes, depth, depth, p.tps, b_rows, N,
                                         tc.block_n, tc.mh, staging, res_bytes) > 227 * 1024)
      --depth;
    p.a_stages = depth;
    p.b_stages = depth;
  } else {
    p.a_stages = tc.block_n >= 256 ? 2 : 3;
    p.b_stages = 4;
    if (tc.pair) p.b_stages = 6;  // half-size weight stages: same bytes in flight
    if (env_as > 0) p.a_stages = env_as;
    if (env_bs > 0) p.b_stages = env_bs;
    while (conv2_smem_bytes(p.a_stage_bytes, p.a_stages, p.b_stages, p.tps, b_rows, N,
                            tc.block_n, tc.mh, staging) > 227 * 1024 && p.b_stages > 2)
      --p.b_stages;
  }
  // a resident weight block that leaves fewer than three activation stages (or does not fit at
  // all) is not worth it: size the rings again for the streaming mode
  if (res_bytes > 0 && (p.a_stages < 3 ||
                        conv2_smem_bytes(p.a_stage_bytes, p.a_stages, p.b_stages, p.tps, b_rows, N,
                                         tc.block_n, tc.mh, staging, res_bytes) > 227 * 1024)) {
    res_bytes = 0;
    continue;
  }
  break;
  }
  p.w_resident = res_bytes > 0 ? 1 : 0;
  p.w_res_bytes = res_bytes;
  // Second issuer by split-K (see conv2_issuers) for one-half tiles that have their SM to themselves
  const bool split_legal = !tc.pair && tc.mh == 1 && !style && p.uni &&
                           (tc.block_n == 128 || tc.block_n == 64);
  // rule: wherever the ring is even-deep anyway (the 128-wide 3-tap ring holds three stages: a
  // two-deep ring costs it more than the second issuer brings)
  const bool split = split_legal && (g_tuning.split < 0 ? (p.a_stages % 2 == 0 && !two_per_sm)
                                                        : g_tuning.split == 1);
  if (split && p.a_stages % 2 != 0) p.a_stages = p.b_stages = p.a_stages - 1;
  const int smem_est =
      conv2_smem_bytes(p.a_stage_bytes, p.a_stages, p.b_stages, p.tps, b_rows, N, tc.block_n,
                       tc.mh, staging, res_bytes);
  const int tile_cols = tc.mh * tc.block_n * ((style || split) ? 2 : 1);
  const int tmem_cols = (2 * tile_cols <= 512 ? 2 : 1) * tile_cols;
  if (style && !(p.uni && p.tps == 3)) return kConvStyleNotFusable;  // tuning overrides
  const int ctas_per_sm = (!split && smem_est <= 113 * 1024 && tmem_cols <= 256) ? 2 : 1;
  const int work_ctas = tc.pair ? 2 * p.tiles_total : p.tiles_total;
  int grid = work_ctas < sms * ctas_per_sm ? work_ctas : sms * ctas_per_sm;
  if (tc.pair) grid &= ~1;

  auto dispatch = [&]() -> int {
  if (style) {
    if (tc.block_n == 64 && tc.mh == 1)
      return launch2<64, 1, 3, false, true>(tx, twm, p, grid, stream, &tf, &ts);
    if (tc.block_n == 64 && !tc.pair)
      return launch2<64, 2, 3, false, true>(tx, twm, p, grid, stream, &tf, &ts);
    if (tc.block_n == 64) return launch2<64, 2, 3, true, true>(tx, twm, p, grid, stream, &tf, &ts);
    return launch2<128, 2, 3, true, true>(tx, twm, p, grid, stream, &tf, &ts);
  }
  if (split) {
    if (tc.block_n == 128) {
      if (p.tps == 3) return launch2<128, 1, 3, false, false, true>(tx, twm, p, grid, stream);
      return launch2<128, 1, 1, false, false, true>(tx, twm, p, grid, stream);
    }
    if (p.tps == 3) return launch2<64, 1, 3, false, false, true>(tx, twm, p, grid, stream);
    return launch2<64, 1, 1, false, false, true>(tx, twm, p, grid, stream);
  }
#define STV_L2(BN, MHV)                                                          \
  if (tc.block_n == BN && tc.mh == MHV && !tc.pair) {                            \
    if (p.tps == 3) return launch2<BN, MHV, 3, false>(tx, twm, p, grid, stream); \
    return launch2<BN, MHV, 1, false>(tx, twm, p, grid, stream);                 \
  }
#define STV_L2P(BN, MHV)                                                        \
  if (tc.block_n == BN && tc.mh == MHV && tc.pair) {                            \
    if (p.tps == 3) return launch2<BN, MHV, 3, true>(tx, twm, p, grid, stream); \
    return launch2<BN, MHV, 1, true>(tx, twm, p, grid, stream);                 \
  }
  STV_L2P(256, 2)
  STV_L2P(256, 1)
  STV_L2P(128, 2)
  STV_L2P(128, 1)
  STV_L2P(64, 2)
  STV_L2P(64, 1)
#undef STV_L2P
  if (tc.block_n == 256 && tc.mh == 2) return launch2<256, 2, 1, false>(tx, twm, p, grid, stream);
  if (tc.block_n == 256 && tc.mh == 1) return launch2<256, 1, 1, false>(tx, twm, p, grid, stream);
  STV_L2(128, 2)
  STV_L2(128, 1)
  STV_L2(64, 2)
  STV_L2(64, 1)
  STV_L2(16, 2)
  STV_L2(16, 1)
#undef STV_L2
  set_error("conv_igemm2: no kernel for N tile %d / M halves %d", tc.block_n, tc.mh);
  return 2;
  };
  if (int rc = dispatch()) return rc;
  // no fused pool for this shape (tw = 32 forced): separate pool kernel
  if (a.out_pool != nullptr && p.out_pool == nullptr) {
    STV_REQUIRE(a.out_post != nullptr, "conv_igemm2: the separate pool kernel needs out_post");
    return maxpool2_fwd_launch(a.out_post, H, W, N, a.out_pool, stream);
  }
  return 0;
}

}  // namespace stv
