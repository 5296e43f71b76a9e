// Row-band sharding of one image over several GPUs (BASELINE configs[4]): halo rows are pushed
// straight into the neighbouring ranks' buffers over NVLink (peer-mapped symmetric memory) by this
// GPU's own kernels, with flag words for the hand-shake -- no NCCL send/recv on the data path.
//
// The reference has no counterpart (single device, image_io.py:49-61 only warns above 3000 px); the
// data dependency is the 1-pixel halo of every 3x3 convolution (core_model.py:316 forward,
// optimization.py:313 backward).
//
// One exchange of a haloed buffer B = [rows + 2][row] (row 0 / rows + 1 = halos), slot s:
//   halo_ready : epoch e = ++epoch[s]; tell both neighbours "my B is final" (flag A = e) and wait
//                for theirs.  Needed because a conv runs over band + halos and so writes (inexact)
//                values into B's halo rows: a neighbour's push must not land before that conv ends.
//   halo_push  : copy my first / last own row into the lower halo of the rank above / the upper
//                halo of the rank below (plain stores to peer memory), zero my own halo at the
//                image boundary (= the conv's zero padding); the last block to finish fences,
//                raises flag B = e at both neighbours and waits for their B: on return my halos
//                hold the neighbours' rows.
// Flags live in symmetric memory as well: word [s][0] / [1] = A from the rank above / below,
// [s][2] / [3] = B from above / below.  Epochs only grow, so nothing is ever reset and a replayed
// CUDA graph needs no host-side values.  Every spin has a watchdog that traps instead of hanging.
#include "stv_common.cuh"
#include "stv_kernels.h"

namespace stv {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch, int slot, int which) {
  const uint64_t t0 = globaltimer_ns();
  unsigned spins = 0;
  while (static_cast<int>(ld_acquire_sys(p) - epoch) < 0) {
    if ((++spins & 0xfff) == 0 && globaltimer_ns() - t0 > 4 * STV_WATCHDOG_NS) {  // ranks may be skewed by host work
      printf("stv: halo flag watchdog fired (slot %d flag %d epoch %u seen %u)\n", slot, which, epoch,
             ld_acquire_sys(p));
      __trap();
    }
  }
}

struct HaloArgs {
  float* mine;
  float* up;    // rank above: base of ITS buffer (peer pointer) or null at the top of the image
  float* down;  // rank below or null
  int rows, rows_up, rows_down;  // own rows of this rank / of the rank above / below
  long row_floats;     // floats per row, multiple of 4
  int planes;          // 1 for NHWC activations, 3 for the NCHW image (rows are then per plane)
  unsigned* flags_mine;
  unsigned* flags_up;
  unsigned* flags_down;
  unsigned* epoch;  // local: [slots]
  unsigned* done;   // local: [slots]
  int slot;
};

__global__ void halo_ready_kernel(const HaloArgs a) {
  if (threadIdx.x != 0) return;
  const unsigned e = a.epoch[a.slot] + 1u;
  a.epoch[a.slot] = e;
  __threadfence_system();  // my conv's stores (stream-ordered before this kernel) precede the flag
  if (a.up) st_release_sys(a.flags_up + 4 * a.slot + 1, e);      // I am the rank BELOW `up`
  if (a.down) st_release_sys(a.flags_down + 4 * a.slot + 0, e);  // I am the rank ABOVE `down`
  if (a.up) wait_flag(a.flags_mine + 4 * a.slot + 0, e, a.slot, 0);
  if (a.down) wait_flag(a.flags_mine + 4 * a.slot + 1, e, a.slot, 1);
}

// own_epoch != 0: no halo_ready kernel ran before this one; every block derives the epoch of this
// exchange from the last completed one and the last block publishes it (all blocks have read
// epoch[slot] before they add to `done`, and the previous exchange of the slot is complete in
// stream order)
__global__ void __launch_bounds__(256) halo_push_kernel(const HaloArgs a, const int own_epoch) {
  const unsigned e = a.epoch[a.slot] + (own_epoch ? 1u : 0u);
  const long row4 = a.row_floats >> 2;
  const long per_plane_mine = static_cast<long>(a.rows + 2) * a.row_floats;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // work item = (direction, plane, float4 index)
  const long total = 2L * a.planes * row4;
  for (long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * 256) {
    const int dir = static_cast<int>(i / (a.planes * row4));
    const long r = i - static_cast<long>(dir) * a.planes * row4;
    const int pl = static_cast<int>(r / row4);
    const long c = r - static_cast<long>(pl) * row4;
    float* base = a.mine + pl * per_plane_mine;
    if (dir == 0) {
      if (a.up) {  // my first own row -> lower halo (row rows_up + 1) of the rank above
        const long per_plane_up = static_cast<long>(a.rows_up + 2) * a.row_floats;
        reinterpret_cast<float4*>(a.up + pl * per_plane_up +
                                  static_cast<long>(a.rows_up + 1) * a.row_floats)[c] =
            reinterpret_cast<const float4*>(base + a.row_floats)[c];
      } else {
        reinterpret_cast<float4*>(base)[c] = zero;  // top of the image: zero padding
      }
    } else {
      if (a.down) {  // my last own row -> upper halo (row 0) of the rank below
        const long per_plane_down = static_cast<long>(a.rows_down + 2) * a.row_floats;
        reinterpret_cast<float4*>(a.down + pl * per_plane_down)[c] =
            reinterpret_cast<const float4*>(base + static_cast<long>(a.rows) * a.row_floats)[c];
      } else {
        reinterpret_cast<float4*>(base + static_cast<long>(a.rows + 1) * a.row_floats)[c] = zero;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned prev = atomicAdd(a.done + a.slot, 1u);
  if (prev != gridDim.x - 1) return;
  a.done[a.slot] = 0;
  if (own_epoch) a.epoch[a.slot] = e;
  __threadfence_system();  // all blocks' peer stores are visible before the flags below
  if (a.up) st_release_sys(a.flags_up + 4 * a.slot + 3, e);
  if (a.down) st_release_sys(a.flags_down + 4 * a.slot + 2, e);
  if (a.up) wait_flag(a.flags_mine + 4 * a.slot + 2, e, a.slot, 2);
  if (a.down) wait_flag(a.flags_mine + 4 * a.slot + 3, e, a.slot, 3);
}

int halo_exchange_launch(float* mine, float* up, float* down, int rows, int rows_up, int rows_down,
                         long row_floats, int planes, unsigned* flags_mine, unsigned* flags_up,
                         unsigned* flags_down, unsigned* epoch, unsigned* done, int slot,
                         int wait_ready, cudaStream_t stream) {
  STV_REQUIRE(mine && flags_mine && epoch && done, "halo_exchange: null buffer");
  STV_REQUIRE(rows >= 1 && row_floats > 0 && row_floats % 4 == 0,
              "halo_exchange: rows %d / row length %ld (must be a multiple of 4 floats)", rows,
              row_floats);
  STV_REQUIRE(planes == 1 || planes == 3, "halo_exchange: planes must be 1 or 3");
  STV_REQUIRE((up == nullptr) == (flags_up == nullptr) && (down == nullptr) == (flags_down == nullptr),
              "halo_exchange: a neighbour needs both its buffer and its flags");
  STV_REQUIRE((up == nullptr || rows_up >= 1) && (down == nullptr || rows_down >= 1),
              "halo_exchange: neighbour band sizes %d / %d", rows_up, rows_down);
  STV_REQUIRE(slot >= 0, "halo_exchange: bad slot");
  HaloArgs a;
  a.mine = mine; a.up = up; a.down = down; a.rows = rows; a.rows_up = rows_up;
  a.rows_down = rows_down;
  a.row_floats = row_floats; a.planes = planes;
  a.flags_mine = flags_mine; a.flags_up = flags_up; a.flags_down = flags_down;
  a.epoch = epoch; a.done = done; a.slot = slot;
  // wait_ready = 0: nothing on this rank or its neighbours writes the halo rows of this buffer
  // except the exchange itself (every producer stores own rows only), so the "my buffer is final"
  // round trip is skipped: ONE launch per exchange, the push kernel advances the epoch itself
  if (wait_ready) {
    halo_ready_kernel<<<1, 32, 0, stream>>>(a);
    STV_CHECK_CUDA(cudaGetLastError());
  }
  const long items = 2L * planes * (row_floats / 4);
  long blocks = (items + 255) / 256;
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  halo_push_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, wait_ready ? 0 : 1);
  STV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace stv
