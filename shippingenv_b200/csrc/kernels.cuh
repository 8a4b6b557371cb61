// kernels.cuh -- the sm_100a kernels: step (+auto-reset, +observation epilogue, +episode
// statistics), reset, observe, sample_action and the MCTS rollout.
//
// Layout in HBM (DESIGN.md "Data layout"): ships[n] 16-byte records, clock[n] u32,
// actions[n] i16, reward[n] f32/f64, done/reset/error ballots ceil(n/32) u32 each.
// One env per thread; a warp moves 512 B of ship records per 128-bit load/store.
// Every CTA is persistent (grid = SMs x resident CTAs), stages the world blob into
// shared memory with one TMA bulk copy and then strides over 256-env tiles.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "transition.cuh"

namespace shipenv {

constexpr int kThreads = 256;  // envs per tile
constexpr int kStages = 4;     // input tiles in flight per CTA
constexpr int kStepBlock = kThreads + 32;  // 8 consumer warps + 1 TMA producer warp

struct alignas(16) TileStage {  // one 256-env tile, filled by three TMA bulk copies
  uint4 ships[kThreads];
  uint32_t clock[kThreads];
  int16_t act[kThreads];
};
static_assert(sizeof(TileStage) == 5632, "stage layout");
constexpr size_t kPipeBytes = sizeof(TileStage) * kStages + 2 * 8 * kStages + 4 * kStages;  // slots + full/empty mbarriers + tile ids

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copies (cp.async.bulk; SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() {  // the 8 consumer warps only (the producer warp has left)
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {   // shared-address forms
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(bar_saddr), "r"(parity) : "memory");
}
// producer-side wait: sleeps ~1 us between polls (a tile takes a CTA about that long and kStages tiles are
// buffered) so that the waiting warp does not eat issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITB_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEB_%=;\n"
      "nanosleep.u32 1024;\n"
      "bra WAITB_%=;\n"
      "DONEB_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int lds_s16(uint32_t saddr) {
  short v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return int(v);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy in the current bulk group
__device__ __forceinline__ void tma_store_bulk(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
// pull a contiguous global range towards L2 (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// streaming 128-bit accesses: ship records are touched once per launch, keep them out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Stage the world blob: thread 0 issues one bulk copy; everybody waits on the mbarrier.
// Returns after the data is visible to all threads of the CTA.
__device__ __forceinline__ void stage_world(void* smem_world, const void* gmem_world, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    tma_load_bulk(smem_world, gmem_world, bytes, bar);
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) <= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

// ------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------
struct StepArgs {
  const void* world;        // device blob
  uint32_t world_bytes;
  uint4* ships;
  uint32_t* clock;
  const int16_t* actions;
  shipenv_step_io io;
  uint64_t seed, env_id_base;
  int64_t n;
  int max_steps;
  uint32_t flags, word0;
  int obs_width;            // 6 + 4P
  int obs_bufs;             // observation tiles resident in shared memory (1 or 2)
  uint32_t round_key[20];   // Philox round keys of `seed`: (k0 + r*W0, k1 + r*W1), r = 0..9
  uint32_t* tile_counter;   // step_fast_kernel: global ticket counter (context-owned, never reset)
  uint32_t tile_base;       // ticket value this launch starts at
};

// Observation row head (utils/preprocessing.py:53-60): x, y, fuel, "cargo" (= fuel, the
// state dict copies it, shipping/environment.py:206), origin | -1, dest | -1.
__device__ __forceinline__ void obs_head(const ShipR& s, float* row) {
  const float f = float(s.fuel);
  row[0] = float(s.x); row[1] = float(s.y); row[2] = f; row[3] = f;
  row[4] = s.origin == kNone ? -1.f : float(s.origin);
  row[5] = s.dest == kNone ? -1.f : float(s.dest);
}

// per-env episode bookkeeping (optional outputs): running return + statistics at episode end
__device__ __forceinline__ void track_episode(const shipenv_step_io& io, int i, double reward, bool was_reset,
                                              uint32_t clock_before) {
  double* er = io.ep_return;
  if (!er) return;
  double ep = __dadd_rn(er[i], reward);
  if (was_reset) {
    if (io.stats) {
      double* st = io.stats;
      atomicAdd(st + 0, ep);
      atomicAdd(st + 1, ep * ep);
      atomicAdd(st + 2, double((clock_before + 1u) & 0xFFFFu));
      atomicAdd(st + 3, 1.0);
      atomic_max_f64(st + 4, ep);
      atomic_min_f64(st + 5, ep);
    }
    ep = 0.0;
  }
  er[i] = ep;
}

enum : uint32_t { kFlagDone = 1u, kFlagReset = 2u, kFlagError = 4u };

// The general path: any action (policy, select, take, move with cargo loss / arrival /
// destination re-draw), any word0.  Out of line on purpose: it reloads its inputs and writes
// every per-env output itself, so nothing of the streaming loop is live across the call and
// the common case keeps a small register footprint.  Returns kFlag* bits for the ballots.
__device__ __forceinline__ uint32_t step_general_inline(const StepArgs& a, const void* world_smem, int i) {
  const WorldView w(world_smem);
  const uint4 rec = a.ships[i];
  const uint32_t c = a.clock[i];
  const int act = (a.flags & SHIPENV_F_ALL_POLICY) ? SHIPENV_ACT_POLICY : int(a.actions[i]);
  ShipR s = unpack(rec);
  Rng g;
  g.seek(a.round_key, c, a.env_id_base + uint64_t(uint32_t(i)), a.word0);
  double reward = 0.0;
  bool done = false;
  int taken = act, status;
  if (act == SHIPENV_ACT_POLICY) {
    taken = -1;
    status = policy_step(w, s, g, reward, done, taken);
  } else {
    int cat, val;
    decode_action(act, w.h->P, cat, val);
    status = do_action(w, s, g, cat, val, reward, done);
  }
  uint32_t flags = 0;
  if (status == SHIPENV_ST_OK) {
    uint32_t c1 = c + 1u;
    const bool ended = done || (a.max_steps > 0 && int(c1 & 0xFFFFu) >= a.max_steps);
    bool was_reset = false;
    if ((a.flags & SHIPENV_F_AUTO_RESET) && ended) {
      env_reset(w, s, g);  // continues the word stream of this (env, clock) cell
      c1 = ((c >> 16) + 1u) << 16;
      was_reset = true;
    }
    track_episode(a.io, i, reward, was_reset, c);
    a.ships[i] = pack(s);
    a.clock[i] = c1;
    flags = (done ? kFlagDone : 0u) | (was_reset ? kFlagReset : 0u);
  } else {
    reward = 0.0;
    flags = kFlagError;
  }
  if (a.io.reward_f32) a.io.reward_f32[i] = float(reward);
  if (a.io.reward_f64) a.io.reward_f64[i] = reward;
  if (a.io.status) a.io.status[i] = uint8_t(status);
  if (a.io.taken) a.io.taken[i] = int16_t(taken);
  return flags;
}

// Out-of-line copy for step_kernel, whose streaming loop calls it for single lanes.
__device__ __noinline__ uint32_t step_general(const StepArgs& a, const void* world_smem, int i) {
  return step_general_inline(a, world_smem, i);
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s16(const int16_t* p) {
  short v;
  asm volatile("ld.global.L1::no_allocate.s16 %0, [%1];" : "=h"(v) : "l"(p));
  return int(v);
}

// kObs:  fused observation epilogue (TMA bulk store of [tile, 6+4P] f32 rows).
// kLean: only reward_f32 and the three ballot masks are produced (io.status, io.taken,
//        io.reward_f64, io.ep_return, io.stats are all NULL) -- the throughput configuration.
#ifndef SHIPENV_STEP_MIN_BLOCKS
#define SHIPENV_STEP_MIN_BLOCKS 5
#endif
template <bool kObs, bool kLean>
__global__ void __launch_bounds__(kStepBlock, SHIPENV_STEP_MIN_BLOCKS) step_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int P = w.h->P, W = w.h->W, H = w.h->H, cap = w.h->capacity;
  const int tid = threadIdx.x;
  const int n = int(a.n);                         // the host guarantees n < 2^31
  const int n_tiles = (n + kThreads - 1) / kThreads;
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;
  const bool fast_enabled = a.word0 == 0u && P >= 2;

  // observation tiles: [obs_bufs][kThreads][width] floats after the world blob; the constant
  // port tail of every row is written once, only the 6-float head changes per tile.
  float* obs_tile = nullptr;
  const int width = a.obs_width;
  if (kObs) {
    obs_tile = reinterpret_cast<float*>(smem + ((a.world_bytes + 127u) & ~127u) + ((kPipeBytes + 127u) & ~size_t(127)));
    for (int b = 0; b < a.obs_bufs && tid < kThreads; ++b) {
      float* row = obs_tile + (size_t(b) * kThreads + tid) * width;
      for (int c = 6; c < width; ++c) row[c] = w.h->obs_tail[c - 6];
    }
  }
  int obs_buf = 0;

  // Input pipeline (warp-specialised): warp 8 is the producer.  Its lane 0 keeps kStages
  // tiles of ship records / clocks / actions in flight with TMA bulk copies that complete on
  // the slot's `full` mbarrier; the 8 consumer warps wait on `full`, copy their env into
  // registers and arrive on `empty`, which lets the producer refill the slot.  No register
  // holds prefetched data and no consumer instruction waits on an HBM load.
  TileStage* stage = reinterpret_cast<TileStage*>(smem + ((a.world_bytes + 127u) & ~127u));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage) + sizeof(TileStage) * kStages);
  uint64_t* empty_bar = full_bar + kStages;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], kThreads / 32); }
    fence_proxy_async();
  }
  __syncthreads();
  if (tid >= kThreads) {             // ---- producer warp ----
    if (tid == kThreads) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int st = it % kStages;
        if (it >= kStages) mbar_wait(&empty_bar[st], uint32_t(it / kStages - 1) & 1u);
        const int base = tile * kThreads;
        const int rows8 = min(kThreads, n - base) & ~7;   // whole 16-byte units of every array
        if (rows8 > 0) {
          mbar_expect_tx(&full_bar[st], uint32_t(rows8) * 22u);
          tma_load_bulk(stage[st].ships, a.ships + base, uint32_t(rows8) * 16u, &full_bar[st]);
          tma_load_bulk(stage[st].clock, a.clock + base, uint32_t(rows8) * 4u, &full_bar[st]);
          tma_load_bulk(stage[st].act, a.actions + base, uint32_t(rows8) * 2u, &full_bar[st]);
        } else {
          mbar_arrive(&full_bar[st]);
        }
      }
    }
    return;
  }
  const int lane = tid & 31;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int i = tile * kThreads + tid;
    const bool live = i < n;
    const int st = it % kStages;
    uint4 rec = make_uint4(0, 0, 0, 0);
    uint32_t c = 0;
    int act = 0;
    mbar_wait(&full_bar[st], uint32_t(it / kStages) & 1u);
    if (tid < (min(kThreads, n - tile * kThreads) & ~7)) {
      rec = stage[st].ships[tid]; c = stage[st].clock[tid]; act = stage[st].act[tid];
    } else if (live) {               // ragged tail of the last tile
      rec = ld_stream_u4(a.ships + i); c = ld_stream_u32(a.clock + i); act = ld_stream_s16(a.actions + i);
    }
    __syncwarp();                    // every lane of this warp has read slot st
    if (lane == 0) mbar_arrive(&empty_bar[st]);
    uint32_t flags = 0;          // kFlag* bits of this env
    bool general = false;
    uint4 rec_out = rec;         // record after the step (kObs only consumer besides the store)
    if (live) {
      // ---- the common case, straight-line: a MOVE with no cargo-loss draw and no arrival ----
      const int x = rec.x & 255u, y = (rec.x >> 8) & 255u, dest = rec.x >> 24;
      const int nx = x + (act == 3) - (act == 1), ny = y + (act == 2) - (act == 0);
      if (!fast_enabled || unsigned(act) >= 4u || dest == kNone) {
        general = true;
      } else if (unsigned(nx) >= unsigned(W) || unsigned(ny) >= unsigned(H)) {
        flags = kFlagError;                                                          // :284, nothing drawn
        if (a.io.reward_f32) a.io.reward_f32[i] = 0.f;
        if (!kLean) {
          if (a.io.reward_f64) a.io.reward_f64[i] = 0.0;
          if (a.io.status) a.io.status[i] = uint8_t(SHIPENV_ST_OFF_MAP);
          if (a.io.taken) a.io.taken[i] = int16_t(act);
        }
      } else {
        uint32_t wd[4];
        const uint64_t id = a.env_id_base + uint64_t(uint32_t(i));
        philox_rk(0u, c, uint32_t(id), uint32_t(id >> 32), a.round_key, wd);
        const double u = __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), unit_double(wd[0])));   // :103-104
        const double cost = __dadd_rn(1.0, u);
        const double fuel = __hiloint2double(int(rec.w), int(rec.z));
        const int cargo = int(rec.y);
        const bool no_fuel = fuel < cost;                                            // :288
        const bool water = w.passable(nx, ny);                                       // :293
        const uint32_t pd = w.h->port[dest];
        const int tx = pd & 255u, ty = (pd >> 8) & 255u;
        const int d2_prev = (x - tx) * (x - tx) + (y - ty) * (y - ty);
        const int d2_new = (nx - tx) * (nx - tx) + (ny - ty) * (ny - ty);
        const int px = water ? nx : x, py = water ? ny : y;
        const bool trigger = cargo >= cap || (cargo >= 0 && wd[1] <= w.h->loss_thr[cargo]);   // :318-320
        const bool arrive = px == tx && py == ty;                                    // :325
        uint32_t c1 = c + 1u;
        const bool ended = no_fuel || (a.max_steps > 0 && int(c1 & 0xFFFFu) >= a.max_steps);
        // reset draws continue the cell: words 2 and 3 (randint(0, P-1) twice, :232-233)
        const int ro = int(__umulhi(wd[2], uint32_t(P))), rd = int(__umulhi(wd[3], uint32_t(P)));
        const bool do_reset = auto_reset && ended;
        if (trigger || arrive || (do_reset && ro == rd)) {
          general = true;   // needs more draws than this block: cargo loss, delivery, destination re-draw
        } else {
          const double reward = w.h->move_reward[(int(no_fuel) << 2) | (int(!water) << 1) | int(d2_prev > d2_new)];
          const double fuel1 = water ? __dsub_rn(fuel, cost) : fuel;                 // :298
          rec_out.x = uint32_t(px) | (uint32_t(py) << 8) | (rec.x & 0xFFFF0000u);
          rec_out.z = uint32_t(__double2loint(fuel1)); rec_out.w = uint32_t(__double2hiint(fuel1));
          if (do_reset) {                                                            // :227-243
            const uint32_t po = w.h->port[ro];
            rec_out.x = (po & 0xFFFFu) | (uint32_t(ro) << 16) | (uint32_t(rd) << 24);
            rec_out.y = uint32_t(w.h->cargo0);
            rec_out.z = uint32_t(__double2loint(w.h->fuel0)); rec_out.w = uint32_t(__double2hiint(w.h->fuel0));
            c1 = ((c >> 16) + 1u) << 16;
          }
          flags = (no_fuel ? kFlagDone : 0u) | (do_reset ? kFlagReset : 0u);
          st_stream_u4(a.ships + i, rec_out);
          a.clock[i] = c1;
          if (a.io.reward_f32) a.io.reward_f32[i] = float(reward);
          if (!kLean) {
            track_episode(a.io, i, reward, do_reset, c);
            if (a.io.reward_f64) a.io.reward_f64[i] = reward;
            if (a.io.status) a.io.status[i] = uint8_t(SHIPENV_ST_OK);
            if (a.io.taken) a.io.taken[i] = int16_t(act);
          }
        }
      }
    }
    if (__any_sync(0xFFFFFFFFu, general)) {
      if (general) {
        flags = step_general(a, smem, i);
        if (kObs) rec_out = a.ships[i];
      }
    }
    // done / reset / error masks: one ballot word per warp
    const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, flags & kFlagDone);
    const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, flags & kFlagReset);
    const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, flags & kFlagError);
    if ((tid & 31) == 0 && live) {
      const int word = i >> 5;
      if (a.io.done_bits) a.io.done_bits[word] = done_w;
      if (a.io.reset_bits) a.io.reset_bits[word] = reset_w;
      if (a.io.error_bits) a.io.error_bits[word] = err_w;
    }
    if (kObs) {
      // head -> shared tile, then one TMA bulk store of the whole contiguous [rows, width] block
      float* tile_base = obs_tile + size_t(obs_buf) * kThreads * width;
      if (a.obs_bufs == 1) {  // the only buffer may still be read by the previous bulk store
        if (tid == 0) tma_store_wait_read<0>();
        consumer_sync();
      }
      if (live) obs_head(unpack(rec_out), tile_base + size_t(tid) * width);
      fence_proxy_async();
      consumer_sync();
      const int64_t row0 = int64_t(tile) * kThreads;
      const int rows = int(min(int64_t(kThreads), a.n - row0));
      float* gdst = a.io.obs + row0 * width;
      const uint32_t bytes = uint32_t(rows) * uint32_t(width) * 4u;
      if ((bytes & 15u) == 0) {
        if (tid == 0) {
          tma_store_bulk(gdst, tile_base, bytes);
          tma_store_commit();
          if (a.obs_bufs == 2) tma_store_wait_read<1>();  // the other buffer is free again
        }
      } else {  // ragged tail whose byte count is not a multiple of 16: plain coalesced stores
        for (int j = tid; j < rows * width; j += kThreads) gdst[j] = tile_base[j];
      }
      if (a.obs_bufs == 2) { obs_buf ^= 1; consumer_sync(); }
    }
  }
  if (kObs && tid == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------
// step, throughput specialisation: reward_f32 + the three masks only, word0 == 0, P >= 2.
//
// Same TMA producer / consumer pipeline as step_kernel, but the streaming loop contains the
// straight-line MOVE case only.  Lanes that need anything else (non-move action, cargo-loss
// draw, arrival, destination re-draw, unset destination) push their env index on a per-warp
// shared-memory list; the warp drains the list through step_general OUTSIDE the streaming
// loop, 32 deferred envs per call, and ORs their done/reset/error bits into the mask words.
// ------------------------------------------------------------------------------------------
constexpr int kDeferCap = 64;   // per-warp deferred-env list per env-per-thread; a tile adds at most 32 of each

// kEpt envs per thread: a tile is 256 * kEpt envs; thread t owns envs t, t + 256, ... of the tile.  Two
// independent env chains per thread interleave (ILP) and share the per-tile bookkeeping.
template <int kEpt>
struct alignas(16) FastStage {
  uint4 ships[kThreads * kEpt];
  uint32_t clock[kThreads * kEpt];
  int16_t act[kThreads * kEpt];
};
template <int kEpt>
constexpr size_t fast_smem_bytes() {   // ring slots + full/empty mbarriers + tile ids + deferred lists
  return sizeof(FastStage<kEpt>) * kStages + 2 * 8 * kStages + 4 * kStages + size_t(kDeferCap) * kEpt * 4 * (kThreads / 32);
}
#ifndef SHIPENV_FAST_EPT
#define SHIPENV_FAST_EPT 1
#endif

#ifndef SHIPENV_FAST_MIN_BLOCKS
#define SHIPENV_FAST_MIN_BLOCKS 3
#endif
// kPolicy: every action of the launch is SHIPENV_ACT_POLICY (SHIPENV_F_ALL_POLICY); the straight-line case is then
// `sample_action` choosing a MOVE (word 0) followed by that move (words 1, 2); `actions` is never read.
// kObs: fused observation rows (utils/preprocessing.py:25-62): lanes that need the general path are resolved on the
// spot (their row must be final before the tile's TMA bulk store), heads go to a shared [256, 6+4P] tile whose
// constant port tail is written once, one bulk store per tile, two tiles in flight.
template <bool kPolicy, bool kObs, int kEpt>
__global__ void __launch_bounds__(kStepBlock, SHIPENV_FAST_MIN_BLOCKS) step_fast_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  const WorldView w(smem);
  const int tid = threadIdx.x;
  const int n = int(a.n);
  static_assert(!kObs || kEpt == 1, "the observation variant keeps one env per thread");
  constexpr int kTile = kThreads * kEpt;
  constexpr int kCap = kDeferCap * kEpt;
  using Stage = FastStage<kEpt>;
  const int full_tiles = n / kTile;                  // the streaming loop takes complete tiles only
  Stage* stage = reinterpret_cast<Stage*>(smem + ((a.world_bytes + 127u) & ~127u));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage) + sizeof(Stage) * kStages);
  uint64_t* empty_bar = full_bar + kStages;
  volatile int* stage_tile = reinterpret_cast<volatile int*>(empty_bar + kStages);   // tile index held by each slot
  int* defer_all = const_cast<int*>(stage_tile) + kStages;
  if (tid == 0) {
    mbar_init(&world_bar, 1);
#pragma unroll
    for (int st = 0; st < kStages; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], kThreads / 32); }
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {                    // the world blob and the first tiles are fetched concurrently
    mbar_expect_tx(&world_bar, a.world_bytes);
    tma_load_bulk(smem, a.world, a.world_bytes, &world_bar);
  }
  // Programmatic dependent launch: let the next step's grid start its prologue while this one runs, and
  // do not touch ships / clock / tickets before the previous grid has completed and flushed its stores.
  // The CTA's first tile is its block index (no ticket needed); pulling it towards L2 before the wait is only a
  // hint -- L2 stays coherent with the stores the previous grid may still be making to those lines.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == kThreads && int(blockIdx.x) < full_tiles) {
    const size_t base = size_t(blockIdx.x) * kTile;
    tma_prefetch_l2(a.ships + base, kTile * 16u);
    tma_prefetch_l2(a.clock + base, kTile * 4u);
    if (!kPolicy) tma_prefetch_l2(a.actions + base, kTile * 2u);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid >= kThreads) {             // ---- producer warp: lane 0 streams tiles in with TMA ----
    if (tid == kThreads) {
      // Tiles are handed out by a global ticket counter, so SMs that run ahead simply take more tiles
      // (no tail imbalance).  The counter is never reset: the host passes the ticket this launch starts at.
      // tickets 0..grid-1 are the block indices; later ones come from the counter (offset by the grid size)
      uint32_t ticket = blockIdx.x;
      for (int it = 0;; ++it) {
        const int st = it % kStages;
        if (it >= kStages) mbar_wait_backoff(&empty_bar[st], uint32_t(it / kStages - 1) & 1u);
        if (ticket >= uint32_t(full_tiles)) {          // out of tiles: publish the end marker
          stage_tile[st] = -1;
          mbar_arrive(&full_bar[st]);
          break;
        }
        const int base = int(ticket) * kTile;
        stage_tile[st] = int(ticket);
        mbar_expect_tx(&full_bar[st], kTile * (kPolicy ? 20u : 22u));
        tma_load_bulk(stage[st].ships, a.ships + base, kTile * 16u, &full_bar[st]);
        tma_load_bulk(stage[st].clock, a.clock + base, kTile * 4u, &full_bar[st]);
        if (!kPolicy) tma_load_bulk(stage[st].act, a.actions + base, kTile * 2u, &full_bar[st]);
        ticket = atomicAdd(a.tile_counter, 1u) - a.tile_base + gridDim.x;   // in flight while the slot is awaited
      }
    }
    return;
  }
  mbar_wait(&world_bar, 0);
  float* obs_tile = nullptr;
  const int width = a.obs_width;
  if (kObs) {
    obs_tile = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(defer_all) + kCap * 4 * (kThreads / 32));
    for (int b = 0; b < 2; ++b) {
      float* row = obs_tile + (size_t(b) * kThreads + tid) * width;
      for (int c = 6; c < width; ++c) row[c] = w.h->obs_tail[c - 6];
    }
  }
  int obs_buf = 0;
  // ---- consumers ----
  // Everything the streaming loop touches in shared memory is addressed through 32-bit shared
  // addresses computed once (ld.shared with immediate offsets, no generic-address arithmetic).
  const int lane = tid & 31;
  int* defer = defer_all + (tid >> 5) * kCap;
  int n_defer = 0;                                   // warp-uniform
  const uint32_t wb = smem_u32(smem);                // world blob
  const uint32_t bits_b = wb + uint32_t(sizeof(WorldHdr));
  const uint32_t W = uint32_t(w.h->W), H = uint32_t(w.h->H), P = uint32_t(w.h->P);
  const uint32_t cap = uint32_t(w.h->capacity);
  const uint32_t wpr = uint32_t(w.h->wpr);
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;
  const uint32_t trunc_at = a.max_steps > 0 ? uint32_t(a.max_steps) : 0x10000u;   // never reached when disabled
  const uint32_t id_lo = uint32_t(a.env_id_base), id_hi = uint32_t(a.env_id_base >> 32);
  // lanes 0..2 each own one mask array: one predicated store writes the warp's three ballot words
  uint32_t* const mask_ptr = lane == 0 ? a.io.done_bits : lane == 1 ? a.io.reset_bits : a.io.error_bits;
  const bool mask_lane = lane < 3;
  const uint32_t reset_rec_y = uint32_t(w.h->cargo0);
  const uint32_t reset_rec_z = uint32_t(__double2loint(w.h->fuel0)), reset_rec_w = uint32_t(__double2hiint(w.h->fuel0));
  const uint32_t stage_b = smem_u32(stage);
  const uint32_t my_ship = stage_b + uint32_t(tid) * 16u;
  const uint32_t my_clock = stage_b + uint32_t(kTile) * 16u + uint32_t(tid) * 4u;
  const uint32_t my_act = stage_b + uint32_t(kTile) * 20u + uint32_t(tid) * 2u;
  const uint32_t full_b = smem_u32(full_bar), empty_b = smem_u32(empty_bar);

  const uint32_t tile_b = smem_u32(const_cast<int*>(stage_tile));
  bool more = true;                                  // warp-uniform: the end marker has not been seen yet
  uint32_t st = 0, parity = 0;                       // next pipeline slot and its phase parity
  while (true) {
    // ------------------------------ streaming loop: no calls in here ------------------------------
    while (more && n_defer <= kCap - 32 * kEpt) {
      mbar_wait_s(full_b + st * 8u, parity);
      const int tile = int(lds_u32(tile_b + st * 4u));
      if (tile < 0) { more = false; break; }
      const uint32_t so = st * uint32_t(sizeof(Stage));
      uint4 rec_e[kEpt];
      uint32_t c_e[kEpt];
      int act_e[kEpt];
#pragma unroll
      for (int e = 0; e < kEpt; ++e) {
        rec_e[e] = lds_u4(my_ship + so + uint32_t(e) * kThreads * 16u);
        c_e[e] = lds_u32(my_clock + so + uint32_t(e) * kThreads * 4u);
        act_e[e] = kPolicy ? SHIPENV_ACT_POLICY : lds_s16(my_act + so + uint32_t(e) * kThreads * 2u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_s(empty_b + st * 8u);
      st = (st + 1u) & uint32_t(kStages - 1);
      parity ^= uint32_t(st == 0u);
#pragma unroll
      for (int e = 0; e < kEpt; ++e) {
      const int i = tile * kTile + e * kThreads + tid;
      const uint4 rec = rec_e[e];
      const uint32_t c = c_e[e];
      const int act = act_e[e];

      uint4 fin = rec;               // kObs: the record the observation row is built from
      const uint32_t x = rec.x & 255u, y = (rec.x >> 8) & 255u, dest = rec.x >> 24;
      // Philox block 0 of the (env, clock) cell.  Explicit action: words 0,1 = fuel-cost and cargo-loss draws,
      // 2,3 = reset draws.  Policy: word 0 = choice([N,E,S,W]) (:165-167), 1,2 = the move's draws.
      uint32_t wd[4];
      {
        const uint32_t lo = id_lo + uint32_t(i);
        philox_rk(0u, c, lo, id_hi + uint32_t(lo < id_lo), a.round_key, wd);
      }
      const uint32_t w_cost = kPolicy ? wd[1] : wd[0], w_loss = kPolicy ? wd[2] : wd[1];
      const double fuel = __hiloint2double(int(rec.w), int(rec.z));
      // N,E,S,W = (0,-1),(-1,0),(0,1),(1,0): odd actions move along x, even ones along y
      const int m = kPolicy ? int(__umulhi(wd[0], 4u)) : (act & 3);
      const int dx = (m & 1) * (m - 2), dy = ((m & 1) ^ 1) * (m - 1);
      const uint32_t nx = x + uint32_t(dx), ny = y + uint32_t(dy);
      bool is_move;
      if (kPolicy) {   // sample_action (:245-263) picks a MOVE unless the ship sits on a port with no cargo / no fuel
        const bool on_port = (lds_u32(bits_b + (uint32_t(w.h->n_bit_words) + x * wpr + (y >> 5)) * 4u) >> (y & 31u)) & 1u;
        is_move = dest != uint32_t(kNone) && !(on_port && (rec.y == 0u || fuel == 0.0));
      } else {
        is_move = unsigned(act) < 4u && dest != uint32_t(kNone);
      }
      const bool in_map = nx < W && ny < H;          // unsigned: -1 wraps to a huge value  (:284)
      const double cost = __dadd_rn(1.0, __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), unit_double(w_cost))));   // :103-104
      const bool no_fuel = fuel < cost;                                                   // :288
      const uint32_t cell = in_map ? nx * wpr + (ny >> 5) : 0u;
      const bool water = (lds_u32(bits_b + cell * 4u) >> (ny & 31u)) & 1u;                // :293
      const uint32_t pd = lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + (dest & 31u) * 4u);
      const uint32_t tx = pd & 255u, ty = (pd >> 8) & 255u;
      // closer <=> d_prev - d_new > 0 <=> (x - tx) * dx + (y - ty) * dy < 0 for a unit move   (:309-312)
      const bool closer = (int(x) - int(tx)) * dx + (int(y) - int(ty)) * dy < 0;
      const uint32_t px = water ? nx : x, py = water ? ny : y;
      // loss_thr[cap] = 0xFFFFFFFF: cargo >= capacity always draws; negative cargo goes the general way
      const bool trigger = w_loss <= lds_u32(wb + uint32_t(offsetof(WorldHdr, loss_thr)) + min(uint32_t(rec.y), cap) * 4u);  // :318-320
      const bool arrive = px == tx && py == ty;                                           // :325
      const uint32_t c1 = c + 1u;
      const bool do_reset = auto_reset && (no_fuel || (c1 & 0xFFFFu) >= trunc_at);
      const uint32_t ro = __umulhi(wd[2], P), rd = __umulhi(wd[3], P);                    // :232-233
      // policy: an off-map choice is re-drawn and a reset needs words 3 and 4 -- both go the general way
      const bool general = kPolicy ? (!is_move || !in_map || trigger || arrive || do_reset)
                                   : (!is_move || (in_map && (trigger || arrive || (do_reset && ro == rd))));
      const bool ok = is_move && in_map && !general;
      const bool off_map = !kPolicy && is_move && !in_map;
      {
        const double fuel1 = water ? __dsub_rn(fuel, cost) : fuel;                        // :298
        uint4 out;
        out.x = px | (py << 8) | (rec.x & 0xFFFF0000u);
        out.y = rec.y;
        out.z = uint32_t(__double2loint(fuel1)); out.w = uint32_t(__double2hiint(fuel1));
        uint32_t c_out = c1;
        if (do_reset) {                                                                   // :227-243
          out.x = (lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + ro * 4u) & 0xFFFFu) | (ro << 16) | (rd << 24);
          out.y = reset_rec_y; out.z = reset_rec_z; out.w = reset_rec_w;
          c_out = ((c >> 16) + 1u) << 16;
        }
        const float r = lds_f32(wb + uint32_t(offsetof(WorldHdr, move_reward_f32)) +
                                ((uint32_t(no_fuel) << 2) | (uint32_t(!water) << 1) | uint32_t(closer)) * 4u);
        if (ok) {
          st_stream_u4(a.ships + i, out);
          a.clock[i] = c_out;
        }
        if (!general) a.io.reward_f32[i] = ok ? r : 0.f;
        if (kObs) fin = ok ? out : rec;
      }
      const uint32_t gen_w = __ballot_sync(0xFFFFFFFFu, general);
      uint32_t gflags = 0;
      if (kObs) {
        if (gen_w && general) {        // resolve now: the observation row of this tile must be final
          gflags = step_general_inline(a, smem, i);
          fin = a.ships[i];
        }
      } else if (gen_w) {
        if (general) defer[n_defer + __popc(gen_w & ((1u << lane) - 1u))] = i;
        n_defer += __popc(gen_w);
      }
      const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, (ok && no_fuel) || (gflags & kFlagDone));
      const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, (ok && do_reset) || (gflags & kFlagReset));
      const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, off_map || (gflags & kFlagError));
      {
        const uint32_t mask_v = lane == 0 ? done_w : lane == 1 ? reset_w : err_w;
        uint32_t* const mp = mask_ptr + (i >> 5);
        asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.global.u32 [%0], %1;\n}\n"
                     :: "l"(mp), "r"(mask_v), "r"(uint32_t(mask_lane)) : "memory");
      }
      if (kObs) {
        float* tile_base = obs_tile + size_t(obs_buf) * kThreads * width;
        obs_head(unpack(fin), tile_base + size_t(tid) * width);
        fence_proxy_async();
        consumer_sync();
        if (tid == 0) {
          tma_store_bulk(a.io.obs + size_t(tile) * kTile * width, tile_base, uint32_t(kThreads) * uint32_t(width) * 4u);
          tma_store_commit();
          tma_store_wait_read<1>();    // the other buffer is free again
        }
        obs_buf ^= 1;
        consumer_sync();
      }
      }  // envs of this thread
    }
    // ------------------------------ drain the deferred envs, a warp-full at a time ------------------------------
    if (n_defer == 0 && !more) break;
    __syncwarp();
    while (n_defer > 0) {
      const int k = min(n_defer, 32);
      if (lane < k) {
        const int i = defer[n_defer - k + lane];
        const uint32_t f = step_general_inline(a, smem, i);   // inlined: an ABI call would pin loop state to the stack
        const uint32_t bit = 1u << (i & 31);
        if (f & kFlagDone) atomicOr(a.io.done_bits + (i >> 5), bit);
        if (f & kFlagReset) atomicOr(a.io.reset_bits + (i >> 5), bit);
        if (f & kFlagError) atomicOr(a.io.error_bits + (i >> 5), bit);
      }
      n_defer -= k;
      __syncwarp();
    }
  }
  if (kObs && tid == 0) tma_store_wait_all();
  // the ragged last tile (n % 256 envs) never enters the pipeline: its owner CTA takes the general path for it
  if (full_tiles * kTile < n && int(blockIdx.x) == full_tiles % int(gridDim.x)) {
#pragma unroll 1
    for (int e = 0; e < kEpt; ++e) {
    const int i = full_tiles * kTile + e * kThreads + tid;
    if (i - lane >= n) break;          // warp-uniform
    const uint32_t f = i < n ? step_general_inline(a, smem, i) : 0u;
    if (kObs && i < n) {               // the ragged tile's rows are written with plain stores
      float* row = a.io.obs + size_t(i) * width;
      obs_head(unpack(a.ships[i]), row);
      for (int c = 6; c < width; ++c) row[c] = w.h->obs_tail[c - 6];
    }
    const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, f & kFlagDone);
    const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, f & kFlagReset);
    const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, f & kFlagError);
    if (lane < 3) mask_ptr[i >> 5] = lane == 0 ? done_w : lane == 1 ? reset_w : err_w;
    }
  }
}

// ------------------------------------------------------------------------------------------
// reset / observe / sample_action
// ------------------------------------------------------------------------------------------
struct ResetArgs {
  const void* world; uint32_t world_bytes;
  uint4* ships; uint32_t* clock; const uint32_t* mask_bits;
  uint64_t seed, env_id_base; int64_t n;
  uint32_t round_key[20];
};

__global__ void __launch_bounds__(kThreads) reset_kernel(const __grid_constant__ ResetArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int64_t n_tiles = (a.n + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * kThreads + threadIdx.x;
    if (i >= a.n) continue;
    if (a.mask_bits && !((a.mask_bits[i >> 5] >> (i & 31)) & 1u)) continue;
    const uint32_t c = a.clock[i];
    ShipR s;
    Rng g;
    g.seek(a.round_key, c, a.env_id_base + uint64_t(i), 0u);
    env_reset(w, s, g);
    st_stream_u4(a.ships + i, pack(s));
    a.clock[i] = ((c >> 16) + 1u) << 16;
  }
}

struct ObserveArgs {
  const void* world; uint32_t world_bytes;
  const uint4* ships; float* obs; int64_t n; int obs_width;
};

__global__ void __launch_bounds__(kThreads) observe_kernel(const ObserveArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int width = a.obs_width, tid = threadIdx.x;
  float* tile = reinterpret_cast<float*>(smem + ((a.world_bytes + 127u) & ~127u));
  for (int c = 6; c < width; ++c) tile[size_t(tid) * width + c] = w.h->obs_tail[c - 6];
  const int64_t n_tiles = (a.n + kThreads - 1) / kThreads;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t i = t * kThreads + tid;
    if (tid == 0) tma_store_wait_read<0>();
    __syncthreads();
    if (i < a.n) obs_head(unpack(a.ships[i]), tile + size_t(tid) * width);
    fence_proxy_async();
    __syncthreads();
    const int64_t row0 = t * kThreads;
    const int rows = int(min(int64_t(kThreads), a.n - row0));
    float* gdst = a.obs + row0 * width;
    const uint32_t bytes = uint32_t(rows) * uint32_t(width) * 4u;
    if ((bytes & 15u) == 0) {
      if (tid == 0) { tma_store_bulk(gdst, tile, bytes); tma_store_commit(); }
    } else {
      for (int j = tid; j < rows * width; j += kThreads) gdst[j] = tile[j];
      __syncthreads();
    }
  }
  if (tid == 0) tma_store_wait_all();
}

struct SampleArgs {
  const void* world; uint32_t world_bytes;
  const uint4* ships; const uint32_t* clock;
  int16_t* actions; uint8_t* words_used; uint8_t* status;
  uint64_t seed, env_id_base; int64_t n; uint32_t word0;
  uint32_t round_key[20];
};

__global__ void __launch_bounds__(kThreads) sample_kernel(const __grid_constant__ SampleArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int64_t n_tiles = (a.n + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * kThreads + threadIdx.x;
    if (i >= a.n) continue;
    const ShipR s = unpack(a.ships[i]);
    Rng g;
    g.seek(a.round_key, a.clock[i], a.env_id_base + uint64_t(i), a.word0);
    int cat = kCatJunk, val = 0;
    int st = w.h->P == 0 ? SHIPENV_ST_NO_PORTS : sample_action(w, s, g, cat, val);
    a.actions[i] = st == SHIPENV_ST_OK ? int16_t(encode_action(cat, val, w.h->P)) : int16_t(-2);
    if (a.words_used) a.words_used[i] = uint8_t(g.i - a.word0);
    if (a.status) a.status[i] = uint8_t(st);
  }
}

// ------------------------------------------------------------------------------------------
// DQN valid-action mask (agents/dqn.py:125-175), one env per thread, ceil(A/32) words per env
// ------------------------------------------------------------------------------------------
struct ValidArgs {
  const void* world; uint32_t world_bytes;
  const uint4* ships; uint32_t* mask; int64_t n; int words;
};

// bits [lo, hi) of a bit string, restricted to word `wi`
__device__ __forceinline__ uint32_t range_bits(int lo, int hi, int wi) {
  const int a = max(lo - wi * 32, 0), b = min(hi - wi * 32, 32);
  if (a >= b) return 0u;
  return (b == 32 ? 0xFFFFFFFFu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
}

__global__ void __launch_bounds__(kThreads) valid_actions_kernel(const ValidArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int P = w.h->P;
  const int64_t n_tiles = (a.n + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * kThreads + threadIdx.x;
    if (i >= a.n) continue;
    const ShipR s = unpack(a.ships[i]);
    const uint32_t key = uint32_t(s.x) | (uint32_t(s.y) << 8);
    uint32_t select = 0;                              // bit k: port k selectable (:141-156)
    for (int k = 0; k < P; ++k)
      if (k != s.origin && (w.h->port[k] & 0xFFFFu) == key) select |= 1u << k;
    const int cur = current_port(w, s);                // :159-173
    const int n_cargo = cur >= 0 ? min(int(w.h->port[cur] >> 24), SHIPENV_N_CARGO_ACTIONS - 1) : 0;
    const int n_fuel = cur >= 0 ? min(int((w.h->port[cur] >> 16) & 255u), SHIPENV_N_FUEL_ACTIONS - 1) : 0;
    const int c0 = 4 + P + 1, f0 = 4 + P + SHIPENV_N_CARGO_ACTIONS + 1;   // first amount is 1
    for (int wi = 0; wi < a.words; ++wi) {
      uint32_t m = range_bits(0, 4, wi);               // the four moves are always valid (:137-139)
      m |= range_bits(c0, c0 + n_cargo, wi) | range_bits(f0, f0 + n_fuel, wi);
      for (int k = 0; k < P; ++k)
        if ((select >> k) & 1u) m |= range_bits(4 + k, 5 + k, wi);
      a.mask[i * a.words + wi] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------
// discretize_state (utils/preprocessing.py:65-90): six int16 per env, no world needed
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int bucket(double v, double width) {   // min(int(v / width), 4), int() truncates
  const double q = __ddiv_rn(v, width);
  const int b = q >= 2147483520.0 ? 2147483647 : q <= -2147483520.0 ? -2147483647 : __double2int_rz(q);
  return max(min(b, 4), -32768);
}

__global__ void __launch_bounds__(kThreads) discretize_kernel(const uint4* __restrict__ ships, int16_t* __restrict__ keys,
                                                              int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kThreads) {
    const ShipR s = unpack(ships[i]);
    int16_t* k = keys + i * 6;
    k[0] = int16_t(s.x); k[1] = int16_t(s.y);
    k[2] = int16_t(bucket(s.fuel, 10.0));      // the state dict's "cargo" is fuel (shipping/environment.py:206)
    k[3] = int16_t(bucket(s.fuel, 20.0));
    k[4] = s.origin == kNone ? int16_t(-1) : int16_t(s.origin);
    k[5] = s.dest == kNone ? int16_t(-1) : int16_t(s.dest);
  }
}

// ------------------------------------------------------------------------------------------
// MCTS rollouts: the whole simulations x max_rollout_steps loop in one kernel
// ------------------------------------------------------------------------------------------
struct RolloutArgs {
  const void* world; uint32_t world_bytes;
  const uint4* roots; int64_t R; int S, L;
  double* returns; uint4* final; int32_t* steps; uint8_t* status;
  uint64_t seed, root_id_base;
  uint32_t round_key[20];   // of the rollout key: seed with k1 ^= "ROLL"
};

constexpr int kRolloutThreads = 128;

__global__ void __launch_bounds__(kRolloutThreads) rollout_kernel(const __grid_constant__ RolloutArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int64_t total = a.R * int64_t(a.S);
  const int64_t n_tiles = (total + kRolloutThreads - 1) / kRolloutThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t j = tile * kRolloutThreads + threadIdx.x;
    if (j >= total) continue;
    const int64_t r = j / a.S;
    const int sim = int(j - r * a.S);
    ShipR s = unpack(__ldg(a.roots + r));
    Rng g;
    g.seek(a.round_key, uint32_t(sim), a.root_id_base + uint64_t(r), 0u);
    double total_reward = 0.0;                     // agents/mcts.py:224
    bool done = false;
    int steps = 0, st = SHIPENV_ST_OK, fails = 0;
    while (!done && steps < a.L) {                 // agents/mcts.py:228
      if (w.h->P == 0) { st = SHIPENV_ST_NO_PORTS; break; }
      if (s.dest == kNone && current_port(w, s) < 0) { st = SHIPENV_ST_NO_DEST; break; }
      int cat, val;
      if (sample_action(w, s, g, cat, val) != SHIPENV_ST_OK) { st = SHIPENV_ST_SAMPLE_CRASH; break; }
      double reward = 0.0;
      const int e = do_action(w, s, g, cat, val, reward, done);
      if (e != SHIPENV_ST_OK) {                    // agents/mcts.py:234-236: try another
        done = false;
        if (++fails >= kMaxTries) { st = e; break; }
        continue;
      }
      fails = 0;
      total_reward = __dadd_rn(total_reward, reward);
      ++steps;
    }
    a.returns[j] = total_reward;
    if (a.final) a.final[j] = pack(s);
    if (a.steps) a.steps[j] = steps;
    if (a.status) a.status[j] = uint8_t(st);
  }
}

// ------------------------------------------------------------------------------------------
// MCTS: the whole choose_action (agents/mcts.py:240-292) per root in one thread, tree in local memory
// ------------------------------------------------------------------------------------------
constexpr int kMctsMaxNodes = SHIPENV_MCTS_MAX_SIMS + 2;
constexpr int kMctsThreads = 128;

struct MctsArgs {
  const void* world; uint32_t world_bytes;
  const uint4* roots; int64_t R; int S, L;
  double c_param;
  const double* log_table;     // device, log_table[n] = log(n)
  int16_t* action; int32_t* root_visits; int16_t* child_action; int32_t* child_visits; double* child_total;
  uint8_t* status;
  uint64_t seed, root_id_base;
  uint32_t round_key[20];
};

struct MctsTree {              // one root's tree (agents/mcts.py:13-40), indexed by node id
  double total[kMctsMaxNodes];
  uint16_t visits[kMctsMaxNodes];
  uint8_t parent[kMctsMaxNodes], action[kMctsMaxNodes], nchild[kMctsMaxNodes], moves_ok[kMctsMaxNodes];
  uint8_t child[kMctsMaxNodes][4];
};

// get_possible_actions (:50-89) is the four moves iff a destination is set and not (fuel == 0 with an origin)
__device__ __forceinline__ bool mcts_moves_ok(const ShipR& s) {
  return s.dest != kNone && !(s.fuel == 0.0 && s.origin != kNone);
}

// best_child (:109-128): UCB1, unvisited children first, np.argmax takes the first maximum
__device__ __forceinline__ int mcts_best_child(const MctsTree& t, int node, double c_param, const double* log_table) {
  int best = 0;
  double best_v = 0.0;
  const double two_log = __dmul_rn(2.0, log_table[t.visits[node]]);
  for (int k = 0; k < t.nchild[node]; ++k) {
    const int ch = t.child[node][k];
    double v;
    if (t.visits[ch] == 0) v = __longlong_as_double(0x7FF0000000000000LL);
    else {
      const double n = double(t.visits[ch]);
      v = __dadd_rn(__ddiv_rn(t.total[ch], n), __dmul_rn(c_param, __dsqrt_rn(__ddiv_rn(two_log, n))));
    }
    if (k == 0 || v > best_v) { best = ch; best_v = v; }
  }
  return best;
}

__global__ void __launch_bounds__(kMctsThreads) mcts_kernel(const __grid_constant__ MctsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  for (int64_t r = blockIdx.x * int64_t(kMctsThreads) + threadIdx.x; r < a.R; r += int64_t(gridDim.x) * kMctsThreads) {
    const ShipR root = unpack(__ldg(a.roots + r));
    MctsTree t;
    int n_nodes = 1, st = SHIPENV_ST_OK;
    t.total[0] = 0.0; t.visits[0] = 0; t.parent[0] = 255; t.action[0] = 255; t.nchild[0] = 0;
    t.moves_ok[0] = mcts_moves_ok(root);
    for (int sim = 0; sim < a.S && st == SHIPENV_ST_OK; ++sim) {
      ShipR s = root;                                                        // _create_env_copy, :257
      Rng g;
      g.seek(a.round_key, uint32_t(sim), a.root_id_base + uint64_t(r), 0u);
      int node = 0;
      bool done = false;
      double rew = 0.0;
      for (;;) {                                                             // selection, :260-268
        if (!t.moves_ok[node]) { st = SHIPENV_ST_MCTS_UNSUPPORTED; break; }
        if (t.nchild[node] != 4) break;
        node = mcts_best_child(t, node, a.c_param, a.log_table);
        const int e = do_action(w, s, g, kCatMove, t.action[node], rew, done);
        if (e != SHIPENV_ST_OK || done) break;
      }
      if (st != SHIPENV_ST_OK) break;
      // expansion, :271-279: untried actions in the order N, S, E, W (:80-87) = flat indices 0, 2, 1, 3
      uint32_t tried = 0;
      for (int k = 0; k < t.nchild[node]; ++k) tried |= 1u << t.action[t.child[node][k]];
      const int n_un = 4 - __popc(tried);
      if (n_un > 0) {
        int pick = g.below(n_un), act = 0;                                   // random.choice(untried), :274
        for (int k = 0; k < 4; ++k) {
          const int cand = (k == 1) ? 2 : (k == 2) ? 1 : k;
          if (!((tried >> cand) & 1u)) { if (pick == 0) { act = cand; break; } --pick; }
        }
        const int e = do_action(w, s, g, kCatMove, act, rew, done);
        if (e != SHIPENV_ST_OK) continue;                                    // :278-279: the simulation is skipped
        const int ch = n_nodes++;
        t.total[ch] = rew; t.visits[ch] = 0; t.parent[ch] = uint8_t(node); t.action[ch] = uint8_t(act);   // :144-146
        t.nchild[ch] = 0; t.moves_ok[ch] = mcts_moves_ok(s);
        t.child[node][t.nchild[node]++] = uint8_t(ch);
        node = ch;
      }
      // simulation: _rollout, :211-238, continuing the same word stream
      double total = 0.0;
      bool rdone = false;
      int steps = 0, fails = 0;
      while (!rdone && steps < a.L) {
        if (s.dest == kNone && current_port(w, s) < 0) { st = SHIPENV_ST_NO_DEST; break; }
        int cat, val;
        if (sample_action(w, s, g, cat, val) != SHIPENV_ST_OK) { st = SHIPENV_ST_SAMPLE_CRASH; break; }
        double rr = 0.0;
        const int e = do_action(w, s, g, cat, val, rr, rdone);
        if (e != SHIPENV_ST_OK) { rdone = false; if (++fails >= kMaxTries) { st = e; break; } continue; }
        fails = 0;
        total = __dadd_rn(total, rr);
        ++steps;
      }
      if (st != SHIPENV_ST_OK) break;
      for (int n = node; n != 255; n = t.parent[n]) { t.visits[n] += 1; t.total[n] = __dadd_rn(t.total[n], total); }   // :149-159
    }
    for (int k = 0; k < 4; ++k) {
      const bool has = k < t.nchild[0];
      const int ch = has ? t.child[0][k] : 0;
      a.child_action[4 * r + k] = has ? int16_t(t.action[ch]) : int16_t(-1);
      a.child_visits[4 * r + k] = has ? int32_t(t.visits[ch]) : 0;
      a.child_total[4 * r + k] = has ? t.total[ch] : 0.0;
    }
    a.root_visits[r] = t.visits[0];
    int act = -1;
    if (st == SHIPENV_ST_OK) {
      if (t.nchild[0] == 0) st = SHIPENV_ST_MCTS_NO_CHILDREN;                // :291-292
      else act = t.action[mcts_best_child(t, 0, 0.0, a.log_table)];          // :288-290
    }
    a.action[r] = int16_t(act);
    a.status[r] = uint8_t(st);
  }
}

}  // namespace shipenv
