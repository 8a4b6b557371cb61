// step_kernels.cuh -- the general step kernel: every output (status, taken action, f64 reward, episode
// statistics, fused observation rows), any action, any word0.  Same TMA producer / consumer ring as the
// throughput kernel; lanes that are not the straight-line MOVE case call the out-of-line general path.
#pragma once
#include "pipeline.cuh"

namespace shipenv {

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
  uint32_t* tile_counter;   // step_fast_kernel: global ticket counter (context-owned, zero between launches)
  uint32_t* chain;          // step_fast_kernel: the env group's chain array (shipenv_step_io.chain) or null
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
  double ep = __dadd_rn(__ldcg(er + i), reward);   // L2: a chained predecessor on another SM may have written it
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
//
// step_general_core is the transition itself on a ship held in registers (any action, any word0, auto-reset); the
// kernels differ only in where the ship comes from and where the outputs go.
struct GeneralOut { double reward; int status, taken; uint32_t flags; };

__device__ __forceinline__ GeneralOut step_general_core(const WorldView& w, const uint32_t* round_key, uint64_t env_id,
                                                        uint32_t word0, int max_steps, uint32_t step_flags, ShipR& s,
                                                        uint32_t& c, int act) {
  Rng g;
  g.seek(round_key, c, env_id, word0);
  GeneralOut o;
  o.reward = 0.0; o.taken = act; o.flags = 0;
  bool done = false;
  ShipR s1 = s;                        // a rejected action leaves the ship untouched
  if (act == SHIPENV_ACT_POLICY) {
    o.taken = -1;
    o.status = policy_step(w, s1, g, o.reward, done, o.taken);
  } else {
    int cat, val;
    decode_action(act, w.h->P, cat, val);
    o.status = do_action(w, s1, g, cat, val, o.reward, done);
  }
  if (o.status == SHIPENV_ST_OK) {
    uint32_t c1 = c + 1u;
    const bool ended = done || (max_steps > 0 && int(c1 & 0xFFFFu) >= max_steps);
    bool was_reset = false;
    if ((step_flags & SHIPENV_F_AUTO_RESET) && ended) {
      env_reset(w, s1, g);  // continues the word stream of this (env, clock) cell
      c1 = ((c >> 16) + 1u) << 16;
      was_reset = true;
    }
    s = s1;
    c = c1;
    o.flags = (done ? kFlagDone : 0u) | (was_reset ? kFlagReset : 0u);
  } else {
    o.reward = 0.0;
    o.flags = kFlagError;
  }
  return o;
}

__device__ __forceinline__ uint32_t step_general_inline(const StepArgs& a, const void* world_smem, int i) {
  const WorldView w(world_smem);
  // read through L2 (ld.global.cg): with chained launches the previous step of this env may have been written by
  // another SM while this SM's L1 still holds the line from the step before
  const uint4 rec = __ldcg(a.ships + i);
  const uint32_t c = __ldcg(a.clock + i);
  const int act = (a.flags & SHIPENV_F_ALL_POLICY) ? SHIPENV_ACT_POLICY : int(__ldcg(a.actions + i));
  ShipR s = unpack(rec);
  uint32_t c1 = c;
  const GeneralOut o = step_general_core(w, a.round_key, a.env_id_base + uint64_t(uint32_t(i)), a.word0, a.max_steps,
                                         a.flags, s, c1, act);
  const double reward = o.reward;
  const int status = o.status, taken = o.taken;
  const uint32_t flags = o.flags;
  if (status == SHIPENV_ST_OK) {
    track_episode(a.io, i, reward, flags & kFlagReset, c);
    a.ships[i] = pack(s);
    a.clock[i] = c1;
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
  const bool all_policy = a.flags & SHIPENV_F_ALL_POLICY;   // `actions` is not read (and may be null)

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
          mbar_expect_tx(&full_bar[st], uint32_t(rows8) * (all_policy ? 20u : 22u));
          tma_load_bulk(stage[st].ships, a.ships + base, uint32_t(rows8) * 16u, &full_bar[st]);
          tma_load_bulk(stage[st].clock, a.clock + base, uint32_t(rows8) * 4u, &full_bar[st]);
          if (!all_policy) tma_load_bulk(stage[st].act, a.actions + base, uint32_t(rows8) * 2u, &full_bar[st]);
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
      rec = stage[st].ships[tid]; c = stage[st].clock[tid];
      act = all_policy ? SHIPENV_ACT_POLICY : int(stage[st].act[tid]);
    } else if (live) {               // ragged tail of the last tile
      rec = ld_stream_u4(a.ships + i); c = ld_stream_u32(a.clock + i);
      act = all_policy ? SHIPENV_ACT_POLICY : ld_stream_s16(a.actions + i);
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
        if (kObs) rec_out = __ldcg(a.ships + i);
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

}  // namespace shipenv
