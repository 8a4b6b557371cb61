// step_fast_kernels.cuh -- the throughput kernels (reward_f32 + masks, optional fused observation rows,
// optional all-policy mode): the kernel bench.py measures.
#pragma once
#include "step_kernels.cuh"

namespace shipenv {

// ------------------------------------------------------------------------------------------
// step, throughput specialisation: reward_f32 + the three masks only, word0 == 0, P >= 2.
//
// Same TMA producer / consumer pipeline as step_kernel, but the streaming loop contains the
// straight-line MOVE case only.  Lanes that need anything else (non-move action, cargo-loss
// draw, arrival, episode end with auto-reset, unset destination) push their env index on a per-warp
// shared-memory list; the warp drains the list through step_general OUTSIDE the streaming
// loop, 32 deferred envs per call, and ORs their done/reset/error bits into the mask words.
// ------------------------------------------------------------------------------------------
// Deferred envs are drained right after the tile that deferred them (kDrainEarly) or when the list is about to
// overflow / the input has ended.
#ifndef SHIPENV_DRAIN_EARLY
#define SHIPENV_DRAIN_EARLY 0
#endif
constexpr bool kDrainEarly = SHIPENV_DRAIN_EARLY;
constexpr int kDeferCap = 64;   // per-warp deferred-env list per env-per-thread; a tile adds at most 32 of each

// ------------------------------------------------------------------------------------------
// Chained launches (shipenv_step_io.chain, SHIPENV_F_CHAINED).
//
// Consecutive step launches on one stream normally serialise at the launch boundary: grid t+1 may start its
// prologue early (programmatic dependent launch) but `griddepcontrol.wait` holds it until every CTA of grid t has
// exited and flushed (measured, tools/trace_launches.py: the go-ahead comes 1.1 us after the last exit, the first
// tile 0.9 us later, and the exits of grid t are spread over 5 us) -- 3-5 of 22 us at 2 Mi envs.  With a chain array
// the tiles are split STATICALLY: CTA b takes tiles b, b + G, b + 2G, ... in every launch with that grid size (and
// the CTA after the last tile's takes the ragged end).  The only dependency between consecutive launches is then
// CTA b on CTA b:
//   chain[0..1]            one 64-bit word: c << 16 | a, c = CTAs of all earlier chain-protocol launches on this env
//                          group, a = CTAs of the current launch that have arrived (the last one adds G to c)
//   chain[kChainDone + b]  c + G once CTA b of the launch that found c has written all of its envs (deferred ones
//                          included): one bar.sync of the consumer warps, then one gpu-scope release by one thread
// A launch with SHIPENV_F_CHAINED skips griddepcontrol.wait: its producer thread waits for chain[kChainDone + b] == c
// before the first TMA load.  CTAs of launch t+1 become resident as CTAs of launch t exit, and the CTAs drift apart
// in time from launch to launch: nothing ever waits for a whole grid again, and that some CTAs carry one tile more
// than others is absorbed by the next launch's CTAs.  (Tried and dropped: per-tile epochs published through a
// publisher warp -- 26 us per launch against 22.8 unchained; shares balanced to 32 envs with a short last tile --
// the per-lane liveness tests cost more than the balance gave; ticketed last rounds behind a mid-launch
// griddepcontrol.wait -- it re-synchronises the grids, 20.8 us against 19.8.)
// The host honours the flag only behind a chain-protocol launch of the same grid size and env count on the same
// array and stream (shipenv_abi.cu), so the split is the same on both sides of every chained boundary.
// c is race-free without host state (CUDA-graph replayable): every CTA learns c from its arrival (one atomic, whose
// result it holds before it triggers its dependents); grid t+1 cannot start before all CTAs of grid t have
// triggered, so all CTAs of grid t arrive on c | 0..G-1, the last of them moves the word on, and all CTAs of grid
// t+1 find c + G.  No deadlock: a CTA of grid t+1 runs only after every CTA of grid t has started, and those wait
// on grid t-1 only.
// ------------------------------------------------------------------------------------------
constexpr int kChainSeq = 0, kChainDone = 32;
constexpr int kChainMaxCtas = 2048;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Timeline instrumentation for tuning builds (-DSHIPENV_TRACE; tools/trace_launches.py): one record per CTA with
// globaltimer stamps of its start, the producer's go-ahead, the consumers' first tile, their last tile and the exit.
#ifdef SHIPENV_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ unsigned int g_trace_count = 0;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int trace_smid() {
  unsigned int v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
#define SHIPENV_TRACE_SET(slot, value) do { if (g_trace_buf) trace_s[slot] = (value); } while (0)
#else
#define SHIPENV_TRACE_SET(slot, value) do { } while (0)
#endif

// Episode end with auto-reset behind a straight-line move (:227-243): the terminal record and clock + 1 have been
// stored; draw the new episode from the words of the (env, clock) cell behind the move's draws -- explicit move:
// words 0, 1 -> 2..; policy: 0, 1, 2 -> 3.. -- as env_reset (transition.cuh) does, and overwrite record and clock.
template <bool kPolicy>
__device__ __forceinline__ uint4 reset_in_place(const StepArgs& a, const WorldView& w, int i, uint32_t c) {
  const uint64_t id = a.env_id_base + uint64_t(uint32_t(i));
  const uint32_t P = uint32_t(w.h->P);               // >= 2 with auto-reset (checked on the host)
  uint32_t wd[4];
  philox_rk(0u, c, uint32_t(id), uint32_t(id >> 32), a.round_key, wd);
  const uint32_t ro = __umulhi(kPolicy ? wd[3] : wd[2], P);                        // randint(0, P-1), :232
  uint32_t rd = kPolicy ? ro : __umulhi(wd[3], P);                                  // :233 (policy: word 4)
  for (uint32_t blk = 1u; rd == ro; ++blk) {         // `while dest == origin` (:236-237): further words of the cell
    philox_rk(blk, c, uint32_t(id), uint32_t(id >> 32), a.round_key, wd);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (rd == ro) rd = __umulhi(wd[k], P);
  }
  uint4 fresh;
  fresh.x = (w.h->port[ro] & 0xFFFFu) | (ro << 16) | (rd << 24);
  fresh.y = uint32_t(w.h->cargo0);
  fresh.z = uint32_t(__double2loint(w.h->fuel0)); fresh.w = uint32_t(__double2hiint(w.h->fuel0));
  st_stream_u4(a.ships + i, fresh);
  a.clock[i] = ((c >> 16) + 1u) << 16;
  return fresh;
}

// kEpt envs per thread: a tile is 256 * kEpt envs; thread t owns envs t, t + 256, ... of the tile.  Two
// independent env chains per thread interleave (ILP) and share the per-tile bookkeeping.
template <int kEpt>
struct alignas(16) FastStage {
  uint4 ships[kThreads * kEpt];
  uint32_t clock[kThreads * kEpt];
  int16_t act[kThreads * kEpt];
};
constexpr int kFastBlock = kStepBlock;          // 8 consumer warps + TMA producer warp
template <int kEpt>
constexpr size_t fast_smem_bytes() {   // ring slots + full/empty mbarriers + tile ids + deferred lists
  return sizeof(FastStage<kEpt>) * kStages + 2 * 8 * kStages + 4 * kStages + size_t(kDeferCap) * kEpt * 4 * (kThreads / 32);
}
#ifndef SHIPENV_FAST_EPT
#define SHIPENV_FAST_EPT 2
#endif

#ifndef SHIPENV_FAST_MIN_BLOCKS
#define SHIPENV_FAST_MIN_BLOCKS 3
#endif
// kPolicy: every action of the launch is SHIPENV_ACT_POLICY (SHIPENV_F_ALL_POLICY); the straight-line case is then
// `sample_action` choosing a MOVE (word 0) followed by that move (words 1, 2); `actions` is never read.
// kObs: fused observation rows (utils/preprocessing.py:25-62): lanes that need the general path are resolved on the
// spot (their row must be final before the tile's TMA bulk store), heads go to a shared [256, 6+4P] tile whose
// constant port tail is written once, one bulk store per tile, two tiles in flight.
// kLean: only reward_f32 and the masks exist; otherwise status / taken / reward_f64 / episode statistics are
// written too (each still nullable), for the straight-line lanes here and by the general path for the others.
template <bool kPolicy, bool kObs, int kEpt, bool kLean>
__global__ void __launch_bounds__(kFastBlock, SHIPENV_FAST_MIN_BLOCKS) step_fast_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  __shared__ uint32_t chain_seq_s;                   // chained: the count c this launch found on its chain array
#ifdef SHIPENV_TRACE
  __shared__ unsigned long long trace_s[8];
  if (threadIdx.x == 0) { trace_s[0] = trace_now(); trace_s[1] = trace_s[2] = trace_s[3] = trace_s[4] = trace_s[5] = trace_s[6] = 0; }
#endif
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
  const bool chain_on = a.chain != nullptr;          // launch-uniform
  if (!chain_on) {
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
  }
  if (tid >= kThreads) {             // ---- producer warp: lane 0 streams tiles in with TMA ----
    if (tid == kThreads) {
      // Unchained: tiles are handed out by a global ticket counter, so SMs that run ahead simply take more tiles.
      // Tickets 0..grid-1 are the block indices; later ones come from the counter, offset by the grid size.  Every
      // tile load is followed by exactly one fetch, so a launch makes exactly `full_tiles` fetches: the thread that
      // draws the last one puts the counter back to zero for the next launch that uses this counter.  No host-side
      // state: the launch is CUDA-graph replayable.
      // Chained: CTA b takes tiles b, b + G, b + 2G, ... in every launch (see above).
      uint32_t stride = 0;                           // chained: G; unchained: tickets
      if (chain_on) {
        const bool run_ahead = a.flags & SHIPENV_F_CHAINED;
        const uint32_t G = gridDim.x;
        const uint32_t* const flag = a.chain + kChainDone + blockIdx.x;
        // a first look at CTA b's flag travels together with the arrival (two L2 round trips side by side)
        uint32_t seen = run_ahead ? ld_relaxed_u32(flag) : 0u;
        unsigned long long* const sa = reinterpret_cast<unsigned long long*>(a.chain + kChainSeq);
        const unsigned long long arrival = atomicAdd(sa, 1ull);
        const uint32_t cum = uint32_t(arrival >> 16);
        if (uint32_t(arrival & 0xFFFFull) == G - 1u) {   // the last CTA to arrive moves the word on to (c + G) << 16 | 0
          atomicAdd(sa, ((unsigned long long)G << 16) - (unsigned long long)G);
          __threadfence();                             // ... and has it performed before it triggers its dependents
        }
        // the arrival has been performed (its result is here): the next launch's CTAs may start, they all read c + G
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        chain_seq_s = cum;                           // the consumers see it behind the first `full` barrier
        if (run_ahead) {                             // CTA b of the previous launch has finished with these tiles
          while (seen != cum) { __nanosleep(64); seen = ld_relaxed_u32(flag); }
          asm volatile("fence.acq_rel.gpu;" ::: "memory");          // relaxed load + fence = acquire
          asm volatile("fence.proxy.async.global;" ::: "memory");   // its generic-proxy stores -> our TMA reads
        } else {
          asm volatile("griddepcontrol.wait;" ::: "memory");
        }
        stride = G;
      }
      SHIPENV_TRACE_SET(1, trace_now());
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
        if (chain_on) {
          ticket += stride;
        } else {
          const uint32_t drawn = atomicAdd(a.tile_counter, 1u);      // in flight while the slot is awaited
          if (drawn == uint32_t(full_tiles) - 1u) atomicExch(a.tile_counter, 0u);
          ticket = drawn + gridDim.x;
        }
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
  const uint32_t W = uint32_t(w.h->W), H = uint32_t(w.h->H);
  const uint32_t cap = uint32_t(w.h->capacity);
  const uint32_t wpr = uint32_t(w.h->wpr);
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;
  const uint32_t trunc_at = a.max_steps > 0 ? uint32_t(a.max_steps) : 0x10000u;   // never reached when disabled
  const uint32_t id_lo = uint32_t(a.env_id_base), id_hi = uint32_t(a.env_id_base >> 32);
  // lanes 0..2 each own one mask array: one predicated store writes the warp's three ballot words
  uint32_t* const mask_ptr = lane == 0 ? a.io.done_bits : lane == 1 ? a.io.reset_bits : a.io.error_bits;
  const bool mask_lane = lane < 3;
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
    while (more && (kDrainEarly ? n_defer == 0 : n_defer <= kCap - 32 * kEpt)) {
      mbar_wait_s(full_b + st * 8u, parity);
      const int tile = int(lds_u32(tile_b + st * 4u));
#ifdef SHIPENV_TRACE
      if (tid == 0 && g_trace_buf && trace_s[2] == 0) trace_s[2] = trace_now();
#endif
      if (tile < 0) { more = false; SHIPENV_TRACE_SET(3, trace_now()); break; }
      const uint32_t so = st * uint32_t(sizeof(Stage));
      uint4 rec_e[kEpt];
      uint32_t c_e[kEpt];
      int act_e[kEpt];
      bool reset_e[kEpt];
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
      const uint32_t wbit = (lds_u32(bits_b + cell * 4u) >> (ny & 31u)) & 1u;             // :293, 1 = not ground
      const bool water = wbit != 0u;
      const uint32_t pd = lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + (dest & 31u) * 4u);
      const int tx = int(pd & 255u), ty = int((pd >> 8) & 255u);
      // closer <=> d_prev - d_new > 0 <=> (x - tx) * dx + (y - ty) * dy < 0 for a unit move   (:309-312)
      const uint32_t closer_bit = uint32_t((int(x) - tx) * dx + (int(y) - ty) * dy) >> 31;
      // positions stay packed as x | y << 8 (the low half of the record's first word)
      const uint32_t pos1 = water ? (nx | (ny << 8)) : (rec.x & 0xFFFFu);
      // loss_thr[cap] = 0xFFFFFFFF: cargo >= capacity always draws; negative cargo goes the general way
      const bool trigger = w_loss <= lds_u32(wb + uint32_t(offsetof(WorldHdr, loss_thr)) + min(uint32_t(rec.y), cap) * 4u);  // :318-320
      const bool arrive = pos1 == (pd & 0xFFFFu);                                         // :325
      const uint32_t c1 = c + 1u;
      // An episode end with auto-reset comes once per few hundred steps of an env.  The step itself is the plain move
      // (the terminal record is stored like any other); the new episode is drawn behind a warp-uniform branch after
      // the stores (taken by one warp-step in ten), by the same code the general path uses, and overwrites record
      // and clock.  Nothing of it rides along on the other steps, and no env waits on the deferred list for it.
      const bool do_reset = auto_reset && (no_fuel || (c1 & 0xFFFFu) >= trunc_at);
      // policy: an off-map choice is re-drawn -- the general way
      const bool general = kPolicy ? (!is_move || !in_map || trigger || arrive)
                                   : (!is_move || (in_map && (trigger || arrive)));
      const bool ok = is_move && in_map && !general;
      const bool reset_here = ok && do_reset;
      const bool off_map = !kPolicy && is_move && !in_map;
      const uint32_t ridx = (no_fuel ? 4u : 0u) | ((wbit ^ 1u) << 1) | closer_bit;        // index of move_reward[]
      {
        const double fuel1 = water ? __dsub_rn(fuel, cost) : fuel;                        // :298
        uint4 out;
        out.x = (rec.x & 0xFFFF0000u) | pos1;
        out.y = rec.y;
        out.z = uint32_t(__double2loint(fuel1)); out.w = uint32_t(__double2hiint(fuel1));
        const float r = lds_f32(wb + uint32_t(offsetof(WorldHdr, move_reward_f32)) + ridx * 4u);
        if (ok) {
          st_stream_u4(a.ships + i, out);
          a.clock[i] = c1;
        }
        if (!general && (kLean || a.io.reward_f32)) a.io.reward_f32[i] = ok ? r : 0.f;
        if (!kLean) {
          if (!general) {              // accepted straight-line move, or an explicit off-map move
            const double r64 = ok ? lds_f64(wb + uint32_t(offsetof(WorldHdr, move_reward)) + ridx * 8u) : 0.0;
            if (a.io.reward_f64) a.io.reward_f64[i] = r64;
            if (a.io.status) a.io.status[i] = uint8_t(ok ? SHIPENV_ST_OK : SHIPENV_ST_OFF_MAP);
            if (a.io.taken) a.io.taken[i] = int16_t(kPolicy ? m : act);
            if (ok) track_episode(a.io, i, r64, reset_here, c);
          }
        }
        if (kObs) fin = ok ? out : rec;
      }
      if (kObs) {                    // the observation row needs the new episode now
        if (__any_sync(0xFFFFFFFFu, reset_here)) {
          if (reset_here) fin = reset_in_place<kPolicy>(a, w, i, c);
        }
      } else {
        reset_e[e] = reset_here;     // behind both envs' straight-line code, which stays interleaved
      }
      const uint32_t gen_w = __ballot_sync(0xFFFFFFFFu, general);
      uint32_t gflags = 0;
      if (kObs) {
        if (gen_w && general) {        // resolve now: the observation row of this tile must be final
          gflags = step_general_inline(a, smem, i);
          fin = __ldcg(a.ships + i);
        }
      } else if (gen_w) {
        if (general) defer[n_defer + __popc(gen_w & ((1u << lane) - 1u))] = i;
        n_defer += __popc(gen_w);
      }
      const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, (ok && no_fuel) || (gflags & kFlagDone));
      const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, reset_here || (gflags & kFlagReset));
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
      if (!kObs) {
        bool any_reset = false;
#pragma unroll
        for (int e = 0; e < kEpt; ++e) any_reset |= reset_e[e];
        if (__any_sync(0xFFFFFFFFu, any_reset)) {
#pragma unroll
          for (int e = 0; e < kEpt; ++e)
            if (reset_e[e]) reset_in_place<kPolicy>(a, w, tile * kTile + e * kThreads + tid, c_e[e]);
        }
      }
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
      obs_head(unpack(__ldcg(a.ships + i)), row);
      for (int c = 6; c < width; ++c) row[c] = w.h->obs_tail[c - 6];
    }
    const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, f & kFlagDone);
    const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, f & kFlagReset);
    const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, f & kFlagError);
    if (lane < 3) mask_ptr[i >> 5] = lane == 0 ? done_w : lane == 1 ? reset_w : err_w;
    }
  }
  if (chain_on) {
    // Every consumer warp has written its share of this CTA's envs (bar.sync orders those writes before thread 0's
    // release, which is cumulative at gpu scope): CTA b of the next launch may start.
    consumer_sync();
    if (tid == 0) {
      SHIPENV_TRACE_SET(5, trace_now());
      st_release_u32(a.chain + kChainDone + blockIdx.x, chain_seq_s + gridDim.x);
      SHIPENV_TRACE_SET(6, trace_now());
    }
  }
#ifdef SHIPENV_TRACE
  if (g_trace_buf) {
    consumer_sync();
    if (tid == 0) {
      const unsigned int slot = atomicAdd(&g_trace_count, 1u);
      unsigned long long* r = g_trace_buf + size_t(slot) * 8;
      r[0] = trace_s[0]; r[1] = trace_s[1]; r[2] = trace_s[2]; r[3] = trace_s[3]; r[4] = trace_now();
      r[5] = trace_s[5]; r[6] = trace_s[6];
      r[7] = (unsigned long long)blockIdx.x | ((unsigned long long)trace_smid() << 16) |
             ((unsigned long long)(chain_on ? chain_seq_s : 0u) << 32);
    }
  }
#endif
}

}  // namespace shipenv
