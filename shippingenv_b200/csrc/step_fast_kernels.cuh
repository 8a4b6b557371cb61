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
constexpr int kDeferCap = 64;   // per-warp deferred-env list per env-per-thread; a tile adds at most 32 of each
constexpr int kPendCap = 64;    // per-warp list of tiles that wait for their deferred envs (chained launches)

// ------------------------------------------------------------------------------------------
// Chained launches (shipenv_step_io.chain, SHIPENV_F_CHAINED).
//
// Consecutive step launches on one stream normally serialise at the launch boundary: grid t+1 may start its
// prologue early (programmatic dependent launch) but `griddepcontrol.wait` holds it until every CTA of grid t has
// exited and flushed -- the ramp of one grid and the tail of the other are exposed on every launch (5 of 23 us at
// 2 Mi envs).  The only true dependency is per tile: tile i of step t+1 needs tile i of step t.  With a chain array
// the kernels track exactly that:
//   chain[0..1]           one 64-bit word: s << 32 | a, s = number of chain-protocol launches on this env group so
//                         far, a = CTAs of the current launch that have arrived (the last one turns it into s+1 | 0)
//   chain[kChainTicket..] a ring of self-resetting tile-ticket counters, slot s % kChainTicketSlots
//   chain[kChainEpoch+u]  completions of 256-env unit u: +1 by each of the 8 consumer warps per step, so the
//                         unit is ready for launch s when the count has reached 8 * s
// A launch with SHIPENV_F_CHAINED skips griddepcontrol.wait: its producer thread polls the units of a tile
// (ld.acquire.gpu) before it issues the tile's TMA loads; the units of a tile are published (gpu-scope release) once
// all stores to it -- and its deferred general-path envs -- are done.  CTAs of launch t+1 become resident as CTAs of
// launch t exit, find most tiles ready and keep the SMs full across the boundary.
// s is race-free without host state (CUDA-graph replayable): every CTA learns s from its arrival (one atomic) and
// triggers its dependents only after that; grid t+1 cannot start before all CTAs of grid t have triggered, so all
// CTAs of grid t arrive on s | 0..G-1, the last of them moves the word on, and all CTAs of grid t+1 find s + 1.
// ------------------------------------------------------------------------------------------
constexpr int kChainSeq = 0, kChainTicket = 64, kChainTicketSlots = 8, kChainEpoch = 64 + 32 * 8;
constexpr int kChainUnit = 256;   // envs per epoch word, whatever the tile size of the kernel variant

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// consumer warp: every lane's stores of unit(s) [u0, u0 + units) are done (call converged, after __syncwarp).
// Direct publication from a consumer warp costs a gpu-scope fence in that warp (it stalls until all of its stores
// have been acknowledged); it is used for the ragged last tile only.
__device__ __forceinline__ void chain_publish(uint32_t* epoch, int u0, int units, int lane) {
  if (lane < units) red_release_add(epoch + u0 + lane, 1u);
}

// In the streaming loop the consumer warps hand their completed tiles to a publisher thread (warp 9) through a queue
// in shared memory (one cta-scope fence and two shared-memory operations per tile and warp); the publisher pops the
// entries, issues ONE gpu-scope fence per batch -- it orders the consumers' stores it has observed through the
// queue (fence cumulativity, the `__syncthreads(); if (t == 0) { __threadfence(); flag }` idiom) -- and adds to the
// epoch words.  A gpu-scope fence takes ~1.5 us under this store load (measured): issued by the consumer warps once
// per tile it made the kernel 30 % slower, issued by the TMA producer thread it starved the input ring (+80 %); the
// publisher thread has nothing else to do.
constexpr int kChainQ = 512;
struct alignas(16) ChainQueue {
  uint32_t tail;                 // entries reserved so far (consumers, atomicAdd)
  uint32_t head_pub;             // entries consumed so far (producer), for the overflow check
  uint32_t pad[2];
  uint32_t entry[kChainQ];       // 0 = empty, tile + 1, or kChainQDone = this consumer warp has finished
};
constexpr uint32_t kChainQDone = 0xFFFFFFFFu;
__device__ __forceinline__ void chainq_push(ChainQueue* q, uint32_t value) {   // one thread; its warp has converged
  const uint32_t pos = atomicAdd(&q->tail, 1u);
  while (pos - *reinterpret_cast<volatile uint32_t*>(&q->head_pub) >= uint32_t(kChainQ)) __nanosleep(100);
  __threadfence_block();
  *reinterpret_cast<volatile uint32_t*>(&q->entry[pos % uint32_t(kChainQ)]) = value;
}

// kEpt envs per thread: a tile is 256 * kEpt envs; thread t owns envs t, t + 256, ... of the tile.  Two
// independent env chains per thread interleave (ILP) and share the per-tile bookkeeping.
template <int kEpt>
struct alignas(16) FastStage {
  uint4 ships[kThreads * kEpt];
  uint32_t clock[kThreads * kEpt];
  int16_t act[kThreads * kEpt];
};
constexpr int kFastBlock = kStepBlock + 32;     // 8 consumer warps + TMA producer warp + publisher warp
template <int kEpt>
constexpr size_t fast_smem_bytes() {   // ring slots + full/empty mbarriers + tile ids + deferred lists + pending-tile lists
  return sizeof(FastStage<kEpt>) * kStages + 2 * 8 * kStages + 4 * kStages +
         (size_t(kDeferCap) * kEpt + kPendCap) * 4 * (kThreads / 32) + sizeof(ChainQueue);
}
#ifndef SHIPENV_FAST_EPT
#define SHIPENV_FAST_EPT 2
#endif

#ifndef SHIPENV_FAST_MIN_BLOCKS
#define SHIPENV_FAST_MIN_BLOCKS 3
#endif
// Chained launches: a consumer warp that holds deferred envs drains them (and publishes their tiles) once it has
// streamed this many further tiles, so that the next launch does not wait for the end of this one; 0 = only when
// the list fills up or the input ends.
#ifndef SHIPENV_CHAIN_DRAIN_TILES
#define SHIPENV_CHAIN_DRAIN_TILES 4
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
  __shared__ uint32_t chain_seq_s;                   // this launch's sequence number on its chain array
  const WorldView w(smem);
  const int tid = threadIdx.x;
  const int n = int(a.n);
  static_assert(!kObs || kEpt == 1, "the observation variant keeps one env per thread");
  constexpr int kTile = kThreads * kEpt;
  constexpr int kCap = kDeferCap * kEpt;
  constexpr int kUnits = kTile / kChainUnit;         // epoch words per tile
  using Stage = FastStage<kEpt>;
  const int full_tiles = n / kTile;                  // the streaming loop takes complete tiles only
  Stage* stage = reinterpret_cast<Stage*>(smem + ((a.world_bytes + 127u) & ~127u));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(stage) + sizeof(Stage) * kStages);
  uint64_t* empty_bar = full_bar + kStages;
  volatile int* stage_tile = reinterpret_cast<volatile int*>(empty_bar + kStages);   // tile index held by each slot
  int* defer_all = const_cast<int*>(stage_tile) + kStages;
  int* pend_all = defer_all + kCap * (kThreads / 32);
  ChainQueue* const chainq = reinterpret_cast<ChainQueue*>(pend_all + kPendCap * (kThreads / 32));
  const bool chain_on = a.chain != nullptr;          // warp-uniform, launch-uniform
  uint32_t* const epoch = chain_on ? a.chain + kChainEpoch : nullptr;
  if (tid == 0) {
    mbar_init(&world_bar, 1);
#pragma unroll
    for (int st = 0; st < kStages; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], kThreads / 32); }
    fence_proxy_async();
  }
  if (chain_on)
    for (int k = tid; k < int(sizeof(ChainQueue) / 4); k += kFastBlock) reinterpret_cast<uint32_t*>(chainq)[k] = 0u;
  __syncthreads();
  if (tid == 0) {                    // the world blob and the first tiles are fetched concurrently
    mbar_expect_tx(&world_bar, a.world_bytes);
    tma_load_bulk(smem, a.world, a.world_bytes, &world_bar);
  }
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
  if (tid >= kThreads + 32) {        // ---- publisher warp: lane 0 publishes completed tiles (chained launches) ----
    if (tid == kThreads + 32 && chain_on) {
      uint32_t q_head = 0, q_done = 0;
      while (q_done < uint32_t(kThreads / 32)) {     // until every consumer warp has handed in its last tile
        uint32_t v[8];
        int nv = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = 0u;
          if (nv == k) {
            volatile uint32_t* e = &chainq->entry[q_head % uint32_t(kChainQ)];
            const uint32_t x = *e;
            if (x != 0u) { *e = 0u; ++q_head; v[k] = x; ++nv; }
          }
        }
        if (nv == 0) { __nanosleep(200); continue; }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");     // the consumers' stores behind these entries, then ...
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (v[k] == kChainQDone) ++q_done;
          else if (v[k] != 0u) {
            uint32_t* e = epoch + (v[k] - 1u) * uint32_t(kUnits);
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(e) : "memory");   // ... the publication
            if (kUnits == 2) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(e + 1) : "memory");
          }
        }
        *reinterpret_cast<volatile uint32_t*>(&chainq->head_pub) = q_head;
      }
    }
    return;
  }
  if (tid >= kThreads) {             // ---- producer warp: lane 0 streams tiles in with TMA ----
    if (tid == kThreads) {
      uint32_t* tickets = a.tile_counter;
      uint32_t ready = 0;
      if (chain_on) {
        // Only a launch the caller marked as chained may run ahead of its predecessor in the stream; any other one
        // first waits for everything before it (then the epochs below are all ready at the first look).
        if (!(a.flags & SHIPENV_F_CHAINED)) asm volatile("griddepcontrol.wait;" ::: "memory");
        unsigned long long* const sa = reinterpret_cast<unsigned long long*>(a.chain + kChainSeq);
        const unsigned long long arrival = atomicAdd(sa, 1ull);
        const uint32_t seq = uint32_t(arrival >> 32);
        if (uint32_t(arrival) == gridDim.x - 1u)     // the last CTA to arrive moves the word on to (seq + 1) << 32 | 0
          atomicAdd(sa, (1ull << 32) - (unsigned long long)gridDim.x);
        __threadfence();
        // only now may the next launch's CTAs start: they all read seq + 1
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        chain_seq_s = seq;                           // the consumers see it behind the first `full` barrier
        ready = seq * 8u;
        tickets = a.chain + kChainTicket + (seq % uint32_t(kChainTicketSlots)) * 32u;
      }
      // Tiles are handed out by a global ticket counter, so SMs that run ahead simply take more tiles (no tail
      // imbalance).  Tickets 0..grid-1 are the block indices; later ones come from the counter, offset by the
      // grid size.  Every tile load is followed by exactly one fetch, so a launch makes exactly `full_tiles`
      // fetches: the thread that draws the last one puts the counter back to zero for the next launch that uses
      // this counter.  No host-side state: the launch is CUDA-graph replayable.  A grid that covers all tiles
      // statically draws nothing (deep chains of small launches never share a counter).
      const bool dynamic = full_tiles > int(gridDim.x);
      uint32_t ticket = blockIdx.x;
      for (int it = 0;; ++it) {
        const int st = it % kStages;
        const bool valid = ticket < uint32_t(full_tiles);
        uint32_t e0 = ready, e1 = ready, drawn = 0xFFFFFFFFu - gridDim.x;
        if (valid) {                                 // both in flight while the slot is awaited
          if (chain_on) {
            e0 = ld_acquire_u32(epoch + ticket * kUnits);
            if (kUnits == 2) e1 = ld_acquire_u32(epoch + ticket * kUnits + 1);
          }
          if (dynamic) drawn = atomicAdd(tickets, 1u);
        }
        if (it >= kStages) mbar_wait_backoff(&empty_bar[st], uint32_t(it / kStages - 1) & 1u);
        if (!valid) {                                // out of tiles: publish the end marker
          stage_tile[st] = -1;
          mbar_arrive(&full_bar[st]);
          break;
        }
        if (chain_on) {                              // every earlier launch has finished with this tile
          while (int(e0 - ready) < 0) { __nanosleep(64); e0 = ld_acquire_u32(epoch + ticket * kUnits); }
          if (kUnits == 2)
            while (int(e1 - ready) < 0) { __nanosleep(64); e1 = ld_acquire_u32(epoch + ticket * kUnits + 1); }
          asm volatile("fence.proxy.async.global;" ::: "memory");   // their generic-proxy stores -> our TMA reads
        }
        const int base = int(ticket) * kTile;
        stage_tile[st] = int(ticket);
        mbar_expect_tx(&full_bar[st], kTile * (kPolicy ? 20u : 22u));
        tma_load_bulk(stage[st].ships, a.ships + base, kTile * 16u, &full_bar[st]);
        tma_load_bulk(stage[st].clock, a.clock + base, kTile * 4u, &full_bar[st]);
        if (!kPolicy) tma_load_bulk(stage[st].act, a.actions + base, kTile * 2u, &full_bar[st]);
        if (dynamic && drawn == uint32_t(full_tiles) - 1u) atomicExch(tickets, 0u);
        ticket = drawn + gridDim.x;                  // static grids: wraps to 0xFFFFFFFF, never valid
      }
    }
    return;
  }
  mbar_wait(&world_bar, 0);
  float* obs_tile = nullptr;
  const int width = a.obs_width;
  if (kObs) {
    obs_tile = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(chainq) + sizeof(ChainQueue));
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
  int* pend = pend_all + (tid >> 5) * kPendCap;      // chained launches: tiles whose deferred envs are not done yet
  int n_defer = 0, n_pend = 0;                       // warp-uniform
  // streamed-tile counter and the count at which the warp leaves the streaming loop to drain its deferred list:
  // moved up when the list is about to overflow and, on chained launches, to a few tiles after the first deferral
  int tiles_done = 0, drain_at = 0x7FFFFFFF;         // warp-uniform
  int obs_prev_tile = -1;                            // kObs, chained: the tile whose bulk store is still in flight
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
    while (more && tiles_done != drain_at) {
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
      uint32_t tile_gen = 0;                         // lanes of this warp that deferred an env of this tile
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
      // An episode end with auto-reset (once per few hundred steps of an env) goes the general way too: the reset
      // draws, the record and clock rewrite and their predicated selects would otherwise ride along on every step.
      const bool do_reset = auto_reset && (no_fuel || (c1 & 0xFFFFu) >= trunc_at);
      // (the fused-observation variant resolves general lanes on the spot, a whole warp waiting for them, so it keeps
      // the reset in line: words 2 and 3 are the two port draws, :232-233; equal draws need more words)
      const uint32_t ro = kObs ? __umulhi(wd[2], uint32_t(w.h->P)) : 0u, rd = kObs ? __umulhi(wd[3], uint32_t(w.h->P)) : 0u;
      const bool reset_general = (kObs && !kPolicy) ? (do_reset && ro == rd) : do_reset;
      // Lean explicit-move variant: a step whose only complication is the reset takes the straight-line path like any
      // other (moved ship, reward, done bit) and leaves a marked entry on the deferred list; the drain then only
      // has to draw the new episode (words 2.. of the cell) instead of redoing the whole step in the general code.
      const bool reset_only = kLean && !kObs && !kPolicy && do_reset && is_move && in_map && !trigger && !arrive;
      // policy: an off-map choice is re-drawn -- the general way as well
      const bool general = kPolicy ? (!is_move || !in_map || trigger || arrive || do_reset)
                                   : (!is_move || (in_map && (trigger || arrive || (reset_general && !reset_only))));
      const bool ok = is_move && in_map && !general;
      const bool off_map = !kPolicy && is_move && !in_map;
      const uint32_t ridx = (no_fuel ? 4u : 0u) | ((wbit ^ 1u) << 1) | closer_bit;        // index of move_reward[]
      {
        const double fuel1 = water ? __dsub_rn(fuel, cost) : fuel;                        // :298
        uint4 out;
        out.x = (rec.x & 0xFFFF0000u) | pos1;
        out.y = rec.y;
        out.z = uint32_t(__double2loint(fuel1)); out.w = uint32_t(__double2hiint(fuel1));
        uint32_t c_out = c1;
        if (kObs && !kPolicy && do_reset) {                                                // :227-243
          out.x = (lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + ro * 4u) & 0xFFFFu) | (ro << 16) | (rd << 24);
          out.y = uint32_t(w.h->cargo0);
          out.z = uint32_t(__double2loint(w.h->fuel0)); out.w = uint32_t(__double2hiint(w.h->fuel0));
          c_out = ((c >> 16) + 1u) << 16;
        }
        const float r = lds_f32(wb + uint32_t(offsetof(WorldHdr, move_reward_f32)) + ridx * 4u);
        if (ok) {
          st_stream_u4(a.ships + i, out);
          a.clock[i] = c_out;
        }
        if (!general && (kLean || a.io.reward_f32)) a.io.reward_f32[i] = ok ? r : 0.f;
        if (!kLean) {
          if (!general) {              // accepted straight-line move, or an explicit off-map move
            const double r64 = ok ? lds_f64(wb + uint32_t(offsetof(WorldHdr, move_reward)) + ridx * 8u) : 0.0;
            if (a.io.reward_f64) a.io.reward_f64[i] = r64;
            if (a.io.status) a.io.status[i] = uint8_t(ok ? SHIPENV_ST_OK : SHIPENV_ST_OFF_MAP);
            if (a.io.taken) a.io.taken[i] = int16_t(kPolicy ? m : act);
            if (ok) track_episode(a.io, i, r64, kObs && !kPolicy && do_reset, c);
          }
        }
        if (kObs) fin = ok ? out : rec;
      }
      const uint32_t gen_w = __ballot_sync(0xFFFFFFFFu, general || reset_only);
      uint32_t gflags = 0;
      if (kObs) {
        if (gen_w && general) {        // resolve now: the observation row of this tile must be final
          gflags = step_general_inline(a, smem, i);
          fin = __ldcg(a.ships + i);
        }
      } else if (gen_w) {
        if (general || reset_only) defer[n_defer + __popc(gen_w & ((1u << lane) - 1u))] = reset_only ? (i | int(0x80000000u)) : i;
        n_defer += __popc(gen_w);
        tile_gen |= gen_w;
        if (n_defer > kCap - 32 * kEpt || n_pend >= kPendCap - 1) drain_at = tiles_done + 1;   // no room for another tile's worth
        else if (SHIPENV_CHAIN_DRAIN_TILES > 0 && chain_on) drain_at = min(drain_at, tiles_done + 1 + SHIPENV_CHAIN_DRAIN_TILES);
      }
      const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, (ok && no_fuel) || (gflags & kFlagDone));
      const uint32_t reset_w = kObs ? __ballot_sync(0xFFFFFFFFu, (!kPolicy && ok && do_reset) || (gflags & kFlagReset))
                                    : 0u;                          // without observations resets are general-path work
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
          // the other buffer is free again; a chained successor may overwrite the previous tile's rows once they
          // have been written out completely, so that tile is published one tile late
          if (chain_on) tma_store_wait<1>(); else tma_store_wait_read<1>();
        }
        obs_buf ^= 1;
        consumer_sync();
        if (chain_on) {
          if (obs_prev_tile >= 0 && lane == 0) chainq_push(chainq, uint32_t(obs_prev_tile) + 1u);
          obs_prev_tile = tile;
        }
      }
      }  // envs of this thread
      if (chain_on && !kObs) {
        // This warp's share of the tile is final (publish it) unless some of its envs wait on the deferred list.
        __syncwarp();
        if (tile_gen) {
          if (lane == 0) pend[n_pend] = tile;
          ++n_pend;
        } else if (lane == 0) {
          chainq_push(chainq, uint32_t(tile) + 1u);
        }
      }
      ++tiles_done;
    }
    // ------------------------------ drain the deferred envs, a warp-full at a time ------------------------------
    if (n_defer == 0 && !more) break;
    __syncwarp();
    while (n_defer > 0) {
      const int k = min(n_defer, 32);
      if (lane < k) {
        const int ent = defer[n_defer - k + lane];
        const int i = ent & 0x7FFFFFFF;
        const uint32_t bit = 1u << (i & 31);
        if (ent < 0) {               // reset only (:227-243): the step itself has been taken and stored in line
          const uint32_t c = __ldcg(a.clock + i) - 1u;        // the cell of the terminal step; its words 0, 1 are used
          Rng g;
          g.seek(a.round_key, c, a.env_id_base + uint64_t(uint32_t(i)), 2u);
          ShipR s;
          env_reset(w, s, g);
          a.ships[i] = pack(s);
          a.clock[i] = ((c >> 16) + 1u) << 16;
          atomicOr(a.io.reset_bits + (i >> 5), bit);
        } else {
          const uint32_t f = step_general_inline(a, smem, i);   // inlined: an ABI call would pin loop state to the stack
          if (f & kFlagDone) atomicOr(a.io.done_bits + (i >> 5), bit);
          if (f & kFlagReset) atomicOr(a.io.reset_bits + (i >> 5), bit);
          if (f & kFlagError) atomicOr(a.io.error_bits + (i >> 5), bit);
        }
      }
      n_defer -= k;
      __syncwarp();
    }
    if (chain_on) {                  // the tiles those envs belonged to are complete for this warp now
      for (int k = lane; k < n_pend; k += 32) chainq_push(chainq, uint32_t(pend[k]) + 1u);
      n_pend = 0;
      __syncwarp();
    }
    drain_at = 0x7FFFFFFF;
  }
  if (kObs) {
    if (tid == 0) tma_store_wait_all();
    if (chain_on) {
      consumer_sync();
      if (obs_prev_tile >= 0 && lane == 0) chainq_push(chainq, uint32_t(obs_prev_tile) + 1u);
    }
  }
  // The ragged last tile (n % tile envs) never enters the pipeline: its owner CTA takes the general path for it.
  // Its epoch words follow those of the full tiles; every consumer warp of the owner publishes them.
  if (full_tiles * kTile < n && int(blockIdx.x) == full_tiles % int(gridDim.x)) {
    const int rag_units = (n - full_tiles * kTile + kChainUnit - 1) / kChainUnit;   // units that hold envs
    if (chain_on) {
      const uint32_t ready = chain_seq_s * 8u;
      if (lane < rag_units) {
        const uint32_t* e = epoch + full_tiles * kUnits + lane;
        while (int(ld_acquire_u32(e) - ready) < 0) __nanosleep(64);
      }
      __syncwarp();
    }
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
    if (chain_on) {
      __syncwarp();
      chain_publish(epoch, full_tiles * kUnits, rag_units, lane);
    }
  }
  if (chain_on) {                    // lets the publisher thread (and with it the CTA) finish
    __syncwarp();
    if (lane == 0) chainq_push(chainq, kChainQDone);
  }
}

}  // namespace shipenv
