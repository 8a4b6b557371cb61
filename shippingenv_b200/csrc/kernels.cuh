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

#include "transition.cuh"

namespace shipenv {

constexpr int kThreads = 256;  // envs per tile

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
};

// Observation row head (utils/preprocessing.py:53-60): x, y, fuel, "cargo" (= fuel, the
// state dict copies it, shipping/environment.py:206), origin | -1, dest | -1.
__device__ __forceinline__ void obs_head(const ShipR& s, float* row) {
  const float f = float(s.fuel);
  row[0] = float(s.x); row[1] = float(s.y); row[2] = f; row[3] = f;
  row[4] = s.origin == kNone ? -1.f : float(s.origin);
  row[5] = s.dest == kNone ? -1.f : float(s.dest);
}

template <bool kObs>
__global__ void __launch_bounds__(kThreads) step_kernel(const StepArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int P = w.h->P;
  const int tid = threadIdx.x;
  const int64_t n_tiles = (a.n + kThreads - 1) / kThreads;
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;

  // observation tiles: [obs_bufs][kThreads][width] floats after the world blob; the constant
  // port tail of every row is written once, only the 6-float head changes per tile.
  float* obs_tile = nullptr;
  const int width = a.obs_width;
  if (kObs) {
    obs_tile = reinterpret_cast<float*>(smem + ((a.world_bytes + 127u) & ~127u));
    for (int b = 0; b < a.obs_bufs; ++b) {
      float* row = obs_tile + (size_t(b) * kThreads + tid) * width;
      for (int c = 6; c < width; ++c) row[c] = w.h->obs_tail[c - 6];
    }
  }
  int obs_buf = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t i = tile * kThreads + tid;
    const bool live = i < a.n;
    int status = SHIPENV_ST_OK;
    bool done = false, was_reset = false;
    double reward = 0.0;
    ShipR s;
    if (live) {
      const uint4 rec = ld_stream_u4(a.ships + i);
      const uint32_t c = a.clock[i];
      const int act = a.actions[i];
      s = unpack(rec);
      Rng g;
      g.seek(a.seed, 0u, c, a.env_id_base + uint64_t(i), a.word0);
      int taken = act;
      if (act == SHIPENV_ACT_POLICY) {
        taken = -1;
        status = policy_step(w, s, g, reward, done, taken);
      } else {
        int cat, val;
        decode_action(act, P, cat, val);
        status = do_action(w, s, g, cat, val, reward, done);
      }
      if (status == SHIPENV_ST_OK) {
        uint32_t c1 = c + 1u;
        const bool ended = done || (a.max_steps > 0 && int(c1 & 0xFFFFu) >= a.max_steps);
        double* er = a.io.ep_return;
        double ep = 0.0;
        if (er) ep = __dadd_rn(er[i], reward);
        if (auto_reset && ended) {
          if (er && a.io.stats) {
            double* st = a.io.stats;
            atomicAdd(st + 0, ep);
            atomicAdd(st + 1, ep * ep);
            atomicAdd(st + 2, double(c1 & 0xFFFFu));
            atomicAdd(st + 3, 1.0);
            atomic_max_f64(st + 4, ep);
            atomic_min_f64(st + 5, ep);
          }
          ep = 0.0;
          env_reset(w, s, g);  // continues the word stream of this (env, clock) cell
          c1 = ((c >> 16) + 1u) << 16;
          was_reset = true;
        }
        if (er) er[i] = ep;
        st_stream_u4(a.ships + i, pack(s));
        a.clock[i] = c1;
      } else {
        reward = 0.0;
        done = false;
      }
      if (a.io.reward_f32) a.io.reward_f32[i] = float(reward);
      if (a.io.reward_f64) a.io.reward_f64[i] = reward;
      if (a.io.status) a.io.status[i] = uint8_t(status);
      if (a.io.taken) a.io.taken[i] = int16_t(taken);
    }
    // done / reset / error masks: one ballot word per warp
    const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, live && done);
    const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, live && was_reset);
    const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, live && status != SHIPENV_ST_OK);
    if ((tid & 31) == 0 && live) {
      const int64_t word = i >> 5;
      if (a.io.done_bits) a.io.done_bits[word] = done_w;
      if (a.io.reset_bits) a.io.reset_bits[word] = reset_w;
      if (a.io.error_bits) a.io.error_bits[word] = err_w;
    }
    if (kObs) {
      // head -> shared tile, then one TMA bulk store of the whole contiguous [rows, width] block
      float* tile_base = obs_tile + size_t(obs_buf) * kThreads * width;
      if (a.obs_bufs == 1) {  // the only buffer may still be read by the previous bulk store
        if (tid == 0) tma_store_wait_read<0>();
        __syncthreads();
      }
      if (live) obs_head(s, tile_base + size_t(tid) * width);
      fence_proxy_async();
      __syncthreads();
      const int64_t row0 = tile * kThreads;
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
      if (a.obs_bufs == 2) { obs_buf ^= 1; __syncthreads(); }
    }
  }
  if (kObs && tid == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------
// reset / observe / sample_action
// ------------------------------------------------------------------------------------------
struct ResetArgs {
  const void* world; uint32_t world_bytes;
  uint4* ships; uint32_t* clock; const uint32_t* mask_bits;
  uint64_t seed, env_id_base; int64_t n;
};

__global__ void __launch_bounds__(kThreads) reset_kernel(const ResetArgs a) {
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
    g.seek(a.seed, 0u, c, a.env_id_base + uint64_t(i), 0u);
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
};

__global__ void __launch_bounds__(kThreads) sample_kernel(const SampleArgs a) {
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
    g.seek(a.seed, 0u, a.clock[i], a.env_id_base + uint64_t(i), a.word0);
    int cat = kCatJunk, val = 0;
    int st = w.h->P == 0 ? SHIPENV_ST_NO_PORTS : sample_action(w, s, g, cat, val);
    a.actions[i] = st == SHIPENV_ST_OK ? int16_t(encode_action(cat, val, w.h->P)) : int16_t(-2);
    if (a.words_used) a.words_used[i] = uint8_t(g.i - a.word0);
    if (a.status) a.status[i] = uint8_t(st);
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
};

constexpr int kRolloutThreads = 128;

__global__ void __launch_bounds__(kRolloutThreads) rollout_kernel(const RolloutArgs a) {
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
    g.seek(a.seed, kRolloutTweak, uint32_t(sim), a.root_id_base + uint64_t(r), 0u);
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

}  // namespace shipenv
