// world.cuh -- the read-only world (map bitmap, port table, reward tables) as one 16-byte
// aligned blob.  It lives once in global memory (owned by the context) and is staged into
// shared memory by every CTA with one TMA bulk copy, so terrain and port lookups never go
// to HBM (shipping/environment.py:293 np_game lookup, :145-153 port scan).
#pragma once
#include <stdint.h>

#include "../../include/shipenv.h"

namespace shipenv {

constexpr int kMaxPorts = SHIPENV_MAX_PORTS;
constexpr int kMaxDim = SHIPENV_MAX_DIM;
constexpr int kMaxCap = SHIPENV_MAX_CAPACITY + 1;
constexpr int kNone = SHIPENV_NONE_PORT;
constexpr int kMaxTries = SHIPENV_MAX_TRIES;

// Fixed-size header followed by two W*wpr-word bitmaps (bit y&31 of word x*wpr + (y>>5)): passable cells,
// then cells that hold a port (so that "is the ship at a port?" is one shared-memory load, not a scan).
struct alignas(16) WorldHdr {
  int32_t W, H, P, wpr;          // wpr = ceil(H / 32)
  int32_t cargo0, capacity;      // Initial.CARGO, Initial.MAX_CARGO_CAPACITY
  int32_t n_bit_words, total_bytes;
  double fuel0;                  // Initial.FUEL
  double r_take_fuel, r_take_cargo, r_deliver, r_arrive, p_loss;
  // Reward of a MOVE before loss/arrival terms, accumulated on the host in the reference's
  // order (shipping/environment.py:288-315); index = out_of_fuel<<2 | ground<<1 | closer.
  double move_reward[8];
  float move_reward_f32[8];      // the same, rounded once to f32 (the reward_f32 output)
  uint32_t port[kMaxPorts];      // x | y<<8 | fuel<<16 | cargo<<24
  float obs_tail[4 * kMaxPorts]; // constant tail of a preprocess_state row
  // loss_thr[c] = floor((c / capacity) * 2^32) evaluated in fp64 on the host: the draw
  // `random() <= cargo / 50` (:318-320) is exactly `word <= loss_thr[cargo]` for cargo < capacity;
  // loss_thr[capacity] = 0xFFFFFFFF (cargo >= capacity: the draw always succeeds).
  uint32_t loss_thr[kMaxCap];
};
static_assert(sizeof(WorldHdr) % 16 == 0, "world header must stay 16-byte sized for the bulk copy");

constexpr size_t kMaxWorldBytes = sizeof(WorldHdr) + 2 * size_t(kMaxDim) * (kMaxDim / 32) * 4;

// View over a staged world (shared or global memory).
struct WorldView {
  const WorldHdr* h;
  const uint32_t* bits;
  __device__ __forceinline__ explicit WorldView(const void* base)
      : h(reinterpret_cast<const WorldHdr*>(base)),
        bits(reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(base) + sizeof(WorldHdr))) {}
  __device__ __forceinline__ bool passable(int x, int y) const {
    return (bits[x * h->wpr + (y >> 5)] >> (y & 31)) & 1u;
  }
  __device__ __forceinline__ bool port_cell(int x, int y) const {
    return (bits[h->n_bit_words + x * h->wpr + (y >> 5)] >> (y & 31)) & 1u;
  }
};

}  // namespace shipenv
