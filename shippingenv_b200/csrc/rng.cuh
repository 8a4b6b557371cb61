// rng.cuh -- the counter-based word stream every stochastic draw of the environment comes from.
//
// The reference draws from CPython's global `random` (shipping/environment.py:3; call sites
// :63,64,104,157,160,167,177,195,320).  A per-env MT19937 is infeasible on device, so the
// contract (DESIGN.md "RNG contract") replaces it by Philox4x32-10 cells:
//     word(i) = Philox4x32-10(ctr = (i >> 2, c1, c2, c3), key = (k0, k1))[i & 3]
// and defines each `random.*` call the reference makes in terms of words.
#pragma once
#include <stdint.h>

namespace shipenv {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kRolloutTweak = 0x524F4C4Cu;  // "ROLL"
constexpr uint32_t kPortTweak = 0x504F5254u;     // "PORT"

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = uint64_t(kPhiloxM0) * c0;
    const uint64_t p1 = uint64_t(kPhiloxM1) * c2;
    const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
    c1 = uint32_t(p1);
    c3 = uint32_t(p0);
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifdef __CUDACC__
// word * 2^-32 as an exact double without an I2F: 0x41300000:word is 2^20 + word * 2^-32.
__device__ __forceinline__ double unit_double(uint32_t w) {
  return __dadd_rn(__hiloint2double(0x41300000, int(w)), -1048576.0);
}

// Philox4x32-10 with the ten round keys (k0 + r*W0, k1 + r*W1) precomputed on the host and passed in the
// kernel parameters: 20 IMAD.WIDE + 20 LOP3 with the keys as constant-bank operands.
__device__ __forceinline__ void philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          const uint32_t* __restrict__ rk, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(kPhiloxM0, c0), l0 = kPhiloxM0 * c0;
    const uint32_t h1 = __umulhi(kPhiloxM1, c2), l1 = kPhiloxM1 * c2;
    c0 = h1 ^ c1 ^ rk[2 * r];
    c2 = h0 ^ c3 ^ rk[2 * r + 1];
    c1 = l1;
    c3 = l0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
  const uint32_t* rk;  // round keys of the stream's key (kernel parameter space)
  uint32_t c1, c2, c3;
  uint32_t i;        // next word of the cell
  uint32_t blk;      // index of the cached block (i >> 2), ~0u = none
  uint32_t w[4];

  __device__ __forceinline__ void seek(const uint32_t* round_keys, uint32_t cell1, uint64_t id, uint32_t word) {
    rk = round_keys;
    c1 = cell1; c2 = uint32_t(id); c3 = uint32_t(id >> 32);
    i = word; blk = ~0u;
  }
  __device__ __forceinline__ uint32_t next() {
    const uint32_t b = i >> 2;
    // key schedule on the fly here: with many inlined call sites, reloading ten round keys per site costs
    // more than the 18 adds (measured on the rollout kernel); the streaming kernel uses philox_rk directly
    if (b != blk) { philox4x32_10(b, c1, c2, c3, rk[0], rk[1], w); blk = b; }
    const uint32_t s = i & 3u;
    ++i;
    return s == 0 ? w[0] : s == 1 ? w[1] : s == 2 ? w[2] : w[3];
  }
  // random.random()
  __device__ __forceinline__ double random() { return unit_double(next()); }
  // random._randbelow(n): one word, (word * n) >> 32
  __device__ __forceinline__ int below(int n) { return int(__umulhi(next(), uint32_t(n))); }
  // random.randint(a, b)
  __device__ __forceinline__ int randint(int a, int b) { return a + below(b - a + 1); }
  // random.betavariate(2, 2): the median of three uniforms
  __device__ __forceinline__ double beta22() {
    const uint32_t a = next(), b = next(), c = next();
    const uint32_t m = max(min(a, b), min(max(a, b), c));
    return unit_double(m);
  }
};

// The same word stream behind a small FIFO in shared memory (16 words per thread, word-interleaved so that the
// threads of a warp never conflict).  Long sequential consumers (rollouts, MCTS) call top_up() + ensure() once per
// environment step: every lane of the warp then refills at ONE code site, in lock-step, instead of each lane
// running Philox at whichever of the many next() call sites it happens to cross a block boundary.  With
// ensure(k) done, peek(0..k-1) reads ahead without consuming, so a step can be written as straight-line code that
// looks at every word it might need and then advance()s by what it used.
// The ring is mapped twice (every word is stored at p and p + 16 of a 32-word column): the words i .. i + 7 are
// then contiguous for any i, and with the column stride a template constant a peek is one ld.shared with an
// immediate offset from window() instead of an add, a mask and a multiply each.
template <uint32_t kStride>
struct RngFifo {
  static constexpr uint32_t kRing = 16;            // words resident per thread
  static constexpr uint32_t kWords = 2 * kRing;    // shared-memory words per thread (the ring is stored twice)
  static constexpr uint32_t kColumnStride = kStride;
  const uint32_t* rk;
  uint32_t* slot;          // this thread's column of the shared [kWords][kStride] word array
  uint32_t c1, c2, c3;
  uint32_t i;              // next word of the cell to hand out
  uint32_t filled;         // words generated so far (multiple of 4); words [filled-16, filled) are resident

  __device__ __forceinline__ void seek(const uint32_t* round_keys, uint32_t* column, uint32_t cell1, uint64_t id) {
    rk = round_keys; slot = column;
    c1 = cell1; c2 = uint32_t(id); c3 = uint32_t(id >> 32);
    i = 0; filled = 0;
  }
  __device__ __forceinline__ void refill() {
    uint32_t w[4];
    philox_rk(filled >> 2, c1, c2, c3, rk, w);                   // round keys from the kernel parameters
    uint32_t* p = slot + (filled & (kRing - 1u)) * kStride;      // 0, 4, 8 or 12
#pragma unroll
    for (int k = 0; k < 4; ++k) { p[k * kStride] = w[k]; p[(k + kRing) * kStride] = w[k]; }
    filled += 4;
  }
  // One block for every lane that has room for it: the lanes of a warp consume 3 to 7 words per step, so refilling
  // on demand runs the Philox code once or twice per step with part of the warp each time; here it runs once with
  // (nearly) all lanes, and ensure() only picks up stragglers.
  __device__ __forceinline__ void top_up() { if (filled - i <= kRing - 4u) refill(); }
  // at least `ahead` (<= 8) words resident beyond the read position
  __device__ __forceinline__ void ensure(uint32_t ahead = 4u) { while (filled - i < ahead) refill(); }
  // the words i, i + 1, ... as window()[k * kStride], k < 8 (valid after ensure(k + 1))
  __device__ __forceinline__ const uint32_t* window() const { return slot + (i & (kRing - 1u)) * kStride; }
  __device__ __forceinline__ uint32_t peek(uint32_t k) const { return window()[k * kStride]; }
  __device__ __forceinline__ void advance(uint32_t k) { i += k; }
  __device__ __forceinline__ uint32_t next() {
    if (i == filled) refill();                         // rare: a step that needs more than the words ahead
    const uint32_t w = slot[(i & (kRing - 1u)) * kStride];
    ++i;
    return w;
  }
  __device__ __forceinline__ double random() { return unit_double(next()); }
  __device__ __forceinline__ int below(int n) { return int(__umulhi(next(), uint32_t(n))); }
  __device__ __forceinline__ int randint(int a, int b) { return a + below(b - a + 1); }
  __device__ __forceinline__ double beta22() {
    const uint32_t a = next(), b = next(), c = next();
    const uint32_t m = max(min(a, b), min(max(a, b), c));
    return unit_double(m);
  }
};
#endif

}  // namespace shipenv
