// step_many_kernels.cuh -- T consecutive steps of every env in ONE launch (shipenv_step_many).
//
// The loops the reference's agents run are `for t: env.step(a_t)` (agents/mcts.py:228-236,
// agents/sarsa.py:222-313).  When the T actions of an env are known up front (a pre-generated
// stream, or the random policy) the ship record and its clock can stay in registers for all T
// steps: per env-step the kernel reads 2 B (the action) and writes 4 B + 3 bits (reward, masks)
// instead of 46 B, and a small batch (BASELINE configs[1], 4,096 envs) is no longer bound by one
// launch per step.  Results are bit-identical to T calls of shipenv_step with word0 = 0.
#pragma once
#include "step_kernels.cuh"

namespace shipenv {

struct StepManyArgs {
  const void* world; uint32_t world_bytes;
  uint4* ships; uint32_t* clock;
  const int16_t* actions;                            // [T, n], or null with SHIPENV_F_ALL_POLICY
  float* reward;                                     // [T, n] or null
  uint32_t *done_bits, *reset_bits, *error_bits;     // [T, ceil(n/32)] each, or null
  uint8_t* status; int16_t* taken;                   // [T, n] each, or null
  double* reward_f64;                                // [T, n] or null
  uint64_t seed, env_id_base; int64_t n;
  int T, max_steps; uint32_t flags;
  uint32_t round_key[20];
};

// kEpt envs per thread (thread t of a group owns envs t, t + blockDim, ...), grid-stride over groups of
// blockDim * kEpt envs; blockDim and kEpt are chosen by the host: one warp per CTA and one env per thread for small
// batches (4,096 envs then spread over 128 SMs), two interleaved env chains per thread for large ones.  The
// straight-line MOVE is the throughput kernel's formulation (step_fast_kernels.cuh: packed record, shared-memory
// lookups through 32-bit addresses, integer closer / loss-threshold tests); an episode end with auto-reset and the
// rare general cases (non-move action, cargo-loss draw, arrival, no destination) sit behind branches.
template <bool kPolicy, int kEpt>
__global__ void __launch_bounds__(128) step_many_kernel(const __grid_constant__ StepManyArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int n = int(a.n), T = a.T;
  const int nw = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t wb = smem_u32(smem);
  const uint32_t bits_b = wb + uint32_t(sizeof(WorldHdr));
  const uint32_t W = uint32_t(w.h->W), H = uint32_t(w.h->H), cap = uint32_t(w.h->capacity), wpr = uint32_t(w.h->wpr);
  const uint32_t P = uint32_t(w.h->P);
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;
  const uint32_t trunc_at = a.max_steps > 0 ? uint32_t(a.max_steps) : 0x10000u;   // never reached when disabled
  const bool fast_enabled = P >= 2u;
  const int span = int(blockDim.x) * kEpt;
  const int groups = (n + span - 1) / span;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    int idx[kEpt];
    bool live[kEpt];
    uint32_t live_w[kEpt];
    uint4 rec[kEpt];
    uint32_t clk[kEpt];
    int act_next[kEpt];
#pragma unroll
    for (int e = 0; e < kEpt; ++e) {
      idx[e] = grp * span + e * int(blockDim.x) + int(threadIdx.x);
      live[e] = idx[e] < n;
      live_w[e] = __ballot_sync(0xFFFFFFFFu, live[e]);
      const int i = live[e] ? idx[e] : 0;
      rec[e] = __ldcg(a.ships + i);
      clk[e] = __ldcg(a.clock + i);
      act_next[e] = kPolicy ? SHIPENV_ACT_POLICY : int(__ldg(a.actions + i));
    }
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      uint32_t flags_e[kEpt];
#pragma unroll
      for (int e = 0; e < kEpt; ++e) {
        const int i = live[e] ? idx[e] : 0;
        const int act = act_next[e];
        if (!kPolicy && t + 1 < T) act_next[e] = int(__ldg(a.actions + size_t(t + 1) * n + i));   // one step ahead
        const uint4 r0 = rec[e];
        const uint32_t c = clk[e];
        const uint64_t id = a.env_id_base + uint64_t(uint32_t(i));
        const uint32_t x = r0.x & 255u, y = (r0.x >> 8) & 255u, dest = r0.x >> 24;
        uint32_t wd[4];
        philox_rk(0u, c, uint32_t(id), uint32_t(id >> 32), a.round_key, wd);
        const uint32_t w_cost = kPolicy ? wd[1] : wd[0], w_loss = kPolicy ? wd[2] : wd[1];
        const double fuel = __hiloint2double(int(r0.w), int(r0.z));
        const int m = kPolicy ? int(__umulhi(wd[0], 4u)) : (act & 3);
        const int dx = (m & 1) * (m - 2), dy = ((m & 1) ^ 1) * (m - 1);
        const uint32_t nx = x + uint32_t(dx), ny = y + uint32_t(dy);
        bool is_move;
        if (kPolicy) {   // sample_action (:245-263): a MOVE unless the ship sits on a port without cargo / fuel
          const bool on_map = x < W && y < H;
          const bool on_port = on_map && ((lds_u32(bits_b + (uint32_t(w.h->n_bit_words) + x * wpr + (y >> 5)) * 4u) >> (y & 31u)) & 1u);
          is_move = dest != uint32_t(kNone) && !(on_port && (r0.y == 0u || fuel == 0.0));
        } else {
          is_move = unsigned(act) < 4u && dest != uint32_t(kNone);
        }
        is_move = is_move && fast_enabled && live[e];
        const bool in_map = nx < W && ny < H;                                             // :284
        const double cost = __dadd_rn(1.0, __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), unit_double(w_cost))));   // :103-104
        const bool no_fuel = fuel < cost;                                                 // :288
        const uint32_t cell = in_map ? nx * wpr + (ny >> 5) : 0u;
        const uint32_t wbit = (lds_u32(bits_b + cell * 4u) >> (ny & 31u)) & 1u;           // :293, 1 = not ground
        const bool water = wbit != 0u;
        const uint32_t pd = lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + (dest & 31u) * 4u);
        const int tx = int(pd & 255u), ty = int((pd >> 8) & 255u);
        const uint32_t closer_bit = uint32_t((int(x) - tx) * dx + (int(y) - ty) * dy) >> 31;   // :309-312
        const uint32_t pos1 = water ? (nx | (ny << 8)) : (r0.x & 0xFFFFu);
        const bool trigger = w_loss <= lds_u32(wb + uint32_t(offsetof(WorldHdr, loss_thr)) + min(r0.y, cap) * 4u);  // :318-320
        const bool arrive = pos1 == (pd & 0xFFFFu);                                       // :325
        const uint32_t c1 = c + 1u;
        const bool do_reset = auto_reset && (no_fuel || (c1 & 0xFFFFu) >= trunc_at);
        // everything but the plain move goes the general way (policy: an off-map choice is re-drawn)
        const bool general = live[e] && (kPolicy ? (!is_move || !in_map || trigger || arrive)
                                                 : (!is_move || (in_map && (trigger || arrive))));
        const bool ok = is_move && in_map && !general;
        const bool off_map = !kPolicy && is_move && !in_map;                              // :284, nothing drawn
        const uint32_t ridx = (no_fuel ? 4u : 0u) | ((wbit ^ 1u) << 1) | closer_bit;
        const double fuel1 = water ? __dsub_rn(fuel, cost) : fuel;                        // :298
        double reward = ok ? lds_f64(wb + uint32_t(offsetof(WorldHdr, move_reward)) + ridx * 8u) : 0.0;
        int status = off_map ? SHIPENV_ST_OFF_MAP : SHIPENV_ST_OK, taken = kPolicy ? m : act;
        uint32_t flags = (ok && no_fuel ? kFlagDone : 0u) | (off_map ? kFlagError : 0u);
        if (ok) {
          rec[e].x = (r0.x & 0xFFFF0000u) | pos1;
          rec[e].z = uint32_t(__double2loint(fuel1)); rec[e].w = uint32_t(__double2hiint(fuel1));
          clk[e] = c1;
        }
        if (__any_sync(0xFFFFFFFFu, ok && do_reset)) {
          if (ok && do_reset) {                                                           // :227-243
            // the cell's words behind the move's draws: explicit move 0, 1 -> 2, 3, ...; policy 0, 1, 2 -> 3, 4, ...
            const uint32_t ro = __umulhi(kPolicy ? wd[3] : wd[2], P);
            uint32_t rd = kPolicy ? ro : __umulhi(wd[3], P);
            for (uint32_t blk = 1u; rd == ro; ++blk) {                                    // `while dest == origin`
              uint32_t w2[4];
              philox_rk(blk, c, uint32_t(id), uint32_t(id >> 32), a.round_key, w2);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (rd == ro) rd = __umulhi(w2[k], P);
            }
            rec[e].x = (lds_u32(wb + uint32_t(offsetof(WorldHdr, port)) + ro * 4u) & 0xFFFFu) | (ro << 16) | (rd << 24);
            rec[e].y = uint32_t(w.h->cargo0);
            rec[e].z = uint32_t(__double2loint(w.h->fuel0)); rec[e].w = uint32_t(__double2hiint(w.h->fuel0));
            clk[e] = ((c >> 16) + 1u) << 16;
            flags |= kFlagReset;
          }
        }
        if (__any_sync(0xFFFFFFFFu, general)) {
          if (general) {
            ShipR s = unpack(r0);
            uint32_t cg = c;
            const GeneralOut o = step_general_core(w, a.round_key, id, 0u, a.max_steps, a.flags, s, cg, act);
            rec[e] = pack(s); clk[e] = cg;
            reward = o.reward; status = o.status; taken = o.taken; flags = o.flags;
          }
        }
        if (live[e]) {
          const size_t at = size_t(t) * n + i;
          if (a.reward) a.reward[at] = float(reward);
          if (a.reward_f64) a.reward_f64[at] = reward;
          if (a.status) a.status[at] = uint8_t(status);
          if (a.taken) a.taken[at] = int16_t(taken);
        }
        flags_e[e] = flags;
      }
#pragma unroll
      for (int e = 0; e < kEpt; ++e) {
        if (live_w[e] == 0u) continue;                 // warp-uniform
        const uint32_t done_w = __ballot_sync(0xFFFFFFFFu, live[e] && (flags_e[e] & kFlagDone));
        const uint32_t reset_w = __ballot_sync(0xFFFFFFFFu, live[e] && (flags_e[e] & kFlagReset));
        const uint32_t err_w = __ballot_sync(0xFFFFFFFFu, live[e] && (flags_e[e] & kFlagError));
        if (lane == 0) {
          const size_t word = size_t(t) * nw + (idx[e] >> 5);
          if (a.done_bits) a.done_bits[word] = done_w;
          if (a.reset_bits) a.reset_bits[word] = reset_w;
          if (a.error_bits) a.error_bits[word] = err_w;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < kEpt; ++e)
      if (live[e]) {
        st_stream_u4(a.ships + idx[e], rec[e]);
        a.clock[idx[e]] = clk[e];
      }
  }
}

}  // namespace shipenv
