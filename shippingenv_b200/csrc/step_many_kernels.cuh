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

// One env per thread, grid-stride over blockDim-sized groups of envs; blockDim is chosen by the host (one warp
// per CTA for small batches so that 4,096 envs spread over 128 SM sub-partitions).
template <bool kPolicy>
__global__ void __launch_bounds__(128) step_many_kernel(const __grid_constant__ StepManyArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  const int P = w.h->P, W = w.h->W, H = w.h->H, cap = w.h->capacity;
  const int n = int(a.n), T = a.T;
  const int nw = (n + 31) >> 5;
  const bool auto_reset = a.flags & SHIPENV_F_AUTO_RESET;
  const int lane = threadIdx.x & 31;
  const int groups = (n + int(blockDim.x) - 1) / int(blockDim.x);
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int i = grp * int(blockDim.x) + int(threadIdx.x);
    const bool live = i < n;
    const uint32_t live_w = __ballot_sync(0xFFFFFFFFu, live);
    if (!live) continue;                             // the ballots below name the live lanes of the warp
    ShipR s = unpack(__ldcg(a.ships + i));
    uint32_t c = __ldcg(a.clock + i);
    const uint64_t id = a.env_id_base + uint64_t(uint32_t(i));
    int act_next = kPolicy ? SHIPENV_ACT_POLICY : int(__ldg(a.actions + i));
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int act = act_next;
      if (!kPolicy && t + 1 < T) act_next = int(__ldg(a.actions + size_t(t + 1) * n + i));   // one step ahead
      double reward = 0.0;
      int status = SHIPENV_ST_OK, taken = act;
      uint32_t flags = 0;
      bool general = true;
      // ---- the common case in line: a MOVE without cargo-loss draw, arrival or colliding reset draws ----
      if (P >= 2 && s.dest != kNone) {
        uint32_t wd[4];
        philox_rk(0u, c, uint32_t(id), uint32_t(id >> 32), a.round_key, wd);
        bool is_move;
        int m;
        if (kPolicy) {        // sample_action (:245-263): a MOVE unless the ship sits on a port without cargo / fuel
          const bool on_port = unsigned(s.x) < unsigned(W) && unsigned(s.y) < unsigned(H) && w.port_cell(s.x, s.y);
          is_move = !(on_port && (s.cargo == 0 || s.fuel == 0.0));
          m = int(__umulhi(wd[0], 4u));              // choice([N, E, S, W]), :165-167
        } else {
          is_move = unsigned(act) < 4u;
          m = act & 3;
        }
        const uint32_t w_cost = kPolicy ? wd[1] : wd[0], w_loss = kPolicy ? wd[2] : wd[1];
        const int nx = s.x + (m == 3) - (m == 1), ny = s.y + (m == 2) - (m == 0);
        const bool in_map = unsigned(nx) < unsigned(W) && unsigned(ny) < unsigned(H);
        if (is_move && !in_map && !kPolicy) {        // :284, nothing drawn, nothing changed
          general = false;
          status = SHIPENV_ST_OFF_MAP;
          flags = kFlagError;
        } else if (is_move && in_map) {
          const double cost = __dadd_rn(1.0, __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), unit_double(w_cost))));   // :103-104
          const bool no_fuel = s.fuel < cost;                                           // :288
          const bool water = w.passable(nx, ny);                                        // :293
          const uint32_t pd = w.h->port[s.dest];
          const int tx = pd & 255u, ty = (pd >> 8) & 255u;
          const int d2_prev = (s.x - tx) * (s.x - tx) + (s.y - ty) * (s.y - ty);
          const int d2_new = (nx - tx) * (nx - tx) + (ny - ty) * (ny - ty);
          const int px = water ? nx : s.x, py = water ? ny : s.y;
          const bool trigger = s.cargo >= cap || (s.cargo >= 0 && w_loss <= w.h->loss_thr[s.cargo]);   // :318-320
          const bool arrive = px == tx && py == ty;                                     // :325
          const uint32_t c1 = c + 1u;
          const bool ended = no_fuel || (a.max_steps > 0 && int(c1 & 0xFFFFu) >= a.max_steps);
          const bool do_reset = auto_reset && ended;
          // explicit moves: the reset draws are words 2 and 3 of the block (:232-233); the policy's are words 3, 4
          const int ro = int(__umulhi(wd[2], uint32_t(P))), rd = int(__umulhi(wd[3], uint32_t(P)));
          if (!(trigger || arrive || (do_reset && (kPolicy || ro == rd)))) {
            general = false;
            reward = w.h->move_reward[(int(no_fuel) << 2) | (int(!water) << 1) | int(d2_prev > d2_new)];
            if (water) s.fuel = __dsub_rn(s.fuel, cost);                                // :298
            s.x = px; s.y = py;
            c = c1;
            if (do_reset) {                                                             // :227-243
              const uint32_t po = w.h->port[ro];
              s.x = po & 255u; s.y = (po >> 8) & 255u; s.origin = ro; s.dest = rd;
              s.cargo = w.h->cargo0; s.fuel = w.h->fuel0;
              c = (((c1 - 1u) >> 16) + 1u) << 16;
            }
            flags = (no_fuel ? kFlagDone : 0u) | (do_reset ? kFlagReset : 0u);
            taken = m;
          }
        }
      }
      if (general) {
        const GeneralOut o = step_general_core(w, a.round_key, id, 0u, a.max_steps, a.flags, s, c, act);
        reward = o.reward; status = o.status; taken = o.taken; flags = o.flags;
      }
      const size_t at = size_t(t) * n + i;
      if (a.reward) a.reward[at] = float(reward);
      if (a.reward_f64) a.reward_f64[at] = reward;
      if (a.status) a.status[at] = uint8_t(status);
      if (a.taken) a.taken[at] = int16_t(taken);
      const uint32_t done_w = __ballot_sync(live_w, flags & kFlagDone);
      const uint32_t reset_w = __ballot_sync(live_w, flags & kFlagReset);
      const uint32_t err_w = __ballot_sync(live_w, flags & kFlagError);
      if (lane == 0) {
        const size_t word = size_t(t) * nw + (i >> 5);
        if (a.done_bits) a.done_bits[word] = done_w;
        if (a.reset_bits) a.reset_bits[word] = reset_w;
        if (a.error_bits) a.error_bits[word] = err_w;
      }
    }
    st_stream_u4(a.ships + i, pack(s));
    a.clock[i] = c;
  }
}

}  // namespace shipenv
