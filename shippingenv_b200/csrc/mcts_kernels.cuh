// mcts_kernels.cuh -- MCTS on the device: the rollout kernel (simulations x max_rollout_steps in one launch)
// and the whole choose_action (UCB1 tree + expansion + rollout + backprop) per root.
#pragma once
#include "pipeline.cuh"

namespace shipenv {

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
  uint32_t* fifo = reinterpret_cast<uint32_t*>(smem + ((a.world_bytes + 127u) & ~127u));   // [8][blockDim.x] words
  const int64_t total = a.R * int64_t(a.S);
  const int64_t n_tiles = (total + kRolloutThreads - 1) / kRolloutThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t j = tile * kRolloutThreads + threadIdx.x;
    if (j >= total) continue;
    const int64_t r = j / a.S;
    const int sim = int(j - r * a.S);
    ShipR s = unpack(__ldg(a.roots + r));
    RngFifo g;
    g.seek(a.round_key, fifo + threadIdx.x, blockDim.x, uint32_t(sim), a.root_id_base + uint64_t(r));
    double total_reward = 0.0;                     // agents/mcts.py:224
    bool done = false;
    int steps = 0, st = SHIPENV_ST_OK, fails = 0;
    while (!done && steps < a.L) {                 // agents/mcts.py:228
      g.ensure();
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
  uint32_t* fifo = reinterpret_cast<uint32_t*>(smem + ((a.world_bytes + 127u) & ~127u));   // [8][blockDim.x] words
  for (int64_t r = blockIdx.x * int64_t(kMctsThreads) + threadIdx.x; r < a.R; r += int64_t(gridDim.x) * kMctsThreads) {
    const ShipR root = unpack(__ldg(a.roots + r));
    MctsTree t;
    int n_nodes = 1, st = SHIPENV_ST_OK;
    t.total[0] = 0.0; t.visits[0] = 0; t.parent[0] = 255; t.action[0] = 255; t.nchild[0] = 0;
    t.moves_ok[0] = mcts_moves_ok(root);
    for (int sim = 0; sim < a.S && st == SHIPENV_ST_OK; ++sim) {
      ShipR s = root;                                                        // _create_env_copy, :257
      RngFifo g;
      g.seek(a.round_key, fifo + threadIdx.x, blockDim.x, uint32_t(sim), a.root_id_base + uint64_t(r));
      int node = 0;
      bool done = false;
      double rew = 0.0;
      for (;;) {                                                             // selection, :260-268
        if (!t.moves_ok[node]) { st = SHIPENV_ST_MCTS_UNSUPPORTED; break; }
        if (t.nchild[node] != 4) break;
        g.ensure();
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
        g.ensure();
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
        g.ensure();
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
