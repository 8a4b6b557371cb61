// mcts_kernels.cuh -- MCTS on the device: the rollout kernel (simulations x max_rollout_steps in one launch)
// and the whole choose_action (UCB1 tree + expansion + rollout + backprop) per root.
#pragma once
#include "pipeline.cuh"

namespace shipenv {

// ------------------------------------------------------------------------------------------
// One attempt of the rollout loop body `a = env.sample_action(); env.step(a)` (agents/mcts.py:229-236) as
// straight-line code, so that the 32 rollouts of a warp stay converged.
//
// Rollouts from roots near a port (BASELINE configs[3]: post-reset ships advanced by 10 random moves) are full of
// what the stepping kernels treat as rare events: TAKE_CARGO every time the ship stands on a port with an empty
// hold, a cargo-loss draw with probability cargo / 50 on every move after that.  Executed as branches, each of them
// idles most of the warp (22 of 32 lanes active on aged roots, fewer near ports).  Here every lane looks at the next
// seven words of its stream (g.ensure(7) has been called: move choice, fuel cost, loss trigger, loss quantile, three
// beta draws), evaluates the whole MOVE with selects, and advances the stream by what the reference would have
// consumed.  Branches remain only for what really is rare: TAKE_CARGO (needs the port scan), an off-map choice
// (re-drawn), an arrival (destination re-draw loop), and states the reference's sample_action treats differently
// (no destination, fuel == 0 at a port), which go to the general code.
// Returns 0 = step accepted, 1 = rejected (the draw is consumed, the agents' loop tries again),
//         2 = not handled here, nothing consumed.
// ------------------------------------------------------------------------------------------
// The header fields every step looks at, read once per thread: the word FIFO is written in the loop, so the compiler
// must assume the staged world may have changed and would otherwise re-read them from shared memory every step, at the
// head of the step's dependency chains (bounds test, bitmap address).
struct RollHdr {
  int W, H, cap, wpr, nbw, P;
  double p_loss;
  __device__ __forceinline__ explicit RollHdr(const WorldView& w)
      : W(w.h->W), H(w.h->H), cap(w.h->capacity), wpr(w.h->wpr), nbw(w.h->n_bit_words), P(w.h->P), p_loss(w.h->p_loss) {}
};

template <class G>
__device__ __forceinline__ int rollout_attempt(const WorldView& w, const RollHdr& k, ShipR& s, G& g, double& reward, bool& done) {
  const int W = k.W, H = k.H, cap = k.cap;
  if (s.dest == kNone || s.cargo < 0 || unsigned(s.x) >= unsigned(W) || unsigned(s.y) >= unsigned(H)) return 2;
  const bool on_port = (w.bits[k.nbw + s.x * k.wpr + (s.y >> 5)] >> (s.y & 31)) & 1u;
  const uint32_t* const win = g.window();              // the next words of the stream: win[k * stride]
  constexpr uint32_t kS = G::kColumnStride;
  const uint32_t w0 = win[0];
  // sample_action looks at the port first (:247-260).  ONE test decides whether a lane leaves the straight-line code:
  // written as nested ifs (on a port -> cargo != 0 -> fuel == 0) the lanes that sit on a port with cargo on board
  // reached the MOVE code below on a path of their own, and the warp ran all of it twice (ncu, round 2: 1.3 passes
  // of 24 lanes per loop iteration on BASELINE configs[3] roots, where ships linger around their land-locked ports).
  if (on_port && (s.cargo == 0 || s.fuel == 0.0)) {
    if (s.cargo != 0) return 2;                        // fuel == 0 with cargo on board: :258-260 -> :163
    // TAKE_CARGO randint(1, stock) (:255-257, :341-348)
    const int stock = int(w.h->port[current_port(w, s)] >> 24);
    if (stock == 0) return 2;                          // randint(1, 0): the general code rejects it
    s.cargo += 1 + int(__umulhi(w0, uint32_t(stock)));
    reward = w.h->r_take_cargo;
    done = false;
    g.advance(1);
    return 0;
  }
  const int m = int(__umulhi(w0, 4u));                 // choice([N, E, S, W]) (:165-167)
  const int dx = (m & 1) * (m - 2), dy = ((m & 1) ^ 1) * (m - 1);
  const int nx = s.x + dx, ny = s.y + dy;
  if (unsigned(nx) >= unsigned(W) || unsigned(ny) >= unsigned(H)) { g.advance(1); return 1; }   // :284
  const uint32_t w1 = win[kS], w2 = win[2 * kS], w3 = win[3 * kS], w4 = win[4 * kS], w5 = win[5 * kS], w6 = win[6 * kS];
  const double cost = __dadd_rn(1.0, __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), unit_double(w1))));   // :287, :103-104
  const bool no_fuel = s.fuel < cost;                                                   // :288
  const bool water = (w.bits[nx * k.wpr + (ny >> 5)] >> (ny & 31)) & 1u;                // :293
  const uint32_t pd = w.h->port[s.dest];
  const int tx = pd & 255u, ty = (pd >> 8) & 255u;
  const bool closer = (s.x - tx) * dx + (s.y - ty) * dy < 0;                            // :307-315, unit move
  s.x = water ? nx : s.x; s.y = water ? ny : s.y;                                       // :297
  s.fuel = water ? __dsub_rn(s.fuel, cost) : s.fuel;                                    // :298
  double r = w.h->move_reward[(int(no_fuel) << 2) | (int(!water) << 1) | int(closer)];
  // cargo loss (:318-323, :169-200): trigger word, quantile word, median of three words for betavariate(2, 2)
  const bool trigger = w2 <= w.h->loss_thr[min(s.cargo, cap)];
  const double q = unit_double(w3);
  const bool lose_none = q < 0.1, lose_all = q > 0.9;
  const uint32_t med = max(min(w4, w5), min(max(w4, w5), w6));
  const int part = __double2int_rz(__dmul_rn(unit_double(med), double(s.cargo)));
  const bool has_cargo = s.cargo != 0;
  const int loss = (!has_cargo || lose_none) ? 0 : lose_all ? s.cargo : part;
  const double r_loss = __dadd_rn(r, __dmul_rn(double(loss), k.p_loss));
  r = trigger ? r_loss : r;
  s.cargo -= trigger ? loss : 0;
  g.advance(3u + (trigger ? 1u : 0u) + ((trigger && has_cargo && !lose_none && !lose_all) ? 3u : 0u));
  if (s.x == tx && s.y == ty) {                                                         // arrival, :325-337
    r = __dadd_rn(r, __dmul_rn(double(s.cargo), w.h->r_deliver));
    s.cargo = 0;
    s.origin = s.dest;
    const int P = w.h->P;
    int d = g.randint(0, P - 1);
    while (d == s.origin && P > 1) d = g.randint(0, P - 1);
    s.dest = P > 1 ? d : kNone;
    r = __dadd_rn(r, w.h->r_arrive);
  }
  reward = r;
  done = no_fuel;
  return 0;
}

// MCTSAgent._rollout (agents/mcts.py:211-238) from ship s on stream g: at most L accepted steps.
// Returns the status (SHIPENV_ST_OK unless the rollout had to stop); total / steps by reference.
template <class G>
__device__ __forceinline__ int rollout_run(const WorldView& w, const RollHdr& k, ShipR& s, G& g, int L, double& total, int& steps) {
  total = 0.0;                                     // agents/mcts.py:224
  steps = 0;
  bool done = false;
  int fails = 0;
  while (!done && steps < L) {                     // agents/mcts.py:228
    g.top_up();
    g.ensure(7);
    double reward = 0.0;
    int rc = k.P == 0 ? 2 : rollout_attempt(w, k, s, g, reward, done);
    if (rc == 2) {                                 // the general code: every state the reference can be in
      if (w.h->P == 0) return SHIPENV_ST_NO_PORTS;
      if (s.dest == kNone && current_port(w, s) < 0) return SHIPENV_ST_NO_DEST;
      int cat, val;
      if (sample_action(w, s, g, cat, val) != SHIPENV_ST_OK) return SHIPENV_ST_SAMPLE_CRASH;
      const int e = do_action(w, s, g, cat, val, reward, done);
      if (e != SHIPENV_ST_OK) {                    // agents/mcts.py:234-236: try another
        done = false;
        if (++fails >= kMaxTries) return e;
        continue;
      }
    } else if (rc == 1) {
      done = false;
      if (++fails >= kMaxTries) return SHIPENV_ST_OFF_MAP;
      continue;
    }
    fails = 0;
    total = __dadd_rn(total, reward);
    ++steps;
  }
  return SHIPENV_ST_OK;
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
  uint32_t* fifo = reinterpret_cast<uint32_t*>(smem + ((a.world_bytes + 127u) & ~127u));   // [32][blockDim.x] words
  const int64_t total = a.R * int64_t(a.S);
  const int64_t n_tiles = (total + kRolloutThreads - 1) / kRolloutThreads;
  const RollHdr hdr(w);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t j = tile * kRolloutThreads + threadIdx.x;
    if (j >= total) continue;
    const int64_t r = j / a.S;
    const int sim = int(j - r * a.S);
    ShipR s = unpack(__ldg(a.roots + r));
    RngFifo<kRolloutThreads> g;
    g.seek(a.round_key, fifo + threadIdx.x, uint32_t(sim), a.root_id_base + uint64_t(r));
    double total_reward;
    int steps;
    const int st = rollout_run(w, hdr, s, g, a.L, total_reward, steps);
    a.returns[j] = total_reward;
    if (a.final) a.final[j] = pack(s);
    if (a.steps) a.steps[j] = steps;
    if (a.status) a.status[j] = uint8_t(st);
  }
}

// ------------------------------------------------------------------------------------------
// MCTS: the whole choose_action (agents/mcts.py:240-292) per root in one thread
// ------------------------------------------------------------------------------------------
constexpr int kMctsThreads = 64;

struct MctsArgs {
  const void* world; uint32_t world_bytes;
  const uint4* roots; int64_t R; int S, L;
  double c_param;
  int16_t* action; int32_t* root_visits; int16_t* child_action; int32_t* child_visits; double* child_total;
  uint8_t* status;
  uint64_t seed, root_id_base;
  uint32_t round_key[20];
  double log_table[SHIPENV_MCTS_MAX_SIMS + 2];   // log_table[n] = log(n) as the caller computes it (numpy.log)
  int explore_table;                             // the exploration-term table fits in shared memory (S <= 64)
};

// One root's tree (agents/mcts.py:13-40), indexed by node id; at most one node per simulation plus the root, so the
// arrays are sized by the launch's S (kNodes = 64 covers the reference's 50 simulations, 128 the ABI maximum).
template <int kNodes>
struct MctsTree {
  double total[kNodes];
  double mean[kNodes];          // total / visits as of the last backpropagation through the node (UCB1's first term)
  uint8_t visits[kNodes];       // <= S <= 126
  uint8_t parent[kNodes];
  uint8_t info[kNodes];         // action (2 bits) | tried actions << 2 (4 bits: the children's actions) | moves_ok << 6
  alignas(4) uint8_t child[kNodes][4];   // in expansion order; read as one word
  __device__ __forceinline__ int action(int n) const { return info[n] & 3; }
  __device__ __forceinline__ uint32_t tried(int n) const { return (info[n] >> 2) & 15u; }
  __device__ __forceinline__ int nchild(int n) const { return __popc(tried(n)); }
  __device__ __forceinline__ bool moves_ok(int n) const { return info[n] & 64; }
  __device__ __forceinline__ uint32_t children(int n) const { return *reinterpret_cast<const uint32_t*>(child[n]); }
};

// get_possible_actions (:50-89) is the four moves iff a destination is set and not (fuel == 0 with an origin)
__device__ __forceinline__ bool mcts_moves_ok(const ShipR& s) {
  return s.dest != kNone && !(s.fuel == 0.0 && s.origin != kNone);
}

// best_child (:109-128): UCB1, unvisited children first, np.argmax takes the first maximum.
// value = total / n + c * sqrt(2 ln N / n), every operation rounded on its own as numpy does.  The divide and the
// square root are long software sequences in fp64 (a selection step over four children cost ~1,200 instructions):
// the first term is cached per node when its statistics change (backpropagation), the second depends on the two
// visit counts only and comes from a table built once per CTA with the same operations,
// explore[N * (N + 1) / 2 + n] for 1 <= n <= N <= S + 1 (null: S too large for the table, computed in place).
// kFull: the node has all four children (the only case selection meets, :261): the loop is unrolled, so the four
// children's statistics are fetched side by side -- the tree lives in local memory, and a loop over a run-time child
// count made every level of the descent eight dependent L2 round trips (child index -> visits / mean, four times).
template <int kNodes, bool kFull>
__device__ __forceinline__ int mcts_best_child(const MctsTree<kNodes>& t, int node, double c_param, const double* log_table,
                                               const double* explore) {
  int best = 0;
  double best_v = 0.0;
  const int N = t.visits[node];
  const uint32_t kids = t.children(node);
  const double two_log = __dmul_rn(2.0, log_table[N]);
  const int nc = kFull ? 4 : t.nchild(node);
  int n_k[4];
  double mean_k[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = (kids >> (8 * k)) & 255u;
    const bool has = kFull || k < nc;
    n_k[k] = has ? t.visits[ch] : 0;
    mean_k[k] = has ? t.mean[ch] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!kFull && k >= nc) break;
    const int n = n_k[k];
    double v;
    if (n == 0) v = __longlong_as_double(0x7FF0000000000000LL);
    else if (explore && n <= N) v = __dadd_rn(mean_k[k], c_param == 0.0 ? 0.0 : explore[N * (N + 1) / 2 + n]);
    else v = __dadd_rn(mean_k[k], __dmul_rn(c_param, __dsqrt_rn(__ddiv_rn(two_log, double(n)))));
    if (k == 0 || v > best_v) { best = int((kids >> (8 * k)) & 255u); best_v = v; }
  }
  return best;
}

template <int kNodes>
__global__ void __launch_bounds__(kMctsThreads) mcts_kernel(const __grid_constant__ MctsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t world_bar;
  stage_world(smem, a.world, a.world_bytes, &world_bar);
  const WorldView w(smem);
  uint32_t* fifo = reinterpret_cast<uint32_t*>(smem + ((a.world_bytes + 127u) & ~127u));   // [32][blockDim.x] words
  double* log_table = reinterpret_cast<double*>(fifo + RngFifo<kMctsThreads>::kWords * kMctsThreads);
  for (int k = threadIdx.x; k < a.S + 2; k += kMctsThreads) log_table[k] = a.log_table[k];
  // the exploration term c * sqrt(2 ln N / n) for every pair of visit counts a launch can meet
  double* explore = a.explore_table ? log_table + (SHIPENV_MCTS_MAX_SIMS + 2) : nullptr;
  if (explore) {
    const int top = a.S + 1;                         // visits never exceed S; the root's reach S
    for (int e = threadIdx.x; e < (top + 1) * (top + 2) / 2; e += kMctsThreads) {
      int N = 0;
      while ((N + 1) * (N + 2) / 2 <= e) ++N;        // e = N (N + 1) / 2 + n, 0 <= n <= N
      const int n = e - N * (N + 1) / 2;
      explore[e] = n == 0 ? 0.0
                          : __dmul_rn(a.c_param, __dsqrt_rn(__ddiv_rn(__dmul_rn(2.0, a.log_table[N]), double(n))));
    }
  }
  __syncthreads();
  const RollHdr hdr(w);
  for (int64_t r = blockIdx.x * int64_t(kMctsThreads) + threadIdx.x; r < a.R; r += int64_t(gridDim.x) * kMctsThreads) {
    const ShipR root = unpack(__ldg(a.roots + r));
    MctsTree<kNodes> t;
    int n_nodes = 1, st = SHIPENV_ST_OK;
    t.total[0] = 0.0; t.mean[0] = 0.0; t.visits[0] = 0; t.parent[0] = 255;
    t.info[0] = uint8_t(mcts_moves_ok(root) ? 64 : 0);
    for (int sim = 0; sim < a.S && st == SHIPENV_ST_OK; ++sim) {
      ShipR s = root;                                                        // _create_env_copy, :257
      RngFifo<kMctsThreads> g;
      g.seek(a.round_key, fifo + threadIdx.x, uint32_t(sim), a.root_id_base + uint64_t(r));
      int node = 0;
      // the last nodes of the descent, leaf first (255 = none): backpropagation then knows its first levels without
      // walking the parent links, and fetches their statistics side by side
      int up1 = 255, up2 = 255, up3 = 255;
      bool done = false;
      double rew = 0.0;
      for (;;) {                                                             // selection, :260-268
        if (!t.moves_ok(node)) { st = SHIPENV_ST_MCTS_UNSUPPORTED; break; }
        if (t.tried(node) != 15u) break;
        g.ensure();
        up3 = up2; up2 = up1; up1 = node;
        node = mcts_best_child<kNodes, true>(t, node, a.c_param, log_table, explore);
        const int e = do_action(w, s, g, kCatMove, t.action(node), rew, done);
        if (e != SHIPENV_ST_OK || done) break;
      }
      if (st != SHIPENV_ST_OK) break;
      // expansion, :271-279: untried actions in the order N, S, E, W (:80-87) = flat indices 0, 2, 1, 3
      const uint32_t tried = t.tried(node);
      const int nc = __popc(tried);
      const int n_un = 4 - nc;
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
        t.total[ch] = rew; t.mean[ch] = 0.0; t.visits[ch] = 0; t.parent[ch] = uint8_t(node);   // :144-146
        t.info[ch] = uint8_t(act | (mcts_moves_ok(s) ? 64 : 0));
        t.child[node][nc] = uint8_t(ch);
        t.info[node] = uint8_t(t.info[node] | (4u << act));                  // one more child: its action is tried
        up3 = up2; up2 = up1; up1 = node;
        node = ch;
      }
      // simulation: _rollout, :211-238, continuing the same word stream
      double total;
      int steps;
      st = rollout_run(w, hdr, s, g, a.L, total, steps);
      if (st != SHIPENV_ST_OK) break;
      {                                                                      // backpropagate, :149-159
        const int path[4] = {node, up1, up2, up3};
        int v_k[4];
        double t_k[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                  // four levels' statistics in one round trip
          const int n = path[k] == 255 ? 0 : path[k];
          v_k[k] = t.visits[n];
          t_k[k] = t.total[n];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (path[k] != 255) {
            const int n = path[k];
            const double tt = __dadd_rn(t_k[k], total);
            t.visits[n] = uint8_t(v_k[k] + 1);
            t.total[n] = tt;
            t.mean[n] = __ddiv_rn(tt, double(v_k[k] + 1));
          }
        }
        for (int n = up3 == 255 ? 255 : t.parent[up3]; n != 255; n = t.parent[n]) {   // deeper trees: the parent links
          t.visits[n] += 1;
          t.total[n] = __dadd_rn(t.total[n], total);
          t.mean[n] = __ddiv_rn(t.total[n], double(t.visits[n]));
        }
      }
    }
    const int nc0 = t.nchild(0);
    for (int k = 0; k < 4; ++k) {
      const bool has = k < nc0;
      const int ch = has ? t.child[0][k] : 0;
      a.child_action[4 * r + k] = has ? int16_t(t.action(ch)) : int16_t(-1);
      a.child_visits[4 * r + k] = has ? int32_t(t.visits[ch]) : 0;
      a.child_total[4 * r + k] = has ? t.total[ch] : 0.0;
    }
    a.root_visits[r] = t.visits[0];
    int act = -1;
    if (st == SHIPENV_ST_OK) {
      if (nc0 == 0) st = SHIPENV_ST_MCTS_NO_CHILDREN;                        // :291-292
      else act = t.action(mcts_best_child<kNodes, false>(t, 0, 0.0, log_table, explore));   // :288-290
    }
    a.action[r] = int16_t(act);
    a.status[r] = uint8_t(st);
  }
}

}  // namespace shipenv
