// transition.cuh -- the ship transition, one env per thread, state in registers.
//
// Device restatement of Environment.step and what it calls (shipping/environment.py:265-376),
// Environment.reset (:227-243) and Environment.sample_action (:245-263).  Every fp64 operation
// that the reference performs as a separate rounded Python operation is a separate __d*_rn
// intrinsic here, in the same order, so fuel and reward are bit-identical (no FMA contraction).
#pragma once
#include "rng.cuh"
#include "world.cuh"

namespace shipenv {

enum : int { kCatMove = 1, kCatSelect = 2, kCatFuel = 3, kCatCargo = 4, kCatJunk = 99 };  // shipping/type.py:1-5
// not a category of the reference: MOVE_SHIP with a move other than the four of ShipMove (any (dx, dy), :278-287);
// the value is the offset k of the SHIPENV_ACT_MOVE_XY encoding, dx = k / 181 - 90, dy = k % 181 - 90
constexpr int kCatMoveXY = 5;

struct ShipR {  // an unpacked shipenv_ship
  int x, y, origin, dest, cargo;
  double fuel;
};

__device__ __forceinline__ ShipR unpack(const uint4 v) {
  ShipR s;
  s.x = v.x & 255u; s.y = (v.x >> 8) & 255u; s.origin = (v.x >> 16) & 255u; s.dest = v.x >> 24;
  s.cargo = int(v.y);
  s.fuel = __hiloint2double(int(v.w), int(v.z));
  return s;
}
__device__ __forceinline__ uint4 pack(const ShipR& s) {
  uint4 v;
  v.x = uint32_t(s.x) | (uint32_t(s.y) << 8) | (uint32_t(s.origin) << 16) | (uint32_t(s.dest) << 24);
  v.y = uint32_t(s.cargo);
  v.z = uint32_t(__double2loint(s.fuel));
  v.w = uint32_t(__double2hiint(s.fuel));
  return v;
}

// _get_current_port_idx (:145-153): first port at the ship's position, -1 if none.
__device__ __forceinline__ int current_port(const WorldView& w, const ShipR& s) {
  if (unsigned(s.x) >= unsigned(w.h->W) || unsigned(s.y) >= unsigned(w.h->H) || !w.port_cell(s.x, s.y)) return -1;
  const uint32_t key = uint32_t(s.x) | (uint32_t(s.y) << 8);
  int found = -1;
  for (int k = w.h->P - 1; k >= 0; --k)
    if ((w.h->port[k] & 0xFFFFu) == key) found = k;
  return found;
}

// reset (:227-243).  P >= 2 is enforced on the host (the reference never terminates for P == 1).
template <class G>
__device__ __forceinline__ void env_reset(const WorldView& w, ShipR& s, G& g) {
  const int P = w.h->P;
  s.cargo = w.h->cargo0;
  s.fuel = w.h->fuel0;
  const int o = g.randint(0, P - 1);
  int d = g.randint(0, P - 1);
  while (o == d && P > 1) d = g.randint(0, P - 1);
  s.origin = o; s.dest = d;
  const uint32_t p = w.h->port[o];
  s.x = p & 255u; s.y = (p >> 8) & 255u;
}

// _move_ship (:273-339) by (dx, dy).  kUnit: one of the four moves of ShipMove, whose distance is exactly 1.
template <bool kUnit, class G>
__device__ __forceinline__ int do_move_xy(const WorldView& w, ShipR& s, G& g, int dx, int dy, double& reward, bool& done) {
  if (s.dest == kNone) return SHIPENV_ST_NO_DEST;                               // :276
  const int sx = s.x, sy = s.y, nx = sx + dx, ny = sy + dy;
  if (unsigned(nx) >= unsigned(w.h->W) || unsigned(ny) >= unsigned(w.h->H)) return SHIPENV_ST_OFF_MAP;  // :284
  // :287 -> :103-104: sqrt(dx^2 + dy^2) * (1 + uniform(-0.1, 0.1)), uniform = a + (b - a) * random().  np.sqrt is
  // correctly rounded, like __dsqrt_rn; for a unit move the distance is 1.0 and the product is its other factor.
  const double u = __dadd_rn(-0.1, __dmul_rn(0.1 - (-0.1), g.random()));
  double cost = __dadd_rn(1.0, u);
  if (!kUnit) cost = __dmul_rn(__dsqrt_rn(double(dx * dx + dy * dy)), cost);
  const bool no_fuel = s.fuel < cost;                                            // :288
  const bool water = w.passable(nx, ny);                                         // :293
  if (water) {                                                                   // :297-298
    s.x = nx; s.y = ny;
    s.fuel = __dsub_rn(s.fuel, cost);
  }
  // :307-315 on the ATTEMPTED cell.  sqrt is strictly monotone on these integers, so the
  // sign of d_prev - d_new is the sign of the squared-distance difference.
  const uint32_t pd = w.h->port[s.dest];
  const int tx = pd & 255u, ty = (pd >> 8) & 255u;
  const int d2_prev = (sx - tx) * (sx - tx) + (sy - ty) * (sy - ty);
  const int d2_new = (nx - tx) * (nx - tx) + (ny - ty) * (ny - ty);
  double r = w.h->move_reward[(int(no_fuel) << 2) | (int(!water) << 1) | int(d2_prev > d2_new)];
  // :318-323 cargo loss
  const uint32_t wt = g.next();
  const bool trigger = s.cargo >= w.h->capacity || (s.cargo >= 0 && wt <= w.h->loss_thr[s.cargo]);
  if (trigger) {
    const double q = g.random();                                                 // :177
    int loss = 0;
    if (s.cargo != 0) {                                                          // :180-198
      if (q < 0.1) loss = 0;
      else if (q > 0.9) loss = s.cargo;
      else loss = __double2int_rz(__dmul_rn(g.beta22(), double(s.cargo)));
    }
    s.cargo -= loss;
    r = __dadd_rn(r, __dmul_rn(double(loss), w.h->p_loss));
  }
  // :325-337 arrival, on the post-move position
  if (s.x == tx && s.y == ty) {
    r = __dadd_rn(r, __dmul_rn(double(s.cargo), w.h->r_deliver));
    s.cargo = 0;
    s.origin = s.dest;
    const int P = w.h->P;
    int d = g.randint(0, P - 1);
    while (d == s.origin && P > 1) d = g.randint(0, P - 1);
    s.dest = P > 1 ? d : kNone;  // P == 1: the reference loops forever (:334); documented divergence
    r = __dadd_rn(r, w.h->r_arrive);
  }
  reward = r;
  done = no_fuel;
  return SHIPENV_ST_OK;
}
// m = 0..3 = N, E, S, W = (0,-1), (-1,0), (0,1), (1,0) (shipping/type.py:13-16)
template <class G>
__device__ __forceinline__ int do_move(const WorldView& w, ShipR& s, G& g, int m, double& reward, bool& done) {
  return do_move_xy<true>(w, s, g, (m == 3) - (m == 1), (m == 2) - (m == 0), reward, done);
}

// step's dispatch (:359-376) with _select_port (:265-271), _take_cargo (:341-348), _take_fuel (:350-357).
template <class G>
__device__ __forceinline__ int do_action(const WorldView& w, ShipR& s, G& g, int cat, int val,
                                         double& reward, bool& done) {
  if (w.h->P == 0) return SHIPENV_ST_NO_PORTS;                                   // :360
  done = false;
  if (cat == kCatMove) return do_move(w, s, g, val, reward, done);
  if (cat == kCatMoveXY) return do_move_xy<false>(w, s, g, val / 181 - 90, val % 181 - 90, reward, done);
  if (cat == kCatSelect) {
    if (val < 0 || val >= w.h->P) return SHIPENV_ST_SEL_RANGE;
    if (s.origin == val) return SHIPENV_ST_SEL_ORIGIN;
    s.dest = val;
    reward = 0.0;
    return SHIPENV_ST_OK;
  }
  if (cat == kCatCargo || cat == kCatFuel) {
    const int k = current_port(w, s);
    if (k < 0) return SHIPENV_ST_NOT_AT_PORT;
    const uint32_t p = w.h->port[k];
    const int stock = cat == kCatCargo ? int(p >> 24) : int((p >> 16) & 255u);
    if (!(0 < val && val <= stock)) return SHIPENV_ST_BAD_AMOUNT;
    if (cat == kCatCargo) { s.cargo += val; reward = w.h->r_take_cargo; }
    else { s.fuel = __dadd_rn(s.fuel, double(val)); reward = w.h->r_take_fuel; }
    return SHIPENV_ST_OK;
  }
  return SHIPENV_ST_UNKNOWN;                                                     // :374
}

// flat index -> (category, value), utils/preprocessing.py:111-137 plus the extended amount forms
__device__ __forceinline__ void decode_action(int a, int P, int& cat, int& val) {
  cat = kCatJunk; val = 0;
  if (a >= 0 && a < 4) { cat = kCatMove; val = a; }
  else if (a >= 4 && a < 4 + P) { cat = kCatSelect; val = a - 4; }
  else if (a >= 4 + P && a < 4 + P + SHIPENV_N_CARGO_ACTIONS) { cat = kCatCargo; val = a - (4 + P); }
  else if (a >= 4 + P + SHIPENV_N_CARGO_ACTIONS && a < 4 + P + SHIPENV_N_CARGO_ACTIONS + SHIPENV_N_FUEL_ACTIONS) {
    cat = kCatFuel; val = a - (4 + P + SHIPENV_N_CARGO_ACTIONS);
  }
  else if (a >= SHIPENV_ACT_EXT_CARGO && a < SHIPENV_ACT_EXT_FUEL) { cat = kCatCargo; val = a - SHIPENV_ACT_EXT_CARGO; }
  else if (a >= SHIPENV_ACT_EXT_FUEL && a < 0x8000) { cat = kCatFuel; val = a - SHIPENV_ACT_EXT_FUEL; }
  else if (a <= SHIPENV_ACT_MOVE_XY_BASE && a >= SHIPENV_ACT_MOVE_XY_BASE - 181 * 181 + 1) { cat = kCatMoveXY; val = SHIPENV_ACT_MOVE_XY_BASE - a; }
}
__device__ __forceinline__ int encode_action(int cat, int val, int P) {
  if (cat == kCatMove) return val;
  if (cat == kCatMoveXY) return SHIPENV_ACT_MOVE_XY_BASE - val;
  if (cat == kCatSelect) return 4 + val;
  if (cat == kCatCargo)
    return (val >= 0 && val < SHIPENV_N_CARGO_ACTIONS) ? 4 + P + val : SHIPENV_ACT_EXT_CARGO + val;
  return (val >= 0 && val < SHIPENV_N_FUEL_ACTIONS) ? 4 + P + SHIPENV_N_CARGO_ACTIONS + val
                                                    : SHIPENV_ACT_EXT_FUEL + val;
}

// sample_action (:245-263).  Returns 0 or SHIPENV_ST_SAMPLE_CRASH (the fuel == 0 branch, :163).
template <class G>
__device__ __forceinline__ int sample_action(const WorldView& w, const ShipR& s, G& g, int& cat, int& val) {
  const int cur = current_port(w, s);
  const int P = w.h->P;
  if (s.dest == kNone && cur >= 0) {                                             // :247-254
    int k = g.randint(0, P - 1);
    while (k == cur && P > 1) k = g.randint(0, P - 1);
    cat = kCatSelect; val = k;
  } else if (s.cargo == 0 && cur >= 0) {                                         // :255-257
    cat = kCatCargo; val = g.randint(1, int(w.h->port[cur] >> 24));
  } else if (s.fuel == 0.0 && cur >= 0) {                                        // :258-260 -> :163
    return SHIPENV_ST_SAMPLE_CRASH;
  } else {                                                                       // :261-263, :165-167
    cat = kCatMove; val = g.below(4);
  }
  return SHIPENV_ST_OK;
}

// `a = sample_action(); try: step(a) except: continue` (agents/mcts.py:228-236), bounded by kMaxTries.
template <class G>
__device__ __forceinline__ int policy_step(const WorldView& w, ShipR& s, G& g, double& reward, bool& done, int& taken) {
  if (w.h->P == 0) return SHIPENV_ST_NO_PORTS;
  if (s.dest == kNone && current_port(w, s) < 0) return SHIPENV_ST_NO_DEST;  // the reference spins forever here
  int last = SHIPENV_ST_OFF_MAP;
  for (int t = 0; t < kMaxTries; ++t) {
    int cat, val;
    if (sample_action(w, s, g, cat, val) != SHIPENV_ST_OK) return SHIPENV_ST_SAMPLE_CRASH;
    const int st = do_action(w, s, g, cat, val, reward, done);
    if (st == SHIPENV_ST_OK) { taken = encode_action(cat, val, w.h->P); return SHIPENV_ST_OK; }
    last = st;
  }
  return last;
}

}  // namespace shipenv
