/* shipenv.h -- C ABI of the B200-native batched ShippingEnv step.
 *
 * The reference (VanshJP/ShippingEnv) is pure Python and has no FFI of its own; these
 * entry points are the surface a maintainer binds with ctypes in place of the Python
 * methods cited on each function (see INTEGRATION.md for the stub).  Citations are
 * file:line in the reference checkout.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SHIPENV_E_* code otherwise;
 *     shipenv_last_error() gives the message of the calling thread's last failure;
 *   - `ships`, `clock`, `actions` and every output pointer are DEVICE pointers owned by
 *     the caller (torch tensors: tensor.data_ptr()); the library allocates nothing per
 *     call.  Arguments of shipenv_set_* are HOST pointers (tiny tables, copied);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued on it (shipenv_step_host also uses internal streams, forked from and
 *     joined to it); nothing blocks the host except shipenv_stream_synchronize;
 *   - two entry points serve callers whose data lives in HOST memory, as the reference's does:
 *     shipenv_step_host (batched: host actions in, host rewards / masks out) and the N = 1 pair
 *     shipenv_host_device_pointer / shipenv_step_sample (the facade's page-locked ship block);
 *   - a context belongs to one device; calls on one context must not race (one host thread at a time).  Launches
 *     of one context may be in flight on several streams at once: the library orders its own world upload before
 *     them and gives each launch its own scheduling state.  Two launches that touch the same ships are the
 *     caller's to order, as with any kernel.
 *
 * Batched contract ("Batched contract"):
 *   env i of a launch has global id env_id_base + i.  Draws of one accepted step come
 *   from the Philox4x32-10 cell (seed, env id, clock[i]), words word0, word0+1, ...;
 *   clock = episode << 16 | step_in_episode; a reset moves it to the next multiple of
 *   65536.  Rejected actions (status != 0) change nothing and consume nothing.
 */
#ifndef SHIPENV_H_
#define SHIPENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHIPENV_ABI_VERSION 2
#define SHIPENV_MAX_PORTS 32      /* ports per world (reference default: 5, utils/constants.py:57-63) */
#define SHIPENV_MAX_DIM 256       /* map side; coordinates are u8 (reference: 100, shipping/environment.py:29) */
#define SHIPENV_MAX_CAPACITY 255  /* Initial.MAX_CARGO_CAPACITY upper bound (reference: 50) */
#define SHIPENV_NONE_PORT 255     /* origin/dest == None (shipping/environment.py:40-41) */
#define SHIPENV_MAX_TRIES 32      /* policy / rollout: consecutive rejected draws before giving up */

/* Flat action index (int16), utils/preprocessing.py:111-137, P = number of ports:
 *   [0,4)            MOVE   N,E,S,W = (0,-1),(-1,0),(0,1),(1,0)   (shipping/type.py:13-16)
 *   [4,4+P)          SELECT_PORT a-4
 *   [4+P,4+P+50)     TAKE_CARGO amount a-(4+P)       (amount 0 is rejected)
 *   [4+P+50,4+P+250) TAKE_FUEL  amount a-(4+P+50)
 *   -1               policy: sample_action() until a step is accepted (shipping/environment.py:245-263,
 *                    agents/mcts.py:228-236)
 *   [0x4000,0x6000)  TAKE_CARGO amount a-0x4000  (amounts the flat space cannot express)
 *   [0x6000,0x8000)  TAKE_FUEL  amount a-0x6000
 *   [-32762,-2]      MOVE by any (dx, dy) with |dx|, |dy| <= 90, a = -2 - ((dx+90)*181 + (dy+90)): the reference's
 *                    _move_ship takes any pair (shipping/environment.py:278-287); the fuel cost is then
 *                    sqrt(dx^2+dy^2) * (1 + uniform(-0.1, 0.1)) and only the target cell's terrain is looked at
 *   anything else    rejected, SHIPENV_ST_UNKNOWN */
#define SHIPENV_ACT_POLICY (-1)
#define SHIPENV_ACT_NONE (-32768) /* output only: shipenv_sample_action / next_action when the draw failed (status says why) */
#define SHIPENV_ACT_MOVE_XY_BASE (-2)
#define SHIPENV_MOVE_XY_MAX 90
#define SHIPENV_ACT_MOVE_XY(dx, dy) (SHIPENV_ACT_MOVE_XY_BASE - (((dx) + 90) * 181 + ((dy) + 90)))
#define SHIPENV_ACT_EXT_CARGO 0x4000
#define SHIPENV_ACT_EXT_FUEL 0x6000
#define SHIPENV_N_CARGO_ACTIONS 50
#define SHIPENV_N_FUEL_ACTIONS 200

/* Per-env status: one code per exception of the reference (shipping/environment.py). */
enum {
  SHIPENV_ST_OK = 0,
  SHIPENV_ST_OFF_MAP = 1,      /* ValueError("Move is out of range")                            :284 */
  SHIPENV_ST_NO_DEST = 2,      /* Exception("Cannot move without destination port")             :276 */
  SHIPENV_ST_SEL_RANGE = 3,    /* IndexError("Port index is out of range")                      :269 */
  SHIPENV_ST_SEL_ORIGIN = 4,   /* Exception("Destination port must be different from current one") :267 */
  SHIPENV_ST_NOT_AT_PORT = 5,  /* Exception("Not currently at port")                            :343,352 */
  SHIPENV_ST_BAD_AMOUNT = 6,   /* ValueError("Invalid fuel amount")                             :346,355 */
  SHIPENV_ST_UNKNOWN = 7,      /* ValueError("Action category unknown")                         :374 */
  SHIPENV_ST_NO_PORTS = 8,     /* Exception("No ports available")                               :360 */
  SHIPENV_ST_SAMPLE_CRASH = 9  /* sample_action with fuel == 0 at a port indexes a float        :163 */
};

enum {
  SHIPENV_E_ARG = -1,      /* bad argument (message says which) */
  SHIPENV_E_CUDA = -2,     /* a CUDA runtime call failed */
  SHIPENV_E_STATE = -3     /* context not ready (no map / no ports) */
};

/* step flags */
#define SHIPENV_F_AUTO_RESET 1u   /* reset an env in the step that ends its episode (done or truncated) */
#define SHIPENV_F_ALL_POLICY 2u   /* every action of this call is SHIPENV_ACT_POLICY; `actions` is not read and may
                                     be NULL when only reward_f32 and the three masks are requested */
#define SHIPENV_F_CHAINED 4u      /* this step may overlap the previous shipenv_step launch of the stream at the launch
                                     boundary (needs shipenv_step_io.chain; see there).  The caller asserts that every
                                     input of this call other than ships / clock -- the `actions` array -- was complete
                                     before that previous launch was enqueued (pre-generated action streams, policy
                                     steps).  Ignored, i.e. the launch waits for all earlier work of the stream as usual,
                                     unless the previous launch of this context was a shipenv_step on the same chain
                                     array and stream, with the same n and output set (i.e. the same kernel and grid),
                                     that itself took the chain protocol. */

/* One ship, 16 bytes, moved as a single 128-bit load/store.  Mirrors the attributes
 * Environment keeps per ship (shipping/environment.py:37-41). */
typedef struct {
  uint8_t x, y;            /* ship_position = [x, y]; np_game[x, y]                 */
  uint8_t origin, dest;    /* origin/destination_port_index; 255 = None             */
  int32_t cargo;           /* Environment.cargo (true cargo, not the state-dict copy of fuel) */
  double fuel;             /* Environment.fuel (float64 after the first water move)  */
} shipenv_ship;

/* Initial / Reward / Penalty (shipping/environment.py:8-26). */
typedef struct {
  int32_t cargo0;          /* Initial.CARGO = 0 */
  int32_t capacity;        /* Initial.MAX_CARGO_CAPACITY = 50 */
  double fuel0;            /* Initial.FUEL = 200 */
  double r_deliver;        /* Reward.CARGO_DELIVER = 2 (per unit) */
  double r_arrive;         /* Reward.REACH_DESTINATION = 10 */
  double r_closer;         /* Reward.CLOSER_TO_DESTINATION = 2 */
  double r_take_fuel;      /* Reward.TAKE_FUEL = 0.05 */
  double r_take_cargo;     /* Reward.TAKE_CARGO = 0.05 */
  double p_no_fuel;        /* Penalty.RUN_OUT_OF_FUEL = -10 */
  double p_ground;         /* Penalty.MOVE_ON_GROUND = -5 */
  double p_water;          /* Penalty.MOVE_ON_WATER = -1 */
  double p_loss;           /* Penalty.CARGO_LOSS = -3 (per unit) */
  double p_farther;        /* Penalty.FARTHER_FROM_DESTINATION = -2 */
  double p_use_fuel;       /* Penalty.USE_FUEL = -0.0001 */
} shipenv_consts;

/* Outputs of one batched step; every pointer is a device pointer or NULL (= not wanted).
 * Bit i%32 of word i/32 belongs to env i; words hold ceil(n/32) entries. */
typedef struct {
  float *reward_f32;       /* [n] reward, computed in fp64 in the reference's order, rounded once */
  double *reward_f64;      /* [n] the fp64 reward itself */
  uint32_t *done_bits;     /* `done` returned by Environment.step (out of fuel, :288-290) */
  uint32_t *reset_bits;    /* env was auto-reset in this step (done, or step count reached max_steps) */
  uint32_t *error_bits;    /* status != 0 */
  uint8_t *status;         /* [n] SHIPENV_ST_* */
  int16_t *taken;          /* [n] the action executed (differs from the input only for policy actions) */
  float *obs;              /* [n, 6+4P] preprocess_state row of the state AFTER the step (and reset) */
  double *ep_return;       /* [n] running return of the current episode (read-modify-write) */
  double *stats;           /* [6] episode statistics accumulated at episode end (needs ep_return):
                              sum return, sum return^2, sum length, episodes, max return, min return */
  uint32_t *chain;         /* NULL, or shipenv_chain_words(n) words of device memory that belong to this (ships, n)
                              pair: zeroed once by the caller, then only touched by shipenv_step.  With it the step
                              kernels split the env range statically over their CTAs and record, per CTA, how many
                              launches have finished their range; a launch flagged SHIPENV_F_CHAINED lets CTA b start
                              as soon as CTA b of the previous launch has finished, instead of waiting for the whole
                              previous grid (consecutive steps of one env group then overlap at the launch boundary;
                              any other kernel still sees normal stream order).  Every shipenv_step on these ships
                              must pass the same array (or none of them may be chained). */
} shipenv_step_io;

typedef struct shipenv_ctx shipenv_ctx;

/* ---- lifetime ------------------------------------------------------------------------ */
int shipenv_abi_version(void);
const char *shipenv_last_error(void);
/* Environment.__init__ (shipping/environment.py:29-43): one context per (process, device). */
int shipenv_ctx_create(int device, shipenv_ctx **out);
int shipenv_ctx_destroy(shipenv_ctx *ctx);

/* ---- world: replaces Environment._initialize_map / add_port state ------------------ */
/* passable[x*H + y] != 0  <=>  np_game[x, y] != Entity.GROUND (shipping/environment.py:45-55,
 * :293).  The JPEG decode/resize/threshold stays on the host (cv2), only the grid comes here. */
int shipenv_set_map(shipenv_ctx *ctx, const uint8_t *passable, int W, int H);
/* port_positions / port_fuel / port_cargo (shipping/environment.py:33-35, :57-65). */
int shipenv_set_ports(shipenv_ctx *ctx, const uint8_t *px, const uint8_t *py, const uint8_t *pfuel,
                      const uint8_t *pcargo, int P);
int shipenv_set_consts(shipenv_ctx *ctx, const shipenv_consts *c);  /* NULL = reference defaults */
int shipenv_get_consts(shipenv_ctx *ctx, shipenv_consts *out);
int shipenv_obs_width(shipenv_ctx *ctx);                            /* 6 + 4P, utils/preprocessing.py:53-62 */
int shipenv_num_actions(shipenv_ctx *ctx);                          /* 4+P+50+200, utils/preprocessing.py:93-108 */

/* ---- the hot path -------------------------------------------------------------------- */
/* Environment.reset (shipping/environment.py:227-243) for every env whose mask bit is set
 * (mask_bits == NULL: all).  Draws from cell (seed, env, clock) then starts a new episode. */
int shipenv_reset(shipenv_ctx *ctx, shipenv_ship *ships, uint32_t *clock, uint64_t seed,
                  uint64_t env_id_base, int64_t n, const uint32_t *mask_bits, void *stream);

/* Environment.step (shipping/environment.py:359-376) and everything it calls (:265-357),
 * for n envs; policy actions also run sample_action (:245-263).  max_steps == 0 disables
 * truncation; word0 is the first word of the cell to use (0 unless the caller already
 * drew from it, e.g. sample_action() followed by step() on the N=1 facade). */
int shipenv_step(shipenv_ctx *ctx, shipenv_ship *ships, uint32_t *clock, const int16_t *actions,
                 const shipenv_step_io *io, uint64_t seed, uint64_t env_id_base, int64_t n,
                 int max_steps, uint32_t flags, uint32_t word0, void *stream);

/* T consecutive Environment.step calls per env in one launch -- the loops of the reference's agents,
 * `for t: env.step(a_t)` (agents/mcts.py:228-236, agents/sarsa.py:222-313), when the T actions of an env are known
 * up front: actions[t * n + i] (int16 [T, n], NULL with SHIPENV_F_ALL_POLICY: the random policy).  The ship and its
 * clock stay in registers for the T steps (auto-reset included), so an env-step moves 6 B instead of 46 B and a small
 * batch is not bound by one launch per step.  Bit-identical to T calls of shipenv_step with word0 = 0.  Outputs, each
 * nullable: reward / reward_f64 / status / taken [T, n]; done / reset / error ballots [T, ceil(n/32)]. */
int shipenv_step_many(shipenv_ctx *ctx, shipenv_ship *ships, uint32_t *clock, const int16_t *actions, int T,
                      float *reward, uint32_t *done_bits, uint32_t *reset_bits, uint32_t *error_bits,
                      uint8_t *status, int16_t *taken, double *reward_f64, uint64_t seed, uint64_t env_id_base,
                      int64_t n, int max_steps, uint32_t flags, void *stream);

/* Words of device memory shipenv_step_io.chain needs for n envs. */
size_t shipenv_chain_words(int64_t n);

/* Environment.sample_action (shipping/environment.py:245-263), one call per env, drawing
 * from word0 of the cell (seed, env, clock).  words_used[i] (nullable) = words consumed;
 * status[i] (nullable) = 0 or SHIPENV_ST_SAMPLE_CRASH / SHIPENV_ST_NO_PORTS. */
int shipenv_sample_action(shipenv_ctx *ctx, const shipenv_ship *ships, const uint32_t *clock,
                          int16_t *actions_out, uint8_t *words_used, uint8_t *status,
                          uint64_t seed, uint64_t env_id_base, int64_t n, uint32_t word0, void *stream);

/* Environment.step for callers whose actions and results live in HOST memory, as the reference's do (Python lists,
 * shipping/environment.py:359-376): h_actions[n] in, h_reward[n] and the three ballot arrays (nullable) out; ships
 * and clock stay on the device.  With chunks == 0, batches of 2 Mi envs and more are cut into up to four chunks (whole
 * 512-env tiles, at least 1 Mi envs each) whose host-to-device copy, kernel and device-to-host copy are pipelined on
 * internal streams, so that one call keeps both PCIe directions and the SMs busy; smaller batches run entirely on
 * `stream`.  chunks = 1..4 fixes the count: 1 is the right choice when the caller already overlaps several
 * independent batches on streams of its own (measured: 2 chunks give +10 % for one 2 Mi-env batch, -8 % for four).  The call is ordered after prior work on `stream`, and `stream` waits for all of it; nothing
 * blocks the host (page-locked host buffers are needed for the copies to be asynchronous).  `workspace` is caller-owned
 * device memory of shipenv_step_host_workspace(n) bytes, 256-byte aligned (staging for actions, rewards, masks).
 * Host outputs that lie back to back in memory -- h_reward[n], then the done, reset and error words -- are fetched
 * with one copy instead of four (each copy costs ~10 us of the copy engine whatever its size). */
size_t shipenv_step_host_workspace(int64_t n);
int shipenv_step_host(shipenv_ctx *ctx, shipenv_ship *ships, uint32_t *clock, const int16_t *h_actions,
                      float *h_reward, uint32_t *h_done_bits, uint32_t *h_reset_bits, uint32_t *h_error_bits,
                      void *workspace, size_t workspace_bytes, uint64_t seed, uint64_t env_id_base, int64_t n,
                      int max_steps, uint32_t flags, int chunks, void *stream);

/* shipenv_step followed by the shipenv_sample_action call the reference's loops make next
 * (`a = env.sample_action(); env.step(a)`: agents/mcts.py:222-230, agents/sarsa.py:211-220), in one launch:
 * next_action[i] / next_words[i] / next_status[i] (the last two nullable) are what shipenv_sample_action would
 * return for env i after this step -- drawn from word 0 of the new cell when the step was accepted, from `word0`
 * of the unchanged cell when it was rejected.  io->obs must be NULL.  General path, one env per thread: meant
 * for the N = 1 facade (sample_action() after step() then costs no launch), correct for any n. */
int shipenv_step_sample(shipenv_ctx *ctx, shipenv_ship *ships, uint32_t *clock, const int16_t *actions,
                        const shipenv_step_io *io, int16_t *next_action, uint8_t *next_words,
                        uint8_t *next_status, uint64_t seed, uint64_t env_id_base, int64_t n, int max_steps,
                        uint32_t flags, uint32_t word0, void *stream);

/* _build_state + preprocess_state (shipping/environment.py:202-225, utils/preprocessing.py:25-62)
 * as float32 rows [n, 6+4P] (the dtype the DQN feeds its network, agents/dqn.py:198). */
int shipenv_observe(shipenv_ctx *ctx, const shipenv_ship *ships, float *obs, int64_t n, void *stream);

/* DQNAgent.is_valid_action for every flat action (agents/dqn.py:125-175): moves always; SELECT_PORT k iff
 * k != origin and the ship stands on port k; TAKE_CARGO / TAKE_FUEL v iff the ship is at a port (first
 * match) and 0 < v <= that port's stock.  mask[i * words + a / 32] bit a % 32, words = ceil(A / 32),
 * A = shipenv_num_actions (9 words for P = 5). */
int shipenv_valid_actions(shipenv_ctx *ctx, const shipenv_ship *ships, uint32_t *mask, int64_t n, void *stream);

/* discretize_state (utils/preprocessing.py:65-90), the key of the tabular agents (agents/sarsa.py:148,187-188):
 * keys[i] = {x, y, min(int("cargo" / 10), 4), min(int(fuel / 20), 4), origin | -1, dest | -1} as int16[6],
 * where "cargo" is the state dict's copy of fuel (shipping/environment.py:206).  Buckets are clamped to the
 * int16 range (fuel below -327,680 only). */
int shipenv_discretize(shipenv_ctx *ctx, const shipenv_ship *ships, int16_t *keys, int64_t n, void *stream);

/* MCTSAgent._rollout (agents/mcts.py:211-238) from copies of R root ships, S rollouts each,
 * at most L accepted steps.  Rollout (r, s) draws one continuous word stream from the cell
 * (seed ^ "ROLL"<<32, root_id_base + r, s).  returns[r*S+s] is the fp64 sum of step rewards;
 * final/steps/status are nullable ([R*S] each). */
int shipenv_rollout(shipenv_ctx *ctx, const shipenv_ship *roots, int64_t R, int S, int L,
                    double *returns, shipenv_ship *final, int32_t *steps, uint8_t *status,
                    uint64_t seed, uint64_t root_id_base, void *stream);

/* MCTSAgent.choose_action (agents/mcts.py:240-292) for R roots, one tree per root, entirely on the device:
 * S simulations of UCB1 selection (:109-128, :260-268), expansion by random.choice of an untried action
 * (:271-279), _rollout of at most L accepted steps (:211-238) and backpropagation (:149-159), then the
 * final pick with c = 0 (:288-290).  Simulation s of root r draws from the rollout cell (root_id_base + r, s):
 * tree-phase draws first, the rollout continues the same stream.  Nodes whose recorded state is not the
 * four-move case (destination unset, or fuel == 0 at a port: get_possible_actions :63-77) end the root with
 * status SHIPENV_ST_MCTS_UNSUPPORTED; a root without children (:291-292) gets SHIPENV_ST_MCTS_NO_CHILDREN and
 * action -1.  log_table is a HOST array with log_table[n] = log(n) for n = 1..S+1 (index 0 unused), or NULL to
 * let the library fill it with the C library's log(); passing numpy.log values makes UCB1 bit-identical to
 * the reference.  S <= SHIPENV_MCTS_MAX_SIMS.  Outputs (device): action[R] flat move index, root_visits[R],
 * child_action / child_visits / child_total [R, 4] in expansion order (-1 / 0 / 0 when absent), status[R]. */
#define SHIPENV_MCTS_MAX_SIMS 126
#define SHIPENV_ST_MCTS_UNSUPPORTED 10
#define SHIPENV_ST_MCTS_NO_CHILDREN 11
int shipenv_mcts(shipenv_ctx *ctx, const shipenv_ship *roots, int64_t R, int S, int L, double c_param,
                 const double *log_table, int16_t *action, int32_t *root_visits, int16_t *child_action,
                 int32_t *child_visits, double *child_total, uint8_t *status, uint64_t seed,
                 uint64_t root_id_base, void *stream);

/* ---- host-side helpers (no GPU) ------------------------------------------------------ */
/* Philox4x32-10 block (the word-stream definition). */
void shipenv_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The two random.randint(5, 20) draws of add_port (shipping/environment.py:63-64) for port
 * `port_index`, from the cell (seed ^ "PORT"<<32, 0xFFFFFFFFFFFFFFFF, port_index). */
void shipenv_port_stock(uint64_t seed, uint32_t port_index, uint8_t *fuel, uint8_t *cargo);
/* kernels launched by this process so far (bench.py's gpu_launches) */
uint64_t shipenv_launch_count(void);

/* ---- host-resident single ship: what the N = 1 drop-in Environment uses ------------------- */
/* The reference's Environment keeps its ship in Python attributes (shipping/environment.py:36-42).  The facade
 * mirrors them in one page-locked host block and lets the N = 1 kernels read and write that block in place over
 * PCIe instead of staging it through device memory.  `host_ptr` must be page-locked (cudaHostAlloc /
 * cudaHostRegister / torch pin_memory); *device_ptr is the address the entry points above accept for it. */
int shipenv_host_device_pointer(void *host_ptr, void **device_ptr);
/* Blocks the caller until the work enqueued on `stream` so far has completed (the facade's step() returns
 * Python values, like shipping/environment.py:376, so it has to wait).  Nothing else in this ABI synchronises. */
int shipenv_stream_synchronize(shipenv_ctx *ctx, void *stream);

/* The N = 1 step without a launch per call: Environment.step (shipping/environment.py:359-376) followed by the draw of
 * the sample_action call that usually comes next (:245-263), exactly what shipenv_step_sample computes for n = 1, served
 * by a resident one-warp kernel that polls a page-locked mailbox owned by the context.  The call writes the request
 * (ship, clock, action, word0, seed, env id, max_steps, flags -- the server keeps no env state, so any number of
 * Environment objects share it), spins until the kernel has answered and copies the answer into `block`; no launch and
 * no stream synchronisation on the way (~6 us instead of ~23 us per step).  The kernel retires by itself after
 * `idle_us` without a request (default 2000; a device-wide synchronisation elsewhere in the process waits at most that
 * long) and is launched again by the next call; it is restarted when the world of the context has changed.
 * `block` is HOST memory (it need not be page-locked) laid out like shipenv_host_block.  Returns 0 also when the step
 * was rejected (block->status says why, nothing changed). */
typedef struct shipenv_host_block {
  shipenv_ship ship;        /*   0  in: the ship; out: the ship after the step */
  double reward;            /*  16  out */
  uint8_t pad0[8];
  uint32_t clock;           /*  32  in / out */
  uint8_t pad1[12];
  int16_t action;           /*  48  in: flat action index */
  uint8_t pad2[14];
  uint32_t done_bits, reset_bits, error_bits;   /* 64, 68, 72  out: bit 0 */
  uint8_t status;           /*  76  out: SHIPENV_ST_* of the step */
  uint8_t pad3;
  int16_t taken;            /*  78  (used by the facade's sample_action path) */
  uint8_t words;            /*  80  (the same) */
  uint8_t pad4[3];
  int16_t next_action;      /*  84  out: what sample_action would return after this step (SHIPENV_ACT_NONE when it fails) */
  uint8_t next_words;       /*  86  out: words that draw consumed */
  uint8_t next_status;      /*  87  out: its SHIPENV_ST_* */
  uint8_t pad5[40];
} shipenv_host_block;       /* 128 bytes */
int shipenv_serve_step(shipenv_ctx *ctx, shipenv_host_block *block, uint64_t seed, uint64_t env_id, int max_steps,
                       uint32_t flags, uint32_t word0);
/* Makes the server kernel exit now (it is launched again by the next shipenv_serve_step) and waits for it. */
int shipenv_serve_stop(shipenv_ctx *ctx);
/* Idle time after which the server kernel retires, in microseconds (applies from its next launch). */
int shipenv_serve_set_idle_us(shipenv_ctx *ctx, uint32_t idle_us);

#ifdef __cplusplus
}
#endif
#endif /* SHIPENV_H_ */
