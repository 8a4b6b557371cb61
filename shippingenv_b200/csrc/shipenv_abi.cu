// shipenv_abi.cu -- the C ABI of include/shipenv.h: context, world upload, kernel launches.
// Host logic only; all arithmetic of the hot path is in kernels.cuh / transition.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "kernels.cuh"

using namespace shipenv;

static_assert(sizeof(shipenv_host_block) == 128 && offsetof(shipenv_host_block, clock) == 32 &&
              offsetof(shipenv_host_block, action) == 48 && offsetof(shipenv_host_block, done_bits) == 64 &&
              offsetof(shipenv_host_block, status) == 76 && offsetof(shipenv_host_block, next_action) == 84,
              "shipenv_host_block is the facade's page-locked block (shippingenv_b200/environment.py)");

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) return fail(SHIPENV_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// Restores the caller's current device on scope exit (torch tracks it through the runtime).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

const shipenv_consts kDefaultConsts = {
    /*cargo0*/ 0, /*capacity*/ 50, /*fuel0*/ 200.0,
    /*r_deliver*/ 2.0, /*r_arrive*/ 10.0, /*r_closer*/ 2.0, /*r_take_fuel*/ 0.05, /*r_take_cargo*/ 0.05,
    /*p_no_fuel*/ -10.0, /*p_ground*/ -5.0, /*p_water*/ -1.0, /*p_loss*/ -3.0, /*p_farther*/ -2.0,
    /*p_use_fuel*/ -0.0001};

}  // namespace

constexpr int kHostStreams = 4;                 // chunks of one shipenv_step_host call in flight
// Tile-ticket counters of step_fast_kernel launches that carry no chain array: a ring, one slot per launch in
// issue order, so that launches of one context that are in flight together (other streams, the chunks of
// shipenv_step_host) never share a counter unless more than kTicketSlots of them overlap.
constexpr int kTicketSlots = 64;
constexpr int kTicketStride = 32;                // uint32 words between counters

struct shipenv_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  // host mirror of the world
  int W = 0, H = 0, P = -1;
  std::vector<uint8_t> passable;
  uint8_t px[kMaxPorts], py[kMaxPorts], pfuel[kMaxPorts], pcargo[kMaxPorts];
  shipenv_consts consts = kDefaultConsts;
  // device blob
  void* d_world = nullptr;
  uint32_t* d_tile_counter = nullptr;   // ticket counters of step_fast_kernel's tile scheduler (self-resetting), a
                                        // ring of kTicketSlots, 128 bytes apart
  std::atomic<uint32_t> ticket_next{0};
  // chained launches: the chain array, stream and capture status of the last launch of this context when that
  // launch took the chain protocol (null otherwise) -- SHIPENV_F_CHAINED is honoured only right after such a launch
  uint32_t* last_chain = nullptr;
  cudaStream_t last_chain_stream = nullptr;
  bool last_chain_capturing = false;
  int last_chain_grid = 0;              // its grid size and env count: a chained launch continues the same static split
  int64_t last_chain_n = 0;
  // the world upload: other streams wait for it until it has completed
  cudaEvent_t world_event = nullptr;
  cudaStream_t world_stream = nullptr;
  bool world_pending = false;
  // shipenv_step_host: internal streams and events, created on first use
  cudaStream_t host_stream[kHostStreams] = {};
  cudaEvent_t host_fork = nullptr, host_join[kHostStreams] = {};
  bool host_ready = false;
  std::vector<unsigned char> h_world;
  uint32_t world_bytes = 0;
  bool dirty = true;
  uint64_t world_gen = 1;        // bumped by every shipenv_set_*: the step server restarts when it has changed
  // shipenv_serve_step: the mailbox (page-locked, device-mapped), the server kernel's stream and what the host knows
  // about the kernel
  volatile uint32_t* srv_mail = nullptr;
  uint32_t* srv_mail_dev = nullptr;
  cudaStream_t srv_stream = nullptr;
  uint32_t srv_seq = 0;
  bool srv_launched = false;
  uint64_t srv_world_gen = 0;
  uint32_t srv_idle_us = 2000;
  bool no_fast_kernel = false;   // SHIPENV_NO_FAST=1: route every step through step_kernel (A/B testing)
  // cached occupancy (resident CTAs per SM) per kernel, keyed by dynamic smem size
  struct Occ { size_t smem = ~size_t(0); int blocks = 0; };
  Occ occ_fast[8], occ_step, occ_step_full, occ_step_obs, occ_step_obs_full, occ_reset, occ_observe, occ_sample, occ_step_sample, occ_rollout, occ_valid, occ_mcts[3], occ_many[6];
};

namespace {

// Build the blob from the host mirror.  The reward table is accumulated exactly as
// _move_ship does (shipping/environment.py:278-315): r = 0; (+no_fuel); (+ground | +use_fuel,
// +water); (+closer | +farther) -- one rounded fp64 add at a time.
int build_world(shipenv_ctx* c) {
  if (c->W <= 0) return fail(SHIPENV_E_STATE, "no map set (shipenv_set_map)");
  if (c->P < 0) return fail(SHIPENV_E_STATE, "no port table set (shipenv_set_ports)");
  const int wpr = (c->H + 31) / 32;
  const size_t n_words = size_t(c->W) * wpr;
  const size_t bytes = (sizeof(WorldHdr) + 2 * n_words * 4 + 15) & ~size_t(15);
  c->h_world.assign(bytes, 0);
  WorldHdr* h = new (c->h_world.data()) WorldHdr();
  uint32_t* bits = reinterpret_cast<uint32_t*>(c->h_world.data() + sizeof(WorldHdr));
  const shipenv_consts& k = c->consts;
  h->W = c->W; h->H = c->H; h->P = c->P; h->wpr = wpr;
  h->cargo0 = k.cargo0; h->capacity = k.capacity;
  h->n_bit_words = int32_t(n_words); h->total_bytes = int32_t(bytes);
  h->fuel0 = k.fuel0;
  h->r_take_fuel = k.r_take_fuel; h->r_take_cargo = k.r_take_cargo;
  h->r_deliver = k.r_deliver; h->r_arrive = k.r_arrive; h->p_loss = k.p_loss;
  for (int idx = 0; idx < 8; ++idx) {
    volatile double r = 0.0;  // volatile: one rounded add per statement, no contraction
    if (idx & 4) r = r + k.p_no_fuel;
    if (idx & 2) { r = r + k.p_ground; }
    else { r = r + k.p_use_fuel; r = r + k.p_water; }
    if (idx & 1) r = r + k.r_closer; else r = r + k.p_farther;
    h->move_reward[idx] = r;
    h->move_reward_f32[idx] = float(r);
  }
  for (int p = 0; p < c->P; ++p) {
    h->port[p] = uint32_t(c->px[p]) | (uint32_t(c->py[p]) << 8) | (uint32_t(c->pfuel[p]) << 16) |
                 (uint32_t(c->pcargo[p]) << 24);
    h->obs_tail[4 * p + 0] = float(c->px[p]); h->obs_tail[4 * p + 1] = float(c->py[p]);
    h->obs_tail[4 * p + 2] = float(c->pfuel[p]); h->obs_tail[4 * p + 3] = float(c->pcargo[p]);
  }
  for (int cg = 0; cg < k.capacity && cg < kMaxCap; ++cg) {
    volatile double q = double(cg) / double(k.capacity);  // normalize(), shipping/util.py:6-8
    h->loss_thr[cg] = uint32_t(q * 4294967296.0);          // exact scaling, then floor
  }
  h->loss_thr[k.capacity] = 0xFFFFFFFFu;
  for (int x = 0; x < c->W; ++x)
    for (int y = 0; y < c->H; ++y)
      if (c->passable[size_t(x) * c->H + y]) bits[size_t(x) * wpr + (y >> 5)] |= 1u << (y & 31);
  // a port cell is passable whatever the map says (add_port stamps Entity.PORT, :65)
  for (int p = 0; p < c->P; ++p)
    if (c->px[p] < c->W && c->py[p] < c->H) {
      const size_t word = size_t(c->px[p]) * wpr + (c->py[p] >> 5);
      bits[word] |= 1u << (c->py[p] & 31);
      bits[n_words + word] |= 1u << (c->py[p] & 31);     // second bitmap: port cells
    }
  c->world_bytes = uint32_t(bytes);
  return 0;
}

// Called by every launch.  `chain` = the chain array when the launch takes the chain protocol, else null (any
// other launch of the context ends a chain: the next SHIPENV_F_CHAINED request is not honoured).
int sync_world(shipenv_ctx* c, cudaStream_t stream, uint32_t* chain = nullptr) {
  c->last_chain = chain;
  c->last_chain_stream = stream;
  if (c->dirty) {
    if (int rc = build_world(c)) return rc;
    // pageable source: the runtime stages the bytes before returning, so h_world may change later
    CUDA_TRY(cudaMemcpyAsync(c->d_world, c->h_world.data(), c->world_bytes, cudaMemcpyHostToDevice, stream));
    // launches on other streams must not read the blob before the upload has landed
    if (!c->world_event) CUDA_TRY(cudaEventCreateWithFlags(&c->world_event, cudaEventDisableTiming));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    c->world_pending = false;
    if (cap == cudaStreamCaptureStatusNone) {
      CUDA_TRY(cudaEventRecord(c->world_event, stream));
      c->world_stream = stream;
      c->world_pending = true;
    }
    c->dirty = false;
  } else if (c->world_pending && stream != c->world_stream) {
    // (not while `stream` is being captured: a capture may not wait on outside work -- and it begins with the device
    // idle, so the upload has landed)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      if (cudaEventQuery(c->world_event) == cudaSuccess) c->world_pending = false;
      else CUDA_TRY(cudaStreamWaitEvent(stream, c->world_event, 0));
    }
  }
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (device, function), not to a context: it only ever rises,
// whatever order contexts with different worlds launch in.
std::mutex g_optin_mutex;
std::map<std::pair<int, const void*>, size_t> g_optin;

template <class Kernel>
int raise_smem_optin(shipenv_ctx* c, Kernel kernel, size_t smem) {
  if (smem <= 48 * 1024) return 0;
  if (smem > size_t(c->max_smem_optin)) return fail(SHIPENV_E_ARG, "needs %zu B of shared memory per CTA", smem);
  std::lock_guard<std::mutex> lock(g_optin_mutex);
  size_t& have = g_optin[{c->device, reinterpret_cast<const void*>(kernel)}];
  if (smem > have) {
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    have = smem;
  }
  return 0;
}

template <class Kernel>
int grid_for(shipenv_ctx* c, Kernel kernel, shipenv_ctx::Occ& occ, int threads, size_t smem, int64_t tiles, int* grid) {
  if (occ.smem != smem) {
    if (int rc = raise_smem_optin(c, kernel, smem)) return rc;
    int blocks = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, threads, smem));
    if (blocks < 1) return fail(SHIPENV_E_CUDA, "kernel does not fit an SM");
    occ.smem = smem;
    occ.blocks = blocks;
  }
  const int64_t resident = int64_t(c->sm_count) * occ.blocks;  // persistent: one wave of CTAs
  *grid = int(tiles < resident ? tiles : resident);
  if (*grid < 1) *grid = 1;
  return 0;
}

// (k0 + r*W0, k1 + r*W1), r = 0..9, for key = seed with k1 ^= tweak
void fill_round_keys(uint32_t* rk, uint64_t seed, uint32_t tweak) {
  for (int r = 0; r < 10; ++r) {
    rk[2 * r] = uint32_t(seed) + uint32_t(r) * kPhiloxW0;
    rk[2 * r + 1] = (uint32_t(seed >> 32) ^ tweak) + uint32_t(r) * kPhiloxW1;
  }
}

int check_ctx(shipenv_ctx* c) {
  if (!c) return fail(SHIPENV_E_ARG, "null context");
  return 0;
}

}  // namespace

extern "C" {

int shipenv_abi_version(void) { return SHIPENV_ABI_VERSION; }
const char* shipenv_last_error(void) { return g_err; }
uint64_t shipenv_launch_count(void) { return g_launches.load(); }

void shipenv_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

void shipenv_port_stock(uint64_t seed, uint32_t port_index, uint8_t* fuel, uint8_t* cargo) {
  uint32_t w[4];
  philox4x32_10(0u, port_index, 0xFFFFFFFFu, 0xFFFFFFFFu, uint32_t(seed), uint32_t(seed >> 32) ^ kPortTweak, w);
  // random.randint(5, 20) twice: fuel then cargo (shipping/environment.py:63-64)
  *fuel = uint8_t(5 + ((uint64_t(w[0]) * 16u) >> 32));
  *cargo = uint8_t(5 + ((uint64_t(w[1]) * 16u) >> 32));
}

int shipenv_host_device_pointer(void* host_ptr, void** device_ptr) {
  if (!host_ptr || !device_ptr) return fail(SHIPENV_E_ARG, "null pointer");
  *device_ptr = nullptr;
  CUDA_TRY(cudaHostGetDevicePointer(device_ptr, host_ptr, 0));
  return 0;
}

int shipenv_stream_synchronize(shipenv_ctx* c, void* stream) {
  if (int rc = check_ctx(c)) return rc;
  DeviceGuard guard(c->device);
  CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return 0;
}

int shipenv_ctx_create(int device, shipenv_ctx** out) {
  if (!out) return fail(SHIPENV_E_ARG, "out is null");
  *out = nullptr;
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(SHIPENV_E_ARG, "device %d out of range (%d visible)", device, count);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SHIPENV_E_CUDA, "cannot select device %d", device);
  int major = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(SHIPENV_E_CUDA, "device %d is sm_%dx; this library is built for sm_100a only", device, major);
  shipenv_ctx* c = new (std::nothrow) shipenv_ctx();
  if (!c) return fail(SHIPENV_E_ARG, "out of host memory");
  c->device = device;
  { const char* e = getenv("SHIPENV_NO_FAST"); c->no_fast_kernel = e && e[0] == '1'; }
  CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  cudaError_t e = cudaMalloc(&c->d_world, kMaxWorldBytes);
  if (e != cudaSuccess) { delete c; return fail(SHIPENV_E_CUDA, "cudaMalloc(world): %s", cudaGetErrorString(e)); }
  e = cudaMalloc(&c->d_tile_counter, kTicketSlots * kTicketStride * 4);
  if (e == cudaSuccess) e = cudaMemset(c->d_tile_counter, 0, kTicketSlots * kTicketStride * 4);
  if (e != cudaSuccess) {
    cudaFree(c->d_world); if (c->d_tile_counter) cudaFree(c->d_tile_counter);
    delete c;
    return fail(SHIPENV_E_CUDA, "tile counter: %s", cudaGetErrorString(e));
  }
  *out = c;
  return 0;
}

int shipenv_ctx_destroy(shipenv_ctx* c) {
  if (!c) return 0;
  shipenv_serve_stop(c);
  {
    DeviceGuard guard(c->device);
    if (c->srv_stream) cudaStreamDestroy(c->srv_stream);
    if (c->srv_mail) cudaFreeHost(const_cast<uint32_t*>(c->srv_mail));
    if (c->d_world) cudaFree(c->d_world);
    if (c->d_tile_counter) cudaFree(c->d_tile_counter);
    if (c->world_event) cudaEventDestroy(c->world_event);
    if (c->host_ready) {
      for (int k = 0; k < kHostStreams; ++k) { cudaStreamDestroy(c->host_stream[k]); cudaEventDestroy(c->host_join[k]); }
      cudaEventDestroy(c->host_fork);
    }
  }
  delete c;
  return 0;
}

int shipenv_set_map(shipenv_ctx* c, const uint8_t* passable, int W, int H) {
  if (int rc = check_ctx(c)) return rc;
  if (!passable) return fail(SHIPENV_E_ARG, "passable is null");
  if (W < 1 || H < 1 || W > kMaxDim || H > kMaxDim) return fail(SHIPENV_E_ARG, "map %dx%d outside 1..%d", W, H, kMaxDim);
  c->W = W; c->H = H;
  c->passable.assign(passable, passable + size_t(W) * H);
  c->dirty = true;
  ++c->world_gen;
  return 0;
}

int shipenv_set_ports(shipenv_ctx* c, const uint8_t* px, const uint8_t* py, const uint8_t* pfuel,
                      const uint8_t* pcargo, int P) {
  if (int rc = check_ctx(c)) return rc;
  if (P < 0 || P > kMaxPorts) return fail(SHIPENV_E_ARG, "P=%d outside 0..%d", P, kMaxPorts);
  if (P > 0 && (!px || !py || !pfuel || !pcargo)) return fail(SHIPENV_E_ARG, "null port table");
  for (int p = 0; p < P; ++p) { c->px[p] = px[p]; c->py[p] = py[p]; c->pfuel[p] = pfuel[p]; c->pcargo[p] = pcargo[p]; }
  c->P = P;
  c->dirty = true;
  ++c->world_gen;
  return 0;
}

int shipenv_set_consts(shipenv_ctx* c, const shipenv_consts* k) {
  if (int rc = check_ctx(c)) return rc;
  const shipenv_consts v = k ? *k : kDefaultConsts;
  if (v.capacity < 1 || v.capacity > SHIPENV_MAX_CAPACITY) return fail(SHIPENV_E_ARG, "capacity %d outside 1..%d", v.capacity, SHIPENV_MAX_CAPACITY);
  c->consts = v;
  c->dirty = true;
  ++c->world_gen;
  return 0;
}

int shipenv_get_consts(shipenv_ctx* c, shipenv_consts* out) {
  if (int rc = check_ctx(c)) return rc;
  if (!out) return fail(SHIPENV_E_ARG, "out is null");
  *out = c->consts;
  return 0;
}

int shipenv_obs_width(shipenv_ctx* c) { return (c && c->P >= 0) ? 6 + 4 * c->P : SHIPENV_E_STATE; }
int shipenv_num_actions(shipenv_ctx* c) {
  return (c && c->P >= 0) ? 4 + c->P + SHIPENV_N_CARGO_ACTIONS + SHIPENV_N_FUEL_ACTIONS : SHIPENV_E_STATE;
}

int shipenv_reset(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, uint64_t seed, uint64_t env_id_base,
                  int64_t n, const uint32_t* mask_bits, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !clock))) return fail(SHIPENV_E_ARG, "reset: null buffer or negative n");
  if (n == 0) return 0;
  if (c->P < 2) return fail(SHIPENV_E_STATE, "reset needs at least 2 ports (the reference loops forever with 1, raises with 0)");
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  ResetArgs a{c->d_world, c->world_bytes, reinterpret_cast<uint4*>(ships), clock, mask_bits, seed, env_id_base, n, {}};
  fill_round_keys(a.round_key, seed, 0u);
  const size_t smem = c->world_bytes;
  int grid = 0;
  if (int rc = grid_for(c, reset_kernel, c->occ_reset, kThreads, smem, (n + kThreads - 1) / kThreads, &grid)) return rc;
  reset_kernel<<<grid, kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// shipenv_step proper.
static int step_impl(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, const int16_t* actions,
                     const shipenv_step_io* io, uint64_t seed, uint64_t env_id_base, int64_t n, int max_steps,
                     uint32_t flags, uint32_t word0, cudaStream_t stream) {
  if (int rc = check_ctx(c)) return rc;
  const bool all_policy = flags & SHIPENV_F_ALL_POLICY;
  if (n < 0 || (n > 0 && (!ships || !clock || (!actions && !all_policy)))) return fail(SHIPENV_E_ARG, "step: null buffer or negative n");
  if (max_steps < 0 || max_steps > 65535) return fail(SHIPENV_E_ARG, "max_steps %d outside 0..65535", max_steps);
  if (n == 0) return 0;
  if (n >= (int64_t(1) << 31)) return fail(SHIPENV_E_ARG, "step: at most 2^31-1 envs per launch");
  if ((reinterpret_cast<uintptr_t>(ships) | reinterpret_cast<uintptr_t>(clock) | reinterpret_cast<uintptr_t>(actions)) & 15u)
    return fail(SHIPENV_E_ARG, "step: ships, clock and actions must be 16-byte aligned (TMA bulk copies)");
  if ((flags & SHIPENV_F_AUTO_RESET) && c->P < 2) return fail(SHIPENV_E_STATE, "auto-reset needs at least 2 ports");
  DeviceGuard guard(c->device);
  // what the previous launch of this context was, before sync_world() forgets it
  uint32_t* const prev_chain = c->last_chain;
  const cudaStream_t prev_chain_stream = c->last_chain_stream;
  const bool prev_chain_capturing = c->last_chain_capturing;
  if (int rc = sync_world(c, stream)) return rc;
  StepArgs a{};
  a.world = c->d_world; a.world_bytes = c->world_bytes;
  a.ships = reinterpret_cast<uint4*>(ships); a.clock = clock; a.actions = actions;
  if (io) a.io = *io; else memset(&a.io, 0, sizeof(a.io));
  a.seed = seed; a.env_id_base = env_id_base; a.n = n; a.max_steps = max_steps; a.flags = flags; a.word0 = word0;
  a.obs_width = 6 + 4 * c->P;
  fill_round_keys(a.round_key, seed, 0u);
  const int64_t tiles = (n + kThreads - 1) / kThreads;
  const bool lean = !a.io.reward_f64 && !a.io.status && !a.io.taken && !a.io.ep_return;
  int grid = 0;
  if (a.io.obs && (reinterpret_cast<uintptr_t>(a.io.obs) & 15u)) return fail(SHIPENV_E_ARG, "obs must be 16-byte aligned");
  // throughput kernels: step_fast_kernel<policy, obs, envs per thread, lean>
  constexpr int kEpt = SHIPENV_FAST_EPT;
  const int ept = a.io.obs ? 1 : kEpt;
  size_t smem = ((size_t(c->world_bytes) + 127) & ~size_t(127)) + (a.io.obs ? fast_smem_bytes<1>() : fast_smem_bytes<kEpt>());
  if (a.io.obs) smem += 2 * size_t(kThreads) * a.obs_width * 4;
  // with many ports two observation tiles of [256, 6+4P] floats do not fit: step_kernel then keeps one
  const bool fast_ok = a.io.done_bits && a.io.reset_bits && a.io.error_bits && (a.io.reward_f32 || !lean) && word0 == 0 &&
                       c->P >= 2 && !c->no_fast_kernel && smem <= size_t(c->max_smem_optin);
  if (fast_ok) {
    void (*kernel)(const StepArgs) = nullptr;
    const int variant = (all_policy ? 4 : 0) | (a.io.obs ? 2 : 0) | (lean ? 1 : 0);
    switch (variant) {
      case 0: kernel = step_fast_kernel<false, false, kEpt, false>; break;
      case 1: kernel = step_fast_kernel<false, false, kEpt, true>; break;
      case 2: kernel = step_fast_kernel<false, true, 1, false>; break;
      case 3: kernel = step_fast_kernel<false, true, 1, true>; break;
      case 4: kernel = step_fast_kernel<true, false, kEpt, false>; break;
      case 5: kernel = step_fast_kernel<true, false, kEpt, true>; break;
      case 6: kernel = step_fast_kernel<true, true, 1, false>; break;
      default: kernel = step_fast_kernel<true, true, 1, true>; break;
    }
    shipenv_ctx::Occ* occ = &c->occ_fast[variant];
    const int64_t fast_tiles = (n + int64_t(kThreads) * ept - 1) / (int64_t(kThreads) * ept);
    if (int rc = grid_for(c, kernel, *occ, kFastBlock, smem, fast_tiles, &grid)) return rc;
    a.tile_counter = c->d_tile_counter + (c->ticket_next.fetch_add(1) % kTicketSlots) * kTicketStride;
    a.chain = a.io.chain;
    if (a.chain) {
      if (reinterpret_cast<uintptr_t>(a.chain) & 7u) return fail(SHIPENV_E_ARG, "step: chain must be 8-byte aligned");
      // SHIPENV_F_CHAINED is honoured only directly behind a chain-protocol launch on the same array and stream with
      // the same grid size and env count (CTA b continues CTA b's static range), and not across the end of a stream
      // capture; otherwise this launch waits for everything before it
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cap);
      const bool capturing = cap != cudaStreamCaptureStatusNone;
      if (grid > kChainMaxCtas) return fail(SHIPENV_E_ARG, "step: chain arrays serve grids of up to %d CTAs", kChainMaxCtas);
      if (!(prev_chain == a.chain && prev_chain_stream == stream && prev_chain_capturing == capturing &&
            c->last_chain_grid == grid && c->last_chain_n == n))
        a.flags &= ~SHIPENV_F_CHAINED;
      c->last_chain = a.chain;
      c->last_chain_capturing = capturing;
      c->last_chain_grid = grid;
      c->last_chain_n = n;
    } else {
      a.flags &= ~SHIPENV_F_CHAINED;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFastBlock); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this grid's prologue
    attr[0].val.programmaticStreamSerializationAllowed = 1;            // with the previous grid's tail
    cfg.attrs = attr; cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, a));
  } else if (a.io.obs) {
    if (!actions) return fail(SHIPENV_E_ARG, "step: this output set needs the `actions` array even with SHIPENV_F_ALL_POLICY");
    const size_t base = ((size_t(c->world_bytes) + 127) & ~size_t(127)) + ((kPipeBytes + 127) & ~size_t(127));
    const size_t tile_bytes = size_t(kThreads) * a.obs_width * 4;
    a.obs_bufs = (base + 2 * tile_bytes <= 96 * 1024) ? 2 : 1;
    const size_t smem = base + a.obs_bufs * tile_bytes;
    if (lean) {
      if (int rc = grid_for(c, step_kernel<true, true>, c->occ_step_obs, kStepBlock, smem, tiles, &grid)) return rc;
      step_kernel<true, true><<<grid, kStepBlock, smem, stream>>>(a);
    } else {
      if (int rc = grid_for(c, step_kernel<true, false>, c->occ_step_obs_full, kStepBlock, smem, tiles, &grid)) return rc;
      step_kernel<true, false><<<grid, kStepBlock, smem, stream>>>(a);
    }
  } else {
    if (!actions) return fail(SHIPENV_E_ARG, "step: this output set needs the `actions` array even with SHIPENV_F_ALL_POLICY");
    const size_t smem = ((size_t(c->world_bytes) + 127) & ~size_t(127)) + kPipeBytes;
    if (lean) {
      if (int rc = grid_for(c, step_kernel<false, true>, c->occ_step, kStepBlock, smem, tiles, &grid)) return rc;
      step_kernel<false, true><<<grid, kStepBlock, smem, stream>>>(a);
    } else {
      if (int rc = grid_for(c, step_kernel<false, false>, c->occ_step_full, kStepBlock, smem, tiles, &grid)) return rc;
      step_kernel<false, false><<<grid, kStepBlock, smem, stream>>>(a);
    }
  }
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_step(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, const int16_t* actions,
                 const shipenv_step_io* io, uint64_t seed, uint64_t env_id_base, int64_t n, int max_steps,
                 uint32_t flags, uint32_t word0, void* stream_) {
  return step_impl(c, ships, clock, actions, io, seed, env_id_base, n, max_steps, flags, word0,
                   static_cast<cudaStream_t>(stream_));
}

#ifdef SHIPENV_TRACE
// tuning builds only: per-CTA timeline records of step_fast_kernel go to `buf` (8 x u64 each); null switches it off
int shipenv_debug_trace(void* buf) {
  unsigned int zero = 0;
  if (cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf)) != cudaSuccess) return SHIPENV_E_CUDA;
  if (cudaMemcpyToSymbol(g_trace_count, &zero, sizeof(zero)) != cudaSuccess) return SHIPENV_E_CUDA;
  return 0;
}
#endif

size_t shipenv_chain_words(int64_t n) {
  if (n <= 0) return 0;
  return size_t(kChainDone) + size_t(kChainMaxCtas);
}

int shipenv_step_many(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, const int16_t* actions, int T,
                      float* reward, uint32_t* done_bits, uint32_t* reset_bits, uint32_t* error_bits, uint8_t* status,
                      int16_t* taken, double* reward_f64, uint64_t seed, uint64_t env_id_base, int64_t n, int max_steps,
                      uint32_t flags, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  const bool all_policy = flags & SHIPENV_F_ALL_POLICY;
  if (n < 0 || T < 0 || (n > 0 && T > 0 && (!ships || !clock || (!actions && !all_policy))))
    return fail(SHIPENV_E_ARG, "step_many: null buffer or negative size");
  if (max_steps < 0 || max_steps > 65535) return fail(SHIPENV_E_ARG, "max_steps %d outside 0..65535", max_steps);
  if (n == 0 || T == 0) return 0;
  if (n >= (int64_t(1) << 31)) return fail(SHIPENV_E_ARG, "step_many: at most 2^31-1 envs per launch");
  if (reinterpret_cast<uintptr_t>(ships) & 15u) return fail(SHIPENV_E_ARG, "step_many: ships must be 16-byte aligned");
  if ((flags & SHIPENV_F_AUTO_RESET) && c->P < 2) return fail(SHIPENV_E_STATE, "auto-reset needs at least 2 ports");
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  StepManyArgs a{};
  a.world = c->d_world; a.world_bytes = c->world_bytes;
  a.ships = reinterpret_cast<uint4*>(ships); a.clock = clock; a.actions = actions;
  a.reward = reward; a.done_bits = done_bits; a.reset_bits = reset_bits; a.error_bits = error_bits;
  a.status = status; a.taken = taken; a.reward_f64 = reward_f64;
  a.seed = seed; a.env_id_base = env_id_base; a.n = n; a.T = T; a.max_steps = max_steps; a.flags = flags;
  fill_round_keys(a.round_key, seed, 0u);
  // one warp per CTA and one env per thread for small batches (4,096 envs then occupy 128 SMs), wider CTAs with two
  // interleaved envs per thread for large ones
  const int threads = n <= (1 << 15) ? 32 : n <= (1 << 18) ? 64 : 128;
  // (the random policy sends a twentieth of its steps the general way -- off-map re-draws, port pick-ups --, which
  // a second env per thread would only serialise further)
  const int ept = (threads == 128 && !all_policy) ? 2 : 1;
  const size_t smem = c->world_bytes;
  void (*kernel)(const StepManyArgs) = ept == 2 ? step_many_kernel<false, 2>
                                       : all_policy ? step_many_kernel<true, 1> : step_many_kernel<false, 1>;
  shipenv_ctx::Occ& occ = c->occ_many[(all_policy ? 3 : 0) + (threads == 32 ? 0 : threads == 64 ? 1 : 2)];
  int grid = 0;
  if (int rc = grid_for(c, kernel, occ, threads, smem, (n + int64_t(threads) * ept - 1) / (int64_t(threads) * ept), &grid)) return rc;
  kernel<<<grid, threads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

size_t shipenv_step_host_workspace(int64_t n) {
  if (n <= 0) return 0;
  const size_t nw = (size_t(n) + 31) / 32;
  // [actions: 2 B per env, padded to 256 B][rewards: 4 B per env][done | reset | error words]: the results are one
  // contiguous block, so that a caller whose host buffers are laid out the same way gets them with one copy
  return ((size_t(n) * 2 + 255) & ~size_t(255)) + size_t(n) * 4 + 3 * nw * 4;
}

int shipenv_step_host(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, const int16_t* h_actions,
                      float* h_reward, uint32_t* h_done_bits, uint32_t* h_reset_bits, uint32_t* h_error_bits,
                      void* workspace, size_t workspace_bytes, uint64_t seed, uint64_t env_id_base, int64_t n,
                      int max_steps, uint32_t flags, int chunks_wanted, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !clock || !h_actions || !h_reward || !workspace)))
    return fail(SHIPENV_E_ARG, "step_host: null buffer or negative n");
  if (n == 0) return 0;
  if (workspace_bytes < shipenv_step_host_workspace(n) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return fail(SHIPENV_E_ARG, "step_host: workspace needs %zu bytes, 256-byte aligned", shipenv_step_host_workspace(n));
  if (flags & SHIPENV_F_ALL_POLICY) return fail(SHIPENV_E_ARG, "step_host: actions come from the host; SHIPENV_F_ALL_POLICY makes no sense here");
  DeviceGuard guard(c->device);
  cudaStream_t user = static_cast<cudaStream_t>(stream_);
  if (!c->host_ready) {
    CUDA_TRY(cudaEventCreateWithFlags(&c->host_fork, cudaEventDisableTiming));
    for (int k = 0; k < kHostStreams; ++k) {
      CUDA_TRY(cudaStreamCreateWithFlags(&c->host_stream[k], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&c->host_join[k], cudaEventDisableTiming));
    }
    c->host_ready = true;
  }
  if (int rc = sync_world(c, user)) return rc;     // before the fork: every chunk sees the uploaded world
  const size_t nw = (size_t(n) + 31) / 32;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int16_t* d_act = reinterpret_cast<int16_t*>(ws);
  float* d_reward = reinterpret_cast<float*>(ws + ((size_t(n) * 2 + 255) & ~size_t(255)));
  uint32_t* d_mask[3];
  d_mask[0] = reinterpret_cast<uint32_t*>(d_reward + n);
  d_mask[1] = d_mask[0] + nw;
  d_mask[2] = d_mask[1] + nw;
  // Chunks of whole 512-env tiles (so mask words, TMA alignment and tile boundaries all fall on chunk boundaries), at
  // least 1 Mi envs each (a copy of less than ~2 MB does not reach PCIe speed).  Chunk 0 runs on the caller's stream,
  // the others on internal streams forked from it and joined back: a one-chunk call involves no other stream.
  constexpr int64_t kGrain = 512, kMinChunk = 1 << 20;
  if (chunks_wanted < 0 || chunks_wanted > kHostStreams) return fail(SHIPENV_E_ARG, "step_host: chunks outside 0..%d", kHostStreams);
  int chunks = chunks_wanted > 0 ? int(std::min<int64_t>(chunks_wanted, (n + kGrain - 1) / kGrain))
                                 : int(std::min<int64_t>(kHostStreams, std::max<int64_t>(1, n / kMinChunk)));
  const int64_t per = ((n + chunks - 1) / chunks + kGrain - 1) / kGrain * kGrain;
  chunks = int((n + per - 1) / per);
  if (chunks > 1) CUDA_TRY(cudaEventRecord(c->host_fork, user));
  // One chunk: enqueue its copy-in, kernel and (with several chunks) its reward copy-out on stream s.
  auto submit = [&](int k, cudaStream_t s) -> int {
    const int64_t lo = int64_t(k) * per, len = std::min<int64_t>(per, n - lo);
    CUDA_TRY(cudaMemcpyAsync(d_act + lo, h_actions + lo, size_t(len) * 2, cudaMemcpyHostToDevice, s));
    shipenv_step_io io{};
    io.reward_f32 = d_reward + lo;
    io.done_bits = d_mask[0] + lo / 32; io.reset_bits = d_mask[1] + lo / 32; io.error_bits = d_mask[2] + lo / 32;
    if (int rc = step_impl(c, ships + lo, clock + lo, d_act + lo, &io, seed, env_id_base + uint64_t(lo), len, max_steps,
                           flags & ~SHIPENV_F_CHAINED, 0u, s))
      return rc;
    if (chunks > 1) CUDA_TRY(cudaMemcpyAsync(h_reward + lo, d_reward + lo, size_t(len) * 4, cudaMemcpyDeviceToHost, s));
    return 0;
  };
  int rc = 0, forked = 0;                            // internal streams first, the caller's stream last
  for (int k = chunks - 1; k >= 1 && rc == 0; --k) {
    cudaStream_t s = c->host_stream[k];
    if (cudaStreamWaitEvent(s, c->host_fork, 0) != cudaSuccess) { rc = fail(SHIPENV_E_CUDA, "step_host: fork failed"); break; }
    forked |= 1 << k;
    rc = submit(k, s);
  }
  if (rc == 0) rc = submit(0, user);
  // join every internal stream that was given work, also on the error path: the caller's stream stays ordered after
  // all of it and the fork / join events are free for the next call
  for (int k = 1; k < chunks; ++k)
    if (forked & (1 << k)) {
      if (cudaEventRecord(c->host_join[k], c->host_stream[k]) != cudaSuccess ||
          cudaStreamWaitEvent(user, c->host_join[k], 0) != cudaSuccess) {
        if (rc == 0) rc = fail(SHIPENV_E_CUDA, "step_host: join failed");
      }
    }
  if (rc) return rc;
  // Masks: one copy per run of host arrays that are laid out back to back like the device block (each copy costs
  // ~10 us of the copy engine whatever its size).  With one chunk the reward copy is still pending and joins the run.
  struct Span { unsigned char* dst; const unsigned char* src; size_t bytes; };
  Span spans[4] = {};
  int n_spans = 0;
  auto add = [&](void* dst, const void* src, size_t bytes) {
    if (!dst) return;
    Span sp{static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), bytes};
    if (n_spans > 0 && spans[n_spans - 1].dst + spans[n_spans - 1].bytes == sp.dst &&
        spans[n_spans - 1].src + spans[n_spans - 1].bytes == sp.src)
      spans[n_spans - 1].bytes += bytes;
    else
      spans[n_spans++] = sp;
  };
  if (chunks == 1) add(h_reward, d_reward, size_t(n) * 4);
  add(h_done_bits, d_mask[0], nw * 4);
  add(h_reset_bits, d_mask[1], nw * 4);
  add(h_error_bits, d_mask[2], nw * 4);
  for (int k = 0; k < n_spans; ++k)
    CUDA_TRY(cudaMemcpyAsync(spans[k].dst, spans[k].src, spans[k].bytes, cudaMemcpyDeviceToHost, user));
  return 0;
}

int shipenv_step_sample(shipenv_ctx* c, shipenv_ship* ships, uint32_t* clock, const int16_t* actions,
                        const shipenv_step_io* io, int16_t* next_action, uint8_t* next_words, uint8_t* next_status,
                        uint64_t seed, uint64_t env_id_base, int64_t n, int max_steps, uint32_t flags, uint32_t word0,
                        void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  const bool all_policy = flags & SHIPENV_F_ALL_POLICY;
  if (n < 0 || (n > 0 && (!ships || !clock || !next_action || (!actions && !all_policy))))
    return fail(SHIPENV_E_ARG, "step_sample: null buffer or negative n");
  if (max_steps < 0 || max_steps > 65535) return fail(SHIPENV_E_ARG, "max_steps %d outside 0..65535", max_steps);
  if (io && io->obs) return fail(SHIPENV_E_ARG, "step_sample: observation rows are not produced here, use shipenv_step");
  if (n == 0) return 0;
  if (n >= (int64_t(1) << 31)) return fail(SHIPENV_E_ARG, "step_sample: at most 2^31-1 envs per launch");
  if (reinterpret_cast<uintptr_t>(ships) & 15u) return fail(SHIPENV_E_ARG, "step_sample: ships must be 16-byte aligned");
  if ((flags & SHIPENV_F_AUTO_RESET) && c->P < 2) return fail(SHIPENV_E_STATE, "auto-reset needs at least 2 ports");
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  StepSampleArgs a{};
  a.s.world = c->d_world; a.s.world_bytes = c->world_bytes;
  a.s.ships = reinterpret_cast<uint4*>(ships); a.s.clock = clock; a.s.actions = actions;
  if (io) a.s.io = *io;
  a.s.seed = seed; a.s.env_id_base = env_id_base; a.s.n = n; a.s.max_steps = max_steps; a.s.flags = flags; a.s.word0 = word0;
  a.s.obs_width = 6 + 4 * c->P;
  fill_round_keys(a.s.round_key, seed, 0u);
  a.next_action = next_action; a.next_words = next_words; a.next_status = next_status;
  const size_t smem = c->world_bytes;
  int grid = 0;
  if (int rc = grid_for(c, step_sample_kernel, c->occ_step_sample, kThreads, smem, (n + kThreads - 1) / kThreads, &grid)) return rc;
  step_sample_kernel<<<grid, kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

// ---- the step server (aux_kernels.cuh: step_server_kernel) ----
static int serve_launch(shipenv_ctx* c) {
  ServerArgs a{c->srv_mail_dev, c->d_world, c->world_bytes, c->srv_seq - 1u, (unsigned long long)c->srv_idle_us * 1000ull};
  c->srv_mail[kSrvAlive] = 1u;
  step_server_kernel<<<1, 32, c->world_bytes, c->srv_stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  c->srv_launched = true;
  ++g_launches;
  return 0;
}

int shipenv_serve_stop(shipenv_ctx* c) {
  if (int rc = check_ctx(c)) return rc;
  if (!c->srv_mail || !c->srv_launched) return 0;
  DeviceGuard guard(c->device);
  if (c->srv_mail[kSrvAlive] != 0u) {               // still polling: ask it to leave
    c->srv_mail[kSrvCmd] = kSrvCmdQuit;
    std::atomic_thread_fence(std::memory_order_release);
    c->srv_mail[kSrvSeq] = ++c->srv_seq;
  }
  CUDA_TRY(cudaStreamSynchronize(c->srv_stream));   // retired (by the request or by its idle limit) and gone
  c->srv_launched = false;
  return 0;
}

int shipenv_serve_set_idle_us(shipenv_ctx* c, uint32_t idle_us) {
  if (int rc = check_ctx(c)) return rc;
  c->srv_idle_us = idle_us < 10u ? 10u : idle_us;
  return 0;
}

int shipenv_serve_step(shipenv_ctx* c, shipenv_host_block* block, uint64_t seed, uint64_t env_id, int max_steps,
                       uint32_t flags, uint32_t word0) {
  if (int rc = check_ctx(c)) return rc;
  if (!block) return fail(SHIPENV_E_ARG, "serve_step: null block");
  if (max_steps < 0 || max_steps > 65535) return fail(SHIPENV_E_ARG, "max_steps %d outside 0..65535", max_steps);
  if ((flags & SHIPENV_F_AUTO_RESET) && c->P < 2) return fail(SHIPENV_E_STATE, "auto-reset needs at least 2 ports");
  DeviceGuard guard(c->device);
  if (!c->srv_mail) {
    void* h = nullptr;
    CUDA_TRY(cudaHostAlloc(&h, 128, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h, 0, 128);
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaFreeHost(h); return fail(SHIPENV_E_CUDA, "serve_step: mailbox is not device-visible"); }
    if (cudaStreamCreateWithFlags(&c->srv_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaFreeHost(h); return fail(SHIPENV_E_CUDA, "serve_step: no stream"); }
    c->srv_mail = static_cast<volatile uint32_t*>(h);
    c->srv_mail_dev = static_cast<uint32_t*>(d);
  }
  if (c->srv_world_gen != c->world_gen || c->dirty) {
    // the world has changed: the resident kernel holds the old one in its shared memory
    if (int rc = shipenv_serve_stop(c)) return rc;
    if (int rc = sync_world(c, c->srv_stream)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->srv_stream));
    c->world_pending = false;                        // (the upload, wherever it was issued, has been waited for)
    c->srv_world_gen = c->world_gen;
  }
  volatile uint32_t* m = c->srv_mail;
  const uint32_t* ship = reinterpret_cast<const uint32_t*>(&block->ship);
  m[0] = ship[0]; m[1] = ship[1]; m[2] = ship[2]; m[3] = ship[3];
  m[4] = block->clock;
  m[5] = uint32_t(uint16_t(block->action));
  m[6] = word0; m[7] = flags; m[8] = uint32_t(max_steps);
  m[9] = uint32_t(seed); m[10] = uint32_t(seed >> 32);
  m[11] = uint32_t(env_id); m[12] = uint32_t(env_id >> 32);
  m[kSrvCmd] = kSrvCmdStep;
  std::atomic_thread_fence(std::memory_order_release);   // the request before its sequence number
  const uint32_t seq = ++c->srv_seq;
  m[kSrvSeq] = seq;
  if (!c->srv_launched || m[kSrvAlive] == 0u) {
    if (c->srv_launched) CUDA_TRY(cudaStreamSynchronize(c->srv_stream));   // the retired kernel has left
    if (int rc = serve_launch(c)) return rc;
  }
  // spin for the answer; a kernel that retires under our eyes is replaced (the request is still in the mailbox)
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t spins = 0; m[kSrvAck] != seq; ++spins) {
    if ((spins & 1023u) == 1023u) {
      if (m[kSrvAlive] == 0u && m[kSrvAck] != seq) {
        CUDA_TRY(cudaStreamSynchronize(c->srv_stream));
        if (m[kSrvAck] == seq) break;
        if (int rc = serve_launch(c)) return rc;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) {
        const cudaError_t e = cudaStreamQuery(c->srv_stream);
        return fail(SHIPENV_E_CUDA, "serve_step: no answer from the server kernel within 10 s (%s)",
                    e == cudaSuccess || e == cudaErrorNotReady ? "kernel alive or retired" : cudaGetErrorString(e));
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  uint32_t* out = reinterpret_cast<uint32_t*>(&block->ship);
  out[0] = m[16]; out[1] = m[17]; out[2] = m[18]; out[3] = m[19];
  block->clock = m[20];
  const uint32_t st = m[21];
  block->status = uint8_t(st); block->next_status = uint8_t(st >> 8); block->next_words = uint8_t(st >> 16);
  block->done_bits = (st >> 24) & 1u; block->reset_bits = (st >> 25) & 1u; block->error_bits = (st >> 26) & 1u;
  const uint64_t rbits = uint64_t(m[22]) | (uint64_t(m[23]) << 32);
  memcpy(&block->reward, &rbits, 8);
  block->next_action = int16_t(uint16_t(m[24]));
#ifdef SHIPENV_SERVER_TIMING
  memcpy(block->pad5, const_cast<const uint32_t*>(&m[25]), 4);      // tuning builds: the kernel's own time for the request, ns
#endif
  return 0;
}

int shipenv_sample_action(shipenv_ctx* c, const shipenv_ship* ships, const uint32_t* clock, int16_t* actions_out,
                          uint8_t* words_used, uint8_t* status, uint64_t seed, uint64_t env_id_base, int64_t n,
                          uint32_t word0, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !clock || !actions_out))) return fail(SHIPENV_E_ARG, "sample_action: null buffer or negative n");
  if (n == 0) return 0;
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  SampleArgs a{c->d_world, c->world_bytes, reinterpret_cast<const uint4*>(ships), clock, actions_out, words_used,
               status, seed, env_id_base, n, word0, {}};
  fill_round_keys(a.round_key, seed, 0u);
  const size_t smem = c->world_bytes;
  int grid = 0;
  if (int rc = grid_for(c, sample_kernel, c->occ_sample, kThreads, smem, (n + kThreads - 1) / kThreads, &grid)) return rc;
  sample_kernel<<<grid, kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_observe(shipenv_ctx* c, const shipenv_ship* ships, float* obs, int64_t n, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !obs))) return fail(SHIPENV_E_ARG, "observe: null buffer or negative n");
  if (reinterpret_cast<uintptr_t>(obs) & 15u) return fail(SHIPENV_E_ARG, "obs must be 16-byte aligned");
  if (n == 0) return 0;
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  ObserveArgs a{c->d_world, c->world_bytes, reinterpret_cast<const uint4*>(ships), obs, n, 6 + 4 * c->P, 2};
  const size_t tile_bytes = size_t(kThreads) * a.obs_width * 4;
  const size_t world_smem = (size_t(c->world_bytes) + 127) & ~size_t(127);
  if (world_smem + 2 * tile_bytes > size_t(c->max_smem_optin) / 2) a.bufs = 1;   // many ports: one tile, and two CTAs per SM
  const size_t smem = world_smem + size_t(a.bufs) * tile_bytes;
  int grid = 0;
  if (int rc = grid_for(c, observe_kernel, c->occ_observe, kThreads, smem, (n + kThreads - 1) / kThreads, &grid)) return rc;
  observe_kernel<<<grid, kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_valid_actions(shipenv_ctx* c, const shipenv_ship* ships, uint32_t* mask, int64_t n, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !mask))) return fail(SHIPENV_E_ARG, "valid_actions: null buffer or negative n");
  if (n == 0) return 0;
  if (c->P < 0) return fail(SHIPENV_E_STATE, "no port table set (shipenv_set_ports)");
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  const int A = 4 + c->P + SHIPENV_N_CARGO_ACTIONS + SHIPENV_N_FUEL_ACTIONS;
  ValidArgs a{c->d_world, c->world_bytes, reinterpret_cast<const uint4*>(ships), mask, n, (A + 31) / 32};
  const size_t smem = c->world_bytes;
  int grid = 0;
  if (int rc = grid_for(c, valid_actions_kernel, c->occ_valid, kThreads, smem, (n + kThreads - 1) / kThreads, &grid)) return rc;
  valid_actions_kernel<<<grid, kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_discretize(shipenv_ctx* c, const shipenv_ship* ships, int16_t* keys, int64_t n, void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (n < 0 || (n > 0 && (!ships || !keys))) return fail(SHIPENV_E_ARG, "discretize: null buffer or negative n");
  if (n == 0) return 0;
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t tiles = (n + kThreads - 1) / kThreads;
  const int grid = int(tiles < int64_t(c->sm_count) * 8 ? tiles : int64_t(c->sm_count) * 8);
  discretize_kernel<<<grid, kThreads, 0, stream>>>(reinterpret_cast<const uint4*>(ships), keys, n);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_rollout(shipenv_ctx* c, const shipenv_ship* roots, int64_t R, int S, int L, double* returns,
                    shipenv_ship* final, int32_t* steps, uint8_t* status, uint64_t seed, uint64_t root_id_base,
                    void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (R < 0 || S < 0 || L < 0) return fail(SHIPENV_E_ARG, "rollout: negative size");
  if (R * int64_t(S) > 0 && (!roots || !returns)) return fail(SHIPENV_E_ARG, "rollout: null buffer");
  if (R == 0 || S == 0) return 0;
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  RolloutArgs a{c->d_world, c->world_bytes, reinterpret_cast<const uint4*>(roots), R, S, L, returns,
                reinterpret_cast<uint4*>(final), steps, status, seed, root_id_base, {}};
  fill_round_keys(a.round_key, seed, kRolloutTweak);
  const size_t smem = ((size_t(c->world_bytes) + 127) & ~size_t(127)) + RngFifo<kRolloutThreads>::kWords * 4 * kRolloutThreads;   // world + word FIFO
  const int64_t tiles = (R * int64_t(S) + kRolloutThreads - 1) / kRolloutThreads;
  int grid = 0;
  if (int rc = grid_for(c, rollout_kernel, c->occ_rollout, kRolloutThreads, smem, tiles, &grid)) return rc;
  rollout_kernel<<<grid, kRolloutThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

int shipenv_mcts(shipenv_ctx* c, const shipenv_ship* roots, int64_t R, int S, int L, double c_param,
                 const double* log_table, int16_t* action, int32_t* root_visits, int16_t* child_action,
                 int32_t* child_visits, double* child_total, uint8_t* status, uint64_t seed, uint64_t root_id_base,
                 void* stream_) {
  if (int rc = check_ctx(c)) return rc;
  if (R < 0 || S < 0 || L < 0) return fail(SHIPENV_E_ARG, "mcts: negative size");
  if (S > SHIPENV_MCTS_MAX_SIMS) return fail(SHIPENV_E_ARG, "mcts: at most %d simulations", SHIPENV_MCTS_MAX_SIMS);
  if (R > 0 && (!roots || !action || !root_visits || !child_action || !child_visits || !child_total || !status))
    return fail(SHIPENV_E_ARG, "mcts: null buffer");
  if (R == 0) return 0;
  DeviceGuard guard(c->device);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = sync_world(c, stream)) return rc;
  MctsArgs a{c->d_world, c->world_bytes, reinterpret_cast<const uint4*>(roots), R, S, L, c_param,
             action, root_visits, child_action, child_visits, child_total, status, seed, root_id_base, {}, {}};
  // the log table travels in the kernel parameters: nothing of one call can be overwritten by the next
  a.log_table[0] = 0.0;
  for (int n = 1; n < S + 2; ++n) a.log_table[n] = log_table ? log_table[n] : std::log(double(n));
  fill_round_keys(a.round_key, seed, kRolloutTweak);
  // world + word FIFO + log table (+ the table of UCB1 exploration terms for every pair of visit counts, S <= 64)
  a.explore_table = S <= 64 ? 1 : 0;
  const size_t smem = ((size_t(c->world_bytes) + 127) & ~size_t(127)) + RngFifo<kMctsThreads>::kWords * 4 * kMctsThreads +
                      sizeof(double) * (SHIPENV_MCTS_MAX_SIMS + 2) +
                      (a.explore_table ? sizeof(double) * size_t(S + 2) * size_t(S + 3) / 2 : 0);
  // one node per simulation plus the root: the tree is a per-thread array in local memory, and the smaller its
  // footprint the more of it stays in L1 (52 nodes = the reference's 50 simulations)
  const int tree = S + 2 <= 52 ? 0 : S + 2 <= 64 ? 1 : 2;
  void (*kernel)(const MctsArgs) = tree == 0 ? mcts_kernel<52> : tree == 1 ? mcts_kernel<64> : mcts_kernel<128>;
  int grid = 0;
  if (int rc = grid_for(c, kernel, c->occ_mcts[tree], kMctsThreads, smem, (R + kMctsThreads - 1) / kMctsThreads, &grid)) return rc;
  kernel<<<grid, kMctsThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  ++g_launches;
  return 0;
}

}  // extern "C"
