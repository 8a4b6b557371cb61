"""ctypes binding of ``libshipenv.so`` (the C ABI declared in ``include/shipenv.h``).

There is no CPU fallback: if the library is missing this module raises at import of the
first symbol, and every call that fails inside the library raises :class:`ShipEnvError`
with the library's message.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHIPENV_LIB") or os.path.join(HERE, "libshipenv.so")   # SHIPENV_LIB: tuning builds only

MAX_PORTS = 32
MAX_DIM = 256
NONE_PORT = 255
ACT_POLICY = -1
ACT_MOVE_XY_BASE = -2        # a = -2 - ((dx + 90) * 181 + (dy + 90)): MOVE_SHIP by any (dx, dy), |dx|, |dy| <= 90
MOVE_XY_MAX = 90
ACT_NONE = -32768            # output only: a failed sample_action draw
ACT_EXT_CARGO = 0x4000
ACT_EXT_FUEL = 0x6000
N_CARGO_ACTIONS = 50
N_FUEL_ACTIONS = 200
F_AUTO_RESET = 1
F_ALL_POLICY = 2
F_CHAINED = 4

(ST_OK, ST_OFF_MAP, ST_NO_DEST, ST_SEL_RANGE, ST_SEL_ORIGIN, ST_NOT_AT_PORT, ST_BAD_AMOUNT,
 ST_UNKNOWN, ST_NO_PORTS, ST_SAMPLE_CRASH) = range(10)
ST_MCTS_UNSUPPORTED, ST_MCTS_NO_CHILDREN, MCTS_MAX_SIMS = 10, 11, 126


class ShipEnvError(RuntimeError):
    pass


class Consts(C.Structure):
    """``shipenv_consts``: Initial / Reward / Penalty of shipping/environment.py:8-26."""
    _fields_ = [
        ("cargo0", C.c_int32), ("capacity", C.c_int32), ("fuel0", C.c_double),
        ("r_deliver", C.c_double), ("r_arrive", C.c_double), ("r_closer", C.c_double),
        ("r_take_fuel", C.c_double), ("r_take_cargo", C.c_double),
        ("p_no_fuel", C.c_double), ("p_ground", C.c_double), ("p_water", C.c_double),
        ("p_loss", C.c_double), ("p_farther", C.c_double), ("p_use_fuel", C.c_double),
    ]


class StepIO(C.Structure):
    """``shipenv_step_io``: nullable device pointers."""
    _fields_ = [
        ("reward_f32", C.c_void_p), ("reward_f64", C.c_void_p),
        ("done_bits", C.c_void_p), ("reset_bits", C.c_void_p), ("error_bits", C.c_void_p),
        ("status", C.c_void_p), ("taken", C.c_void_p), ("obs", C.c_void_p),
        ("ep_return", C.c_void_p), ("stats", C.c_void_p), ("chain", C.c_void_p),
    ]


_SIGNATURES = {
    "shipenv_abi_version": (C.c_int, []),
    "shipenv_last_error": (C.c_char_p, []),
    "shipenv_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "shipenv_ctx_destroy": (C.c_int, [C.c_void_p]),
    "shipenv_set_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "shipenv_set_ports": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "shipenv_set_consts": (C.c_int, [C.c_void_p, C.POINTER(Consts)]),
    "shipenv_get_consts": (C.c_int, [C.c_void_p, C.POINTER(Consts)]),
    "shipenv_obs_width": (C.c_int, [C.c_void_p]),
    "shipenv_num_actions": (C.c_int, [C.c_void_p]),
    "shipenv_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64,
                                C.c_void_p, C.c_void_p]),
    "shipenv_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(StepIO), C.c_uint64,
                               C.c_uint64, C.c_int64, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]),
    "shipenv_chain_words": (C.c_size_t, [C.c_int64]),
    "shipenv_step_many": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                    C.c_int64, C.c_int, C.c_uint32, C.c_void_p]),
    "shipenv_step_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(StepIO), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_uint32,
                                      C.c_uint32, C.c_void_p]),
    "shipenv_step_host_workspace": (C.c_size_t, [C.c_int64]),
    "shipenv_serve_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32]),
    "shipenv_serve_stop": (C.c_int, [C.c_void_p]),
    "shipenv_serve_set_idle_us": (C.c_int, [C.c_void_p, C.c_uint32]),
    "shipenv_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int64, C.c_int,
                                    C.c_uint32, C.c_int, C.c_void_p]),
    "shipenv_sample_action": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p]),
    "shipenv_observe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "shipenv_valid_actions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "shipenv_discretize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "shipenv_rollout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "shipenv_mcts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                               C.c_uint64, C.c_void_p]),
    "shipenv_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "shipenv_port_stock": (None, [C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]),
    "shipenv_launch_count": (C.c_uint64, []),
    "shipenv_host_device_pointer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "shipenv_stream_synchronize": (C.c_int, [C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load ``libshipenv.so`` (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ShipEnvError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.shipenv_abi_version() != 2:
            raise ShipEnvError("libshipenv.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise ShipEnvError(lib().shipenv_last_error().decode() or f"shipenv error {rc}")


def philox4x32_10(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().shipenv_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return tuple(out)


def port_stock(seed, port_index):
    """(fuel, cargo) that ``add_port`` draws for port ``port_index`` (shipping/environment.py:63-64)."""
    f, c = C.c_uint8(), C.c_uint8()
    lib().shipenv_port_stock(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_uint32(port_index), C.byref(f), C.byref(c))
    return f.value, c.value


def launch_count():
    return int(lib().shipenv_launch_count())


def host_device_pointer(host_ptr):
    """Device-visible address of a page-locked host allocation (e.g. ``tensor.pin_memory().data_ptr()``)."""
    out = C.c_void_p()
    check(lib().shipenv_host_device_pointer(C.c_void_p(host_ptr), C.byref(out)))
    return int(out.value)


class Context:
    """One ``shipenv_ctx``: the world (map, ports, constants) of one device."""

    def __init__(self, device):
        self.device = int(device)
        h = C.c_void_p()
        check(lib().shipenv_ctx_create(self.device, C.byref(h)))
        self._h = h
        self._h_int = h.value                  # the handle as a plain int (read by the C fast path of the facade)
        self._serve = lib().shipenv_serve_step
        self.P = -1
        self.shape = None

    def close(self):
        if getattr(self, "_h", None):
            lib().shipenv_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, passable):
        """passable: uint8 ndarray [W, H], nonzero = np_game[x, y] != GROUND."""
        import numpy as np
        g = np.ascontiguousarray(passable, dtype=np.uint8)
        check(lib().shipenv_set_map(self._h, g.ctypes.data, g.shape[0], g.shape[1]))
        self.shape = g.shape

    def set_ports(self, xy, fuel, cargo):
        import numpy as np
        P = len(xy)
        xy = np.asarray(xy, dtype=np.uint8).reshape(P, 2)
        px, py = np.ascontiguousarray(xy[:, 0]), np.ascontiguousarray(xy[:, 1])
        pf, pc = np.asarray(fuel, dtype=np.uint8), np.asarray(cargo, dtype=np.uint8)
        check(lib().shipenv_set_ports(self._h, px.ctypes.data, py.ctypes.data, pf.ctypes.data, pc.ctypes.data, P))
        self.P = P

    def set_consts(self, **kw):
        c = Consts()
        check(lib().shipenv_set_consts(self._h, None))
        check(lib().shipenv_get_consts(self._h, C.byref(c)))
        for k, v in kw.items():
            if not hasattr(c, k):
                raise TypeError(f"unknown constant {k!r}")
            setattr(c, k, v)
        check(lib().shipenv_set_consts(self._h, C.byref(c)))

    def consts(self):
        c = Consts()
        check(lib().shipenv_get_consts(self._h, C.byref(c)))
        return c

    @property
    def obs_width(self):
        return 6 + 4 * self.P

    @property
    def num_actions(self):
        return 4 + self.P + N_CARGO_ACTIONS + N_FUEL_ACTIONS

    # raw launches: pointers are ints (tensor.data_ptr()) or None
    def reset(self, ships, clock, seed, env_id_base, n, mask_bits, stream):
        check(lib().shipenv_reset(self._h, ships, clock, seed, env_id_base, n, mask_bits, stream))

    def step(self, ships, clock, actions, io, seed, env_id_base, n, max_steps, flags, word0, stream):
        check(lib().shipenv_step(self._h, ships, clock, actions, C.byref(io), seed, env_id_base, n, max_steps,
                                 flags, word0, stream))

    def step_sample(self, ships, clock, actions, io, next_action, next_words, next_status, seed, env_id_base, n,
                    max_steps, flags, word0, stream):
        check(lib().shipenv_step_sample(self._h, ships, clock, actions, C.byref(io), next_action, next_words,
                                        next_status, seed, env_id_base, n, max_steps, flags, word0, stream))

    def step_many(self, ships, clock, actions, T, reward, done, reset, error, status, taken, reward64, seed,
                  env_id_base, n, max_steps, flags, stream):
        check(lib().shipenv_step_many(self._h, ships, clock, actions, T, reward, done, reset, error, status, taken,
                                      reward64, seed, env_id_base, n, max_steps, flags, stream))

    def step_host(self, ships, clock, h_actions, h_reward, h_done, h_reset, h_error, workspace, workspace_bytes, seed,
                  env_id_base, n, max_steps, flags, chunks, stream):
        check(lib().shipenv_step_host(self._h, ships, clock, h_actions, h_reward, h_done, h_reset, h_error, workspace,
                                      workspace_bytes, seed, env_id_base, n, max_steps, flags, chunks, stream))

    def synchronize(self, stream):
        check(lib().shipenv_stream_synchronize(self._h, stream))

    def serve_step(self, block_host_ptr, seed, env_id, max_steps, flags, word0):
        """``shipenv_serve_step``: the N = 1 step + look-ahead sample answered by the resident server kernel."""
        rc = self._serve(self._h, block_host_ptr, seed, env_id, max_steps, flags, word0)
        if rc:
            check(rc)

    def serve_stop(self):
        check(lib().shipenv_serve_stop(self._h))

    def serve_set_idle_us(self, idle_us):
        check(lib().shipenv_serve_set_idle_us(self._h, int(idle_us)))

    def sample_action(self, ships, clock, actions, words_used, status, seed, env_id_base, n, word0, stream):
        check(lib().shipenv_sample_action(self._h, ships, clock, actions, words_used, status, seed, env_id_base,
                                          n, word0, stream))

    def observe(self, ships, obs, n, stream):
        check(lib().shipenv_observe(self._h, ships, obs, n, stream))

    def valid_actions(self, ships, mask, n, stream):
        check(lib().shipenv_valid_actions(self._h, ships, mask, n, stream))

    def discretize(self, ships, keys, n, stream):
        check(lib().shipenv_discretize(self._h, ships, keys, n, stream))

    def mcts(self, roots, R, S, L, c_param, log_table, action, root_visits, child_action, child_visits, child_total,
             status, seed, root_id_base, stream):
        check(lib().shipenv_mcts(self._h, roots, R, S, L, c_param, log_table, action, root_visits, child_action,
                                 child_visits, child_total, status, seed, root_id_base, stream))

    def rollout(self, roots, R, S, L, returns, final, steps, status, seed, root_id_base, stream):
        check(lib().shipenv_rollout(self._h, roots, R, S, L, returns, final, steps, status, seed, root_id_base,
                                    stream))
