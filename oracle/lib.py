"""TEST INFRASTRUCTURE -- ctypes/numpy front end of the C oracle (oracle/shipenv_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
MAX_PORTS = 32
NONE_PORT = 255

SHIP_DTYPE = np.dtype([("x", "u1"), ("y", "u1"), ("origin", "u1"), ("dest", "u1"),
                       ("cargo", "<i4"), ("fuel", "<f8")])
assert SHIP_DTYPE.itemsize == 16


class World(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32), ("P", C.c_int32),
        ("cargo0", C.c_int32), ("capacity", C.c_int32),
        ("fuel0", C.c_double),
        ("r_deliver", C.c_double), ("r_arrive", C.c_double), ("r_closer", C.c_double),
        ("r_take_fuel", C.c_double), ("r_take_cargo", C.c_double),
        ("p_no_fuel", C.c_double), ("p_ground", C.c_double), ("p_water", C.c_double),
        ("p_loss", C.c_double), ("p_farther", C.c_double), ("p_use_fuel", C.c_double),
        ("px", C.c_uint8 * MAX_PORTS), ("py", C.c_uint8 * MAX_PORTS),
        ("pfuel", C.c_uint8 * MAX_PORTS), ("pcargo", C.c_uint8 * MAX_PORTS),
        ("passable", C.c_void_p),
    ]


# reference constants, shipping/environment.py:8-26
DEFAULT_CONSTS = dict(
    cargo0=0, capacity=50, fuel0=200.0,
    r_deliver=2.0, r_arrive=10.0, r_closer=2.0, r_take_fuel=0.05, r_take_cargo=0.05,
    p_no_fuel=-10.0, p_ground=-5.0, p_water=-1.0, p_loss=-3.0, p_farther=-2.0, p_use_fuel=-0.0001,
)


def build(force=False):
    import hashlib
    src = os.path.join(HERE, "shipenv_oracle.c")
    digest = hashlib.sha256(open(src, "rb").read() + open(os.path.join(HERE, "Makefile"), "rb").read()).hexdigest()
    stamp = SO + ".srchash"
    if force or not os.path.exists(SO) or not os.path.exists(stamp) or open(stamp).read().strip() != digest:
        subprocess.run(["make", "-C", HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(digest)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        assert _lib.orc_sizeof_world() == C.sizeof(World), "oracle World layout mismatch"
        assert _lib.orc_sizeof_ship() == 16
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """A world (map + ports + constants) and the batched reference semantics over numpy arrays."""

    def __init__(self, passable, port_xy, port_fuel, port_cargo, **consts):
        self.passable = np.array(passable, dtype=np.uint8, order="C")          # own copy
        for x, y in port_xy:                                                   # add_port stamps Entity.PORT on the
            self.passable[int(x), int(y)] = 1                                  # cell (env:65): never GROUND afterwards
        w = World()
        w.W, w.H = self.passable.shape
        w.P = len(port_xy)
        for k, v in {**DEFAULT_CONSTS, **consts}.items():
            setattr(w, k, v)
        for k in range(w.P):
            w.px[k], w.py[k] = int(port_xy[k][0]), int(port_xy[k][1])
            w.pfuel[k], w.pcargo[k] = int(port_fuel[k]), int(port_cargo[k])
        w.passable = self.passable.ctypes.data
        self.w = w
        self.P = w.P

    @classmethod
    def from_npz(cls, z, **consts):
        return cls(z["passable"], z["port_xy"], z["port_fuel"], z["port_cargo"], **consts)

    def reset(self, ships, clock, seed, env_base=0, mask=None):
        lib().orc_reset(C.byref(self.w), _p(ships), _p(clock), C.c_uint64(seed), C.c_uint64(env_base),
                        C.c_int64(len(ships)), _p(mask))

    def step(self, ships, clock, actions, seed, env_base=0, max_steps=0, auto_reset=False, word0=0):
        n = len(ships)
        out = dict(reward=np.zeros(n, np.float64), done=np.zeros(n, np.uint8),
                   status=np.zeros(n, np.uint8), reset=np.zeros(n, np.uint8),
                   taken=np.zeros(n, np.int16))
        actions = np.ascontiguousarray(actions, dtype=np.int16)
        lib().orc_step(C.byref(self.w), _p(ships), _p(clock), _p(actions), _p(out["reward"]),
                       _p(out["done"]), _p(out["status"]), _p(out["reset"]), _p(out["taken"]),
                       C.c_uint64(seed), C.c_uint64(env_base), C.c_int64(n), C.c_int(max_steps),
                       C.c_int(int(auto_reset)), C.c_uint32(word0))
        return out

    def sample_action(self, ships, clock, seed, env_base=0, word0=0):
        """Environment.sample_action per env -> (actions int16, words_used uint8, status uint8)."""
        n = len(ships)
        a, wu, st = np.zeros(n, np.int16), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        lib().orc_sample_action(C.byref(self.w), _p(ships), _p(clock), _p(a), _p(wu), _p(st), C.c_uint64(seed),
                                C.c_uint64(env_base), C.c_int64(n), C.c_uint32(word0))
        return a, wu, st

    def observe(self, ships):
        obs = np.zeros((len(ships), 6 + 4 * self.P), np.float32)
        lib().orc_observe(C.byref(self.w), _p(ships), _p(obs), C.c_int64(len(ships)))
        return obs

    def valid_actions(self, ships):
        valid = np.zeros((len(ships), 4 + self.P + 250), np.uint8)
        lib().orc_valid_actions(C.byref(self.w), _p(ships), _p(valid), C.c_int64(len(ships)))
        return valid

    def rollout(self, roots, S, L, seed, root_base=0):
        R = len(roots)
        ret = np.zeros((R, S), np.float64)
        fin = np.zeros((R, S), SHIP_DTYPE)
        steps = np.zeros((R, S), np.int32)
        status = np.zeros((R, S), np.uint8)
        lib().orc_rollout(C.byref(self.w), _p(roots), C.c_int64(R), C.c_int(S), C.c_int(L), _p(ret),
                          _p(fin), _p(steps), _p(status), C.c_uint64(seed), C.c_uint64(root_base))
        return ret, fin, steps, status


def mcts(oracle, roots, S, L, seed, root_base=0, c_param=1.4):
    """MCTSAgent.choose_action per root -> dict(action, root_visits, child_action/visits/total [R,4], status)."""
    R = len(roots)
    out = dict(action=np.zeros(R, np.int16), root_visits=np.zeros(R, np.int32),
               child_action=np.zeros((R, 4), np.int16), child_visits=np.zeros((R, 4), np.int32),
               child_total=np.zeros((R, 4), np.float64), status=np.zeros(R, np.uint8))
    log_table = np.concatenate([[0.0], np.log(np.arange(1, max(S, 1) + 2))])
    lib().orc_mcts(C.byref(oracle.w), _p(roots), C.c_int64(R), C.c_int(S), C.c_int(L), C.c_double(c_param),
                   _p(log_table), _p(out["action"]), _p(out["root_visits"]), _p(out["child_action"]),
                   _p(out["child_visits"]), _p(out["child_total"]), _p(out["status"]), C.c_uint64(seed),
                   C.c_uint64(root_base))
    return out


def discretize(ships):
    keys = np.zeros((len(ships), 6), np.int16)
    lib().orc_discretize(_p(ships), _p(keys), C.c_int64(len(ships)))
    return keys


def ships_from(pos, cargo, fuel_bits):
    """Build a ship array from golden-fixture columns (pos u8[N,4], cargo i32, fuel u64 bits)."""
    s = np.zeros(len(cargo), SHIP_DTYPE)
    s["x"], s["y"], s["origin"], s["dest"] = pos[:, 0], pos[:, 1], pos[:, 2], pos[:, 3]
    s["cargo"] = cargo
    s["fuel"] = np.asarray(fuel_bits, np.uint64).view(np.float64)
    return s


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return tuple(out)
