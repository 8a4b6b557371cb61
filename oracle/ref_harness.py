"""TEST INFRASTRUCTURE -- not part of the product path.

Drives the UNMODIFIED reference (``shipping/environment.py``, ``agents/mcts.py``,
``utils/preprocessing.py``) with ``shipping.environment.random`` rebound to the
Philox shim, under the batched semantics this repo adds on top (clock, status
codes, auto-reset, policy retry loop -- DESIGN.md "Batched contract").  It is
only runnable where a reference checkout exists (this container); its outputs
are committed as fixtures under ``tests/golden/`` by ``tests/golden/make_golden.py``
so that the GPU box never needs the reference.

Reference location: ``$SHIPPINGENV_REF`` or ``/root/reference``.
"""
import os
import struct
import sys
import types

import numpy as np

from .philox_shim import PhiloxRandom

NONE_PORT = 255
MAX_TRIES = 32  # policy/rollout: consecutive rejected draws before giving up

# status codes (one per reference exception, shipping/environment.py)
ST_OK, ST_OFF_MAP, ST_NO_DEST, ST_SEL_RANGE, ST_SEL_ORIGIN = 0, 1, 2, 3, 4
ST_NOT_AT_PORT, ST_BAD_AMOUNT, ST_UNKNOWN, ST_NO_PORTS, ST_SAMPLE_CRASH = 5, 6, 7, 8, 9

_MSG_TO_STATUS = {
    "Move is out of range": ST_OFF_MAP,                                        # :284
    "Cannot move without destination port": ST_NO_DEST,                        # :276
    "Port index is out of range": ST_SEL_RANGE,                                # :269
    "Destination port must be different from current one": ST_SEL_ORIGIN,      # :267
    "Not currently at port": ST_NOT_AT_PORT,                                   # :343,352
    "Invalid fuel amount": ST_BAD_AMOUNT,                                      # :346,355
    "Action category unknown": ST_UNKNOWN,                                     # :374
    "No ports available": ST_NO_PORTS,                                         # :360
}

EXT_CARGO, EXT_FUEL = 0x4000, 0x6000  # extended action encodings (amount = a - base)
POLICY = -1
MOVE_XY_BASE = -2                     # a = -2 - ((dx + 90) * 181 + (dy + 90)), |dx|, |dy| <= 90


def move_xy(dx, dy):
    return MOVE_XY_BASE - ((int(dx) + 90) * 181 + (int(dy) + 90))


def ref_root():
    return os.environ.get("SHIPPINGENV_REF", "/root/reference")


def import_reference():
    """Import the reference packages without touching them.  Returns a namespace."""
    root = ref_root()
    if not os.path.isdir(os.path.join(root, "shipping")):
        raise FileNotFoundError(f"no reference checkout at {root}")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    if "matplotlib" not in sys.modules:  # agents/sarsa.py:7 imports it; absent here
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    import shipping.environment as envmod
    import utils.preprocessing as prep
    import utils.constants as consts
    import agents.mcts as mcts
    ns = types.SimpleNamespace(envmod=envmod, prep=prep, consts=consts, mcts=mcts, root=root)
    ns.map_path = os.path.join(root, consts.DEFAULT_MAP_FILE)
    return ns


def f64_bits(v):
    return struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def status_of(exc):
    msg = str(exc)
    if msg in _MSG_TO_STATUS:
        return _MSG_TO_STATUS[msg]
    return ST_SAMPLE_CRASH  # TypeError/IndexError out of _sample_random_fuel (:163)


class RefBatch:
    """N reference ``Environment`` objects stepped under the batched contract."""

    def __init__(self, ns, n_envs, seed, ports, port_fuel=None, port_cargo=None,
                 env_id_base=0, map_path=None):
        self.ns = ns
        self.shim = PhiloxRandom(seed)
        ns.envmod.random = self.shim  # the only hook: a module attribute rebind
        self.n = n_envs
        self.base = env_id_base
        self.envs = []
        for _ in range(n_envs):
            e = ns.envmod.Environment(map_path or ns.map_path)
            for k, p in enumerate(ports):
                self.shim.seek_port(k)
                e.add_port(list(p))
            if port_fuel is not None:
                e.port_fuel = list(port_fuel)
            if port_cargo is not None:
                e.port_cargo = list(port_cargo)
            self.envs.append(e)
        self.clock = np.zeros(n_envs, dtype=np.uint32)

    # -- helpers ---------------------------------------------------------
    def decode(self, env, a):
        P = len(env.port_positions)
        AT = self.ns.envmod.ActionType
        if 0 <= a < 4 + P + 250:
            return self.ns.prep.map_action_to_env_action(int(a), env)
        if EXT_CARGO <= a < EXT_FUEL:
            return [AT.TAKE_CARGO, int(a) - EXT_CARGO]
        if EXT_FUEL <= a < 0x8000:
            return [AT.TAKE_FUEL, int(a) - EXT_FUEL]
        if MOVE_XY_BASE - 181 * 181 < a <= MOVE_XY_BASE:      # any (dx, dy): _move_ship takes it as it is (:278-287)
            k = MOVE_XY_BASE - int(a)
            return [AT.MOVE_SHIP, (k // 181 - 90, k % 181 - 90)]
        return [99, 0]

    def encode(self, env, act):
        """[category, value] -> flat int16 index (inverse of decode for in-range values)."""
        AT = self.ns.envmod.ActionType
        P = len(env.port_positions)
        cat, val = act
        if cat == AT.MOVE_SHIP:
            return {(0, -1): 0, (-1, 0): 1, (0, 1): 2, (1, 0): 3}[tuple(val)]
        if cat == AT.SELECT_PORT:
            return 4 + val
        if cat == AT.TAKE_CARGO:
            return 4 + P + val if 0 <= val < 50 else EXT_CARGO + val
        if cat == AT.TAKE_FUEL:
            return 4 + P + 50 + val if 0 <= val < 200 else EXT_FUEL + val
        raise ValueError(act)

    def record(self, i):
        e = self.envs[i]
        pos = e.ship_position if len(e.ship_position) == 2 else [0, 0]
        o = NONE_PORT if e.origin_port_index is None else e.origin_port_index
        d = NONE_PORT if e.destination_port_index is None else e.destination_port_index
        return (int(pos[0]), int(pos[1]), o, d, int(e.cargo), f64_bits(e.fuel))

    def set_state(self, i, x, y, origin, dest, cargo, fuel):
        e = self.envs[i]
        e.ship_position = [int(x), int(y)]
        e.origin_port_index = None if origin == NONE_PORT else int(origin)
        e.destination_port_index = None if dest == NONE_PORT else int(dest)
        e.cargo = int(cargo)
        e.fuel = fuel

    # -- batched contract ------------------------------------------------
    def reset(self, mask=None):
        for i, e in enumerate(self.envs):
            if mask is not None and not mask[i]:
                continue
            self.shim.seek_env(self.base + i, int(self.clock[i]))
            e.reset()
            self.clock[i] = ((int(self.clock[i]) >> 16) + 1) << 16

    def step(self, actions, max_steps=0, auto_reset=False, word0=0):
        """One batched step.  Returns dict of per-env arrays."""
        n = self.n
        out = dict(
            reward=np.zeros(n, np.float64), done=np.zeros(n, np.uint8),
            status=np.zeros(n, np.uint8), reset=np.zeros(n, np.uint8),
            taken=np.full(n, -1, np.int16),
        )
        for i, e in enumerate(self.envs):
            a = int(actions[i])
            c = int(self.clock[i])
            self.shim.seek_env(self.base + i, c, word=word0)
            status, reward, done, taken = self._one(e, a)
            out["status"][i] = status
            out["taken"][i] = taken
            if status != ST_OK:
                continue
            out["reward"][i] = reward
            out["done"][i] = done
            c1 = (c + 1) & 0xFFFFFFFF
            if auto_reset and (done or (max_steps > 0 and (c1 & 0xFFFF) >= max_steps)):
                e.reset()  # continues the same (env, clock) word stream
                c1 = (((c >> 16) + 1) << 16) & 0xFFFFFFFF
                out["reset"][i] = 1
            self.clock[i] = c1
        return out

    def _one(self, e, a):
        if a != POLICY:
            act = self.decode(e, a)
            try:
                _, reward, done, _ = e.step(act)
            except Exception as exc:  # reference raises before any draw/mutation
                return status_of(exc), 0.0, 0, a
            return ST_OK, reward, int(done), a
        # policy: the reference's `a = sample_action(); try step(a) except: continue`
        # loop (agents/mcts.py:228-236), bounded.
        if len(e.port_positions) == 0:
            return ST_NO_PORTS, 0.0, 0, -1
        if e.destination_port_index is None and e._get_current_port_idx() is None:
            return ST_NO_DEST, 0.0, 0, -1  # the reference would spin forever here
        last = ST_OFF_MAP
        for _ in range(MAX_TRIES):
            try:
                act = e.sample_action()
            except Exception:
                return ST_SAMPLE_CRASH, 0.0, 0, -1
            try:
                _, reward, done, _ = e.step(act)
            except Exception as exc:
                last = status_of(exc)
                continue
            return ST_OK, reward, int(done), self.encode(e, act)
        return last, 0.0, 0, -1

    def observe(self):
        """preprocess_state rows as the DQN consumes them: float32 [N, 6+4P]."""
        rows = [self.ns.prep.preprocess_state(e._build_state()).astype(np.float32)[0]
                for e in self.envs]
        return np.stack(rows)

    def states(self):
        recs = [self.record(i) for i in range(self.n)]
        return (np.array([r[:4] for r in recs], np.uint8),
                np.array([r[4] for r in recs], np.int32),
                np.array([r[5] for r in recs], np.uint64))


def ref_rollout(ns, shim, env, root_id, sim, max_steps):
    """Unmodified ``MCTSAgent._rollout`` (agents/mcts.py:211-238) from a copy of ``env``."""
    agent = ns.mcts.MCTSAgent.__new__(ns.mcts.MCTSAgent)
    agent.env = env
    agent.max_rollout_steps = max_steps
    agent.map_file = ns.map_path
    copy = agent._create_env_copy()
    shim.seek_rollout(root_id, sim)
    total = agent._rollout(copy, None)
    return total, copy


def ref_mcts_decide(ns, shim, env, root_id, sims, max_rollout_steps, c_param=1.4):
    """Unmodified ``MCTSAgent.choose_action`` (agents/mcts.py:240-292) on ``env``.

    The agent's own ``random`` (``random.choice(untried)``, agents/mcts.py:274) and the environment's draws
    both come from the shim; simulation ``s`` of root ``root_id`` reads the stream cell (root_id, s) -- the
    same cell the rollout kernel uses -- which is arranged by wrapping ``_create_env_copy`` (called once at
    the top of every simulation, agents/mcts.py:257) with a seek.  No reference file is edited.
    Returns (action, root) with ``root`` the reference's MCTSNode tree.
    """
    ns.mcts.random = shim
    agent = ns.mcts.MCTSAgent.__new__(ns.mcts.MCTSAgent)
    agent.env = env
    agent.num_simulations = sims
    agent.exploration_param = c_param
    agent.max_rollout_steps = max_rollout_steps
    agent.map_file = ns.map_path
    counter = [0]
    original = ns.mcts.MCTSAgent._create_env_copy

    def positioned_copy():
        shim.seek_rollout(root_id, counter[0])
        counter[0] += 1
        return original(agent)

    agent._create_env_copy = positioned_copy
    captured = {}
    node_cls = ns.mcts.MCTSNode
    real_init = node_cls.__init__

    def spy_init(self, state, action=None, parent=None):
        real_init(self, state, action=action, parent=parent)
        if parent is None and action is None:
            captured["root"] = self

    node_cls.__init__ = spy_init
    try:
        action = agent.choose_action(env._build_state())
    finally:
        node_cls.__init__ = real_init
    return action, captured["root"]
