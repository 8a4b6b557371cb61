#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Runs only where a reference checkout exists ($SHIPPINGENV_REF or /root/reference).
Every array written here is the output of the reference's own
``Environment.step/reset/sample_action`` (``shipping/environment.py``),
``MCTSAgent._rollout`` (``agents/mcts.py:211-238``) or ``preprocess_state``
(``utils/preprocessing.py:25-62``), driven through ``oracle/ref_harness.py`` with
``shipping.environment.random`` rebound to the Philox shim.  The fixtures pin
``oracle/shipenv_oracle.c`` (tests, ``-m "not gpu"``) and the CUDA path (``-m gpu``).
"""
import collections
import hashlib
import os
import random as stock_random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_harness import (RefBatch, import_reference, ref_rollout, f64_bits,  # noqa: E402
                                NONE_PORT, POLICY, EXT_CARGO, EXT_FUEL, move_xy)
from oracle.philox_shim import PhiloxRandom  # noqa: E402

SEED = 0x5EED
STOCK_FUEL = [20, 14, 18, 12, 16]   # random.seed(2024) draw, SURVEY App. C
STOCK_CARGO = [10, 11, 13, 20, 18]


def run(batch, actions, max_steps, auto_reset, word0=0):
    """Step ``batch`` over actions[T, N]; returns the per-step trace as arrays."""
    T, N = actions.shape
    tr = dict(
        actions=actions.astype(np.int16),
        pos=np.zeros((T, N, 4), np.uint8), cargo=np.zeros((T, N), np.int32),
        fuel=np.zeros((T, N), np.uint64), reward=np.zeros((T, N), np.float64),
        done=np.zeros((T, N), np.uint8), status=np.zeros((T, N), np.uint8),
        reset=np.zeros((T, N), np.uint8), taken=np.zeros((T, N), np.int16),
        clock=np.zeros((T, N), np.uint32),
    )
    p0, c0, f0 = batch.states()
    tr["init_pos"], tr["init_cargo"], tr["init_fuel"] = p0, c0, f0
    tr["init_clock"] = batch.clock.copy()
    for t in range(T):
        o = batch.step(actions[t], max_steps=max_steps, auto_reset=auto_reset, word0=word0)
        p, c, f = batch.states()
        tr["pos"][t], tr["cargo"][t], tr["fuel"][t] = p, c, f
        for k in ("reward", "done", "status", "reset", "taken"):
            tr[k][t] = o[k]
        tr["clock"][t] = batch.clock
    tr["max_steps"] = np.int64(max_steps)
    tr["auto_reset"] = np.int64(auto_reset)
    tr["word0"] = np.int64(word0)
    tr["env_id_base"] = np.uint64(batch.base)
    tr["seed"] = np.uint64(batch.shim.seed)
    return tr


def port_meta(batch):
    e = batch.envs[0]
    return dict(
        port_xy=np.array(e.port_positions, np.uint8).reshape(-1, 2),
        port_fuel=np.array(e.port_fuel, np.uint8), port_cargo=np.array(e.port_cargo, np.uint8),
        passable=(e.np_game != 0).astype(np.uint8),
    )


def bfs_next_move(passable, src, dst):
    """First move of a shortest passable path src->dst (moves N,E,S,W); None if unreachable."""
    moves = [(0, -1), (-1, 0), (0, 1), (1, 0)]
    W, H = passable.shape
    prev = {tuple(dst): None}
    q = collections.deque([tuple(dst)])
    while q:  # search backwards from dst so prev[] gives the forward step
        cur = q.popleft()
        if cur == tuple(src):
            break
        for k, (dx, dy) in enumerate(moves):
            nb = (cur[0] - dx, cur[1] - dy)  # nb --move k--> cur
            if 0 <= nb[0] < W and 0 <= nb[1] < H and passable[nb] and nb not in prev:
                prev[nb] = k
                q.append(nb)
    return prev.get(tuple(src))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def valid_action_golden(ns):
    """g9: the reference's DQNAgent.is_valid_action (agents/dqn.py:125-175) over all 259 actions."""
    import agents.dqn as dqn
    PORTS = ns.consts.DEFAULT_PORTS
    b = RefBatch(ns, 64, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    agent = dqn.DQNAgent.__new__(dqn.DQNAgent)
    pos, cargo, fuel, valid = [], [], [], []

    def snapshot():
        p, c, f = b.states()
        pos.append(p); cargo.append(c); fuel.append(f)
        rows = []
        for e in b.envs:
            agent.env = e
            rows.append([agent.is_valid_action(a) for a in range(259)])
        valid.append(np.array(rows, np.uint8))

    snapshot()                                        # right after reset: on the origin port
    for t in range(40):
        b.step(np.full(64, POLICY), 0, False)
        if t in (0, 1, 5, 39):
            snapshot()
    # crafted: ship on a port that is not its origin / origin None / on the destination port
    for i, (x, y, o, d) in enumerate([(60, 22, 4, 1), (60, 22, NONE_PORT, NONE_PORT), (41, 40, 0, 3), (78, 29, 1, 2)]):
        b.set_state(i, x, y, o, d, 3, 150.5)
    snapshot()
    save("g9_valid_actions", pos=np.concatenate(pos), cargo=np.concatenate(cargo), fuel=np.concatenate(fuel),
         valid=np.concatenate(valid), **port_meta(b))
    print("  valid-action rows:", sum(len(v) for v in valid), " mean valid per row:",
          float(np.concatenate(valid).sum(1).mean()))


def mcts_golden(ns):
    """g11: the reference's MCTSAgent.choose_action (agents/mcts.py:240-292), 50 simulations x 100-step rollouts
    (and a 12 x 20 variant), on roots reached by the random policy plus roots at the map border / low on fuel."""
    from oracle.ref_harness import ref_mcts_decide
    PORTS = ns.consts.DEFAULT_PORTS
    R = 40
    b = RefBatch(ns, R, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    for t in range(10):
        b.step(np.full(R, POLICY), 0, False)
    b.set_state(0, 0, 50, 1, 2, 4, 150.0)          # on the western border: one move is off-map for ever
    b.set_state(1, 99, 0, 1, 2, 0, 88.5)           # in a corner
    b.set_state(2, 61, 72, 4, 1, 9, 2.5)           # about to run out of fuel
    b.set_state(3, 60, 23, 4, 1, 12, 120.0)        # next to its destination
    rp, rc, rf = b.states()
    out = dict(root_pos=rp, root_cargo=rc, root_fuel=rf)
    move_idx = {(0, -1): 0, (-1, 0): 1, (0, 1): 2, (1, 0): 3}
    for S, L in ((50, 100), (12, 20)):
        act = np.full(R, -1, np.int16); rv = np.zeros(R, np.int32)
        ca = np.full((R, 4), -1, np.int16); cv = np.zeros((R, 4), np.int32); ct = np.zeros((R, 4), np.uint64)
        for r in range(R):
            a, root = ref_mcts_decide(ns, b.shim, b.envs[r], 7000 + r, S, L)
            act[r] = move_idx[tuple(a[1])]
            rv[r] = root.visits
            for k, ch in enumerate(root.children):
                ca[r, k] = move_idx[tuple(ch.action[1])]; cv[r, k] = ch.visits; ct[r, k] = f64_bits(ch.total_reward)
        tag = f"S{S}_L{L}"
        out[f"action_{tag}"], out[f"root_visits_{tag}"] = act, rv
        out[f"child_action_{tag}"], out[f"child_visits_{tag}"], out[f"child_total_{tag}"] = ca, cv, ct
        print(f"  mcts {tag}: action histogram", np.bincount(act, minlength=4), " mean root visits", rv.mean())
    out["root_id_base"] = np.uint64(7000)
    out["seed"] = np.uint64(SEED)
    save("g11_mcts", **out, **port_meta(b))


def discretize_golden(ns):
    """g10: the reference's discretize_state (utils/preprocessing.py:65-90) on trajectory and edge states."""
    PORTS = ns.consts.DEFAULT_PORTS
    b = RefBatch(ns, 64, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    pos, cargo, fuel, keys = [], [], [], []

    def snapshot():
        p, c, f = b.states()
        pos.append(p); cargo.append(c); fuel.append(f)
        rows = []
        for e in b.envs:
            (x, y), cb, fb, prev, dest = ns.prep.discretize_state(e._build_state())
            rows.append([x, y, max(cb, -32768), max(fb, -32768), prev, dest])
        keys.append(np.array(rows, np.int16))

    snapshot()
    for t in range(230):
        b.step(np.full(64, POLICY), 0, False)
        if t % 23 == 0:
            snapshot()
    edge = [0.0, -0.0, 9.999999999999998, 10.0, 19.999, 20.0, 39.99999, 40.0, 49.9, 50.0, 79.99, 80.0, 99.9, 100.0,
            1e6, -0.5, -9.99, -10.0, -19.9, -20.0, -20.00001, -399.9, -1e4, 1e300, 199.99, 200, 5e-324, 30, 60.0]
    for i, f in enumerate(edge):
        b.set_state(i, 10 + i, 20, NONE_PORT if i % 3 == 0 else i % 5, NONE_PORT if i % 4 == 0 else (i + 1) % 5, i, f)
    snapshot()
    save("g10_discretize", pos=np.concatenate(pos), cargo=np.concatenate(cargo), fuel=np.concatenate(fuel),
         keys=np.concatenate(keys))


def free_moves_golden(ns):
    """g13: MOVE_SHIP with moves outside ShipMove -- the reference's _move_ship takes any (dx, dy)
    (shipping/environment.py:278-287): fuel cost sqrt(dx^2+dy^2) * (1 + uniform), terrain of the target cell only,
    closer/farther on the real distances (ties count as farther).  Includes (0, 0), the four unit moves written
    as pairs, jumps over land, off-map targets, and crafted states (a null move on the destination port, negative
    fuel, a jump onto the destination with cargo)."""
    PORTS = ns.consts.DEFAULT_PORTS
    rng = np.random.default_rng(1313)
    T, N = 320, 64
    b = RefBatch(ns, N, SEED + 13, PORTS, STOCK_FUEL, STOCK_CARGO, env_id_base=77)
    b.reset()
    crafted = [(60, 22, 4, 1, 7, 150.0), (60, 22, 4, 1, 0, -2.5), (57, 26, 4, 1, 9, 99.5), (62, 72, 4, 1, 3, 0.4),
               (0, 0, 4, 1, 0, 200), (99, 99, 0, 2, 55, 200)]
    for i, c in enumerate(crafted):
        b.set_state(i, *c)
    first = [move_xy(0, 0), move_xy(0, 0), move_xy(3, -4), move_xy(-1, 0), move_xy(90, 90), move_xy(-90, -90)]
    kind = rng.random((T, N))
    small = np.vectorize(move_xy)(rng.integers(-3, 4, (T, N)), rng.integers(-3, 4, (T, N)))
    medium = np.vectorize(move_xy)(rng.integers(-12, 13, (T, N)), rng.integers(-12, 13, (T, N)))
    large = np.vectorize(move_xy)(rng.integers(-90, 91, (T, N)), rng.integers(-90, 91, (T, N)))
    acts = rng.integers(0, 4, (T, N))
    acts = np.where(kind > 0.25, small, acts)
    acts = np.where(kind > 0.80, medium, acts)
    acts = np.where(kind > 0.95, large, acts)
    acts[0, :len(first)] = first
    acts[1, :len(first)] = first[::-1]
    tr = run(b, acts, 120, True)
    xy = acts <= -2
    print("  free moves: accepted =", int(((tr["status"] == 0) & xy).sum()), " off-map =", int(((tr["status"] == 1) & xy).sum()),
          " arrivals =", int((tr["reward"] > 5).sum()), " dones =", int(tr["done"].sum()), " resets =", int(tr["reset"].sum()))
    save("g13_free_moves", **tr, **port_meta(b))


def facade_trace_golden(ns):
    """g14: the reference's single-env API driven call by call -- Environment(map, game_size), add_port, reset,
    step([category, value]) -- on the default map with moves of any (dx, dy) (lists, tuples, numpy integers, (0, 0),
    a float component, a triple) and on NON-SQUARE game sizes, where cv2.resize makes the grid size_game[1] x
    size_game[0] while _is_within_map bounds x by size_game[0] (shipping/environment.py:53, :99-100): moves that pass
    the bounds check but leave the array raise numpy's IndexError at :293.  Per call the trace holds the exception
    (class name + message) or the new attributes, the reward bits, done, and a hash of np_game (the paint).
    The shim is positioned the way the facade draws: cell (env id, clock), clock + 1 per accepted step, reset draws
    from the current cell and moves the clock to the next episode."""
    AT = ns.envmod.ActionType
    out = {}
    scenarios = [("sq", (100, 100), 5, 260), ("wide", (120, 80), 4, 260), ("tall", (80, 120), 4, 260)]
    for tag, size, n_ports, T in scenarios:
        rng = np.random.default_rng(1400 + len(tag) + size[0])
        shim = PhiloxRandom(SEED + 14)
        ns.envmod.random = shim
        env_id = 5 + size[0]
        e = ns.envmod.Environment(ns.map_path, size)
        out[f"{tag}_grid"] = e.np_game.astype(np.uint8)
        out[f"{tag}_size"] = np.array(size, np.int64)
        lim = min(size)
        if tag == "sq":
            ports = [list(p) for p in ns.consts.DEFAULT_PORTS]
        else:      # ports on water cells of the part of the grid a ship can reach, spread out
            cand = [(x, y) for x in range(4, lim - 4) for y in range(4, lim - 4)
                    if e.np_game[x, y] != 0 and e.np_game[x - 1, y] != 0 and e.np_game[x + 1, y] != 0]
            ports = [list(map(int, cand[k])) for k in rng.choice(len(cand), n_ports, replace=False)]
        for k, p in enumerate(ports):
            shim.seek_port(k)
            e.add_port(list(p))
        out[f"{tag}_ports"] = np.array(ports, np.int64)
        out[f"{tag}_env_id"] = np.int64(env_id)
        clock = 0
        rec = dict(kind=[], cat=[], v0=[], v1=[], exc=[], state=[], fuel=[], reward=[], done=[], grid=[])

        def snap(kind, cat, v0, v1, exc, reward, done):
            pos = e.ship_position if len(e.ship_position) == 2 else [255, 255]
            o = NONE_PORT if e.origin_port_index is None else e.origin_port_index
            d = NONE_PORT if e.destination_port_index is None else e.destination_port_index
            rec["kind"].append(kind); rec["cat"].append(cat); rec["v0"].append(v0); rec["v1"].append(v1)
            rec["exc"].append(exc)
            rec["state"].append([int(pos[0]), int(pos[1]), o, d, int(e.cargo), int(isinstance(e.fuel, int))])
            rec["fuel"].append(f64_bits(e.fuel)); rec["reward"].append(f64_bits(reward)); rec["done"].append(int(done))
            rec["grid"].append(int.from_bytes(hashlib.sha256(e.np_game.astype(np.uint8).tobytes()).digest()[:8], "little"))

        def do_reset():
            nonlocal clock
            shim.seek_env(env_id, clock)
            e.reset()
            clock = ((clock >> 16) + 1) << 16
            snap(0, 0, 0, 0, "", 0.0, 0)

        def do_step(cat, value, kind=1):
            nonlocal clock
            shim.seek_env(env_id, clock)
            v0, v1 = (value[0], value[1]) if cat == AT.MOVE_SHIP and len(value) == 2 else (value if cat != AT.MOVE_SHIP else 0, 0)
            try:
                _, r, d, _ = e.step([cat, value])
            except Exception as exc:
                snap(kind, cat, v0, v1, f"{type(exc).__name__}|{exc}", 0.0, 0)
                return
            clock += 1
            snap(kind, cat, v0, v1, "", r, d)

        do_reset()
        # kinds: 1 list move, 2 tuple move, 3 numpy-integer move, 4 float component, 5 triple
        do_step(AT.MOVE_SHIP, [0, 0]); do_step(AT.MOVE_SHIP, (0, 0), 2)
        do_step(AT.MOVE_SHIP, [0.5, 0], 4); do_step(AT.MOVE_SHIP, [200.5, 0], 4); do_step(AT.MOVE_SHIP, [1, 0, 0], 5)
        do_step(AT.MOVE_SHIP, [-200, 3]); do_step(AT.MOVE_SHIP, [np.int64(1), np.int64(-1)], 3)
        for t in range(T):
            cur = e._get_current_port_idx()
            u = rng.random()
            if cur is not None and e.cargo == 0 and u < 0.7:
                do_step(AT.TAKE_CARGO, int(rng.integers(1, e.port_cargo[cur] + 1)))
            elif u < 0.04:
                do_reset()
            elif u < 0.10:
                do_step(AT.SELECT_PORT, int(rng.integers(0, len(ports) + 1)))
            else:
                tx, ty = e.port_positions[e.destination_port_index]
                x, y = e.ship_position
                if u < 0.55:       # towards the destination, a few cells at a time (may jump over land)
                    dx = int(np.clip(tx - x, -3, 3)) + int(rng.integers(-1, 2))
                    dy = int(np.clip(ty - y, -3, 3)) + int(rng.integers(-1, 2))
                elif u < 0.85:     # anywhere nearby
                    dx, dy = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
                else:              # far: off the map, or (non-square) inside size_game but outside the grid
                    dx, dy = int(rng.integers(-60, 61)), int(rng.integers(-60, 61))
                    if tag == "wide" and u > 0.93:
                        dx = int(rng.integers(80, 120)) - x
                    if tag == "tall" and u > 0.93:
                        dy = int(rng.integers(80, 120)) - y
                kind = 1 + int(rng.integers(0, 3))
                value = [dx, dy] if kind == 1 else (dx, dy) if kind == 2 else [np.int64(dx), np.int64(dy)]
                do_step(AT.MOVE_SHIP, value, kind)
        excs = [x.split("|")[0] for x in rec["exc"] if x]
        print(f"  facade trace {tag} {size}: calls {len(rec['kind'])}, exceptions {dict(collections.Counter(excs))}, "
              f"arrivals {sum(1 for r in rec['reward'] if np.array(r, np.uint64).view(np.float64) > 5)}")
        for k in ("kind", "cat", "v1", "done"):
            out[f"{tag}_{k}"] = np.array(rec[k], np.int64)
        out[f"{tag}_v0"] = np.array([float(v) for v in rec["v0"]], np.float64)
        out[f"{tag}_state"] = np.array(rec["state"], np.int64)
        for k, name in (("fuel", "fuel"), ("reward", "reward"), ("grid", "paint")):
            out[f"{tag}_{name}"] = np.array(rec[k], np.uint64)
        out[f"{tag}_exc"] = np.array(rec["exc"])
    out["seed"] = np.uint64(SEED + 14)
    save("g14_facade_trace", **out)


def main():
    ns = import_reference()
    PORTS = ns.consts.DEFAULT_PORTS
    if "free_moves" in sys.argv[1:]:
        free_moves_golden(ns)
        return
    if "facade_trace" in sys.argv[1:]:
        facade_trace_golden(ns)
        return
    if "valid" in sys.argv[1:]:
        valid_action_golden(ns)
        return
    if "discretize" in sys.argv[1:]:
        discretize_golden(ns)
        return
    if "mcts" in sys.argv[1:]:
        mcts_golden(ns)
        return

    # --- self-check: the reference imported here is the program SURVEY App. C hashed ----
    ns.envmod.random = stock_random
    stock_random.seed(2024)
    e = ns.envmod.Environment(ns.map_path)
    h_map = hashlib.sha256(e.np_game.astype(np.uint8).tobytes()).hexdigest()
    assert h_map == "49253a2cbaf1c86b8ae4f4b400c082eac97da02718ee2237b1734783a12622e8", h_map
    for p in PORTS:
        e.add_port(list(p))
    assert e.port_fuel == STOCK_FUEL and e.port_cargo == STOCK_CARGO
    h = hashlib.sha256()
    e.reset()
    h.update(repr((e.ship_position, e.origin_port_index, e.destination_port_index)).encode())
    n = 0
    while n < 1000:
        a = e.sample_action()
        try:
            _, r, d, _ = e.step(a)
        except Exception:
            continue
        n += 1
        rec = (a[0], a[1], tuple(e.ship_position), float(e.fuel).hex(), e.cargo,
               e.origin_port_index, e.destination_port_index, float(r).hex(), d)
        h.update(repr(rec).encode())
        if d:
            e.reset()
            h.update(repr((e.ship_position, e.origin_port_index, e.destination_port_index)).encode())
    print("stock-RNG trajectory sha256:", h.hexdigest())

    rng = np.random.default_rng(1234)

    # --- default map ------------------------------------------------------------------
    b = RefBatch(ns, 1, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    meta = port_meta(b)
    assert hashlib.sha256(meta["passable"].tobytes()).hexdigest() == \
        "2774b34f9bc98e9806ad059529d413d6af2685fe5270e3817ebe19e5c6c5b0d0"
    np_game_before = ns.envmod.Environment(ns.map_path).np_game.astype(np.uint8)
    save("default_map", np_game=np_game_before, ports=np.array(PORTS, np.uint8),
         stock_fuel=np.array(STOCK_FUEL, np.uint8), stock_cargo=np.array(STOCK_CARGO, np.uint8),
         **{k: v for k, v in meta.items() if k == "passable"})

    # --- port stock drawn by add_port under the shim ----------------------------------
    stock = {}
    for seed in (SEED, 0, 0xDEADBEEFCAFEF00D):
        bb = RefBatch(ns, 1, seed, PORTS)
        stock[f"fuel_{seed:x}"] = np.array(bb.envs[0].port_fuel, np.uint8)
        stock[f"cargo_{seed:x}"] = np.array(bb.envs[0].port_cargo, np.uint8)
    save("port_stock", **stock)

    # --- G1: uniform random moves, auto-reset, truncation at 300 ----------------------
    b = RefBatch(ns, 64, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    save("g1_moves", **run(b, rng.integers(0, 4, (640, 64)), 300, True), **port_meta(b))

    # --- G2: on-device policy (sample_action retry loop), truncation at 500 -----------
    b = RefBatch(ns, 64, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    tr = run(b, np.full((640, 64), POLICY), 500, True)
    obs = b.observe()
    save("g2_policy", **tr, obs=obs, **port_meta(b))

    # --- G3: the whole flat action space incl. rejected actions -----------------------
    b = RefBatch(ns, 64, SEED, PORTS, STOCK_FUEL, STOCK_CARGO, env_id_base=(1 << 33) + 5)
    b.reset()
    T, N = 400, 64
    kind = rng.random((T, N))
    acts = rng.integers(0, 4, (T, N))
    acts = np.where(kind > 0.55, rng.integers(4, 9, (T, N)), acts)
    acts = np.where(kind > 0.65, rng.integers(9, 59, (T, N)), acts)
    acts = np.where(kind > 0.80, rng.integers(59, 259, (T, N)), acts)
    junk = rng.choice(np.array([-2, -300, 259, 260, 1000, EXT_CARGO, EXT_CARGO + 7, EXT_CARGO + 300,
                                EXT_FUEL, EXT_FUEL + 3, EXT_FUEL + 8191, 0x3FFF, -32768, POLICY]), (T, N))
    acts = np.where(kind > 0.93, junk, acts)
    save("g3_mixed", **run(b, acts, 200, True), **port_meta(b))

    # --- G4: navigator -- take cargo, sail the shortest route, deliver; no auto-reset --
    b = RefBatch(ns, 32, SEED ^ 0xABCDEF, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    T, N = 700, 32
    passable = port_meta(b)["passable"].astype(bool)
    acts = np.zeros((T, N), np.int64)
    tr_parts = []
    # G12: the reference's own grid paint (np_game: BOAT on the ship, TRAVEL behind it, PORT cells painted over;
    # shipping/environment.py:228, 240, 303-304) of a few of these envs after reset and after every step
    paint_envs = [0, 3, 9, 30]
    frames = np.zeros((T + 1, len(paint_envs), 100, 100), np.uint8)
    for k, i in enumerate(paint_envs):
        frames[0, k] = b.envs[i].np_game
    for t in range(T):
        row = np.zeros(N, np.int64)
        for i, e in enumerate(b.envs):
            cur = e._get_current_port_idx()
            if cur is not None and e.cargo == 0:
                row[i] = 4 + 5 + min(e.port_cargo[cur], 3 + (i % 17))       # TAKE_CARGO
            elif cur is not None and e.fuel < 60 and (t + i) % 3 == 0:
                row[i] = 4 + 5 + 50 + e.port_fuel[cur]                        # TAKE_FUEL
            else:
                k = bfs_next_move(passable, e.ship_position,
                                  e.port_positions[e.destination_port_index])
                row[i] = k if k is not None and (t * 7 + i) % 11 else int(rng.integers(0, 4))
        acts[t] = row
        tr_parts.append(run(b, row[None, :], 0, False))
        for k, i in enumerate(paint_envs):
            frames[t + 1, k] = b.envs[i].np_game
    tr = {k: np.concatenate([p[k] for p in tr_parts]) for k in
          ("actions", "pos", "cargo", "fuel", "reward", "done", "status", "reset", "taken", "clock")}
    for k in ("init_pos", "init_cargo", "init_fuel", "init_clock", "max_steps", "auto_reset",
              "word0", "env_id_base", "seed"):
        tr[k] = tr_parts[0][k]
    print("  navigator: arrivals =", int((tr["reward"] > 5).sum()), " dones =", int(tr["done"].sum()),
          " min fuel =", np.frombuffer(tr["fuel"][-1].tobytes(), np.float64).min())
    save("g4_navigator", **tr, **port_meta(b))
    # stored as the frame after reset plus the cells each step changed: (step, env slot, x, y, new value)
    changed = [(t, k, x, y, frames[t + 1, k, x, y]) for t in range(T) for k in range(len(paint_envs))
               for x, y in zip(*np.nonzero(frames[t + 1, k] != frames[t, k]))]
    save("g12_paint", env_index=np.array(paint_envs, np.int64), frame0=frames[0],
         changes=np.array(changed, np.int32).reshape(-1, 5))

    # --- G5: all-water synthetic map: edges, off-map rejections, 2 ports --------------
    import cv2
    img = np.full((355, 533), 255, np.uint8)
    water_png = "/tmp/shipenv_allwater.png"
    cv2.imwrite(water_png, img)
    for tag, ports, acts_fn in (
        ("g5_edges_moves", [[0, 0], [99, 99], [0, 99], [50, 50]],
         lambda: rng.integers(0, 4, (500, 48))),
        ("g5_edges_policy", [[0, 0], [1, 0]], lambda: np.full((500, 48), POLICY)),
    ):
        b = RefBatch(ns, 48, SEED + 1, ports, map_path=water_png)
        b.reset()
        tr = run(b, acts_fn(), 120, True, word0=0)
        print(f"  {tag}: off-map rejections =", int((tr['status'] == 1).sum()),
              " arrivals =", int((tr["reward"] > 5).sum()))
        save(tag, **tr, **port_meta(b))

    # --- G6: crafted states (each env starts from a hand-built state) -----------------
    crafted = [
        # x, y, origin, dest, cargo, fuel, actions...
        (62, 72, 4, 1, 0, 0.5, [1, 1, 0, 2]),            # out of fuel: done, move proceeds, fuel < 0
        (62, 72, 4, 1, 0, 0.95, [1, 1, 1, 1]),           # fuel ~ cost: done depends on the jitter
        (41, 40, 0, NONE_PORT, 0, 200, [POLICY, POLICY, POLICY, POLICY]),  # policy selects a port
        (41, 40, NONE_PORT, NONE_PORT, 0, 200, [0, 4, 5, 2]),   # move w/o dest (2), select ok
        (50, 50, 0, NONE_PORT, 0, 200, [POLICY, 3, 9, 60]),     # policy w/o dest off port -> 2
        (61, 72, 4, 1, 60, 200, [1, 3, 1, 3]),           # cargo > capacity: loss every move
        (61, 72, 4, 1, 1, 200, [1, 3, 1, 3]),            # cargo 1
        (60, 22, 4, 1, 7, 200, [0, 1, 2, 3]),            # sits on its destination: blocked-move arrival
        (60, 23, 4, 1, 9, 150.25, [0, 0, 2, 0]),         # sail onto the destination with cargo
        (49, 72, 3, 0, 5, 0, [POLICY, POLICY, 2, 2]),    # fuel == 0 at port, cargo != 0: sample crash (9)
        (49, 72, 3, 0, 0, 0, [POLICY, POLICY, POLICY, 2]),  # fuel == 0 but cargo == 0: takes cargo
        (62, 72, 4, 1, 0, 200, [4 + 4, 4 + 1, 4 + 5, 4 + 0]),  # select origin (4), ok, range.. (flat 9 = cargo 0 -> 6)
        (62, 72, 4, 1, 0, 200, [9 + 18, 9 + 19, 59 + 16, 59 + 17]),  # stock bounds 18/16 at port 4
        (10, 10, 4, 1, 0, 200, [9 + 1, 59 + 1, 4 + 2, 4 + 7]),       # not at port (5), select range (3)
        (62, 72, 4, 1, 2_000_000, 1e6, [1, 3, 1, 3]),    # large cargo: int(beta*cargo) in fp64
        (62, 72, 4, 1, 0, -3.5, [1, 1, 59 + 5, 1]),      # negative fuel keeps stepping
    ]
    N = len(crafted)
    b = RefBatch(ns, N, SEED + 2, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.clock[:] = np.arange(N, dtype=np.uint32) * 65536 * 3 + 17
    for i, c in enumerate(crafted):
        b.set_state(i, *c[:6])
    acts = np.array([c[6] for c in crafted]).T  # [4, N]
    acts = np.concatenate([acts, rng.integers(0, 4, (60, N))])
    tr = run(b, acts, 0, False)
    print("  crafted status rows:\n", tr["status"][:4])
    save("g6_crafted", **tr, **port_meta(b))

    # --- G7: word0 offset (facade: sample_action() then step()) -----------------------
    b = RefBatch(ns, 16, SEED + 3, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    save("g7_word0", **run(b, rng.integers(0, 4, (64, 16)), 0, False, word0=1), **port_meta(b))

    # --- G8: rollouts = the reference's MCTSAgent._rollout ----------------------------
    R, S = 48, 6
    b = RefBatch(ns, R, SEED, PORTS, STOCK_FUEL, STOCK_CARGO)
    b.reset()
    for t in range(10):
        b.step(np.full(R, POLICY), 0, False)
    # make some roots short on fuel so that rollouts terminate by `done`
    for i in range(0, R, 6):
        b.envs[i].fuel = 3.0 + i
    rp, rc, rf = b.states()
    out = dict(root_pos=rp, root_cargo=rc, root_fuel=rf)
    for L in (100, 25):
        ret = np.zeros((R, S), np.uint64)
        fin_pos = np.zeros((R, S, 4), np.uint8)
        fin_cargo = np.zeros((R, S), np.int32)
        fin_fuel = np.zeros((R, S), np.uint64)
        for r in range(R):
            for s in range(S):
                total, copy = ref_rollout(ns, b.shim, b.envs[r], 1000 + r, s, L)
                ret[r, s] = f64_bits(total)
                fin_pos[r, s] = (copy.ship_position[0], copy.ship_position[1],
                                 copy.origin_port_index, copy.destination_port_index)
                fin_cargo[r, s] = copy.cargo
                fin_fuel[r, s] = f64_bits(copy.fuel)
        out[f"ret_L{L}"], out[f"fin_pos_L{L}"] = ret, fin_pos
        out[f"fin_cargo_L{L}"], out[f"fin_fuel_L{L}"] = fin_cargo, fin_fuel
    out["root_id_base"] = np.uint64(1000)
    out["seed"] = np.uint64(SEED)
    save("g8_rollouts", **out, **port_meta(b))

    valid_action_golden(ns)
    discretize_golden(ns)
    mcts_golden(ns)
    free_moves_golden(ns)
    facade_trace_golden(ns)


if __name__ == "__main__":
    main()
