"""GPU: the drop-in ``Environment`` (the reference's Python API, shipping/environment.py) on the kernels.
Replays reference-generated traces one env at a time and exercises the call patterns of the
reference's agents (agents/mcts.py, agents/dqn.py, agents/sarsa.py)."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import lib as orc

pytestmark = pytest.mark.gpu
MAP = os.path.join(GOLDEN, "default_map.npz")
PORTS = [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]
STOCK_FUEL, STOCK_CARGO = [20, 14, 18, 12, 16], [10, 11, 13, 20, 18]


def bits(v):
    return struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def make_env(seed, env_id=0, ports=PORTS, stock=True):
    from shippingenv_b200 import Environment
    env = Environment(MAP, seed=seed, env_id=env_id)
    for p in ports:
        env.add_port(list(p))
    if stock:
        env.port_fuel, env.port_cargo = list(STOCK_FUEL), list(STOCK_CARGO)
    return env


def test_construction_and_errors():
    from shippingenv_b200 import Environment, ShipMove, ActionType
    with pytest.raises(FileNotFoundError):
        Environment("does_not_exist.jpg")
    env = Environment(MAP, seed=0x5EED)
    assert env.np_game.shape == (100, 100) and env.np_game.dtype == int and env.fuel == 200 and env.cargo == 0
    with pytest.raises(Exception, match="No ports available"):
        env.step([ActionType.MOVE_SHIP, ShipMove.NORTH])
    with pytest.raises(ValueError, match="Coordinates not within map"):
        env.add_port([100, 5])
    z = golden("port_stock")
    for k, p in enumerate(PORTS):
        env.add_port(p)
        assert env.np_game[p[0], p[1]] == 2
    assert env.port_fuel == z["fuel_5eed"].tolist() and env.port_cargo == z["cargo_5eed"].tolist()
    with pytest.raises(TypeError):
        env.remove_port(0)                      # the reference's remove_port is broken the same way (:68)
    with pytest.raises(Exception, match="Cannot move without destination port"):
        env.ship_position = [41, 40]
        env.step([ActionType.MOVE_SHIP, ShipMove.NORTH])
    state = env.reset()
    assert state["ship"]["cargo"] == state["ship"]["fuel"] == 200          # observation quirk (:206)
    assert state["ship"]["position"] is env.port_positions[env.origin_port_index]   # aliasing (:241)
    assert env.np_game[env.ship_position[0], env.ship_position[1]] == 5
    for action, exc, msg in (
            ([ActionType.SELECT_PORT, 9], IndexError, "Port index is out of range"),
            ([ActionType.SELECT_PORT, env.origin_port_index], Exception, "Destination port must be different"),
            ([ActionType.TAKE_CARGO, 0], ValueError, "Invalid fuel amount"),
            ([ActionType.TAKE_FUEL, 999], ValueError, "Invalid fuel amount"),
            ([7, 1], ValueError, "Action category unknown"),
            ([ActionType.MOVE_SHIP, (1, 2, 3)], ValueError, "Move needs to be of format")):
        before = (list(env.ship_position), env.fuel, env.cargo, env._clock)
        with pytest.raises(exc, match=msg):
            env.step(action)
        assert before == (list(env.ship_position), env.fuel, env.cargo, env._clock)
    s, r, d, info = env.step([ActionType.TAKE_FUEL, 3])
    assert (r, d, info) == (0.05, False, {}) and env.fuel == 203 and isinstance(env.fuel, int)


def test_facade_replays_navigator_trace():
    """g4: take cargo, sail shortest routes, deliver, run out of fuel and keep stepping (no auto-reset)."""
    from shippingenv_b200 import ActionType
    z = golden("g4_navigator")
    T, N = z["actions"].shape
    for i in range(0, N, 3):
        env = make_env(int(z["seed"]), env_id=i)
        env.reset()
        assert (env.ship_position[0], env.ship_position[1], env.origin_port_index, env.destination_port_index) == \
            tuple(z["init_pos"][i]) and env._clock == z["init_clock"][i]
        for t in range(T):
            a = int(z["actions"][t, i])
            P = 5
            act = ([ActionType.MOVE_SHIP, [(0, -1), (-1, 0), (0, 1), (1, 0)][a]] if a < 4 else
                   [ActionType.SELECT_PORT, a - 4] if a < 4 + P else
                   [ActionType.TAKE_CARGO, a - 9] if a < 59 else [ActionType.TAKE_FUEL, a - 59])
            try:
                state, reward, done, _ = env.step(act)
                status = 0
            except Exception:
                status, reward, done = 1, 0.0, False
            assert (status != 0) == (z["status"][t, i] != 0), (i, t)
            assert bits(reward) == bits(z["reward"][t, i]) and done == bool(z["done"][t, i]), (i, t)
            assert (env.ship_position[0], env.ship_position[1], env.origin_port_index,
                    env.destination_port_index) == tuple(z["pos"][t, i]), (i, t)
            assert env.cargo == z["cargo"][t, i] and bits(env.fuel) == z["fuel"][t, i], (i, t)
            assert env._clock == z["clock"][t, i]


def test_facade_paints_np_game_like_the_reference():
    """g12: the reference's np_game (BOAT / TRAVEL paint, port cells painted over, reset wiping the trail) after reset
    and after each of the 700 navigator steps of four envs, compared cell by cell."""
    from paint_trace import replay_paint_trace
    assert replay_paint_trace() == 4 * 700


def test_facade_free_moves_and_nonsquare_maps_like_the_reference():
    """g14: moves of any (dx, dy) and non-square game sizes through the facade on the real kernels (the served step and,
    with SHIPENV_NO_SERVER semantics, the launch path are the same general transition): exception class + message or
    attributes, reward bits, done and np_game paint after each of 804 calls of the reference's API."""
    from facade_trace import replay_facade_trace
    assert replay_facade_trace() == 804


def test_c_fast_path_equals_the_python_path_on_the_kernels():
    """tests/fastpath_diff.py with the real server kernel behind both: 3,000 random API calls on two facades, one through
    the C fast path, one through Python only -- same results, exceptions and attributes after every call."""
    import shippingenv_b200.environment as E
    from fastpath_diff import diff
    from shippingenv_b200 import Environment
    Environment(os.path.join(GOLDEN, "default_map.npz"), seed=1)       # loads and binds the C module
    if E._fast is None:
        pytest.skip("the C fast path is not built or is disabled (SHIPENV_NO_FASTPATH / SHIPENV_NO_SERVER)")
    assert diff(calls=3000, seed=5) == 3000


def test_facade_crafted_states_and_policy_loop():
    """g6: attributes assigned from outside (as agents/mcts.py:199-208 does), then the reference's
    `a = sample_action(); try: step(a) except: continue` loop for policy rows."""
    z = golden("g6_crafted")
    T, N = z["actions"].shape
    for i in range(N):
        env = make_env(int(z["seed"]), env_id=i)
        x, y, o, d = (int(v) for v in z["init_pos"][i])
        env.ship_position = [x, y]
        env.origin_port_index = None if o == 255 else o
        env.destination_port_index = None if d == 255 else d
        env.cargo = int(z["init_cargo"][i])
        env.fuel = float(np.uint64(z["init_fuel"][i]).view(np.float64))
        env._clock = int(z["init_clock"][i])
        for t in range(T):
            a = int(z["actions"][t, i])
            want_status = int(z["status"][t, i])
            status, reward, done = 0, 0.0, False
            try:
                if a == -1:
                    if env.destination_port_index is None and env._get_current_port_idx() is None:
                        raise Exception("Cannot move without destination port")   # the reference spins here
                    for _ in range(32):
                        act = env.sample_action()
                        try:
                            _, reward, done, _ = env.step(act)
                            break
                        except Exception:
                            continue
                else:
                    from shippingenv_b200.preprocessing import map_action_to_env_action
                    _, reward, done, _ = env.step(map_action_to_env_action(a, env))
            except Exception:
                status = 1
            assert (status != 0) == (want_status != 0), (i, t, a, want_status)
            if want_status == 0:
                assert bits(reward) == bits(z["reward"][t, i]), (i, t)
                assert done == bool(z["done"][t, i])
            pos = (env.ship_position[0], env.ship_position[1],
                   255 if env.origin_port_index is None else env.origin_port_index,
                   255 if env.destination_port_index is None else env.destination_port_index)
            assert pos == tuple(z["pos"][t, i]), (i, t)
            assert env.cargo == z["cargo"][t, i] and bits(env.fuel) == z["fuel"][t, i], (i, t)
            if want_status != 0 and a == -1:
                env._word = 0     # the batched contract rolls the cell back on a failed policy step
            assert env._clock == z["clock"][t, i], (i, t)


def test_agent_call_patterns(default_map):
    """What the three reference agents do with an env: MCTS copies attributes onto a fresh instance and
    rolls out; DQN preprocesses states and probes _build_state / _get_current_port_idx; SARSA retries."""
    from shippingenv_b200 import Environment, ActionType
    from shippingenv_b200.preprocessing import (preprocess_state, discretize_state, get_action_space_size,
                                                map_action_to_env_action, get_state_size)
    env = make_env(123)
    assert get_state_size(env) == (1, 26) and get_action_space_size(env) == 259
    state = env.reset()
    obs = preprocess_state(state)
    assert obs.shape == (1, 26) and obs[0, 2] == obs[0, 3] == 200
    assert discretize_state(state)[0] == tuple(env.ship_position)
    # MCTS: _create_env_copy (agents/mcts.py:192-209) + _rollout (:211-238)
    returns = []
    for sim in range(4):
        copy = Environment(MAP, seed=99, env_id=sim)
        copy.port_positions = env.port_positions.copy()
        copy.port_fuel = env.port_fuel.copy()
        copy.port_cargo = env.port_cargo.copy()
        copy.ship_position = env.ship_position.copy()
        copy.cargo, copy.fuel = env.cargo, env.fuel
        copy.origin_port_index, copy.destination_port_index = env.origin_port_index, env.destination_port_index
        copy.np_game = env.np_game.copy()
        total, done, steps = 0.0, False, 0
        while not done and steps < 30:
            action = copy.sample_action()
            try:
                _, reward, done, _ = copy.step(action)
                total += reward
                steps += 1
            except Exception:
                continue
        returns.append(total)
        assert steps == 30 and copy.fuel < 200
    assert len(set(returns)) > 1                     # different simulations draw different streams
    assert env.fuel == 200 and env.cargo == 0        # the copies never touched the original
    # DQN: flat actions through map_action_to_env_action, "Destination port must be different" is string-matched
    n_ok = 0
    for a in range(259):
        try:
            env.step(map_action_to_env_action(a, env))
            n_ok += 1
        except Exception as e:
            assert str(e) in ("Destination port must be different from current one", "Invalid fuel amount",
                              "Not currently at port", "Move is out of range")
        assert env._build_state()["ship"]["fuel"] == env.fuel
    assert n_ok >= 8        # 4 moves + 4 selectable ports; takes are rejected once the ship left the port


def test_sample_action_lookahead_is_invisible():
    """step() draws the following sample_action() in the same launch (shipenv_step_sample).  The sequence of actions,
    states and rewards must be the one obtained with a launch per sample_action(), also when attributes are edited
    between the two calls (the look-ahead is then stale and must not be used) and across rejected steps."""
    from shippingenv_b200 import _native
    def run(lookahead, edits):
        env = make_env(77, env_id=5)
        env.reset()
        trace, launches0 = [], _native.launch_count()
        for t in range(300):
            if not lookahead:
                env._spec = None
            if t in edits:
                env.fuel = 150.5 if t % 2 else 33
                env.cargo = 7
            a = env.sample_action()
            if t % 7 == 3 and env._get_current_port_idx() is None:
                a = [4, 5] if t % 2 else [3, 5]           # TAKE_CARGO / TAKE_FUEL away from a port: rejected
            try:
                _, r, done, _ = env.step(a)
            except Exception as e:
                trace.append((t, repr(a), str(e)))
                continue
            trace.append((t, repr(a), bits(r), done, tuple(env.ship_position), env.cargo, bits(env.fuel), env._clock))
            if done:
                env.reset()
        return trace, _native.launch_count() - launches0
    edits = {40, 41, 90, 200}
    plain, n_plain = run(False, edits)
    fast, n_fast = run(True, edits)
    assert plain == fast
    assert any(len(row) == 3 for row in plain)            # rejected steps occurred (take at a port, off-map ...)
    assert n_fast < 0.62 * n_plain                        # about one launch per pair instead of two


def test_facade_mcts_extensions_match_batched_kernels():
    """Environment.rollout_returns / mcts_action are the R = 1 view of shipenv_rollout / shipenv_mcts."""
    from shippingenv_b200 import ActionType
    z = golden("g11_mcts")
    move = [(0, -1), (-1, 0), (0, 1), (1, 0)]
    for r in (0, 3, 7, 20):
        env = make_env(int(z["seed"]))
        x, y, o, d = (int(v) for v in z["root_pos"][r])
        env.ship_position, env.origin_port_index, env.destination_port_index = [x, y], o, d
        env.cargo = int(z["root_cargo"][r])
        env.fuel = float(np.uint64(z["root_fuel"][r]).view(np.float64))
        a = env.mcts_action(50, 100, root_id=int(z["root_id_base"]) + r)
        assert a == [ActionType.MOVE_SHIP, list(move[int(z["action_S50_L100"][r])])]
        before = (list(env.ship_position), env.fuel, env.cargo, env._clock)
        rets = env.rollout_returns(6, 25, root_id=1000)
        assert len(rets) == 6 and len(set(rets)) > 1
        assert before == (list(env.ship_position), env.fuel, env.cargo, env._clock)   # planning never moves the env


def test_step_server_lifecycle_and_equivalence():
    """step() is answered by the resident server kernel (shipenv_serve_step).  Two envs that share it, interleaved,
    against two envs stepped by one shipenv_step_sample launch per call: same states, rewards, dones, exceptions and
    look-ahead draws -- across server retirement (idle limit), an explicit stop, a device-wide synchronisation and a
    change of the world (the server restarts with the new one)."""
    import time
    import torch
    import shippingenv_b200.environment as E
    from shippingenv_b200 import ActionType

    def drive(use_server, log):
        E._USE_SERVER = use_server
        a, b = make_env(11, env_id=3), make_env(12, env_id=4)
        ctx = a._sync_world()
        ctx.serve_set_idle_us(300)
        a.reset(); b.reset()
        for t in range(400):
            for env in (a, b):
                act = env.sample_action() if t % 7 else [ActionType.MOVE_SHIP, (0, 1)]
                try:
                    s, r, d, _ = env.step(act)
                    log.append((t, bits(r), d, tuple(env.ship_position), env.cargo, bits(env.fuel), env._clock))
                    if d:
                        env.reset()
                except Exception as exc:
                    log.append((t, type(exc).__name__, str(exc)))
            if t == 50:
                time.sleep(0.01)                     # the server retires (idle limit 300 us) and comes back
            if t == 100 and use_server:
                ctx.serve_stop()
            if t == 150:
                torch.cuda.synchronize()             # returns: the server leaves after its idle limit
            if t == 200:                             # another world on the same context: one more port
                c = make_env(13, env_id=5, ports=PORTS + [[30, 30]], stock=False)
                c.reset()
                for _ in range(20):
                    try:
                        s, r, d, _ = c.step(c.sample_action())
                        log.append(("c", bits(r), d, tuple(c.ship_position), c.cargo, bits(c.fuel), c._clock))
                    except Exception as exc:
                        log.append(("c", type(exc).__name__))
        ctx.serve_set_idle_us(2000)

    try:
        served, launched = [], []
        drive(True, served)
        drive(False, launched)
    finally:
        E._USE_SERVER = os.environ.get("SHIPENV_NO_SERVER") != "1"
    assert len(served) > 800 and served == launched


REF = os.environ.get("SHIPPINGENV_REF", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "agents")),
                    reason="no reference checkout on this box ($SHIPPINGENV_REF or /root/reference)")
def test_unmodified_reference_agents_on_the_cuda_facade():
    """The reference's agents/mcts.py, agents/sarsa.py and agents/dqn.py, unmodified, import ``shipping`` from compat/
    and train on the drop-in facade with every launch on the GPU (the CPU suite runs the same script with the launches
    computed by the oracle).  Runs wherever a reference checkout travels with the GPU."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, os.path.join(ROOT, "tests"), REF])
    env["SHIPENV_AGENTS_ON"] = "cuda"
    env["SHIPPINGENV_REF"] = REF
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_agents.py")], capture_output=True,
                         text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert r["mcts"]["episodes"] >= 1 and r["sarsa"]["episodes"] >= 1 and r["dqn"]["episodes"] >= 1
    assert r["sarsa"]["q_entries"] > 10 and r["launches"] > 500
