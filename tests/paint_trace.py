"""Test scaffolding (not a test module): replays tests/golden/g12_paint.npz -- the reference's own ``np_game`` after
``reset`` and after every step of four envs of the g4 navigator trace (BOAT on the ship, TRAVEL behind it, port cells
painted over; /root/reference/shipping/environment.py:228, 240, 303-304) -- through the drop-in ``Environment``.
Used in-process by tests/test_gpu_facade.py (real kernels) and, as a script, by tests/test_oracle_golden.py on a CPU box
(launches computed by the oracle, tests/facade_on_oracle.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
PORTS = [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]
STOCK_FUEL, STOCK_CARGO = [20, 14, 18, 12, 16], [10, 11, 13, 20, 18]


def replay_paint_trace():
    """Returns the number of (env, step) grids compared; raises AssertionError at the first difference."""
    from shippingenv_b200 import ActionType, Environment
    g4, g12 = np.load(os.path.join(GOLDEN, "g4_navigator.npz")), np.load(os.path.join(GOLDEN, "g12_paint.npz"))
    T = g4["actions"].shape[0]
    changes = g12["changes"]
    compared = 0
    for slot, i in enumerate(g12["env_index"].tolist()):
        env = Environment(os.path.join(GOLDEN, "default_map.npz"), seed=int(g4["seed"]), env_id=i)
        for p in PORTS:
            env.add_port(list(p))
        env.port_fuel, env.port_cargo = list(STOCK_FUEL), list(STOCK_CARGO)
        env.reset()
        want = g12["frame0"][slot].astype(np.int64)
        assert (np.asarray(env.np_game) == want).all(), f"env {i}: grid after reset"
        mine = changes[changes[:, 1] == slot]
        for t in range(T):
            a = int(g4["actions"][t, i])
            act = ([ActionType.MOVE_SHIP, [(0, -1), (-1, 0), (0, 1), (1, 0)][a]] if a < 4 else
                   [ActionType.SELECT_PORT, a - 4] if a < 9 else
                   [ActionType.TAKE_CARGO, a - 9] if a < 59 else [ActionType.TAKE_FUEL, a - 59])
            try:
                env.step(act)
            except Exception:
                assert g4["status"][t, i] != 0, (i, t)
            for _, _, x, y, v in mine[mine[:, 0] == t]:
                want[x, y] = v
            assert (np.asarray(env.np_game) == want).all(), f"env {i}: grid after step {t}"
            compared += 1
    return compared


if __name__ == "__main__":              # CPU: the facade's launches computed by the oracle
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import facade_on_oracle
    facade_on_oracle.install()
    print("PAINT_OK", replay_paint_trace())
