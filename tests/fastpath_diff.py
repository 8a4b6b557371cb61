"""Test scaffolding (also a script): the C fast path of the facade (shippingenv_b200/csrc/facade_fast.c) against the
Python implementation it shortcuts.  Two ``Environment`` objects with the same seed receive the same random sequence of
API calls -- moves as lists / tuples / numpy integers / floats / bools, every action category with in- and out-of-range
values, malformed actions, ``sample_action``, ``reset``, attributes overwritten from outside the way the agents do
(agents/mcts.py:199-208) -- one through the C code (where it accepts the call), one through Python only; after every
call the result or the exception (class and message) and every attribute, including ``np_game``, must be the same.
The launches are computed by the oracle (tests/facade_on_oracle.py), so this runs on a CPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _state(env):
    return (list(env.ship_position), env.origin_port_index, env.destination_port_index, env.cargo, repr(env.fuel),
            type(env.fuel).__name__, env._clock, env._word, env.np_game.tobytes())


def diff(calls=6000, seed=11):
    import shippingenv_b200.environment as E
    from shippingenv_b200 import ActionType, Environment
    fast = E._fast
    assert fast is not None, "the C fast path is not loaded"
    rng = np.random.default_rng(seed)
    envs = []
    for _ in range(2):
        e = Environment(os.path.join(HERE, "golden", "default_map.npz"), seed=1234, env_id=3)
        for p in [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]:
            e.add_port(list(p))
        envs.append(e)
    # a third facade with another world on the same device: its calls replace the world the context holds
    intruder = Environment(os.path.join(HERE, "golden", "default_map.npz"), seed=99)
    for p in [[41, 40], [60, 22], [49, 72]]:
        intruder.add_port(list(p))
    intruder.reset()
    served = declined = 0

    def both(fn):
        nonlocal served, declined
        out = []
        for k, e in enumerate(envs):
            E._fast = fast if k == 0 else None
            try:
                out.append(("ok", fn(e)))
            except BaseException as exc:      # noqa: BLE001 -- the comparison is the point
                out.append((type(exc).__name__, str(exc)))
        E._fast = fast
        assert repr(out[0]) == repr(out[1]), (out[0], out[1])
        assert _state(envs[0]) == _state(envs[1])
        return out[0]

    both(lambda e: e.reset())
    moves = [(0, -1), (-1, 0), (0, 1), (1, 0)]
    for c in range(calls):
        u = rng.random()
        if u < 0.45:                                   # the agents' loop
            r = both(lambda e: e.sample_action())
            if r[0] == "ok":
                a = r[1]
                both(lambda e: e.step(a))
        elif u < 0.70:                                 # explicit unit moves in every spelling
            m = moves[int(rng.integers(0, 4))]
            k = int(rng.integers(0, 6))
            v = [list(m), tuple(m), [np.int64(m[0]), np.int64(m[1])], (float(m[0]), m[1]), [bool(m[0] == 1), m[1]],
                 np.array(m)][k]
            both(lambda e: e.step([ActionType.MOVE_SHIP, v]))
        elif u < 0.80:                                 # the other categories, in and out of range
            cat = int(rng.choice([2, 3, 4, 2, 3, 4, 0, 5, 99]))
            val = [int(rng.integers(-2, 8)), int(rng.integers(0, 30)), int(rng.integers(190, 300)), 9000, -5, 2.0, 2.5,
                   np.int64(3), True, None][int(rng.integers(0, 10))]
            both(lambda e: e.step([cat, val]))
        elif u < 0.85:                                 # malformed actions
            bad = [[1], [1, (0, 1), 2], (1, [0, 1]), [1, (0, 1, 0)], [1, 5], "ab", [1.0, (0, 1)], [True, (0, 1)],
                   [1, (2, 0)], [1, (0, 0)], [1, (-3, 4)]][int(rng.integers(0, 11))]
            both(lambda e: e.step(bad))
        elif u < 0.90:
            both(lambda e: e.reset())
        else:                                          # attributes assigned from outside
            k = int(rng.integers(0, 11))
            port = [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]][int(rng.integers(0, 5))]
            for e in envs:
                if k == 0:
                    e.fuel = [0, 3, 200, 0.5, 17.25, -2.0][int(c) % 6]
                elif k == 1:
                    e.cargo = [0, 1, 7, 49, 50, 60][int(c) % 6]
                elif k == 2:
                    e.destination_port_index = [None, 0, 1, 2, 3, 4][int(c) % 6]
                elif k == 3:
                    e.origin_port_index = [None, 0, 1, 2, 3, 4][int(c) % 6]
                elif k == 4:
                    e.ship_position = list(port)
                elif k == 5:
                    e.ship_position = tuple(port)          # a tuple: the C code declines, Python handles it
                elif k == 6:
                    e.ship_position = [int(rng.integers(0, 100)), int(rng.integers(0, 100))] if c % 2 else list(port)
                elif k == 7:
                    e.port_cargo = list(e.port_cargo)      # a new list object, same contents
                elif k == 8:
                    e.port_cargo[int(c) % 5] = 5 + int(c) % 16      # the stock changed in place: a new world
                elif k == 9:
                    e.port_fuel = [20, 14, 18, 12, 16 + int(c) % 3]
            if k == 10:                                # another facade steps in between: the device's world changes
                try:
                    intruder.step(intruder.sample_action())
                except Exception:                      # noqa: BLE001
                    pass
            if k == 6:                                 # both envs must see the same random position
                envs[1].ship_position = list(envs[0].ship_position)
            assert _state(envs[0]) == _state(envs[1])
    return calls


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    import facade_on_oracle
    facade_on_oracle.install(fast=True)
    print("FASTPATH_DIFF_OK", diff())
