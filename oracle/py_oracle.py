"""TEST INFRASTRUCTURE -- a scalar pure-Python restatement of the path, NOT part of the product.

Two uses: (1) a third, independently written implementation that the golden fixtures pin (small cases only --
``tests/test_oracle_golden.py``); (2) the "Python step loop" CPU figure ``bench.py`` reports next to the C
oracle: the reference is an interpreted per-step loop, and this is what such a loop costs on the box's cores
(the reference itself cannot travel to the GPU box).

State is a plain list ``[x, y, origin, dest, cargo, fuel]`` (``None`` for unset ports); ``rng`` is any object with
``random() / uniform(a, b) / randint(a, b) / choice(seq) / betavariate(a, b)`` -- the Philox shim for parity,
the stdlib ``random`` module for timing.  Citations: shipping/environment.py unless noted.
"""
import math

MOVES = ((0, -1), (-1, 0), (0, 1), (1, 0))          # N, E, S, W  (shipping/type.py:13-16)
X, Y, ORIGIN, DEST, CARGO, FUEL = range(6)


class World:
    def __init__(self, passable, ports, port_fuel, port_cargo, capacity=50, fuel0=200, cargo0=0):
        self.passable = passable                      # [W][H] truthy = not ground
        self.W, self.H = len(passable), len(passable[0])
        self.ports = [tuple(p) for p in ports]
        self.port_fuel, self.port_cargo = list(port_fuel), list(port_cargo)
        self.capacity, self.fuel0, self.cargo0 = capacity, fuel0, cargo0


class Rejected(Exception):
    def __init__(self, status):
        self.status = status


def port_at(w, s):                                    # :145-153
    pos = (s[X], s[Y])
    for k, p in enumerate(w.ports):
        if p == pos:
            return k
    return None


def reset(w, s, rng):                                 # :227-243
    s[CARGO], s[FUEL] = w.cargo0, w.fuel0
    n = len(w.ports)
    s[ORIGIN] = rng.randint(0, n - 1)
    s[DEST] = rng.randint(0, n - 1)
    while s[DEST] == s[ORIGIN]:
        s[DEST] = rng.randint(0, n - 1)
    s[X], s[Y] = w.ports[s[ORIGIN]]


def sample_action(w, s, rng):                         # :245-263 -> (category, value)
    cur = port_at(w, s)
    if s[DEST] is None and cur is not None:
        k = rng.randint(0, len(w.ports) - 1)
        while k == cur:
            k = rng.randint(0, len(w.ports) - 1)
        return 2, k
    if s[CARGO] == 0 and cur is not None:
        return 4, rng.randint(1, w.port_cargo[cur])
    if s[FUEL] == 0 and cur is not None:
        raise Rejected(9)                             # :163 indexes a float
    return 1, rng.choice((0, 1, 2, 3))


def step(w, s, rng, cat, val):
    """-> (reward, done); raises Rejected(status) before any draw or mutation, like the reference's exceptions."""
    if not w.ports:
        raise Rejected(8)
    if cat == 1:                                      # _move_ship :273-339
        if s[DEST] is None:
            raise Rejected(2)
        dx, dy = MOVES[val]
        x0, y0 = s[X], s[Y]
        nx, ny = x0 + dx, y0 + dy
        if not (0 <= nx < w.W and 0 <= ny < w.H):
            raise Rejected(1)
        cost = math.sqrt(dx * dx + dy * dy) * (1 + rng.uniform(-0.1, 0.1))
        r, done = 0, False
        if s[FUEL] < cost:
            r += -10
            done = True
        if not w.passable[nx][ny]:
            r += -5
        else:
            s[X], s[Y] = nx, ny
            s[FUEL] -= cost
            r += -0.0001
            r += -1
        tx, ty = w.ports[s[DEST]]
        before = math.sqrt((tx - x0) ** 2 + (ty - y0) ** 2)
        after = math.sqrt((tx - nx) ** 2 + (ty - ny) ** 2)
        r += 2 if before - after > 0 else -2
        if rng.random() <= s[CARGO] / w.capacity:     # :318-323 with _calculate_cargo_loss :169-200
            kind = rng.random()
            if s[CARGO] == 0 or kind < 0.1:
                loss = 0
            elif kind > 0.9:
                loss = s[CARGO]
            else:
                loss = int(rng.betavariate(2, 2) * s[CARGO])
            s[CARGO] -= loss
            r += loss * -3
        if (s[X], s[Y]) == (tx, ty):                  # :325-337
            r += s[CARGO] * 2
            s[CARGO] = 0
            s[ORIGIN] = s[DEST]
            s[DEST] = rng.randint(0, len(w.ports) - 1)
            while s[DEST] == s[ORIGIN]:
                s[DEST] = rng.randint(0, len(w.ports) - 1)
            r += 10
        return r, done
    if cat == 2:                                      # _select_port :265-271
        if not 0 <= val < len(w.ports):
            raise Rejected(3)
        if s[ORIGIN] == val:
            raise Rejected(4)
        s[DEST] = val
        return 0, False
    if cat in (3, 4):                                 # _take_fuel :350-357 / _take_cargo :341-348
        cur = port_at(w, s)
        if cur is None:
            raise Rejected(5)
        stock = w.port_fuel[cur] if cat == 3 else w.port_cargo[cur]
        if not 0 < val <= stock:
            raise Rejected(6)
        if cat == 3:
            s[FUEL] += val
        else:
            s[CARGO] += val
        return 0.05, False
    raise Rejected(7)


def random_policy_loop(w, rng, n_steps, max_steps=1000):
    """The reference's CPU-runnable loop (BASELINE configs[0]): sample_action + step, retry on rejection, reset on
    done or after max_steps.  Returns the number of accepted steps (== n_steps)."""
    s = [0, 0, None, None, 0, 0]
    reset(w, s, rng)
    done_steps = ep = 0
    while done_steps < n_steps:
        cat, val = sample_action(w, s, rng)
        try:
            _, done = step(w, s, rng, cat, val)
        except Rejected:
            continue
        done_steps += 1
        ep += 1
        if done or ep >= max_steps:
            reset(w, s, rng)
            ep = 0
    return done_steps
