#!/usr/bin/env python3
"""Where the N = 1 step's time goes: the bare shipenv_serve_step round trip against the facade's Python around it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shippingenv_b200 import Environment
import shippingenv_b200.environment as E
MAP = os.path.join(ROOT, "tests", "golden", "default_map.npz")
e = Environment(MAP, seed=1)
for p in [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]:
    e.add_port(list(p))
e.reset()
e.step(e.sample_action())
ctx = e._sync_world()
e._pack(0)
N = 20000
t = time.perf_counter()
for _ in range(N):
    ctx.serve_step(e._host_ptr, e._seed, e._env_id, 0, 0, 0)
dt = time.perf_counter() - t
print(f"bare serve_step round trip (ctypes call included): {dt / N * 1e6:.2f} us")
import cProfile, pstats
pr = cProfile.Profile()
def loop():
    for _ in range(5000):
        a = e.sample_action()
        try:
            s, r, d, _ = e.step(a)
        except Exception:
            continue
        if d:
            e.reset()
t = time.perf_counter(); loop(); print(f"pair: {(time.perf_counter() - t) / 5000 * 1e6:.2f} us")
pr.enable(); loop(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
