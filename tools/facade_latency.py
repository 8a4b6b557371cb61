#!/usr/bin/env python3
"""Latency of the N=1 drop-in facade (construct / reset / sample_action / step), for INTEGRATION.md."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shippingenv_b200 import Environment
MAP = os.path.join(ROOT, "tests", "golden", "default_map.npz")
PORTS = [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]

def make():
    e = Environment(MAP, seed=1)
    for p in PORTS: e.add_port(list(p))
    return e

e = make(); e.reset()
t = time.perf_counter()
for _ in range(200): make()
print(f"construct + 5 add_port: {(time.perf_counter()-t)/200*1e6:.0f} us")
t = time.perf_counter(); n = 0
for _ in range(3000):
    a = e.sample_action()
    try:
        s, r, d, _ = e.step(a); n += 1
    except Exception:
        continue
    if d: e.reset()
dt = time.perf_counter() - t
print(f"sample_action + step: {dt/3000*1e6:.0f} us per pair  ({n/dt:.0f} accepted steps/s)")
t = time.perf_counter()
for _ in range(2000):
    try: e.step([1, (0, 1)])
    except Exception: pass
print(f"step only: {(time.perf_counter()-t)/2000*1e6:.0f} us")
