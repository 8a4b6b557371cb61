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

# the shape of MCTSAgent.choose_action in drop-in mode (agents/mcts.py:192-238): 50 simulations, each a fresh
# instance with the attributes copied onto it, then up to 100 x (sample_action, step)
e = make(); e.reset()
t = time.perf_counter(); steps = 0
for sim in range(50):
    c = Environment(MAP, seed=sim)
    c.port_positions = e.port_positions.copy(); c.port_fuel = e.port_fuel.copy(); c.port_cargo = e.port_cargo.copy()
    c.ship_position = e.ship_position.copy(); c.cargo, c.fuel = e.cargo, e.fuel
    c.origin_port_index, c.destination_port_index = e.origin_port_index, e.destination_port_index
    c.np_game = e.np_game.copy()
    done, k = False, 0
    while not done and k < 100:
        a = c.sample_action()
        try:
            _, r, done, _ = c.step(a); k += 1
        except Exception:
            continue
    steps += k
dt = time.perf_counter() - t
print(f"MCTS-shaped decision through the facade (50 copies x <=100 sample+step): {dt*1e3:.0f} ms ({steps} steps); "
      f"reference: 180-280 ms; Environment.mcts_action(): one launch")
t = time.perf_counter()
for _ in range(20): e.mcts_action()
print(f"Environment.mcts_action(): {(time.perf_counter()-t)/20*1e3:.2f} ms per decision")
t = time.perf_counter()
for _ in range(50): e.rollout_returns()
print(f"Environment.rollout_returns(): {(time.perf_counter()-t)/50*1e3:.2f} ms per 50 rollouts")
