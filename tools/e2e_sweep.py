#!/usr/bin/env python3
"""End-to-end throughput of shipenv_step_host (host actions in, host reward + masks out) for G env groups of n envs on
G streams, for different chunk counts, to see what the pipelining inside the call buys."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--groups", type=int, default=4)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--chunks", type=int, default=0)
ap.add_argument("--classic", action="store_true", help="torch copies around shipenv_step instead of shipenv_step_host")
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n, G, K = args.envs, args.groups, args.steps
nw = (n + 31) // 32
envs = [BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device="cuda:0", env_id_base=g * n) for g in range(G)]
for e in envs: e.reset()
streams = [torch.cuda.Stream() for _ in range(G)]
h_act = [torch.randint(0, 4, (n,), dtype=torch.int16).pin_memory() for _ in range(G)]
h_out = [e.host_buffers() for e in envs]       # contiguous reward + masks: one copy back per step
h_rew = [o[0] for o in h_out]
h_m = [list(o[1:]) for o in h_out]
d_act = [torch.empty(n, dtype=torch.int16, device="cuda") for _ in range(G)]

def step(t):
    g = t % G
    with torch.cuda.stream(streams[g]):
        if args.classic:
            d_act[g].copy_(h_act[g], non_blocking=True)
            r = envs[g].step(d_act[g])
            h_rew[g].copy_(r.reward, non_blocking=True)
            for dst, src in zip(h_m[g], (r.done_bits, r.reset_bits, r.error_bits)):
                dst.copy_(src, non_blocking=True)
        else:
            envs[g].step_host(h_act[g], h_rew[g], *h_m[g], chunks=args.chunks)

for t in range(2 * G): step(t)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
t0.record()
for s in streams: s.wait_event(t0)
import time
w0 = time.perf_counter()
for t in range(K): step(t)
host = time.perf_counter() - w0
for g, s in enumerate(streams): ends[g].record(s)
torch.cuda.synchronize()
ms = max(t0.elapsed_time(e) for e in ends)
print(f"chunks={args.chunks} classic={args.classic} groups={G} envs={n}: "
      f"{n * K / (ms * 1e-3) / 1e9:.2f} G env-steps/s, {ms / K * 1e3:.0f} us per step, host {host / K * 1e6:.0f} us per step")
