#!/usr/bin/env python3
"""A/B of the headline step kernel between source trees on the same population.

    python tools/ab_bench.py [--root TREE] [--envs N] [--steps K] [--age A] [--groups G]

Imports shippingenv_b200 from TREE (default: this repo), ages every env group with A plain shipenv_step launches
(so any tree can run it), then prints us per launch for 1 group / G groups round-robin on one stream / G groups on
G streams."""
import argparse, os, sys
ap = argparse.ArgumentParser()
ap.add_argument("--root", default=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--age", type=int, default=1536)
ap.add_argument("--groups", type=int, default=4)
ap.add_argument("--tag", default="")
args = ap.parse_args()
ROOT = os.path.abspath(args.root)
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n, G, K = args.envs, args.groups, args.steps
dev = torch.device("cuda:0")
envs = [BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device=dev, env_id_base=g * n, max_steps=1000)
        for g in range(G)]
gen = torch.Generator(device=dev).manual_seed(1234)
block = torch.randint(0, 4, (32, n), generator=gen, device=dev, dtype=torch.int16)
for e in envs:
    e.reset()
    for t in range(args.age):
        e.step(block[t % 32])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(groups, streams):
    for g in range(groups):
        envs[g].use_stream(streams[g] if streams else None)
    for t in range(4 * groups):
        envs[t % groups].step(block[t % 32])
    torch.cuda.synchronize()
    ends = []
    e0.record()
    if streams:
        for s in streams[:groups]:
            s.wait_event(e0)
    for t in range(K):
        envs[t % groups].step(block[t % 32])
    if streams:
        for s in streams[:groups]:
            ev = torch.cuda.Event(enable_timing=True); ev.record(s); ends.append(ev)
    else:
        e1.record(); ends = [e1]
    torch.cuda.synchronize()
    for g in range(groups):
        envs[g].use_stream(None)
    return max(e0.elapsed_time(ev) for ev in ends) / K * 1e3


streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
tag = args.tag or os.path.basename(ROOT)
for name, groups, st in (("1 group, 1 stream", 1, None), (f"{G} groups, 1 stream", G, None),
                         (f"{G} groups, {G} streams", G, streams)):
    us = min(run(groups, st) for _ in range(3))
    print(f"{tag:16s} age {args.age:5d} {name:22s} {us:7.2f} us/launch  {n / us / 1e3:7.1f} G env-steps/s  "
          f"frac {n * 46 / us / 1e3 / 6552.6:.3f}", flush=True)
err = sum(int(e.error_bits.ne(0).sum().item()) for e in envs)
print(f"{tag:16s} age {args.age:5d} population: {err} error words in the last step")
