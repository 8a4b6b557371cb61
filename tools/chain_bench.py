#!/usr/bin/env python3
"""A/B of the launch-boundary behaviour of the step kernel on a steady-state population (tuning builds via SHIPENV_LIB).

    python tools/chain_bench.py [--envs N] [--steps K] [--age A] [--groups G]

Prints us per launch for: one env group / G groups round-robin on one stream / G groups on G streams, each
unchained and chained (SHIPENV_F_CHAINED).  The population is aged with shipenv_step_many first so that episode ends,
auto-resets and off-map rejections occur at their steady rates."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--age", type=int, default=1536)
ap.add_argument("--groups", type=int, default=4)
ap.add_argument("--policy", action="store_true")
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n, G, K = args.envs, args.groups, args.steps
dev = torch.device("cuda:0")
envs = [BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device=dev, env_id_base=g * n, max_steps=1000)
        for g in range(G)]
gen = torch.Generator(device=dev).manual_seed(1234)
block = torch.randint(0, 4, (32, n), generator=gen, device=dev, dtype=torch.int16)
for e in envs:
    e.reset()
    for _ in range(args.age // 32):
        e.step_many(None if args.policy else block, steps=32, want_masks=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(groups, streams, chained):
    def issue(t):
        e = envs[t % groups]
        if args.policy:
            e._chain_ok = e._chain_ok and chained
            e.step_policy()
        else:
            e.step(block[t % 32], chained=chained)
    for g in range(groups):
        envs[g].use_stream(streams[g] if streams else None)
    for t in range(4 * groups):
        issue(t)
    torch.cuda.synchronize()
    ends = []
    e0.record()
    if streams:
        for s in streams[:groups]:
            s.wait_event(e0)
    for t in range(K):
        issue(t)
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
lib = os.path.basename(os.environ.get("SHIPENV_LIB", "default"))
res = {}
for name, groups, st in (("1 group, 1 stream", 1, None), (f"{G} groups, 1 stream", G, None),
                         (f"{G} groups, {G} streams", G, streams)):
    for chained in (False, True):
        us = min(run(groups, st, chained) for _ in range(3))
        res[(name, chained)] = us
        print(f"{lib:28s} {name:22s} {'chained' if chained else 'plain  '} {us:7.2f} us/launch  "
              f"{n / us / 1e3:7.1f} G env-steps/s  frac {n * 46 / us / 1e3 / 6552.6:.3f}", flush=True)
err = sum(int(e.error_bits.ne(0).sum().item()) for e in envs)
rst = sum(int(sum(bin(int(w) & 0xFFFFFFFF).count("1") for w in e.reset_bits[:2048].cpu().tolist())) for e in envs)
print(f"{lib:28s} population: {err} error words in last step, {rst} resets in the first 65,536 envs of each group")
