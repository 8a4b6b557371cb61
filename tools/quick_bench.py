#!/usr/bin/env python3
"""Time the step kernel alone (2,097,152 envs, random moves) -- for A/B-ing tuning builds via SHIPENV_LIB."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--mode", default="moves")
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n = args.envs
env = BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device="cuda:0", max_steps=1000)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
acts = [torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.int16) for _ in range(8)]
if args.mode == "policy":
    env.step_policy()
    acts = [env._policy]          # the env's own policy tensor: step() then sets SHIPENV_F_ALL_POLICY
kw = dict(want_obs=True) if args.mode == "obs" else {}
for t in range(300):
    env.step(acts[t % len(acts)], **kw)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
K = args.steps
st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
torch.cuda.synchronize()
for t in range(K):
    flush.fill_(t & 255)
    st[t].record(); env.step(acts[t % len(acts)], **kw); en[t].record()
torch.cuda.synchronize()
ms = sorted(s.elapsed_time(e) for s, e in zip(st, en))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(K):
    env.step(acts[t % len(acts)], **kw)
e1.record(); torch.cuda.synchronize()
warm = e0.elapsed_time(e1) / K
mean = sum(ms) / K
print(f"{os.environ.get('SHIPENV_LIB', 'default'):40s} mode={args.mode} flushed mean {mean*1e3:7.2f} us  median {ms[K//2]*1e3:7.2f} us  "
      f"warm {warm*1e3:7.2f} us   -> {n/mean/1e6:8.1f} Gsteps/s flushed, {n/warm/1e6:8.1f} warm")
