#!/usr/bin/env python3
"""BASELINE config 4: 65,536 root states x 50 simulations x <=100-step random rollouts in one launch."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

ap = argparse.ArgumentParser()
ap.add_argument("--roots", type=int, default=65536)
ap.add_argument("--sims", type=int, default=50)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
R = args.roots
env = BatchedEnvironment(R, passable=z["passable"], seed=0x5EED, device="cuda:0", auto_reset=False)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
for t in range(10):   # roots = post-reset states advanced by 10 random legal moves
    env.step(torch.randint(0, 4, (R,), generator=gen, device="cuda", dtype=torch.int16))
out = env.rollout(args.sims, args.steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    out = env.rollout(args.sims, args.steps)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
steps = int(out["steps"].sum().item())
print(f"rollouts: {R} roots x {args.sims} sims x <={args.steps} steps: {ms:.3f} ms/launch, {steps/ms/1e6:.2f} G rollout-steps/s, "
      f"{R/ms*1e3:.3e} roots/s, mean return {out['returns'].mean().item():.4f}, status!=0: {int((out['status'] != 0).sum())}")
