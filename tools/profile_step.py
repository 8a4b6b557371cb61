#!/usr/bin/env python3
"""A short run of the bench workload (2,097,152 envs, random moves, auto-reset) for ncu.

    python tools/profile_step.py [--envs N] [--steps K] [--mode moves|policy|obs|rollout|mcts]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from shippingenv_b200 import BatchedEnvironment

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--mode", default="moves")
ap.add_argument("--groups", type=int, default=1, help="moves: env groups stepped round-robin (working set > L2 from 2)")
ap.add_argument("--chained", action="store_true")
args = ap.parse_args()

z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
env = BatchedEnvironment(args.envs, passable=z["passable"], seed=0x5EED, device="cuda:0", max_steps=1000)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
acts = [torch.randint(0, 4, (args.envs,), generator=gen, device="cuda", dtype=torch.int16) for _ in range(4)]
# age the population so that resets, blocked moves and edge rejections occur at their steady rates
others = [BatchedEnvironment(args.envs, passable=z["passable"], seed=0x5EED, device="cuda:0", max_steps=1000,
                             env_id_base=g * args.envs) for g in range(1, args.groups)]
for o in others:
    o.reset()
for t in range(300):
    for o in others:
        o.step(acts[t % 4])
    if args.mode == "policy":
        env.step_policy()         # the random policy picks cargo up: cargo-loss draws, deliveries, takes at their rates
    else:
        env.step(acts[t % 4])
torch.cuda.synchronize()
if args.mode == "mcts":
    roots = env.ships[: 1 << 16].contiguous()
    for t in range(args.steps):
        env.mcts(50, 100, roots=roots)
elif args.mode == "rollout":
    roots = env.ships[: 1 << 16].contiguous()
    for t in range(args.steps):
        env.rollout(50, 100, roots=roots)
else:
    for t in range(args.steps):
        if args.mode == "moves":
            ([env] + others)[t % args.groups].step(acts[t % 4], chained=args.chained)
        elif args.mode == "policy":
            env.step_policy()
        elif args.mode == "obs":
            env.step(acts[t % 4], want_obs=True)
torch.cuda.synchronize()
print("ok", args.mode, args.envs, args.steps)
