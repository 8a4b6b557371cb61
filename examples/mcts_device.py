#!/usr/bin/env python3
"""BASELINE config 4 as an agent: the reference's MCTSAgent (agents/mcts.py) playing N environments at once.
Every environment step asks `shipenv_mcts` for the decision of `MCTSAgent.choose_action` (UCB1 tree, 50
simulations, 100-step random rollouts) for all N ships in one launch, then steps them; roots the tree cannot
handle (no destination yet, agents/mcts.py:291-292) fall back to `sample_action`, as the reference does.

    python examples/mcts_device.py --envs 4096 --steps 200
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from shippingenv_b200 import BatchedEnvironment


def play(envs=4096, steps=200, simulations=50, rollout_steps=100, seed=0, device="cuda:0", planner=True, map_npz=None):
    z = np.load(map_npz or os.path.join(ROOT, "tests", "golden", "default_map.npz"))
    env = BatchedEnvironment(envs, passable=z["passable"], seed=seed, device=device, max_steps=1000, auto_reset=True,
                             track_stats=True)
    env.reset()
    policy = torch.full((envs,), -1, dtype=torch.int16, device=device)
    total = torch.zeros((), dtype=torch.float64, device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        if planner:
            # one tree per ship; (root id, simulation) selects the rollout stream, so every step uses fresh cells
            out = env.mcts(simulations, rollout_steps, root_id_base=(t + 1) * envs)
            actions = torch.where(out["action"] >= 0, out["action"], policy)
        else:
            actions = policy
        total += env.step(actions).reward.sum(dtype=torch.float64)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"envs": envs, "steps": steps, "planner": planner, "simulations": simulations if planner else 0,
            "decisions_per_s": envs * steps / dt, "mean_reward_per_step": float(total) / (envs * steps),
            "episodes": env.episode_summary()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--simulations", type=int, default=50)
    ap.add_argument("--rollout-steps", type=int, default=100)
    ap.add_argument("--random", action="store_true", help="the random policy instead of the planner, for comparison")
    args = ap.parse_args()
    print(json.dumps(play(args.envs, args.steps, args.simulations, args.rollout_steps, planner=not args.random)))
