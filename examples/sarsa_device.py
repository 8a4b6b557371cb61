#!/usr/bin/env python3
"""the reference's tabular SARSA agent (agents/sarsa.py) over N environments with everything on
the device -- `discretize_state` keys from `shipenv_discretize`, a dense Q-table in HBM, the agent's
"illegal action -> keep trying random ones" rule (`_try_action`, agents/sarsa.py:211-220) through the policy action.

Hyper-parameters are the reference's (utils/constants.py:40-46): alpha 0.1, gamma 0.9, epsilon 1.0 -> 1e-4 x 0.999
per finished episode.  The action set is `_get_possible_actions` (agents/sarsa.py:80-129): every port selection,
the four moves and, at a port, cargo / fuel amounts 1..min(20, stock), expressed in the flat action index of
utils/preprocessing.py:93-137 so that the table is [states, 4 + P + 250] float32 (9.3 GB for the default
100 x 100 map with 5 ports -- a size that only makes sense with the table resident in HBM).

Differences from the reference loop, both forced by running N envs in lock-step: the update of step t is made
one iteration later, when the action actually taken at t+1 is known (the reference draws `next_action` first and
executes it afterwards; same rule, same data); a transition that ended its episode (auto-reset) uses 0 for the
next Q-value, because the state observed after it already belongs to the next episode.

    python examples/sarsa_device.py --envs 4096 --steps 2000
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from shippingenv_b200 import BatchedEnvironment, unpack_bits

NOOP = 32767      # not an action: the step is rejected and the env left untouched (used for "no second try needed")


def train(envs=4096, steps=2000, seed=0, lr=0.1, gamma=0.9, eps_min=1e-4, eps_decay=0.999, max_steps=1000,
          device="cuda:0", map_npz=None, graph=False):
    torch.manual_seed(seed)
    z = np.load(map_npz or os.path.join(ROOT, "tests", "golden", "default_map.npz"))
    env = BatchedEnvironment(envs, passable=z["passable"], seed=seed, device=device, max_steps=max_steps,
                             auto_reset=True, track_stats=True)
    env.reset()
    W, H = env.passable.shape
    P, A, N = env.P, env.num_actions, envs
    n_states = W * H * 5 * 5 * (P + 1) * (P + 1)
    q = torch.zeros(n_states, A, dtype=torch.float32, device=device)
    ports = torch.tensor(env.ports, device=device)                                   # [P, 2]
    cargo_max = torch.tensor(env.port_cargo, device=device).clamp(max=20)            # agents/sarsa.py:118,125
    fuel_max = torch.tensor(env.port_fuel, device=device).clamp(max=20)
    amounts_c = torch.arange(50, device=device)
    amounts_f = torch.arange(200, device=device)
    ar = torch.arange(N, device=device)
    noop = torch.full((N,), NOOP, dtype=torch.int16, device=device)
    policy = torch.full((N,), -1, dtype=torch.int16, device=device)
    # the transition waiting for its next action: state index, action, reward, "episode ended", "is valid"
    prev_s = torch.zeros(N, dtype=torch.long, device=device)
    prev_a = torch.zeros(N, dtype=torch.long, device=device)
    prev_r = torch.zeros(N, device=device)
    prev_end = torch.zeros(N, dtype=torch.bool, device=device)
    prev_ok = torch.zeros(N, dtype=torch.bool, device=device)
    td_abs = torch.zeros((), device=device)

    def state_index():                      # discretize_state key -> row of the table
        k = env.discretize().long()
        cb, fb = k[:, 2].clamp(0, 4), k[:, 3].clamp(0, 4)
        return ((((k[:, 0] * H + k[:, 1]) * 5 + cb) * 5 + fb) * (P + 1) + (k[:, 4] + 1)) * (P + 1) + (k[:, 5] + 1)

    def possible_actions():                 # _get_possible_actions in the flat index space, bool [N, A]
        at = (env.ships[:, None, :2] == ports[None].to(torch.uint8)).all(2)          # [N, P]
        on_port = at.any(1)
        k = at.float().argmax(1)
        m = torch.zeros(N, A, dtype=torch.bool, device=device)
        m[:, :4 + P] = True
        m[:, 4 + P:4 + P + 50] = on_port[:, None] & (amounts_c[None] >= 1) & (amounts_c[None] <= cargo_max[k][:, None])
        m[:, 4 + P + 50:] = on_port[:, None] & (amounts_f[None] >= 1) & (amounts_f[None] <= fuel_max[k][:, None])
        return m

    def iteration():
        s = state_index()
        eps = torch.clamp(eps_decay ** env.stats[3], min=eps_min).float()            # one decay per finished episode
        greedy = q[s].masked_fill(~possible_actions(), -torch.inf).argmax(1)
        explore = torch.rand(N, device=device) < eps
        first = env.step(torch.where(explore, policy, greedy.to(torch.int16)), want_taken=True)
        r1, taken = first.reward.clone(), first.taken.long().clone()
        rejected = unpack_bits(first.error_bits, N).clone()
        end1 = unpack_bits(first.reset_bits, N).clone()
        # _try_action: a rejected action is replaced by random ones until one is accepted
        second = env.step(torch.where(rejected, policy, noop))
        reward = torch.where(rejected, second.reward, r1)
        ended = end1 | (rejected & unpack_bits(second.reset_bits, N))
        action = torch.where(explore, taken, greedy)                                 # what update() is given
        ok = action >= 0
        action = action.clamp(min=0)
        # SARSA update of the previous transition, now that its next action is known (agents/sarsa.py:161-190)
        cur = q[prev_s, prev_a]
        nxt = torch.where(prev_end, torch.zeros_like(cur), q[s, action])
        td = lr * (prev_r + gamma * nxt - cur)
        q[prev_s, prev_a] = torch.where(prev_ok, cur + td, cur)
        td_abs.copy_((td.abs() * prev_ok).sum() / prev_ok.sum().clamp(min=1))
        prev_s.copy_(s); prev_a.copy_(action); prev_r.copy_(reward); prev_end.copy_(ended); prev_ok.copy_(ok)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it, replay = 0, iteration
    if graph and steps > 8:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                iteration()
        torch.cuda.current_stream(device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            iteration()
        it, replay = 4, g.replay
    while it < steps:
        replay()
        it += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    visited = int((q != 0).any(1).sum())
    return {"envs": envs, "iterations": steps, "cuda_graph": replay is not iteration, "agent_steps_per_s": envs * steps / dt,
            "iterations_per_s": steps / dt, "table_gb": q.numel() * 4 / 1e9, "states_visited": visited,
            "mean_abs_td_last": float(td_abs), "epsilon": float(torch.clamp(eps_decay ** env.stats[3], min=eps_min)),
            "episodes": env.episode_summary()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    import json
    print(json.dumps(train(args.envs, args.steps, args.seed, graph=args.graph)))
