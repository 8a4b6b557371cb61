#!/usr/bin/env python3
"""BASELINE config 5: the reference's DQN loop (agents/dqn.py) with the environment, the observation
preprocessing and the valid-action mask on the device, feeding a stock torch Q-network.

Hyper-parameters are the reference's (utils/constants.py:29-38, agents/dqn.py:21-33,54): MLP
26 -> 128 -> 128 -> 259, batch 32, memory 2000, gamma 0.95, lr 1e-3, epsilon 1.0 -> 0.01 x 0.995,
target sync every 10 episodes.  Per agent step the reference rebuilds the state dict 259 times on the
host (`is_valid_action`, agents/dqn.py:189); here the step kernel writes the float32 observation rows into
a torch tensor, `shipenv_valid_actions` produces the mask, and nothing leaves the GPU.

    python examples/dqn_device.py --envs 1 --steps 2000        # the reference's loop shape: one env
    python examples/dqn_device.py --envs 4096 --steps 200      # the same agent over 4,096 envs
    python examples/dqn_device.py --envs 1 --steps 5000 --graph   # one CUDA-graph replay per agent step
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn as nn

from shippingenv_b200 import BatchedEnvironment, unpack_bits


class QNet(nn.Module):                      # agents/dqn.py:21-33
    def __init__(self, n_obs, n_act, hidden=128):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(n_obs, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                 nn.Linear(hidden, n_act))

    def forward(self, x):
        return self.net(x)


def train(envs=1, steps=2000, seed=0, memory=2000, batch=32, gamma=0.95, lr=1e-3, eps_min=0.01, eps_decay=0.995,
          max_steps=1000, target_every=10, device="cuda:0", map_npz=None, log=None, graph=False):
    """Runs ``steps`` agent iterations over ``envs`` environments.  All loop state (replay ring head and fill,
    epsilon, the current observation) lives in device tensors, so one iteration is a fixed sequence of launches:
    with ``graph=True`` it is captured once in a CUDA graph (environment step, mask, action selection, replay write,
    Q-update with a capturable Adam) and replayed, one host call per agent step."""
    torch.manual_seed(seed)
    z = np.load(map_npz or os.path.join(ROOT, "tests", "golden", "default_map.npz"))
    env = BatchedEnvironment(envs, passable=z["passable"], seed=seed, device=device, max_steps=max_steps,
                             auto_reset=True, track_stats=True)
    env.reset()
    n_obs, n_act = env.obs_width, env.num_actions
    q, target = QNet(n_obs, n_act).to(device), QNet(n_obs, n_act).to(device)
    target.load_state_dict(q.state_dict())
    opt = torch.optim.Adam(q.parameters(), lr=lr, capturable=True)
    cap = max(memory, envs)                 # replay ring on the device
    mem_s = torch.zeros(cap, n_obs, device=device); mem_s2 = torch.zeros(cap, n_obs, device=device)
    mem_a = torch.zeros(cap, dtype=torch.long, device=device)
    mem_r = torch.zeros(cap, device=device); mem_d = torch.zeros(cap, device=device)
    head = torch.zeros((), dtype=torch.long, device=device)
    filled = torch.zeros((), dtype=torch.long, device=device)
    eps = torch.ones((), device=device)
    loss_out = torch.zeros((), device=device)
    obs = env.observe().clone()
    ar = torch.arange(envs, device=device)

    def act_and_store():
        mask = env.valid_action_mask()                                   # [N, 259] bool, agents/dqn.py:189
        with torch.no_grad():
            qv = q(obs).masked_fill(~mask, -torch.inf)
            greedy = qv.argmax(1)
            rnd = torch.multinomial(mask.float(), 1).squeeze(1)          # random.choice(valid), :194-195
            explore = torch.rand(envs, device=device) <= eps
            action = torch.where(explore, rnd, greedy)
        res = env.step(action.to(torch.int16), want_obs=True)            # obs of the next state, on device
        done = unpack_bits(res.done_bits, envs)
        ok = ~unpack_bits(res.error_bits, envs)                          # e.g. select == origin: skipped, :309-311
        idx = (head + ar) % cap
        mem_s[idx], mem_a[idx], mem_r[idx], mem_s2[idx], mem_d[idx] = obs, action, res.reward, res.obs, done.float()
        head.add_(envs).remainder_(cap)
        filled.add_(envs).clamp_(max=cap)
        obs.copy_(torch.where(ok[:, None], res.obs, obs))

    def update():                                                        # update(), :206-245
        j = (torch.rand(batch, device=device) * filled).long()
        cur = q(mem_s[j]).gather(1, mem_a[j, None]).squeeze(1)
        with torch.no_grad():
            tgt = mem_r[j] + gamma * target(mem_s2[j]).max(1)[0] * (1 - mem_d[j])
        loss = nn.functional.mse_loss(cur, tgt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        eps.mul_(eps_decay).clamp_(min=eps_min)
        loss_out.copy_(loss.detach())

    def iteration():
        act_and_store()
        update()

    losses, episodes_seen = [], 0
    fill_iters = min(steps, -(-batch // envs))     # iterations before the replay memory holds one batch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(fill_iters):
        act_and_store()
    it = fill_iters
    replay = iteration
    if graph and it + 3 < steps:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):              # warm-up on a side stream, as torch's capture recipe asks
            for _ in range(3):
                iteration()
        torch.cuda.current_stream(device).wait_stream(side)
        it += 3
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            iteration()
        it += 1
        replay = g.replay
    while it < steps:
        replay()
        it += 1
        if it % 50 == 0:
            losses.append(loss_out.clone())
        if it % 100 == 0:                                                # target sync every `target_every` episodes
            n_ep = int(env.stats[3].item())
            if n_ep // target_every > episodes_seen // target_every:
                target.load_state_dict(q.state_dict())
            episodes_seen = n_ep
    if not losses:
        losses.append(loss_out.clone())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"envs": envs, "iterations": steps, "cuda_graph": bool(graph and replay is not iteration),
           "agent_steps_per_s": envs * steps / dt, "iterations_per_s": steps / dt,
           "final_epsilon": float(eps), "episodes": env.episode_summary(),
           "loss_first_last": [float(losses[0]), float(losses[-1])]}
    if log:
        log(out)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="capture one agent iteration in a CUDA graph and replay it")
    args = ap.parse_args()
    import json
    print(json.dumps(train(args.envs, args.steps, args.seed, graph=args.graph)))
