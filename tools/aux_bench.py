#!/usr/bin/env python3
"""us per 2 Mi-env launch of the small kernels around the step (reset, observe, sample_action, valid-action mask,
discretize) with the bytes they move, against the measured HBM rate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n = 1 << 21
env = BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device="cuda:0")
env.reset()
for _ in range(8):
    env.step_policy()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


W = env.obs_width
for name, fn, nbytes in (("observe", env.observe, 16 + 4 * W), ("valid_action_bits", env.valid_action_bits, 16 + 36),
                         ("discretize", env.discretize, 16 + 12), ("sample_action", env.sample_action, 20 + 4),
                         ("reset", env.reset, 20)):
    ms = timed(fn)
    print(f"{name:18s} {ms * 1e3:7.1f} us per 2 Mi envs  {n / ms / 1e6:7.1f} G envs/s  {nbytes * n / ms / 1e6:6.0f} GB/s of 6553")
