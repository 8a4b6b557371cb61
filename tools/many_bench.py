#!/usr/bin/env python3
"""shipenv_step_many: env-steps/s for BASELINE configs[1] (4,096 envs) and for the 2 Mi-env batch."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment

z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
dev = torch.device("cuda:0")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gen = torch.Generator(device=dev).manual_seed(1234)
for n, T, reps in ((4096, 256, 20), (4096, 1024, 10), (65536, 256, 10), (1 << 21, 32, 10), (1 << 21, 8, 20)):
    env = BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device=dev, max_steps=1000)
    env.reset()
    acts = torch.randint(0, 4, (T, n), generator=gen, device=dev, dtype=torch.int16)
    for policy in (False, True):
        for _ in range(2):
            env.step_many(None if policy else acts, steps=T)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            env.step_many(None if policy else acts, steps=T)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"step_many n={n:8d} T={T:5d} {'policy' if policy else 'moves '}: {ms*1e3:9.1f} us/launch, "
              f"{ms*1e3/T:7.3f} us/step, {n*T/ms/1e6:8.2f} G env-steps/s", flush=True)
