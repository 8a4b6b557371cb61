import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment
z = np.load('/root/repo/tests/golden/default_map.npz')
n = 1 << 21
env = BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device="cuda:0")
env.reset()
for _ in range(3): env.step_policy() if hasattr(env, 'step_policy') else None
for _ in range(5): env.valid_action_bits()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): m = env.valid_action_bits()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"valid_action_bits: {ms*1e3:.1f} us per 2 Mi envs, {n/ms/1e6:.1f} G envs/s, {(36+16)*n/ms/1e6:.0f} GB/s, popcount {int(torch.sum(m != 0))}")
