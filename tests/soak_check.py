#!/usr/bin/env python3
"""Test scaffolding (also a script): long randomized comparison of the CUDA path with the C oracle: N envs stepped T
times with a mixed action stream -- the four moves, the random policy, the whole flat action space, moves of any
(dx, dy), junk codes -- with auto-reset, alternating plain, chained and many-step launches; every reward (f64 bits),
status, taken action and mask of every step and the full state every few steps must be identical.

    python tests/soak_check.py [--envs 1048653] [--steps 256] [--seed 1]     (tests/test_gpu_parity.py runs a short one)
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import lib as orc
from shippingenv_b200 import BatchedEnvironment, unpack_bits

def soak(envs=(1 << 20) + 77, steps=256, seed=1, max_steps=50):
    """Returns the number of env-steps compared; raises AssertionError at the first difference."""
    class args:
        pass
    args.envs, args.steps, args.seed, args.max_steps = envs, steps, seed, max_steps
    z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
    n, T = args.envs, args.steps
    rng = np.random.default_rng(args.seed)
    base = int(rng.integers(0, 1 << 40))
    env = BatchedEnvironment(n, passable=z["passable"], seed=0xC0FFEE + args.seed, device="cuda:0", env_id_base=base,
                             max_steps=args.max_steps, auto_reset=True)
    o = orc.Oracle(z["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    env.reset(); o.reset(ships, clock, env.seed, base)
    n_act = env.num_actions
    checked = 0
    t0 = time.time()
    t = 0
    while t < T:
        mode = int(rng.integers(0, 4))                   # 0 plain, 1 chained, 2 step_many (4 steps), 3 lean plain
        k = 4 if mode == 2 else 1
        kind = rng.random((k, n))
        acts = rng.integers(0, 4, (k, n))
        acts = np.where(kind > 0.60, -1, acts)
        acts = np.where(kind > 0.75, rng.integers(0, n_act + 2, (k, n)), acts)
        free = -2 - ((rng.integers(-5, 6, (k, n)) + 90) * 181 + (rng.integers(-5, 6, (k, n)) + 90))
        acts = np.where(kind > 0.85, free, acts)
        acts = np.where(kind > 0.95, rng.integers(-32768, 32768, (k, n)), acts).astype(np.int16)
        a = torch.from_numpy(acts).cuda()
        if mode == 2:
            out = env.step_many(a, want_status=True, want_taken=True, want_reward64=True)
            got = [(out["reward64"][j].cpu().numpy(), out["status"][j].cpu().numpy(), out["taken"][j].cpu().numpy(),
                    out["done_bits"][j], out["reset_bits"][j], out["error_bits"][j]) for j in range(k)]
        elif mode == 3:
            r = env.step(a[0])
            got = [(None, None, None, r.done_bits.clone(), r.reset_bits.clone(), r.error_bits.clone(), r.reward.cpu().numpy())]
        else:
            r = env.step(a[0], want_status=True, want_taken=True, want_reward64=True, chained=(mode == 1))
            got = [(r.reward64.cpu().numpy(), r.status.cpu().numpy(), r.taken.cpu().numpy(), r.done_bits.clone(),
                    r.reset_bits.clone(), r.error_bits.clone())]
        for j in range(k):
            ref = o.step(ships, clock, acts[j], env.seed, base, args.max_steps, True)
            g = got[j]
            if g[0] is not None:
                assert (g[0].view(np.uint64) == ref["reward"].view(np.uint64)).all(), f"reward, step {t + j}"
                assert (g[1] == ref["status"]).all(), f"status, step {t + j}"
                ok = ref["status"] == 0
                assert (g[2][ok] == ref["taken"][ok]).all(), f"taken, step {t + j}"
            else:
                assert (g[6] == ref["reward"].astype(np.float32)).all(), f"reward32, step {t + j}"
            assert (unpack_bits(g[3], n).cpu().numpy() == ref["done"].astype(bool)).all(), f"done, step {t + j}"
            assert (unpack_bits(g[4], n).cpu().numpy() == ref["reset"].astype(bool)).all(), f"reset, step {t + j}"
            assert (unpack_bits(g[5], n).cpu().numpy() == (ref["status"] != 0)).all(), f"error, step {t + j}"
            checked += n
        t += k
        if t % 16 < k:
            gs, gc = env.state_numpy()
            assert gs.tobytes() == ships.tobytes() and (gc == clock).all(), f"state after step {t}"
    gs, gc = env.state_numpy()
    assert gs.tobytes() == ships.tobytes() and (gc == clock).all(), "final state"
    print(f"SOAK_OK {checked} env-steps compared bit for bit in {time.time() - t0:.0f} s "
          f"({n} envs x {t} steps, env_id_base {base}, episodes up to {int(clock.max() >> 16)})")

    return checked

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=(1 << 20) + 77)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-steps", type=int, default=50)
    args = ap.parse_args()
    soak(args.envs, args.steps, args.seed, args.max_steps)
