#!/usr/bin/env python3
"""Per-CTA timeline of consecutive step launches (tuning build with -DSHIPENV_TRACE, see tools/build_variant.sh).

    SHIPENV_LIB=gpurun_variants/libshipenv_trace.so python tools/trace_launches.py [--chained] [--launches 6]

For every launch: when its CTAs started, when their producer got the go-ahead, when the consumers saw their first
tile, when they ran out of tiles and when the CTA left -- all relative to the first CTA start of the first launch --
and, per SM, the gap between a CTA leaving and the next one starting."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shippingenv_b200 import BatchedEnvironment, _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--age", type=int, default=640)
ap.add_argument("--groups", type=int, default=1)
ap.add_argument("--chained", action="store_true")
ap.add_argument("--dump", default="")
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
n, G = args.envs, args.groups
dev = torch.device("cuda:0")
envs = [BatchedEnvironment(n, passable=z["passable"], seed=0x5EED, device=dev, env_id_base=g * n, max_steps=1000)
        for g in range(G)]
gen = torch.Generator(device=dev).manual_seed(1234)
block = torch.randint(0, 4, (32, n), generator=gen, device=dev, dtype=torch.int16)
for e in envs:
    e.reset()
    for _ in range(args.age // 32):
        e.step_many(block, steps=32, want_masks=False)
for t in range(8 * G):
    envs[t % G].step(block[t % 32], chained=args.chained)
torch.cuda.synchronize()
lib = nat.lib()
fn = lib.shipenv_debug_trace
fn.argtypes, fn.restype = [C.c_void_p], C.c_int
buf = torch.zeros((1 << 16, 8), dtype=torch.int64, device=dev)
assert fn(buf.data_ptr()) == 0
K = args.launches * G
for t in range(K):
    envs[t % G].step(block[t % 32], chained=args.chained)
torch.cuda.synchronize()
assert fn(None) == 0
rec = buf.cpu().numpy().view(np.uint64)
rec = rec[rec[:, 0] != 0]
t0 = rec[:, 0].min()
start, go, first, last, end, w0, w1 = [(rec[:, k].astype(np.int64) - int(t0)) / 1e3 for k in range(7)]
meta = rec[:, 7]
blk, sm, cum = (meta & 0xFFFF).astype(int), ((meta >> 16) & 0xFFFF).astype(int), (meta >> 32).astype(np.int64)
# launches: CTAs sorted by start time, cut into groups of the grid size
grid = int(blk.max()) + 1
order = np.argsort(start, kind="stable")
print(f"{len(rec)} CTA records, grid {grid}, {len(rec) / grid:.2f} launches; times in us")
if args.dump:
    np.savez(args.dump, rec=rec)
q = lambda v: f"{np.min(v):7.2f} {np.percentile(v, 50):7.2f} {np.max(v):7.2f}"
print("launch |        CTA start (min med max) |   producer go-ahead |   first tile seen  |  out of tiles      |  CTA exit           | life med")
if args.chained:
    keys = sorted(set(cum.tolist()))
    groups = [np.where(cum == k)[0] for k in keys]
else:
    groups = [order[k * grid:(k + 1) * grid] for k in range(len(rec) // grid)]
for k, idx in enumerate(groups):
    print(f"{k:6d} | {q(start[idx])} | {q(go[idx])} | {q(first[idx])} | {q(last[idx])} | {q(end[idx])} | "
          f"{np.median(end[idx] - start[idx]):6.2f}   publish {np.median((w1 - w0)[idx]):5.2f}")
# per-SM slot gaps: sort CTAs of an SM by start; for each start, the latest exit before it among the 3 resident slots
gaps = []
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    ev = sorted([(end[i], 0) for i in idx] + [(start[i], 1) for i in idx])
    free = []
    for tm, kind in ev:
        if kind == 0:
            free.append(tm)
        elif free:
            gaps.append(tm - free.pop(0))
gaps = np.array(gaps)
if len(gaps):
    print(f"per-SM exit -> next start gap: median {np.median(gaps):.2f} us, p90 {np.percentile(gaps, 90):.2f}, max {gaps.max():.2f}  ({len(gaps)} pairs)")
span = end.max() - start.min()
print(f"span {span:.2f} us for {len(rec) / grid:.2f} launches = {span / (len(rec) / grid):.2f} us per launch")
