#!/usr/bin/env python3
"""Tuning builds (-DSHIPENV_SERVER_TIMING): how long the server kernel itself works on one request (globaltimer, from
seeing the request to having the answer in shared memory), next to the whole round trip seen by the host."""
import os, sys, time, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SHIPENV_NO_FASTPATH"] = "1"
from shippingenv_b200 import Environment
MAP = os.path.join(ROOT, "tests", "golden", "default_map.npz")
e = Environment(MAP, seed=1)
for p in [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]:
    e.add_port(list(p))
e.reset()
e.step(e.sample_action())
ctx = e._sync_world()
for name, act in (("move", 2), ("junk action", 9999)):
    ns = []
    t = time.perf_counter()
    for _ in range(5000):
        e._pack(act)
        ctx.serve_step(e._host_ptr, e._seed, e._env_id, 0, 0, 0)
        ns.append(struct.unpack_from("<I", e._buf, 88)[0])
    dt = (time.perf_counter() - t) / 5000
    ns.sort()
    print(f"{name:12s}: round trip {dt * 1e6:.2f} us (with _pack + ctypes), kernel's own work median {ns[len(ns) // 2]} ns, p90 {ns[int(len(ns) * .9)]} ns")
