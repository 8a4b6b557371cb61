#!/usr/bin/env python3
"""bench.py -- env-steps/s of the batched ShippingEnv step on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 2,097,152 envs
per launch per GPU (16M over 8), default map, 5 default ports, uniform random move actions, auto-reset on
done or after 1,000 steps, Philox seed 0x5EED, global env ids rank * 2^21 + i (weak scaling, no
collective on the step path).  One "step" = one shipenv_step launch over every env of the rank.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM (4 env groups
per GPU stepped round-robin so that the working set exceeds L2); `e2e` goes through the C-ABI call that takes
HOST buffers (shipenv_step_host): the step's actions come from pinned host memory and reward + done / reset /
error masks go back to it, copies inside the timed region.  `--impl reference` times the CPU oracle port (oracle/shipenv_oracle.c, the reference
is pure Python and cannot travel to the GPU box) on all host cores on the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 21
SEED = 0x5EED
MAX_STEPS = 1000
BYTES_PER_STEP = 46          # 16+4+2 read, 16+4+4 written, + 3 ballot bits
GROUPS = 4                   # env groups per GPU, stepped round-robin (working set > L2)
ACTION_POOL = 32             # distinct pre-generated action vectors, cycled
FLUSH_BYTES = 256 << 20      # > 126 MB L2
METRIC = "env-steps/sec (whole box)"


def world():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
    return z["passable"]


# ---------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def pin_to_gpu_numa_node(index):
    """Restrict this rank to the CPUs NVML reports as local to GPU `index` (first-touch then places its pinned
    host buffers on that NUMA node).  Returns the number of CPUs kept, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port on host threads (ctypes releases the GIL inside the C call)
# ---------------------------------------------------------------------------------------------
def cpu_steps_per_s(n_envs, budget_s, threads, min_steps=4):
    """Random-move auto-reset stepping of `n_envs` envs with the C oracle on `threads` threads."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import lib as orc
    from shippingenv_b200.batched import DEFAULT_PORTS, DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO
    o = orc.Oracle(world(), DEFAULT_PORTS, DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO)
    ships, clock = np.zeros(n_envs, orc.SHIP_DTYPE), np.zeros(n_envs, np.uint32)
    o.reset(ships, clock, SEED)
    rng = np.random.default_rng(1234)
    pool = [rng.integers(0, 4, n_envs).astype(np.int16) for _ in range(8)]
    bounds = np.linspace(0, n_envs, threads + 1).astype(np.int64)
    reward = np.zeros(n_envs, np.float64)
    done = np.zeros(n_envs, np.uint8)
    import ctypes as C
    L = orc.lib()

    def work(k, acts):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        L.orc_step(C.byref(o.w), C.c_void_p(ships.ctypes.data + 16 * lo), C.c_void_p(clock.ctypes.data + 4 * lo),
                   C.c_void_p(acts.ctypes.data + 2 * lo), C.c_void_p(reward.ctypes.data + 8 * lo),
                   C.c_void_p(done.ctypes.data + lo), None, None, None, C.c_uint64(SEED), C.c_uint64(lo),
                   C.c_int64(hi - lo), C.c_int(MAX_STEPS), C.c_int(1), C.c_uint32(0))

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: work(k, pool[0]), range(threads)))      # warm-up step
        t0 = time.perf_counter()
        steps = 0
        while steps < min_steps or time.perf_counter() - t0 < budget_s:
            list(ex.map(lambda k: work(k, pool[steps % len(pool)]), range(threads)))
            steps += 1
        dt = time.perf_counter() - t0
    return n_envs * steps / dt, steps, dt


def run_reference(args):
    """The reference arm: the path's CPU implementation on all host cores, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import ctypes as C
    from oracle import lib as orc
    from shippingenv_b200.batched import DEFAULT_PORTS, DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO
    threads = os.cpu_count() or 1
    n = ENVS_PER_GPU * args.gpus
    sample = min(n, ENVS_PER_GPU)          # each step: one pass over a 2M-env sample of the workload
    o = orc.Oracle(world(), DEFAULT_PORTS, DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO)
    ships, clock = np.zeros(sample, orc.SHIP_DTYPE), np.zeros(sample, np.uint32)
    o.reset(ships, clock, SEED)
    rng = np.random.default_rng(1234)
    pool = [rng.integers(0, 4, sample).astype(np.int16) for _ in range(8)]
    reward, done = np.zeros(sample, np.float64), np.zeros(sample, np.uint8)
    bounds = np.linspace(0, sample, threads + 1).astype(np.int64)
    L = orc.lib()

    def work(k, acts):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        L.orc_step(C.byref(o.w), C.c_void_p(ships.ctypes.data + 16 * lo), C.c_void_p(clock.ctypes.data + 4 * lo),
                   C.c_void_p(acts.ctypes.data + 2 * lo), C.c_void_p(reward.ctypes.data + 8 * lo),
                   C.c_void_p(done.ctypes.data + lo), None, None, None, C.c_uint64(SEED), C.c_uint64(lo),
                   C.c_int64(hi - lo), C.c_int(MAX_STEPS), C.c_int(1), C.c_uint32(0))

    steps = min(args.steps, 256)            # bounded: the whole run stays within a few minutes
    warm = min(args.warmup, 8)
    with ThreadPoolExecutor(threads) as ex:
        for t in range(warm):
            list(ex.map(lambda k: work(k, pool[t % 8]), range(threads)))
        t0 = time.perf_counter()
        for t in range(steps):
            list(ex.map(lambda k: work(k, pool[t % 8]), range(threads)))
        dt = time.perf_counter() - t0
    value = sample * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} envs x {steps} steps, oracle/shipenv_oracle.c (C restatement of the "
                                   "pure-Python reference, which cannot travel to this box) on all host threads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(gpus):
    return {"workload": f"{ENVS_PER_GPU} envs/GPU x {gpus} GPU(s) (BASELINE configs[2]: 16M envs over 8 GPUs), "
                        "default 100x100 map, 5 default ports, uniform random N/E/S/W actions (int16, pre-generated on "
                        "device), auto-reset on done or 1000 steps, Philox seed 0x5EED",
            "envs_per_launch": ENVS_PER_GPU, "env_groups_per_gpu": GROUPS, "max_steps": MAX_STEPS,
            "l2": "inputs larger than L2: 4 env groups of 2,097,152 envs per GPU stepped round-robin (216 MB working "
                  "set vs 126 MB L2); group g launches on stream g; timed from one start event to the latest of "
                  "the per-stream end events",
            "bytes_per_env_step": BYTES_PER_STEP}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from shippingenv_b200 import BatchedEnvironment, launch_count, unpack_bits

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if world_size != args.gpus:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = pin_to_gpu_numa_node(local)     # pinned host buffers of the e2e leg land next to the GPU's PCIe root
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as entry
    if local == 0:
        entry.build()                     # no-op when the shipped library matches the sources
    if world_size > 1:
        dist.barrier()

    n = args.envs
    G = args.groups
    # G independent env groups per GPU, stepped round-robin: one launch touches 46 B x n = 96 MB and the
    # next G-1 launches stream 3 x 96 MB through the 126 MB L2 before the group comes round again, so
    # every timed launch reads its inputs from HBM ("inputs larger than L2").
    envs = [BatchedEnvironment(n, passable=world(), seed=SEED, device=dev, env_id_base=(rank * G + g) * n,
                               max_steps=MAX_STEPS, auto_reset=True) for g in range(G)]
    for e in envs:
        e.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [torch.randint(0, 4, (n,), generator=gen, device=dev, dtype=torch.int16) for _ in range(ACTION_POOL)]
    K, W = args.steps, args.warmup

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident: K launches, group g on stream g, between CUDA events ----
    # The groups are independent, so each gets its own stream: the tail of one group's launch overlaps the head
    # of the next group's (within a group, consecutive launches still serialise -- and overlap their prologue
    # through programmatic dependent launch).  Timed on the device: one start event that every stream waits on,
    # one end event per stream, duration = start -> latest end.
    gstreams = [torch.cuda.Stream(device=dev) for _ in range(G)]

    def run_steps(count):
        for t in range(count):
            envs[t % G].step(pool[t % ACTION_POOL])       # env g launches on its own stream (use_stream below)

    for g, st in enumerate(gstreams):
        st.wait_stream(torch.cuda.current_stream(dev))
        envs[g].use_stream(st)
    run_steps(W)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    g_end = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
    barrier()
    launches0 = launch_count()
    with ClockSampler(local) as clocks:
        e0.record()
        for st in gstreams:
            st.wait_event(e0)
        t_host = time.perf_counter()
        run_steps(K)
        host_us = (time.perf_counter() - t_host) / K * 1e6     # host cost of one launch (must stay below the GPU time)
        for g, st in enumerate(gstreams):
            g_end[g].record(st)
        barrier()
    launches = launch_count() - launches0
    total_ms = max(e0.elapsed_time(e) for e in g_end)
    err = int(sum(int(e.error_bits.ne(0).sum().item()) for e in envs))

    # ---- the same K launches on ONE stream (strictly back to back), reported separately ----
    for e in envs:
        e.use_stream(None)
    barrier()
    e0.record()
    for t in range(K):
        envs[t % G].step(pool[t % ACTION_POOL])
    e1.record()
    barrier()
    one_stream_ms = e0.elapsed_time(e1)

    # ---- single group, L2 flushed by a 256 MiB fill between steps, one event pair per step (reported separately) ----
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    Kf = min(K, 256)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(Kf)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(Kf)]
    barrier()
    for t in range(Kf):
        flush.fill_(t & 255)
        starts[t].record()
        envs[0].step(pool[t % ACTION_POOL])
        stops[t].record()
    barrier()
    flushed_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops)) * K / Kf
    del flush

    # ---- end to end: actions from pinned host memory, reward + masks back to pinned host memory ----
    # one stream per group: copies of one group overlap the kernel of another
    nw = (n + 31) // 32
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    h_act = [torch.randint(0, 4, (n,), dtype=torch.int16).pin_memory() for _ in range(G)]
    h_out = [envs[g].host_buffers() for g in range(G)]      # reward + three masks, contiguous: one copy back per step
    h_reward = [o[0] for o in h_out]
    Ke = max(8, min(K, 128))

    def e2e_step(t):        # shipenv_step_host: host actions in, host reward + done / reset / error masks out
        g = t % G
        with torch.cuda.stream(streams[g]):
            envs[g].step_host(h_act[g], *h_out[g], chunks=1)   # the groups already overlap each other

    for t in range(2 * G):
        e2e_step(t)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
    t0.record()                                  # on the current stream; the group streams wait for it
    for st in streams:
        st.wait_event(t0)
    for t in range(Ke):
        e2e_step(t)
    for g, st in enumerate(streams):
        ends[g].record(st)
    barrier()
    e2e_ms = max(t0.elapsed_time(e) for e in ends)
    checksum = float(sum(float(h.sum()) for h in h_reward))

    # ---- the other device paths of the scope table, N = 1 only, a few launches each (reported as extras) ----
    extras = {}
    if world_size == 1 and not args.no_extras:
        def timed(fn, reps):
            fn()
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps
        g0 = envs[0]
        ms = timed(lambda: g0.step_policy(), 50)
        extras["policy_step"] = {"env_steps_per_s": n / (ms * 1e-3), "ms_per_launch": ms,
                                 "what": "action -1: on-device sample_action + retry (Environment.sample_action loop)"}
        ms = timed(lambda: g0.step(pool[0], want_obs=True), 50)
        extras["step_with_observation"] = {"env_steps_per_s": n / (ms * 1e-3), "ms_per_launch": ms,
                                           "bytes_per_env_step": 150, "hbm_gbs": n * 150 / (ms * 1e-3) / 1e9,
                                           "what": "fused preprocess_state rows, f32 [n, 26], TMA bulk stores"}
        ms = timed(lambda: g0.valid_action_bits(), 50)
        extras["valid_action_mask"] = {"envs_per_s": n / (ms * 1e-3), "ms_per_launch": ms}
        # BASELINE configs[1]: 4,096 envs -- launch-bound; eager launches vs one CUDA-graph replay of 256 steps
        small = BatchedEnvironment(4096, passable=world(), seed=SEED, device=dev, max_steps=MAX_STEPS)
        small.reset()
        sact = torch.randint(0, 4, (256, 4096), generator=gen, device=dev, dtype=torch.int16)
        ms = timed(lambda: [small.step(sact[t]) for t in range(256)], 3)
        graph = small.capture_steps(sact)
        msg = timed(lambda: graph.replay(), 10)
        extras["config_4096_envs"] = {"eager_env_steps_per_s": 4096 * 256 / (ms * 1e-3),
                                      "cuda_graph_env_steps_per_s": 4096 * 256 / (msg * 1e-3),
                                      "us_per_step_eager": ms * 1e3 / 256, "us_per_step_graph": msg * 1e3 / 256,
                                      "what": "BASELINE configs[1]: 4,096 envs on one GPU, random moves, auto-reset"}
        R, S, L = 65536, 50, 100                      # BASELINE configs[3]
        roots = g0.ships[:R].contiguous()
        holder = {}
        ms = timed(lambda: holder.update(out=g0.rollout(S, L, roots=roots)), 5)
        rsteps = int(holder["out"]["steps"].sum().item())
        extras["mcts_rollouts"] = {"roots": R, "simulations": S, "max_rollout_steps": L, "ms_per_launch": ms,
                                   "rollout_steps_per_s": rsteps / (ms * 1e-3), "roots_per_s": R / (ms * 1e-3),
                                   "what": "MCTSAgent._rollout x 50 per root in one kernel (reference: 5.6 roots/s/core)"}
        ms = timed(lambda: g0.mcts(S, L, roots=roots), 2)
        extras["mcts_choose_action"] = {"roots": R, "simulations": S, "max_rollout_steps": L, "ms_per_launch": ms,
                                        "decisions_per_s": R / (ms * 1e-3),
                                        "what": "MCTSAgent.choose_action (UCB1 tree + expansion + rollouts + backprop) per "
                                                "root in one kernel; reference: 0.18 s per decision per core"}

    # ---- max over ranks ----
    times = torch.tensor([total_ms, flushed_ms, e2e_ms, one_stream_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, flushed_ms, e2e_ms, one_stream_ms = times.tolist()

    if rank == 0:
        n_total = n * world_size
        value = n_total * K / (total_ms * 1e-3)
        kernel_ms = total_ms / K
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = n * BYTES_PER_STEP / (kernel_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world_size, "steps": K,
            "warmup": W, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world_size),
            "value_one_stream": n_total * K / (one_stream_ms * 1e-3),
            "value_l2_flushed_per_step_events": n_total * K / (flushed_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "shipenv::step_fast_kernel",
                         "algorithmic_bytes_per_launch": n * BYTES_PER_STEP,
                         "launch_us": kernel_ms * 1e3,
                         "how": "elapsed / K over the timed region; launches of independent env groups sit on "
                                "different streams and overlap at their boundaries",
                         "frac_one_stream": n * BYTES_PER_STEP / (one_stream_ms / K * 1e-3) / 1e9 / peak},
            "e2e": {"value": n_total * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s",
                    "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n + 12 * nw, "steps": Ke,
                    "what": "shipenv_step_host (BatchedEnvironment.step_host): the C-ABI call takes int16 actions from "
                            "page-locked host memory and delivers f32 reward + done/reset/error masks to page-locked "
                            "host memory; env groups round-robin on one stream each, so the copies of one group overlap the "
                            "kernel of another (chunks=1; a lone batch would let the call pipeline 1 Mi-env chunks)"},
            "gpu_launches": launches, "host_us_per_launch": host_us, "clocks": clocks.summary(), "cpus_pinned_to_gpu_numa_node": numa,
            "rejected_actions_last_step": err, "reward_checksum": checksum,
        }
        if extras:
            line["extras"] = extras
        if world_size == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)          # the CPU baseline uses every host core again
            threads = os.cpu_count() or 1
            v, s, dt = cpu_steps_per_s(1 << 18, args.cpu_seconds, threads)
            v1, s1, dt1 = cpu_steps_per_s(1 << 16, min(3.0, args.cpu_seconds), 1)
            import random as _random
            from oracle import py_oracle as _py
            zz = __import__("numpy").load(os.path.join(ROOT, "tests", "golden", "default_map.npz"))
            pw = _py.World(zz["passable"].astype(bool).tolist(), zz["ports"].tolist(), zz["stock_fuel"].tolist(),
                           zz["stock_cargo"].tolist())
            _random.seed(0)
            tp = time.perf_counter()
            npy = _py.random_policy_loop(pw, _random, 300_000)
            py_rate = npy / (time.perf_counter() - tp)
            line["cpu_baseline"] = {
                "value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                "single_thread_value": v1,
                "python_loop_value_per_core": py_rate,
                "python_loop_what": "oracle/py_oracle.py: the reference's sample_action+step loop (BASELINE configs[0]) as "
                                    "a scalar pure-Python loop on one core, stdlib random; the reference itself measured "
                                    "6.2e4 steps/s/core at survey time (BASELINE.md)",
                "sample": f"{1 << 18} envs x {s} steps ({dt:.1f} s) of the same workload with oracle/shipenv_oracle.c "
                          f"on {threads} threads; the reference itself is pure Python (~6e4 steps/s/core, BASELINE.md)"}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per launch")
    ap.add_argument("--groups", type=int, default=GROUPS, help="env groups per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            from oracle import lib as oracle_lib
            oracle_lib.build()
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
