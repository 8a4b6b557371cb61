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
import math
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
CHAINED_ONE_STREAM = True


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
                  "set vs 126 MB L2); group g launches on stream g (forked from / joined to the timing stream)",
            "bytes_per_env_step": BYTES_PER_STEP}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def timed_regions(torch, dev, issue, K, min_total_ms=50.0, max_reps=400, barrier=None):
    """Time the K-launch region `issue()` enqueues (a CUDA-graph replay, or K eager launches) until at least
    `min_total_ms` of device time has been measured; one CUDA-event pair per region on the current stream.
    Returns the list of region times in ms."""
    issue()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); issue(); e1.record()
    torch.cuda.synchronize(dev)
    est = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(max_reps, max(5, math.ceil(min_total_ms / est))))
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    if barrier:
        barrier()
    for r in range(reps):
        starts[r].record()
        issue()
        stops[r].record()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize(dev)
    return [s.elapsed_time(e) for s, e in zip(starts, stops)]


def copy_link_rates(torch, dev, n, G, iters=24):
    """Pinned-memory copy rates of this rank's host link with the e2e leg's own transfer sizes: G streams, each
    alternating one host-to-device copy of 2n bytes and one device-to-host copy of 4n + 12 ceil(n/32) bytes."""
    nw = (n + 31) // 32
    h_in = [torch.empty(2 * n, dtype=torch.uint8).pin_memory() for _ in range(G)]
    h_out = [torch.empty(4 * n + 12 * nw, dtype=torch.uint8).pin_memory() for _ in range(G)]
    d_in = [torch.empty(2 * n, dtype=torch.uint8, device=dev) for _ in range(G)]
    d_out = [torch.empty(4 * n + 12 * nw, dtype=torch.uint8, device=dev) for _ in range(G)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    out = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(dev)
        t0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
        t0.record()
        for g, st in enumerate(streams):
            st.wait_event(t0)
            with torch.cuda.stream(st):
                for _ in range(iters):
                    if mode != "d2h":
                        d_in[g].copy_(h_in[g], non_blocking=True)
                    if mode != "h2d":
                        h_out[g].copy_(d_out[g], non_blocking=True)
                ends[g].record(st)
        torch.cuda.synchronize(dev)
        ms = max(t0.elapsed_time(e) for e in ends)
        nbytes = G * iters * ((2 * n if mode != "d2h" else 0) + (4 * n + 12 * nw if mode != "h2d" else 0))
        out[mode] = nbytes / (ms * 1e-3) / 1e9
    return out


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
    K, W = args.steps, args.warmup
    # G independent env groups per GPU, stepped round-robin: one launch touches 46 B x n = 96 MB and the
    # next G-1 launches stream 3 x 96 MB through the 126 MB L2 before the group comes round again, so
    # every timed launch reads its inputs from HBM ("inputs larger than L2").
    envs = [BatchedEnvironment(n, passable=world(), seed=SEED, device=dev, env_id_base=(rank * G + g) * n,
                               max_steps=MAX_STEPS, auto_reset=True) for g in range(G)]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    block = torch.randint(0, 4, (ACTION_POOL, n), generator=gen, device=dev, dtype=torch.int16)
    pool = [block[k] for k in range(ACTION_POOL)]
    # Steady state before anything is timed (untimed, independent of --warmup): every group is aged with the fused
    # T-step kernel so that ships have diffused, fuel runs out at its steady rate (an episode lasts ~280 steps), the
    # deferred auto-resets and the off-map rejections are in the timed mix.
    for e in envs:
        e.reset()
        for _ in range(max(0, args.age) // ACTION_POOL):
            e.step_many(block, want_masks=False)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident: K launches, group g on stream g ----
    # The groups are independent, so each gets its own stream: the tail of one group's launch overlaps the head
    # of the next group's (within a group, consecutive launches still serialise -- and overlap their prologue
    # through programmatic dependent launch).  The K-launch region is captured ONCE in a CUDA graph (fork from the
    # capture stream to the group streams, join back), so that a timed region costs the host one graph launch and
    # host scheduling noise cannot reach the device timeline; the region is replayed until >= 50 ms have been timed
    # and the MEDIAN region time is reported (min / max alongside).
    gstreams = [torch.cuda.Stream(device=dev) for _ in range(G)]

    def issue_steps(count, chained=False):
        for t in range(count):
            envs[t % G].step(pool[t % ACTION_POOL], chained=chained)

    def capture(multi_stream, chained=False, count=None):
        for g in range(G):
            envs[g].use_stream(gstreams[g] if multi_stream else None)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            cur = torch.cuda.current_stream(dev)
            if multi_stream:
                for st in gstreams:
                    st.wait_stream(cur)
            issue_steps(K if count is None else count, chained)
            if multi_stream:
                for st in gstreams:
                    cur.wait_stream(st)
        for g in range(G):
            envs[g].use_stream(None)
        return graph

    for g, st in enumerate(gstreams):
        st.wait_stream(torch.cuda.current_stream(dev))
    issue_steps(W)                                    # W untimed warm-up steps (eager, current stream)
    torch.cuda.synchronize(dev)
    how_issue = "cuda_graph"
    launches0 = launch_count()
    try:
        graph_multi = capture(True)
        issue_multi = graph_multi.replay
    except Exception as exc:                           # capture not possible: eager launches, host cost reported
        how_issue = f"eager ({type(exc).__name__})"

        def issue_multi():
            cur = torch.cuda.current_stream(dev)
            for g, st in enumerate(gstreams):
                st.wait_stream(cur)
                envs[g].use_stream(st)
            issue_steps(K)
            for g, st in enumerate(gstreams):
                cur.wait_stream(st)
                envs[g].use_stream(None)
    launches_per_region = launch_count() - launches0 if how_issue == "cuda_graph" else K
    with ClockSampler(local) as clocks:
        t_host = time.perf_counter()
        region_ms = timed_regions(torch, dev, issue_multi, K, barrier=barrier)
        host_wall_s = time.perf_counter() - t_host
    total_ms = median(region_ms)
    err = int(sum(int(e.error_bits.ne(0).sum().item()) for e in envs))
    resets = int(sum(int(unpack_bits(e.reset_bits, n).sum().item()) for e in envs))

    # ---- the same K launches on ONE stream (strictly in stream order), reported separately ----
    # (a chain starts with one unchained launch per group: regions of at least 256 launches, reported per launch)
    Ko = max(K, 256)
    try:
        graph_one = capture(False, chained=args.chained, count=Ko)
        issue_one = graph_one.replay
    except Exception:
        def issue_one():
            issue_steps(Ko, args.chained)
    one_ms = median(timed_regions(torch, dev, issue_one, Ko, barrier=barrier)) * K / Ko

    # ---- single group, L2 flushed by a 256 MiB fill between steps, one event pair per step: the isolated launch ----
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    Kf = max(16, min(K, 128))
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(Kf)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(Kf)]
    barrier()
    for t in range(Kf):
        flush.fill_(t & 255)
        starts[t].record()
        envs[0].step(pool[t % ACTION_POOL])
        stops[t].record()
    barrier()
    isolated_ms = median([s.elapsed_time(e) for s, e in zip(starts, stops)])
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

    def e2e_region():
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        for t in range(Ke):
            e2e_step(t)
        for st in streams:
            cur.wait_stream(st)

    e2e_ms = median(timed_regions(torch, dev, e2e_region, Ke, min_total_ms=200.0, max_reps=40, barrier=barrier))
    checksum = float(sum(float(h.sum()) for h in h_reward))
    link = copy_link_rates(torch, dev, n, G)
    if world_size > 1:                                       # the ranks share the host: rates with all of them copying
        dist.barrier()
        link_all = copy_link_rates(torch, dev, n, G)
        dist.barrier()
    else:
        link_all = link

    # ---- episode statistics: the one collective of the path (NCCL all-reduce of 6 doubles), checked against a gather
    stats_check = None
    if world_size > 1:
        from shippingenv_b200.stats import crosscheck
        tracked = BatchedEnvironment(1 << 16, passable=world(), seed=SEED, device=dev, env_id_base=rank << 16,
                                     max_steps=40, auto_reset=True, track_stats=True)
        tracked.reset()
        for t in range(64):
            tracked.step(pool[t % ACTION_POOL][:1 << 16].contiguous())
        stats_check = crosscheck(tracked.stats)             # all-reduce over NCCL vs all-gather + host sum

    # ---- the other device paths of the scope table, N = 1 only, a few launches each (reported as extras) ----
    extras = {}
    if world_size == 1 and not args.no_extras:
        extras = run_extras(torch, dev, envs[0], pool, block, gen, n)

    # ---- max over ranks ----
    times = torch.tensor([total_ms, isolated_ms, e2e_ms, one_ms, min(region_ms), max(region_ms)], dtype=torch.float64,
                         device=dev)
    rates = torch.tensor([link_all["h2d"], link_all["d2h"], link_all["both"]], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(rates, op=dist.ReduceOp.SUM)
    total_ms, isolated_ms, e2e_ms, one_ms, region_min, region_max = times.tolist()

    if rank == 0:
        n_total = n * world_size
        value = n_total * K / (total_ms * 1e-3)
        kernel_ms = total_ms / K
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        algo = n * BYTES_PER_STEP
        achieved = algo / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src, ncu_us = None, None, None
        prof = os.path.join(ROOT, "profiles", "r2_step_kernel_traffic.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            traffic, traffic_src = pj.get("dram_bytes_per_launch"), pj.get("how")
            ncu_us = pj.get("ncu_us_per_launch_serialised")
        e2e_bytes = 2 * n + 4 * n + 12 * nw
        e2e_value = n_total * Ke / (e2e_ms * 1e-3)
        link_total = rates.tolist()
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world_size, "steps": K,
            "warmup": W, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world_size),
            "timing": {"how": f"{how_issue}: the K-step region (group g on stream g, forked from and joined to one stream) "
                              "is issued as one unit and timed with one CUDA-event pair; repeated until >= 50 ms were "
                              "timed; ms_per_step = median region time / K, max over ranks",
                       "regions": len(region_ms), "region_ms_median": total_ms, "region_ms_min": region_min,
                       "region_ms_max": region_max, "host_wall_s": host_wall_s,
                       "population": f"every group aged {args.age} steps (untimed) before the warm-up: "
                                     f"{resets} auto-resets and {err} mask words with rejected (off-map) actions in the "
                                     f"last timed step of the {G} groups"},
            "value_one_stream": n_total * K / (one_ms * 1e-3),
            "value_isolated_launch_l2_flushed": n_total / (isolated_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": "shipenv::step_fast_kernel<false,false,2,true>",
                         "algorithmic_bytes_per_launch": algo,
                         "launch_us": kernel_ms * 1e3,
                         "how": "algorithmic bytes / (median region time / K); launches of independent env groups sit "
                                "on different streams and overlap at their boundaries",
                         "frac_one_stream": algo / (one_ms / K * 1e-3) / 1e9 / peak,
                         "one_stream_how": f"max(K, 256) = {max(K, 256)} launches in ONE stream, groups round-robin"
                                           + (", SHIPENV_F_CHAINED (pre-generated actions)" if args.chained else ""),
                         "traffic_over_algorithmic": traffic / algo if traffic else None,
                         "frac_ncu_serialised": algo / (ncu_us * 1e-6) / 1e9 / peak if ncu_us else None,
                         "ncu_serialised_how": "mean gpu__time_duration of the same capture (kernels serialised by the "
                                               "profiler: every launch pays its own ramp and tail)",
                         "frac_isolated": algo / (isolated_ms * 1e-3) / 1e9 / peak,
                         "isolated_how": "one launch alone between CUDA events, L2 flushed by a 256 MiB fill before it "
                                         "(median of per-launch times): ramp and tail fully exposed"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s",
                    "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n + 12 * nw, "steps": Ke,
                    "what": "shipenv_step_host (BatchedEnvironment.step_host): the C-ABI call takes int16 actions from "
                            "page-locked host memory and delivers f32 reward + done/reset/error masks to page-locked "
                            "host memory; env groups round-robin on one stream each, so the copies of one group overlap the "
                            "kernel of another (chunks=1; a lone batch would let the call pipeline 1 Mi-env chunks)",
                    "roofline": {"bound": "host link (pinned-memory copies)", "unit": "GB/s",
                                 "achieved": e2e_value * e2e_bytes / n / 1e9,
                                 "peak": link_total[2],
                                 "frac": (e2e_value * e2e_bytes / n / 1e9) / link_total[2] if link_total[2] else None,
                                 "peak_how": "measured in this run: all ranks at once copy the e2e leg's own transfer "
                                             "sizes in both directions on the same number of streams, no kernels",
                                 "h2d_only_gbs": link_total[0], "d2h_only_gbs": link_total[1],
                                 "rank0_alone_both_gbs": link["both"]}},
            "gpu_launches": launches_per_region, "timed_regions": len(region_ms),
            "clocks": clocks.summary(), "cpus_pinned_to_gpu_numa_node": numa,
            "rejected_actions_last_step": err, "auto_resets_last_step": resets, "reward_checksum": checksum,
        }
        if stats_check:
            line["episode_stats_nccl"] = stats_check
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


def run_extras(torch, dev, g0, pool, block, gen, n):
    """The other device paths of SURVEY 8 on one GPU, a few launches each."""
    import importlib.util
    from shippingenv_b200 import BatchedEnvironment
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    extras = {}

    def timed(fn, reps):
        # warm-up: at least three calls (the caching allocator then has the output buffers of two calls) and at least
        # 100 ms of GPU work -- after host-side set-up the SM clock has dropped to idle and takes tens of
        # milliseconds to come back (measured: 120 -> 330 -> 990 -> 1965 MHz)
        t_warm, calls = time.perf_counter(), 0
        while calls < 3 or time.perf_counter() - t_warm < 0.1:
            fn()
            calls += 1
            if calls % 8 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: g0.step_policy(), 50)
    extras["policy_step"] = {"env_steps_per_s": n / (ms * 1e-3), "ms_per_launch": ms,
                             "what": "action -1: on-device sample_action + retry (Environment.sample_action loop)"}
    ms = timed(lambda: g0.step(pool[0], want_obs=True), 50)
    extras["step_with_observation"] = {"env_steps_per_s": n / (ms * 1e-3), "ms_per_launch": ms,
                                       "bytes_per_env_step": 150, "hbm_gbs": n * 150 / (ms * 1e-3) / 1e9,
                                       "what": "fused preprocess_state rows, f32 [n, 26], TMA bulk stores"}
    ms = timed(lambda: g0.valid_action_bits(), 50)
    extras["valid_action_mask"] = {"envs_per_s": n / (ms * 1e-3), "ms_per_launch": ms}
    # the fused T-step entry point on the headline batch: 6 + 40/T bytes per env-step, bound by instruction issue
    T = block.shape[0]
    ms = timed(lambda: g0.step_many(block), 5)
    extras["step_many_2Mi_envs"] = {
        "env_steps_per_s": n * T / (ms * 1e-3), "ms_per_launch": ms, "steps_per_launch": T,
        "roofline": {"bound": "issue", "algorithmic_bytes_per_env_step": 6 + 40 / T,
                     "hbm_gbs": n * T * (6 + 40 / T) / (ms * 1e-3) / 1e9},
        "what": "shipenv_step_many: T steps per launch, ship record and clock in registers; never part of the 46 B headline"}
    # BASELINE configs[1]: 4,096 envs -- one launch per step is launch-bound; eager vs CUDA graph vs shipenv_step_many
    small = BatchedEnvironment(4096, passable=world(), seed=SEED, device=dev, max_steps=MAX_STEPS)
    small.reset()
    sact = torch.randint(0, 4, (256, 4096), generator=gen, device=dev, dtype=torch.int16)
    ms = timed(lambda: [small.step(sact[t]) for t in range(256)], 3)
    graph = small.capture_steps(sact)
    msg = timed(lambda: graph.replay(), 10)
    msm = timed(lambda: small.step_many(sact), 20)
    extras["config_4096_envs"] = {"eager_env_steps_per_s": 4096 * 256 / (ms * 1e-3),
                                  "cuda_graph_env_steps_per_s": 4096 * 256 / (msg * 1e-3),
                                  "step_many_env_steps_per_s": 4096 * 256 / (msm * 1e-3),
                                  "us_per_step_eager": ms * 1e3 / 256, "us_per_step_graph": msg * 1e3 / 256,
                                  "us_per_step_step_many": msm * 1e3 / 256,
                                  "what": "BASELINE configs[1]: 4,096 envs on one GPU, random moves, auto-reset; 256 steps as "
                                          "256 launches, as one graph replay, and as one shipenv_step_many launch"}
    # BASELINE configs[3]: roots = env ids 0..R-1 after reset, advanced by 10 random legal moves
    R, S, L = 65536, 50, 100
    rootenv = BatchedEnvironment(R, passable=world(), seed=SEED, device=dev, env_id_base=0, auto_reset=False)
    rootenv.reset()
    rgen = torch.Generator(device=dev).manual_seed(4321)
    rejected = 0
    for _ in range(10):
        r = rootenv.step(torch.randint(0, 4, (R,), generator=rgen, device=dev, dtype=torch.int16))
        rejected += int(r.error_bits.ne(0).sum().item())
    roots = rootenv.ships.clone()
    holder = {}
    ms = timed(lambda: holder.update(out=rootenv.rollout(S, L, roots=roots)), 10)
    rsteps = int(holder["out"]["steps"].sum().item())
    extras["mcts_rollouts"] = {"roots": R, "simulations": S, "max_rollout_steps": L, "ms_per_launch": ms,
                               "rollout_steps_per_s": rsteps / (ms * 1e-3), "roots_per_s": R / (ms * 1e-3),
                               "roots_how": f"fresh reset of env ids 0..{R - 1}, then 10 uniform random N/E/S/W moves "
                                            f"({rejected} mask words with rejected moves)",
                               "what": "MCTSAgent._rollout x 50 per root in one kernel (reference: 5.6 roots/s/core)"}
    ms = timed(lambda: rootenv.mcts(S, L, roots=roots), 5)
    extras["mcts_choose_action"] = {"roots": R, "simulations": S, "max_rollout_steps": L, "ms_per_launch": ms,
                                    "decisions_per_s": R / (ms * 1e-3),
                                    "what": "MCTSAgent.choose_action (UCB1 tree + expansion + rollouts + backprop) per "
                                            "root in one kernel; reference: 0.18 s per decision per core"}
    # BASELINE configs[4]: the reference's DQN loop with env + observation + valid-action mask on the device
    try:
        spec = importlib.util.spec_from_file_location("dqn_device", os.path.join(ROOT, "examples", "dqn_device.py"))
        dqn = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(dqn)
        runs = {}
        for name, envs_, steps_, graph_ in (("1_env_eager", 1, 300, False), ("1_env_cuda_graph", 1, 1500, True),
                                            ("4096_envs_eager", 4096, 100, False), ("4096_envs_cuda_graph", 4096, 300, True)):
            out = dqn.train(envs=envs_, steps=steps_, graph=graph_, device=str(dev))
            runs[name] = {"agent_steps_per_s": out["agent_steps_per_s"], "iterations_per_s": out["iterations_per_s"],
                          "cuda_graph": out["cuda_graph"]}
        extras["dqn_loop"] = dict(runs, what="BASELINE configs[4] (examples/dqn_device.py): reference hyper-parameters "
                                             "(batch 32, memory 2000, MLP 26-128-128-259), env step + fused observation + "
                                             "valid-action mask on the device feeding a stock torch Q-network and Adam; "
                                             "the reference runs ~220 agent steps/s on CPU")
    except Exception as exc:                          # an example failing must not take the headline down
        extras["dqn_loop"] = {"error": f"{type(exc).__name__}: {exc}"}
    return extras


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per launch")
    ap.add_argument("--groups", type=int, default=GROUPS, help="env groups per GPU")
    ap.add_argument("--age", type=int, default=1536, help="untimed steps every env group is aged by before timing")
    ap.add_argument("--chained", action="store_true", default=CHAINED_ONE_STREAM,
                    help="one-stream leg: SHIPENV_F_CHAINED launches")
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
