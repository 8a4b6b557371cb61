"""Episode statistics across ranks.

The step path shards by env id and never communicates; the only collective of the whole system is
this one: each rank's kernel accumulates ``[sum return, sum return^2, sum length, episodes, max
return, min return]`` (``shipenv_step_io.stats``, float64[6]) and the ranks reduce those six
numbers with NCCL (gloo on CPU) when somebody asks for a report -- the counterpart of
``TrainingResult.get_summary`` (agents/base.py:119-137).
"""
import math

import torch


def shard_range(total_envs, rank, world_size):
    """Contiguous block of global env ids owned by ``rank``: [start, stop).  Env ids are global, so a
    given env's trajectory does not depend on the world size (the Philox cell carries the global id)."""
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def new_stats(device):
    return torch.tensor([0.0, 0.0, 0.0, 0.0, -math.inf, math.inf], dtype=torch.float64, device=device)


def reduce_stats(stats, group=None):
    """All-reduce a stats vector over the process group (no-op without one).  Returns a new tensor."""
    import torch.distributed as dist
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        sums, mx, mn = out[:4].clone(), out[4:5].clone(), out[5:6].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        out = torch.cat([sums, mx, mn])
    return out


def summarize(stats):
    """dict with the keys of ``TrainingResult.get_summary`` (agents/base.py:119-137) plus counts."""
    s = stats.detach().cpu().tolist()
    n = s[3]
    if n == 0:
        return {"total_episodes": 0, "mean_reward": 0.0, "std_reward": 0.0, "max_reward": 0.0, "min_reward": 0.0,
                "mean_length": 0.0, "total_steps": 0}
    mean = s[0] / n
    var = max(s[1] / n - mean * mean, 0.0)
    return {"total_episodes": int(n), "mean_reward": mean, "std_reward": math.sqrt(var), "max_reward": s[4],
            "min_reward": s[5], "mean_length": s[2] / n, "total_steps": int(s[2])}


def crosscheck(stats, group=None):
    """The all-reduced statistics next to the same numbers gathered rank by rank and summed on the host (bench.py and
    the gloo test use it to show that the one collective of the path computes what it should)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    red = reduce_stats(stats, group)
    if world > 1:
        every = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(every, stats, group=group)
        st = torch.stack(every).cpu()
    else:
        st = stats.detach().cpu()[None]
    want = torch.cat([st[:, :4].sum(0), st[:, 4:5].max(0).values, st[:, 5:6].min(0).values])
    r = red.detach().cpu()
    same = bool(r[3] == want[3] and r[4] == want[4] and r[5] == want[5] and r[2] == want[2]
                and abs(float(r[0] - want[0])) <= 1e-9 * max(1.0, abs(float(want[0]))))
    summary = summarize(red)
    return {"episodes": summary["total_episodes"], "mean_reward": summary["mean_reward"],
            "mean_length": summary["mean_length"], "reduced_equals_gathered": same}
