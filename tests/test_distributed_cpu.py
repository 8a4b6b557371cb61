"""CPU, world_size 2 over gloo: the only collective of the system (episode-statistics reduction) and the
rank -> env-id sharding that bench.py and BatchedEnvironment use under torchrun."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shippingenv_b200.stats import crosscheck, new_stats, reduce_stats, shard_range, summarize
    lo, hi = shard_range(10, rank, world)
    # episode returns owned by this rank: env id e contributes return e - 4.5, length 100 + e
    st = new_stats("cpu")
    for e in range(lo, hi):
        r = e - 4.5
        st[0] += r; st[1] += r * r; st[2] += 100 + e; st[3] += 1
        st[4] = max(st[4].item(), r); st[5] = min(st[5].item(), r)
    red = reduce_stats(st)
    check = crosscheck(st)                 # what bench.py prints at N > 1: all-reduce next to all-gather + host sum
    if rank == 0:
        torch.save({"summary": summarize(red), "local": summarize(st), "range": (lo, hi), "check": check}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    s = res["summary"]
    assert res["range"] == (0, 5) and res["local"]["total_episodes"] == 5
    assert s["total_episodes"] == 10 and s["total_steps"] == sum(100 + e for e in range(10))
    assert s["max_reward"] == 4.5 and s["min_reward"] == -4.5 and abs(s["mean_reward"]) < 1e-12
    assert abs(s["std_reward"] - (sum((e - 4.5) ** 2 for e in range(10)) / 10) ** 0.5) < 1e-12
    assert s["mean_length"] == 104.5
    assert res["check"] == {"episodes": 10, "mean_reward": s["mean_reward"], "mean_length": 104.5,
                            "reduced_equals_gathered": True}


def test_reduce_without_process_group_is_identity():
    sys.path.insert(0, ROOT)
    from shippingenv_b200.stats import new_stats, reduce_stats, summarize
    st = new_stats("cpu")
    assert summarize(st)["total_episodes"] == 0
    st[:4] = torch.tensor([6.0, 14.0, 30.0, 3.0], dtype=torch.float64); st[4] = 3.0; st[5] = 1.0
    assert torch.equal(reduce_stats(st), st)
    s = summarize(st)
    assert s["mean_reward"] == 2.0 and s["mean_length"] == 10.0 and abs(s["std_reward"] - (14 / 3 - 4) ** 0.5) < 1e-12
