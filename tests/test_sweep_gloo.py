"""The N>1 host path on CPU: world_size-2 gloo process group, contiguous sharding by global chain id
and the final gather of per-chain averages."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "polymer-stats_b200"))
    from polymc import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sweep.shard_range(total, rank, world)
        # fake per-chain results that encode the global chain id (the compute needs a GPU)
        local = np.stack([np.arange(lo, hi, dtype=np.float64) * 10 + k for k in range(18)], axis=1)
        full = sweep.gather_rows(local, total, lo)
        want = np.stack([np.arange(total, dtype=np.float64) * 10 + k for k in range(18)], axis=1)
        ok = bool(np.array_equal(full, want))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item()), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 4096])
def test_world2_shard_and_gather(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total % 7
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _, _ in res)
    assert all(t == 2.0 for _, _, t, _, _ in res)
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == total
