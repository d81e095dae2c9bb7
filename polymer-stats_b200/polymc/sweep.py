"""Sweeps across GPUs (north_star "Sweeps across GPUs").

The reference fans cases out with `pmap`, one fresh `julia` process per case
(run/interacting_dielectric_study.jl:37-47), and "gathers" through files on a shared filesystem.
Here the chains of a sweep (cases × replicas) are independent units: each rank (one process per
GPU, torch.distributed) takes a contiguous block of global chain ids, runs it with no data-path
collective, and only the final per-chain averages are gathered (NCCL all_gather over NVLink).
Philox streams are keyed by the GLOBAL chain id, so the result does not depend on the GPU count.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import lib


def shard_range(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `total` units owned by `rank` (SURVEY §8e partitioning)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi


def bucket_cases(cases):
    """Group case indices by (n, energy_type): one handle/kernel per bucket (include/polymc.h)."""
    buckets = OrderedDict()
    for i, c in enumerate(cases):
        buckets.setdefault((int(c.n), int(c.energy_type)), []).append(i)
    return buckets


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def gather_rows(local: np.ndarray, total_rows: int, lo: int, device=None) -> np.ndarray:
    """Assemble [total_rows][k] from each rank's [hi-lo][k] block.  One all_gather of equal-sized
    padded blocks (NCCL when the process group is NCCL, gloo on CPU); identity without a group."""
    dist = _dist()
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        assert local.shape[0] == total_rows
        return local
    import torch
    world = dist.get_world_size()
    k = local.shape[1]
    maxrows = max(shard_range(total_rows, r, world)[1] - shard_range(total_rows, r, world)[0] for r in range(world))
    pad = np.zeros((maxrows, k))
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.empty((total_rows, k))
    for r in range(world):
        a, b = shard_range(total_rows, r, world)
        full[a:b] = outs[r].cpu().numpy()[: b - a]
    return full


def run_shard(cases, replicas: int, nsteps: int, stepout: int = 0, seed: int = 0, device: int = 0,
              rank: int = 0, world: int = 1):
    """Run this rank's contiguous block of every (n, energy) bucket.  Returns a list of
    (global chain ids of the bucket, lo, block [hi-lo][35]) with block columns = 16 averages,
    acceptance rate, normaliser, 17 raw sums.  No communication."""
    cases = list(cases)
    out = []
    for (_, _), idxs in bucket_cases(cases).items():
        # global chain ids of this bucket, case-major
        gids = np.concatenate([np.arange(i * replicas, (i + 1) * replicas) for i in idxs])
        lo, hi = shard_range(len(gids), rank, world)
        mine = gids[lo:hi]
        block = np.zeros((hi - lo, 35))
        # a rank's block may start/end inside a case: one handle per run of consecutive chains of
        # one case, with chain_id_base = global id of its first chain.
        pos = 0
        while pos < len(mine):
            case_i = mine[pos] // replicas
            end = pos
            while end < len(mine) and mine[end] // replicas == case_i and mine[end] == mine[pos] + (end - pos):
                end += 1
            with lib.Ensemble(cases[case_i], replicas=end - pos, seed=seed, device=device,
                              chain_id_base=int(mine[pos])) as ens:
                ens.run(nsteps, stepout, fetch_rows=False)
                avg, ar, nrm = ens.averages()
                block[pos:end, :16] = avg
                block[pos:end, 16] = ar
                block[pos:end, 17] = nrm
                block[pos:end, 18:35] = ens.accumulators()
            pos = end
        out.append((gids, lo, block))
    return out


def assemble(total: int, parts):
    """dict of result arrays from [(gids, full [len(gids)][35]), ...]."""
    res = {"avg": np.full((total, 16), np.nan), "acc_rate": np.full(total, np.nan),
           "normalizer": np.full(total, np.nan), "sums": np.full((total, 17), np.nan)}
    for gids, full in parts:
        res["avg"][gids] = full[:, :16]
        res["acc_rate"][gids] = full[:, 16]
        res["normalizer"][gids] = full[:, 17]
        res["sums"][gids] = full[:, 18:35]
    return res


def run_sweep(cases, replicas: int, nsteps: int, stepout: int = 0, seed: int = 0, device: int = 0,
              torch_device=None):
    """Run every (case, replica) chain of a sweep for nsteps trials, sharded over the ranks of the
    current torch.distributed group (or a single process), then gather the final per-chain results on
    every rank.  Chain order = case-major: avg [ncases*replicas][16], acc_rate, normalizer, sums [..][17]."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    cases = list(cases)
    parts = []
    for gids, lo, block in run_shard(cases, replicas, nsteps, stepout, seed, device, rank, world):
        parts.append((gids, gather_rows(block, len(gids), lo, device=torch_device)))
    return assemble(len(cases) * replicas, parts)
