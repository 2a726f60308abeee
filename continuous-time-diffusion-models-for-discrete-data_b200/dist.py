"""Batch sharding of the samplers across ranks (one process per GPU, torch.distributed).

The reverse step has no cross-sample dependence, so the only collective is the final gather of
the (N/G, D) integer states. Philox counters are keyed on the GLOBAL row index (`row_offset`), which makes the
gathered result independent of the number of ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as tdist


def shard_bounds(n_total: int, world: int, rank: int, multiple: int = 8):
    """[first, last) samples of `rank`; every shard but the last is a multiple of `multiple` samples so that each
    rank's first global row (first * D) is 8-aligned for any D."""
    per = -(-n_total // world)
    per = -(-per // multiple) * multiple
    first = min(rank * per, n_total)
    return first, min(first + per, n_total)


def sample_sharded(sampler, model, n_total: int, *args, group=None):
    """Run `sampler.sample(model, n_local, ...)` on this rank's slice and all-gather the samples.

    Returns a tuple: the gathered (n_total, D) int array followed by whatever else the local `sample` call returned
    (diagnostic lists stay rank-local; a rank whose shard is empty returns only the gathered array)."""
    if tdist.is_available() and tdist.is_initialized():
        world, rank = tdist.get_world_size(group), tdist.get_rank(group)
    else:
        world, rank = 1, 0
    first, last = shard_bounds(n_total, world, rank)
    n_local = last - first
    D = getattr(sampler, "D", None)
    if D is None:
        cfg = sampler.cfg
        D = cfg.data.shape[0] - getattr(cfg.sampler, "condition_dim", 0) if hasattr(cfg.sampler, "condition_dim") \
            else cfg.model.concat_dim
    sampler.row_offset = first * D
    local_args = tuple(a[first:last] if (torch.is_tensor(a) and a.shape[0] == n_total) else a for a in args)
    res = sampler.sample(model, n_local, *local_args) if n_local > 0 else None
    tup = res if isinstance(res, tuple) else (res,)
    x_local = np.asarray(tup[0]) if res is not None else np.zeros((0, 0), dtype=np.int64)
    if world == 1:
        return tup
    backend = tdist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = torch.tensor([x_local.shape[1] if x_local.size else 0], device=dev)
    tdist.all_reduce(width, op=tdist.ReduceOp.MAX, group=group)
    width = int(width.item())
    per = shard_bounds(n_total, world, 0)[1]
    buf = torch.zeros((per, width), dtype=torch.int64, device=dev)
    if n_local:
        buf[:n_local] = torch.from_numpy(x_local.astype(np.int64)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    tdist.all_gather(out, buf, group=group)
    parts = []
    for r, t in enumerate(out):
        f, l = shard_bounds(n_total, world, r)
        parts.append(t[: l - f].cpu().numpy())
    x_all = np.concatenate(parts, axis=0).astype(int)
    return (x_all,) + tuple(tup[1:])
