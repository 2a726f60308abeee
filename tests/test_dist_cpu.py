"""CPU, world_size-2 gloo: the batch-sharding helper gathers exactly the unsharded result."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from ctdd_b200 import dist as cdist


class FakeSampler:
    """Stands in for a CUDA sampler: row i of the global batch is a deterministic function of its global row index,
    exactly the property the Philox keying gives the real samplers."""
    D = 6
    row_offset = 0

    def sample(self, model, n):
        first_sample = self.row_offset // self.D
        assert self.row_offset % 8 == 0
        g = (np.arange(n)[:, None] + first_sample) * self.D + np.arange(self.D)[None, :]
        return (g * 2654435761 % 251).astype(int), [float(n)]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    res = cdist.sample_sharded(FakeSampler(), None, n_total)
    q.put((rank, res[0], len(res)))
    tdist.destroy_process_group()


@pytest.mark.parametrize("n_total", [21, 5])
def test_sample_sharded_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, _ = FakeSampler().sample(None, n_total)
    for rank, x, diag in got:
        np.testing.assert_array_equal(x, want)


def test_shard_bounds_are_aligned_and_cover():
    for n, w in ((1024, 8), (1000, 8), (7, 4), (64, 3)):
        spans = [cdist.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c
        for a, b in spans:
            assert b == a or a % 8 == 0      # every non-empty shard starts on an 8-sample boundary
