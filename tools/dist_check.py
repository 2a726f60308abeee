"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): the batch-sharded TauL sampler gathers exactly the
samples a single GPU produces for the whole batch (Philox is keyed on the global row), S = 256 tensor path.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctdd_b200 import make_config, dist as cdist  # noqa: E402
from ctdd_b200.lib.models import forward_model as fm  # noqa: E402
from ctdd_b200.lib.sampling import sampling_utils  # noqa: E402
import ctdd_b200.lib.sampling.sampling  # noqa: E402,F401
from oracle import ref_harness as rh  # noqa: E402  (StubNet only: a deterministic stand-in score network)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    tdist.init_process_group("nccl", device_id=torch.device(dev))
    S, D, N, seed = 256, 96, 40, 77
    cfg = make_config(data=dict(S=S, shape=[D], name="DiscreteCIFAR10"),
                      model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0, concat_dim=D),
                      training=dict(max_t=1.0),
                      sampler=dict(name="TauL", num_steps=12, min_t=0.01, eps_ratio=1e-9, initial_dist="gaussian",
                                   num_corrector_steps=0, corrector_step_size_multiplier=1.5, corrector_entry_time=0.0,
                                   is_ordinal=True),
                      loss=dict(name="CTElboLambda", eps_ratio=1e-9, logit_type="reverse_prob"), device=dev)

    class M(rh.StubNet, fm.GaussianTargetRate):
        def __init__(self):
            rh.StubNet.__init__(self, S, D, seed, 0.3, 12.0)
            fm.GaussianTargetRate.__init__(self, cfg, dev)

        def forward(self, x, t):
            return self.net(x, t)

    model = M().to(dev)
    model.device = dev
    sampler = sampling_utils.get_sampler(cfg)
    sampler.seed = seed
    res = cdist.sample_sharded(sampler, model, N)
    sharded = np.asarray(res[0])
    single = sampling_utils.get_sampler(cfg)
    single.seed = seed
    whole = np.asarray(single.sample(model, N)[0])
    same = np.array_equal(sharded, whole)
    flag = torch.tensor([1 if same else 0], device=dev)
    tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_check world={world}: sharded == single-GPU samples on every rank: {bool(flag.item())} "
              f"(N={N}, D={D}, shards {[cdist.shard_bounds(N, world, r) for r in range(world)]})", flush=True)
    tdist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
