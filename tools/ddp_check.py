"""Data-parallel training check (north_star: "NCCL ... for the DDP gradient allreduce"; reference stub
lib/models/models.py:104-107 wraps the network inside the model in DistributedDataParallel).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py

Every rank takes its slice of one global minibatch and runs ONE `Standard.step` (fused-kernel loss behind its
autograd.Function, backward, DDP gradient all-reduce over NCCL, SGD step, one-launch EMA) for CTElbo and SDDMElbo.
Rank 0 repeats the step single-process on the concatenated batch.  Checked: parameters after the step (= gradients, the
optimiser is plain SGD) and EMA shadows agree to fp32 reduction noise, and are identical on both ranks.
The time draw and the noising draws are keyed on the GLOBAL sample (loss.ts_override / loss.batch_offset), so the
sharded run sees exactly the samples of the single-process run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models import forward_model as fm
    from ctdd_b200.lib.models.models import EMA
    from ctdd_b200.lib.losses import losses_utils
    import ctdd_b200.lib.losses.losses  # noqa: F401
    import ctdd_b200.lib.training.training as tr

    S, D, B = 32, 24, 16 * world
    Bl = B // world
    results = {}
    for loss_name in ("CTElbo", "SDDMElbo"):
        cfg = make_config(data=dict(S=S, shape=[D], name="DiscreteCIFAR10"),
                          model=dict(rate_sigma=3.0, Q_sigma=20.0, time_exp=50.0, time_base=1.0, concat_dim=D, ema_decay=0.99),
                          training=dict(max_t=1.0, n_iters=100, clip_grad=False, grad_norm=1.0, warmup=0),
                          optimizer=dict(lr=0.05),
                          loss=dict(name=loss_name, eps_ratio=1e-9, nll_weight=0.01, min_time=0.01, one_forward_pass=True,
                                    logit_type="reverse_prob", loss_type="rm", ce_coeff=0.0), device=str(dev))

        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                g = torch.Generator().manual_seed(5)
                self.A = nn.Parameter(0.5 * torch.randn((S, S), generator=g))
                self.Bv = nn.Parameter(0.5 * torch.randn((D, S), generator=g))

            def forward(self, x, t):
                return self.A[x.long()] + t.view(-1, 1, 1) * self.Bv.unsqueeze(0)

        class Model(EMA, nn.Module, fm.GaussianTargetRate):
            def __init__(self, ddp):
                nn.Module.__init__(self)
                EMA.__init__(self, cfg)
                fm.GaussianTargetRate.__init__(self, cfg, str(dev))
                net = Net().to(dev)
                # the reference wraps the network INSIDE the model (models.py:104-107)
                self.net = nn.parallel.DistributedDataParallel(net, device_ids=[local]) if ddp else net
                self.init_ema()

            def forward(self, x, t):
                return self.net(x, t)

        g = np.random.Generator(np.random.PCG64(11))
        x0 = torch.from_numpy(g.integers(0, S, (B, D))).to(dev)
        ts = torch.from_numpy(g.uniform(0.05, 0.95, B).astype(np.float32)).to(dev)

        def run(ddp, lo, hi):
            m = Model(ddp)
            m.device = str(dev)
            loss = losses_utils.get_loss(cfg)
            loss.seed, loss.ts_override, loss.batch_offset = 77, ts[lo:hi], lo
            state = {"model": m, "optimizer": torch.optim.SGD(m.parameters(), lr=0.05), "n_iter": 1}
            step = tr.Standard(cfg)
            mb = x0[lo:hi]
            l = step.step(state, loss, mb) if loss_name == "CTElbo" else step.step(state, mb, loss)
            params = [p.detach().clone() for p in m.parameters()]
            return float(l), params, [s.detach().clone() for s in m.shadow_params]

        l_d, p_d, s_d = run(world > 1, rank * Bl, (rank + 1) * Bl)
        ok = True
        if world > 1:   # both ranks hold the same parameters after the all-reduced step
            for p in p_d + s_d:
                q = p.clone()
                dist.broadcast(q, 0)
                ok &= bool(torch.equal(p, q))
        if rank == 0:
            l_s, p_s, s_s = run(False, 0, B)
            err_p = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(p_d, p_s))
            err_s = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(s_d, s_s))
            results[loss_name] = dict(loss_rank0_shard=l_d, loss_full_batch=l_s, param_rel_err=err_p, shadow_rel_err=err_s,
                                      ranks_identical=ok)
            assert ok, "ranks diverged after the DDP step"
            assert err_p < 2e-5 and err_s < 2e-5, (loss_name, err_p, err_s)
    if rank == 0:
        import json
        print(json.dumps({"ddp_check": "ok", "world": world, "results": results}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
