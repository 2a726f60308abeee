"""Diagnostic: host-side profile of TauL.sample at the C4 shape with a stub network (where does the per-step time go?)."""
import cProfile, io, os, pstats, sys, time
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ctdd_b200 import make_config
from ctdd_b200.lib.models import forward_model as fm
from ctdd_b200.lib.sampling import sampling_utils
import ctdd_b200.lib.sampling.sampling  # noqa: F401

w = bench.WORKLOADS["C4"]; S, D, B = w["S"], w["D"], w["B"]; dev = "cuda:0"
steps = int(os.environ.get("STEPS", 16))
mcfg = dict(w["model"], concat_dim=D)
cfg = make_config(data=dict(S=S, shape=[D], name="DiscreteCIFAR10"), model=mcfg, training=dict(max_t=1.0),
                  sampler=dict(name="TauL", num_steps=steps, min_t=0.01, eps_ratio=1e-9, initial_dist="gaussian", num_corrector_steps=0,
                               corrector_step_size_multiplier=1.5, corrector_entry_time=0.0, is_ordinal=True),
                  loss=dict(name="CTElboLambda", eps_ratio=1e-9, logit_type="reverse_prob"), device=dev)
lg, _ = bench.synth_logits(B, D, S, 1, torch.device(dev))

class Stub(nn.Module, fm.GaussianTargetRate):
    def __init__(self):
        nn.Module.__init__(self); fm.GaussianTargetRate.__init__(self, cfg, dev)
    def forward(self, x, t):
        return lg

m = Stub(); m.device = dev
s = sampling_utils.get_sampler(cfg); s.seed = 1
s.sample(m, B); torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter(); pr.enable(); s.sample(m, B); torch.cuda.synchronize(); pr.disable()
print("total ms", (time.perf_counter() - t0) * 1e3, "per step", (time.perf_counter() - t0) * 1e3 / steps)
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(18); print(out.getvalue()[-3500:])
