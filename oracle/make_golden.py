"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference has no tests / golden vectors of its own; these fixtures are outputs of the reference's own code
(lib.models.forward_model, lib.sampling.sampling, lib.losses.losses) driven on CPU with injected randomness
(oracle/ref_harness.py).  tests/test_oracle_golden.py pins oracle/ctmc_oracle.py against them on CPU;
tests/test_gpu_*.py compare the CUDA path with the same fixtures on the B200.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import cases, ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _fwd_cfg(name):
    f = cases.FORWARD[name]
    model = dict(f["model"])
    model.setdefault("Q_sigma", 20.0)
    return rh.make_cfg(data=dict(S=f["S"]), model=model, device="cpu")


def golden_forward(ref):
    out = {}
    for name, f in cases.FORWARD.items():
        m = getattr(ref.fm, f["mixin"])(_fwd_cfg(name), "cpu")
        t = torch.tensor(cases.FORWARD_TIMES[name], dtype=torch.float32)
        out[f"{name}/transition"] = m.transition(t).numpy()
        out[f"{name}/rate"] = m.rate(t).numpy()
        base = getattr(m, "base_rate", None)
        if base is None:
            base = m.rate_matrix
        out[f"{name}/base_rate"] = base.numpy()
        if hasattr(m, "transit_between"):
            t1 = 0.5 * t
            out[f"{name}/transit_between"] = m.transit_between(t1, t).numpy()
        if hasattr(m, "rate_mat"):
            g = np.random.Generator(np.random.PCG64(3))
            y = torch.from_numpy(g.integers(0, f["S"], (t.shape[0], 6)))
            out[f"{name}/rate_mat_y"] = y.numpy()
            out[f"{name}/rate_mat"] = m.rate_mat(y, t).numpy()
    np.savez_compressed(os.path.join(OUT, "forward.npz"), **out)
    return out


def rates_inputs(case):
    name, fwd, N, D, loss_name, logit_type, stub, t = case
    S = cases.FORWARD[fwd]["S"]
    g = np.random.Generator(np.random.PCG64(sum(map(ord, name))))
    x0 = g.integers(0, S, (N, D))
    s = np.arange(S)
    logits = stub[0] * 4.0 * g.standard_normal((N, D, S))
    if stub[1] is not None:
        logits = logits - (s[None, None, :] - x0[:, :, None]) ** 2 / (2.0 * stub[1] ** 2)
    x = np.clip(x0 + g.integers(-3, 4, (N, D)), 0, S - 1)
    return torch.from_numpy(logits.astype(np.float32)), torch.from_numpy(x), S


def golden_rates(ref):
    out = {}
    for case in cases.RATES:
        name, fwd, N, D, loss_name, logit_type, stub, t = case
        logits, x, S = rates_inputs(case)
        cfg = _fwd_cfg(fwd)
        cfg["loss"] = rh.Cfg(name=loss_name, logit_type=logit_type)
        cfg["sampler"] = rh.Cfg(eps_ratio=1e-9)
        m = rh.make_ref_model(ref, cases.FORWARD[fwd]["mixin"], cfg, S, D, 0)
        t_ones = t * torch.ones((N,))
        with torch.no_grad():
            rr, ratio = ref.ss.get_reverse_rates(m, logits, x, t_ones, cfg, N, D, S)
        out[f"{name}/rr"] = rr.numpy()
        out[f"{name}/ratio"] = ratio.numpy()
    np.savez_compressed(os.path.join(OUT, "rates.npz"), **out)
    return out


def golden_samplers(ref):
    out = {}
    for case in cases.SAMPLERS:
        name, cls, fwd, N, D, loss_name, logit_type, stub, over, max_t, seed = case
        cfg = cases.sampler_cfg(rh.make_cfg, case)
        S = cfg.data.S
        cond = over.get("condition_dim", 0)
        m = rh.make_ref_model(ref, cases.FORWARD[fwd]["mixin"], cfg, S, D, seed, stub[0], stub[1])
        sampler = getattr(ref.ss, cls)(cfg)
        args = ()
        if cond:
            g = np.random.Generator(np.random.PCG64(seed))
            conditioner = torch.from_numpy(g.integers(0, S, (N, cond)))
            args = (conditioner,)
            out[f"{name}/conditioner"] = conditioner.numpy()
        with rh.Injector(ref, seed=seed):
            res = sampler.sample(m, N, *args)
        if not isinstance(res, tuple):
            res = (res,)
        out[f"{name}/x"] = np.asarray(res[0]).astype(np.int64)
        for i, extra in enumerate(res[1:]):
            out[f"{name}/diag{i}"] = np.asarray(extra, dtype=np.float64)
        print(f"  {name}: x mean {out[f'{name}/x'].mean():.3f}")
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), **out)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)  # fixed summation order inside MKL for reproducible fixtures
    ref = rh.import_reference()
    print("forward ..."); golden_forward(ref)
    print("rates ..."); golden_rates(ref)
    print("samplers ..."); golden_samplers(ref)
    from . import make_golden_losses
    print("losses ..."); make_golden_losses.main(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
