"""TEST INFRASTRUCTURE ONLY — loss fixtures (tests/golden/losses.npz) from the UNMODIFIED reference loss classes.

Called by oracle/make_golden.py.  Each case runs `calc_loss` of the reference class (lib/losses/losses.py) on CPU with
  * the time draw injected (torch.rand patched to return a fixed uniform vector),
  * the three Categorical draws of the noising (x_t, jump dimension, jump value; losses.py:46-101) replaced by the
    shared inverse-CDF map on Philox uniforms (streams NOISE_XT / TILDE_DIM / TILDE_VAL of oracle/rng.py),
  * a stub score network (ref_harness.StubNet) whose logits are retained so d loss / d logits can be stored.
Stored per case: the uniform vector, the times the network saw, x_t, x~, the loss value, d loss / d logits of every
forward pass and d loss / d w (the stub's only parameter).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import cases, ref_harness as rh, rng

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def minibatch_for(case):
    name, cls, fwd, B, D, over, t_hi, seed, n_iter = case
    S = cases.FORWARD[fwd]["S"]
    g = np.random.Generator(np.random.PCG64(seed))
    total_D = D + over.get("condition_dim", 0)
    x0 = g.integers(0, S, (B, total_D))
    u = g.uniform(0.02, 0.98, (B,)).astype(np.float32)
    label = g.integers(0, 10, (B,))
    return torch.from_numpy(x0), torch.from_numpy(u), torch.from_numpy(label)


def run_reference_loss(ref, case):
    name, cls, fwd, B, D, over, t_hi, seed, n_iter = case
    cfg = cases.loss_cfg(rh.make_cfg, case)
    S = cfg.data.S
    cd = over.get("condition_dim", 0)
    x0, u, label = minibatch_for(case)
    mixin = getattr(ref.fm, cases.FORWARD[fwd]["mixin"])
    seen = dict(logits=[], x=[], t=[])

    class M(rh.StubNet, mixin):
        def __init__(self):
            rh.StubNet.__init__(self, S, D + cd, seed, 1.0, 3.0 if S > 8 else None)
            mixin.__init__(self, cfg, "cpu")

        def forward(self, x, t, label=None):
            out = self.net(x, t)
            out.retain_grad()
            seen["logits"].append(out); seen["x"].append(x.clone()); seen["t"].append(t.clone())
            return out

    model = M()
    model.device = "cpu"
    loss_obj = getattr(ref.ll, cls)(cfg)
    state = {"model": model, "optimizer": None, "n_iter": n_iter}
    cat_log = []
    inj = rh.Injector(ref, seed=seed, ts=u, cat_offset=0,
                      cat_schedule=[rng.STREAM_NOISE_XT, rng.STREAM_TILDE_DIM, rng.STREAM_TILDE_VAL], cat_log=cat_log)
    with inj:
        if cls in cases.LOSS_MINIBATCH_FIRST:
            loss = loss_obj.calc_loss(x0, state)
        elif cls == "NLLOriginal":
            loss = loss_obj.calc_loss(state, x0, label)
        else:
            loss = loss_obj.calc_loss(state, x0)
    loss.backward()
    out = {f"{name}/loss": np.float32(loss.item()), f"{name}/u": u.numpy(), f"{name}/ts": seen["t"][0].numpy(),
           f"{name}/x0": x0.numpy(), f"{name}/label": label.numpy(),
           f"{name}/xt": cat_log[0].reshape(B, D).numpy(), f"{name}/grad_w": model.w.grad.numpy()}
    if len(cat_log) >= 3:
        xtil = cat_log[0].reshape(B, D).clone()
        xtil[torch.arange(B), cat_log[1]] = cat_log[2]
        out[f"{name}/x_tilde"] = xtil.numpy()
    for i, lg in enumerate(seen["logits"]):
        out[f"{name}/model_x{i}"] = seen["x"][i].numpy()
        out[f"{name}/grad_logits{i}"] = (lg.grad if lg.grad is not None else torch.zeros_like(lg)).numpy()
    return out


def oracle_loss_f64(case, ts):
    """fp64 evaluation of the same loss by oracle/loss_oracle.py (noising on the fp32 tables, everything after in
    double): the yardstick for the fp32 cancellation noise of the gradients (softmax Jacobian, 1/(q+eps) up to 1e9).
    -> (loss, grad_w, [grad_logits per forward pass])"""
    from . import loss_oracle as lo
    from . import ctmc_oracle as oc
    name, cls, fwd, B, D, over, t_hi, seed, n_iter = case
    cfg = cases.loss_cfg(rh.make_cfg, case)
    S, cd = cfg.data.S, over.get("condition_dim", 0)
    x0, u, label = minibatch_for(case)
    f = cases.FORWARD[fwd]
    fp = oc.ForwardProcess(f["kind"], f["S"], **dict(f["model"]))
    Q, R = fp.transition(ts), fp.rate(ts)

    class FP64:
        S = fp.S

        def transition(self, t):
            return Q.double()

        def rate(self, t):
            return R.double()

    net = rh.StubNet(S, D + cd, seed, 1.0, 3.0 if S > 8 else None).double()
    seen = []

    def model(x, t, label=None):
        out = net.net(x, t.double())
        out.retain_grad()
        seen.append(out)
        return out

    xd = x0[:, cd:]
    noised = lo.noise_xt(Q, R, xd, seed, 0)
    real = lo.noise_xt
    lo.noise_xt = lambda *a, **k: noised
    try:
        L = cfg.loss
        loss = lo.loss_value(cls, FP64(), model, x0, ts, seed=seed, eps=L.eps_ratio, nll_weight=L.nll_weight,
                             logit_type=L.logit_type, loss_type=L.loss_type, ce_coeff=L.ce_coeff,
                             one_forward_pass=L.one_forward_pass, n_iter=n_iter, n_iters=cfg.training.n_iters,
                             condition_dim=cd, label=label)
    finally:
        lo.noise_xt = real
    loss.backward()
    return loss.item(), net.w.grad.numpy(), [(lg.grad if lg.grad is not None else torch.zeros_like(lg)).numpy() for lg in seen]


def main(ref=None):
    ref = ref or rh.import_reference()
    torch.set_num_threads(1)
    out = {}
    for case in cases.LOSSES:
        r = run_reference_loss(ref, case)
        out.update(r)
        l64, gw64, gl64 = oracle_loss_f64(case, torch.from_numpy(r[case[0] + "/ts"]))
        assert abs(l64 - float(r[case[0] + "/loss"])) <= 2e-5 * abs(l64), (case[0], l64)
        out[f"{case[0]}/grad_w_f64"] = gw64
        for i, gl in enumerate(gl64):
            out[f"{case[0]}/grad_logits{i}_f64"] = gl
        print(f"  {case[0]}: loss {r[case[0] + '/loss']:.6f}")
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    return out


if __name__ == "__main__":
    main()
