"""CPU: oracle/ema_oracle.py against the fixture the reference's own EMA.update_ema produced (tests/golden/ema.npz), and
the host-side logic of the EMA mixin / Standard train step (state-dict keys, error behaviour, no CPU fallback)."""
import numpy as np
import pytest
import torch

from oracle import ema_oracle as eo
from oracle.make_golden_ema import CASES, SHAPES


def test_ema_oracle_matches_reference_fixture(golden):
    g = golden["ema"]
    for name, decay, steps, seed in CASES:
        traj = eo.ema_inputs(seed, SHAPES, steps)
        shadows, n = [p.copy() for p in traj[0]], 0
        for k in range(1, steps + 1):
            shadows, n = eo.ema_update(shadows, traj[k], decay, n)
        assert n == steps
        for i, s in enumerate(shadows):
            assert np.array_equal(s, g[f"{name}/shadow{i}"]), (name, i)       # bit-exact


def test_effective_decay_warmup():
    assert eo.effective_decay(0.9999, 1) == 2 / 11
    assert eo.effective_decay(0.9999, 10 ** 6) == 0.9999
    assert eo.effective_decay(0.0, 5) == 0.0


def _model(decay=0.9, device="cpu"):
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models.models import EMA

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(3, 2)
            self.frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)

    class M(EMA, Net):
        def __init__(self, cfg):
            EMA.__init__(self, cfg)
            Net.__init__(self)
            self.init_ema()

    return M(make_config(model=dict(ema_decay=decay), device=device))


def test_ema_mixin_host_contract():
    """Same attributes, state-dict keys and errors as the reference mixin (lib/models/models.py:729-826)."""
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models.models import EMA
    with pytest.raises(ValueError):
        EMA(make_config(model=dict(ema_decay=1.5), device="cpu"))
    m = _model()
    assert len(m.shadow_params) == 2 and m.num_updates == 0            # the frozen tensor has no shadow
    sd = m.state_dict()
    assert {"ema_decay", "ema_num_updates", "ema_shadow_params"} <= set(sd)
    m2 = _model(decay=0.5)
    m2.load_state_dict(sd)
    assert m2.decay == 0.9 and m2.shadow_params is sd["ema_shadow_params"]
    bad = dict(sd)
    del bad["ema_decay"]
    with pytest.raises(ValueError):
        m2.load_state_dict(bad)
    bad = {k: v for k, v in sd.items() if k != "lin.weight"}
    with pytest.raises(ValueError):
        m2.load_state_dict(bad)
    with pytest.raises(ValueError):
        m.train(True)                                                   # same mode twice
    w = m.lin.weight.detach().clone()
    m.shadow_params[0].add_(1.0)
    m.train(False)                                                      # eval: shadows move into the model
    assert torch.equal(m.lin.weight, m.shadow_params[0])
    m.train(True)                                                       # back: collected parameters restored
    assert torch.equal(m.lin.weight, w)
    empty = _model()
    empty.shadow_params = []
    with pytest.raises(ValueError):
        empty.update_ema()


def test_ema_update_has_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        _model().update_ema()


def test_train_step_registry_and_call_orders():
    from ctdd_b200 import make_config
    from ctdd_b200.lib.training import training_utils
    import ctdd_b200.lib.training.training as tr
    cfg = make_config(model=dict(name="x"), training=dict(train_step_name="Standard", clip_grad=True, grad_norm=1.0, warmup=10),
                      optimizer=dict(lr=0.1), device="cpu")
    step = training_utils.get_train_step(cfg)
    assert isinstance(step, tr.Standard) and step.do_ema is False
    with pytest.raises(ValueError):
        training_utils.register_train_step(tr.Standard)
    net = torch.nn.Linear(3, 1)
    state = {"model": net, "optimizer": torch.optim.SGD(net.parameters(), lr=0.1), "n_iter": 5}

    class Loss:
        def __init__(self, bad=False):
            self.bad = bad

        def calc_loss(self, state, minibatch, label=None):
            out = state["model"](minibatch).pow(2).mean()
            return out * float("nan") if self.bad else out

    x = torch.randn(8, 3)
    w0 = net.weight.detach().clone()
    l1 = step.step(state, Loss(), x)                                    # train_image.py order
    assert state["optimizer"].param_groups[0]["lr"] == pytest.approx(0.05)
    l2 = step.step(state, x, Loss())                                    # train_maze.py / train_synthetic.py order
    assert l1.dim() == 0 and l2.dim() == 0 and not torch.equal(net.weight, w0)
    w1 = net.weight.detach().clone()
    l3 = step.step(state, Loss(bad=True), x)                            # NaN loss: 1e9, update skipped
    assert float(l3) == 1e9 and torch.equal(net.weight, w1)
