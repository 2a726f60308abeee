"""GPU: `ctdd_ema_update` (one launch for every trainable tensor) against the fixture the reference's EMA.update_ema
produced — bit-exact — and the Standard train step with the EMA mixin against a plain torch restatement of the loop."""
import numpy as np
import pytest
import torch

from oracle import ema_oracle as eo
from oracle.make_golden_ema import CASES, SHAPES

pytestmark = pytest.mark.gpu


def _ema_model(traj0, decay):
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models.models import EMA

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(p.copy())) for p in traj0])
            self.frozen = torch.nn.Parameter(torch.ones(5), requires_grad=False)

    class M(EMA, Net):
        def __init__(self, cfg):
            EMA.__init__(self, cfg)
            Net.__init__(self)
            self.init_ema()

    return M(make_config(model=dict(ema_decay=decay), device="cuda")).cuda()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_ema_kernel_matches_reference_fixture(golden, case):
    from ctdd_b200 import _native as nat
    g = golden["ema"]
    name, decay, steps, seed = case
    traj = eo.ema_inputs(seed, SHAPES, steps)
    m = _ema_model(traj[0], decay)
    before = nat.launch_count()
    for k in range(1, steps + 1):
        with torch.no_grad():
            for p, v in zip(m.ps, traj[k]):
                p.copy_(torch.from_numpy(v))
        m.update_ema()
    assert nat.launch_count() - before == steps                        # one launch per update, whatever the tensor count
    assert m.num_updates == steps
    for i, s in enumerate(m.shadow_params):
        assert np.array_equal(s.cpu().numpy(), g[f"{name}/shadow{i}"]), (name, i)   # bit-exact


def test_ema_table_follows_replaced_tensors_and_unaligned_views():
    """load_state_dict swaps the shadow list; odd offsets into a flat buffer give 4-byte-aligned (not 16) pointers."""
    from ctdd_b200 import ops
    g = torch.Generator().manual_seed(5)
    flat_s = torch.randn(100003, generator=g).cuda()
    flat_p = torch.randn(100003, generator=g).cuda()
    cuts = [0, 1, 6, 70001, 100003]
    sh = [flat_s[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    pa = [flat_p[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    want, _ = eo.ema_update([s.cpu().numpy() for s in sh], [p.cpu().numpy() for p in pa], 0.75, 100)
    t = ops.EmaTable(sh, pa)
    assert t.elements == 100003 and t.matches(sh, pa) and not t.matches(pa, sh)
    t.update(1.0 - 0.75)
    for s, w in zip(sh, want):
        assert np.array_equal(s.cpu().numpy(), w)
    with pytest.raises(ValueError):
        ops.EmaTable(sh, pa[:-1])
    with pytest.raises(RuntimeError):
        ops.EmaTable([s.double() for s in sh], pa)


def test_standard_train_step_with_ema_matches_torch_loop():
    """Two optimiser steps through Standard.step (fused loss not involved: plain quadratic loss) — parameters and shadows
    equal a hand-written torch loop of the reference's step (lib/training/training.py:17-40) bit for bit."""
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models.models import EMA
    import ctdd_b200.lib.training.training as tr

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(16, 33)
            self.b = torch.nn.Linear(33, 1)

        def forward(self, x):
            return self.b(torch.tanh(self.a(x)))

    class M(EMA, Net):
        def __init__(self, cfg):
            EMA.__init__(self, cfg)
            Net.__init__(self)
            self.init_ema()

    cfg = make_config(model=dict(ema_decay=0.99), training=dict(clip_grad=True, grad_norm=0.5, warmup=4),
                      optimizer=dict(lr=0.01), device="cuda")
    torch.manual_seed(3)
    m = M(cfg).cuda()
    m.init_ema()                                                       # shadows on the device, as after create_model(...).to(device)
    ref = Net().cuda()
    ref.load_state_dict({k: v for k, v in m.state_dict().items() if not k.startswith("ema_")})
    ref_shadow = [p.detach().clone() for p in ref.parameters()]
    opt, ropt = torch.optim.Adam(m.parameters(), lr=0.01), torch.optim.Adam(ref.parameters(), lr=0.01)

    class Loss:
        def calc_loss(self, state, minibatch, label=None):
            return state["model"](minibatch).pow(2).mean()

    step = tr.Standard(cfg)
    assert step.do_ema
    state = {"model": m, "optimizer": opt, "n_iter": 0}
    x = torch.randn(64, 16, generator=torch.Generator().manual_seed(4)).cuda()
    for it in range(1, 4):
        state["n_iter"] = it
        l = step.step(state, Loss(), x)
        ropt.zero_grad()
        rl = ref(x).pow(2).mean()
        rl.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        for gq in ropt.param_groups:
            gq["lr"] = 0.01 * min(it / 4, 1.0)
        ropt.step()
        omd = 1.0 - min(0.99, (1 + it) / (10 + it))
        with torch.no_grad():
            for s, p in zip(ref_shadow, ref.parameters()):
                s.sub_(omd * (s - p))
        assert float(l) == float(rl)
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.equal(p, q)
    for s, q in zip(m.shadow_params, ref_shadow):
        assert torch.equal(s, q)
