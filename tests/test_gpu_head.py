"""GPU: the truncated-logistic head (standalone kernel and fused into the tcgen05 reverse step) against the fixture the
reference's sample_logistic + get_reverse_rates produced, and against the oracle on the same injected uniforms."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho, ctmc_oracle as oc
from oracle import cases, ref_harness as rh
from oracle.make_golden_head import HEAD_CASES, HEAD_SAMPLERS, case_inputs, FWD
from helpers import oracle_forward, mismatch_fraction

pytestmark = pytest.mark.gpu


def _nat():
    from ctdd_b200 import _native as nat
    return nat


def _tables(fp, t, branch):
    from ctdd_b200 import ops
    tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
    Q = fp.transition(tt)[0]
    tb = dict(Q=Q.cuda().contiguous(), QT=Q.t().contiguous().cuda(), Rb=fp.base_rate.cuda().contiguous(),
              RbT=fp.base_rate.t().contiguous().cuda(), beta=float(fp.beta(tt)[0]))
    tct = ops.prep_tc_tables(tb["Q"][None], tb["QT"][None], tb["Rb"], 1e-9, branch)[0]
    return tb, tct, ops.prep_tc_static(tb["Rb"])


@pytest.mark.parametrize("case", HEAD_CASES, ids=[c[0] for c in HEAD_CASES])
def test_logistic_logits_kernel_matches_reference(golden, case):
    """Standalone head kernel: probabilities within the reference's own fp32 noise of the fp64 value; logits agree
    wherever they are above the 1e-6 guard."""
    from ctdd_b200 import ops
    g = golden["head"]
    name, N, D, fix = case[:4]
    mu, ls, x, S = case_inputs(case)
    got = ops.logistic_logits(mu.cuda(), ls.cuda(), S, fix).cpu()
    ref = torch.from_numpy(g[f"{name}/logits"])
    p64 = torch.from_numpy(g[f"{name}/p64"])
    ref_err = float((torch.softmax(ref, -1).double() - p64).abs().max())
    err = float((torch.softmax(got, -1).double() - p64).abs().max())
    assert err <= max(2e-7, 2.0 * ref_err), (err, ref_err)
    l64 = ho.truncated_logistic_logits(mu.double(), ls.double(), S, fix)
    assert float((got.double() - l64).abs().max()) <= 1e-4 * float(l64.abs().max())      # full range, no floor
    live = ref > -8.0        # closer to log(1e-6) the reference's own 1 - exp(.) + 1e-6 carries percent-level fp32 noise
    assert float((got - ref)[live].abs().max()) <= 2e-3


@pytest.mark.parametrize("case", HEAD_CASES, ids=[c[0] for c in HEAD_CASES])
def test_fused_head_rates_match_reference(golden, case):
    """Reverse rates with the head evaluated inside the tcgen05 kernel (no logits tensor): 1e-4 relative against what the
    reference's get_reverse_rates makes of the reference head's logits; exact zeros preserved."""
    from ctdd_b200 import ops
    nat = _nat()
    g = golden["head"]
    name, N, D, fix, _, loss_name, logit_type, t = case
    mu, ls, x, S = case_inputs(case)
    fp = oracle_forward(FWD)
    branch = nat.branch_for(loss_name, logit_type)
    tb, tct, tcs = _tables(fp, t, branch)
    out = ops.reverse_step(nat.MODE_RATES_ONLY, branch, None, x.to(torch.int32).cuda(), tb["Q"], tb["QT"], tb["Rb"], tb["RbT"],
                           tb["beta"], 0.0, 1e-9, N=N, D=D, S=S, impl=nat.IMPL_TC, tc_tables=tct, tc_static=tcs,
                           want_rr=True, want_ratio=True, head=(mu.cuda(), ls.cuda(), fix))
    for key in ("rr", "ratio"):
        ref, r64 = g[f"{name}/{key}"], g[f"{name}/{key}64"]
        got = out[key].cpu().numpy()
        big = np.abs(ref) > 1e-30
        rel = np.abs(got - ref)[big] / np.abs(ref)[big]
        rel64 = np.abs(got - r64)[big] / np.abs(r64)[big]
        assert rel.max() <= 1e-4, (key, rel.max(), rel64.max())
        assert rel64.max() <= 1e-4, (key, rel64.max())
        assert np.all(np.abs(got[~big]) <= 1e-30)


HEAD_STEP_CASES = [
    # N, D, fix, (scale_lo, scale_hi), loss, logit_type, t, h
    (6, 40, False, (-2.0, 2.0), "CTElboLambda", None, 0.9, 0.004),
    (6, 40, True, (-3.0, 1.0), "CatRM", "reverse_prob", 0.3, 0.01),
    (3, 75, False, (-1.0, 3.0), "NLL", None, 0.05, 0.02),          # 225 rows: ragged last tile
]


@pytest.mark.parametrize("sc", HEAD_STEP_CASES, ids=[f"{c[4]}-fix{int(c[2])}-t{c[6]}" for c in HEAD_STEP_CASES])
def test_fused_head_step_modes_match_oracle(sc):
    """Every update mode with the fused head on identical inputs and injected uniforms: integer states bit-exact modulo
    threshold ties against the oracle stepping from the oracle head's logits."""
    from ctdd_b200 import ops
    nat = _nat()
    N, D, fix, (lo, hi), loss_name, logit_type, t, h = sc
    S = 256
    mu, ls = ho.head_inputs(N * D, 17 + N, lo, hi)
    mu, ls = mu.view(N, D), ls.view(N, D)
    g = np.random.Generator(np.random.PCG64(3))
    centre = np.clip(np.round((mu.numpy() + 1.0) * S / 2.0 - 0.5), 0, S - 1).astype(np.int64)
    x = torch.from_numpy(np.clip(centre + g.integers(-4, 5, (N, D)), 0, S - 1))
    fp = oracle_forward(FWD)
    branch = nat.branch_for(loss_name, logit_type)
    tb, tct, tcs = _tables(fp, t, branch)
    tt = torch.tensor([t], dtype=torch.float64).to(torch.float32)
    Qo, Ro = fp.transition(tt), fp.rate(tt)
    logits = ho.truncated_logistic_logits(mu, ls, S, fix)
    rr, _ = oc.reverse_rates(logits, x, Qo, Ro, loss_name, logit_type or "reverse_prob", 1e-9)
    rz = oc._zero_at(rr, x)
    rz_corr = oc._zero_at(Ro.expand(N, -1, -1)[torch.arange(N).view(N, 1), x.long()] + rz, x)
    xe = x.to(torch.int32).cuda()
    # the two halves of a chunked (N, 2D) network output: batch stride 2D, like unet.py:451
    both = torch.cat([mu, ls], dim=1).cuda()
    mu_v, ls_v = torch.chunk(both, 2, dim=1)
    seed = 777

    def run(mode, offset, reject=False, x_base=None):
        out = ops.reverse_step(mode, branch, None, xe, tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], h, 1e-9,
                               N=N, D=D, S=S, impl=nat.IMPL_TC, tc_tables=tct, tc_static=tcs, reject_multi=reject,
                               seed=seed, offset=offset, x_base=x_base, head=(mu_v, ls_v, fix))
        return out["x"].cpu().numpy().astype(np.int64)

    tol = 3e-3
    for reject in (False, True):
        want, _ = oc.tau_leap_update(rz, x, x, h, S, reject, 3, seed)
        assert mismatch_fraction(run(nat.MODE_TAU_LEAP, 3, reject), want.numpy()) <= tol
    want, _ = oc.tau_leap_update(rz_corr, x, x, h, S, False, 5, seed)
    assert mismatch_fraction(run(nat.MODE_TAU_LEAP_CORR, 5), want.numpy()) <= tol
    assert mismatch_fraction(run(nat.MODE_MIDPOINT_DRIFT, 0), oc.midpoint_drift(rz, x, h, S).numpy()) <= tol
    want, _ = oc.euler_update(rz, x, h, S, 9, seed)
    assert mismatch_fraction(run(nat.MODE_EULER, 9), want.numpy()) <= tol
    want, _ = oc.euler_update(rz_corr, x, h, S, 10, seed)
    assert mismatch_fraction(run(nat.MODE_EULER_CORR, 10), want.numpy()) <= tol


def test_fused_head_edge_rows():
    """Rows outside what tanh emits (|mu| >= 1), extreme scales: the fused numerators and the standalone logits stay
    finite and agree with the fp64 oracle after the softmax; fused == dense path fed with the standalone logits."""
    from ctdd_b200 import ops
    nat = _nat()
    S = 256
    mu = torch.tensor([1.0, -1.0, 0.0, 2.5, -3.0, 0.999, -0.2, 0.3, 1.5, -1.5, 0.7, -0.7, 0.1, 40.0, -40.0, 0.0])
    ls = torch.tensor([0.0, 0.0, -6.0, -2.0, -2.0, 6.0, 8.0, -8.0, 3.0, 3.0, -4.0, 1.0, 2.0, -1.0, -1.0, 12.0])
    N, D = 1, mu.numel()
    fp = oracle_forward(FWD)
    x = torch.full((N, D), 128, dtype=torch.int64)
    for fix in (False, True):
        got = ops.logistic_logits(mu.view(N, D).cuda(), ls.view(N, D).cuda(), S, fix).cpu()
        assert torch.isfinite(got).all()
        l64 = ho.truncated_logistic_logits(mu.double().view(N, D), ls.double().view(N, D), S, fix)
        p64 = torch.softmax(l64, -1)
        assert float((torch.softmax(got.double(), -1) - p64).abs().max()) <= 2e-6
        for branch, loss_name, lt in ((nat.BRANCH_TAULDR, "CTElbo", None), (nat.BRANCH_SDDM_REVERSE_PROB, "CatRM", "reverse_prob")):
            tb, tct, tcs = _tables(fp, 0.5, branch)
            kw = dict(N=N, D=D, S=S, impl=nat.IMPL_TC, tc_tables=tct, tc_static=tcs, want_rr=True)
            args = (tb["Q"], tb["QT"], tb["Rb"], tb["RbT"], tb["beta"], 0.0, 1e-9)
            fused = ops.reverse_step(nat.MODE_RATES_ONLY, branch, None, x.to(torch.int32).cuda(), *args,
                                     head=(mu.view(N, D).cuda(), ls.view(N, D).cuda(), fix), **kw)["rr"].cpu().double()
            tt = torch.tensor([0.5], dtype=torch.float64).to(torch.float32)
            r64, _ = oc.reverse_rates(l64, x, fp.transition(tt).double(), fp.rate(tt).double(), loss_name, lt or "reverse_prob")
            assert torch.isfinite(fused).all()
            scale = r64.abs().amax(-1, keepdim=True)
            assert float(((fused - r64).abs() / scale).max()) <= 1e-4, (fix, branch)


def test_head_needs_tc_or_materialises():
    """Shapes the tcgen05 path does not take are served by materialising the logits once (same results as passing them);
    the C entry point itself refuses a head on the SIMT path instead of silently ignoring it."""
    from ctdd_b200 import ops
    nat = _nat()
    fp = oracle_forward("gauss32")
    S, N, D = 32, 4, 20
    mu, ls = ho.head_inputs(N * D, 9, -1.0, 2.0)
    mu, ls = mu.view(N, D).cuda(), ls.view(N, D).cuda()
    x = torch.randint(0, S, (N, D), dtype=torch.int32, device="cuda")
    tt = torch.tensor([0.4], dtype=torch.float32)
    Q = fp.transition(tt)[0]
    a = (Q.cuda().contiguous(), Q.t().contiguous().cuda(), fp.base_rate.cuda().contiguous(), fp.base_rate.t().contiguous().cuda(),
         float(fp.beta(tt)[0]), 0.02, 1e-9)
    via_head = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, None, x, *a, N=N, D=D, S=S, seed=5, head=(mu, ls, False))["x"]
    logits = ops.logistic_logits(mu, ls, S, False)
    direct = ops.reverse_step(nat.MODE_TAU_LEAP, nat.BRANCH_TAULDR, logits, x, *a, N=N, D=D, S=S, seed=5)["x"]
    assert torch.equal(via_head, direct)
    p = nat.StepParams(mode=nat.MODE_TAU_LEAP, branch=nat.BRANCH_TAULDR, impl=nat.IMPL_SIMT, N=N, D=D, S=S,
                       x_eval=x.data_ptr(), Q=a[0].data_ptr(), QT=a[1].data_ptr(), Rb=a[2].data_ptr(), RbT=a[3].data_ptr(),
                       beta=a[4], h=0.02, eps=1e-9, x_out=via_head.data_ptr(), head=nat.HEAD_LOGISTIC,
                       head_mu=mu.data_ptr(), head_log_scale=ls.data_ptr(), head_batch_stride=D)
    assert nat.lib().ctdd_reverse_step(p, nat.stream()) != 0
    assert b"logistic head" in nat.lib().ctdd_last_error()


def test_logistic_logits_small_state_spaces():
    """Scalar (S % 4 != 0) and vector paths of the standalone head kernel against the fp64 oracle."""
    from ctdd_b200 import ops
    for S in (2, 3, 5, 32, 100):
        mu, ls = ho.head_inputs(77, S, -2.0, 2.0)
        for fix in (False, True):
            got = ops.logistic_logits(mu.view(7, 11).cuda(), ls.view(7, 11).cuda(), S, fix).cpu().double()
            l64 = ho.truncated_logistic_logits(mu.double().view(7, 11), ls.double().view(7, 11), S, fix)
            assert float((got - l64).abs().max()) <= 1e-4 * max(1.0, float(l64.abs().max()))


def _head_product_model(cfg, S, D, seed, fix):
    from ctdd_b200.lib.models import forward_model as fm, models
    mixin = getattr(fm, cases.FORWARD[FWD]["mixin"])

    class M(rh.HeadStubNet, mixin, models.TruncatedLogisticHead):
        def __init__(self):
            rh.HeadStubNet.__init__(self, S, D, seed)
            mixin.__init__(self, cfg, "cuda")
            self.fix_logistic = fix

        def forward(self, x, t):
            B, Dx = x.shape
            return self.head_forward(self.head_params(x, t), B, Dx)

    m = M().to("cuda")
    m.device = "cuda"
    return m


@pytest.mark.parametrize("case", HEAD_SAMPLERS, ids=[c[0] for c in HEAD_SAMPLERS])
def test_samplers_with_fused_head_match_reference_fixtures(golden, case):
    """Whole reverse process with a model whose forward returns the un-expanded head (ops.LogisticHead): the samplers
    feed it to the fused kernel; final states equal the reference's (its sample_logistic + its sampler) up to ties."""
    from ctdd_b200 import make_config, ops
    from ctdd_b200.lib.sampling import sampling_utils
    import ctdd_b200.lib.sampling.sampling  # noqa: F401
    name, cls, fwd, N, D, loss_name, logit_type, fix, over, max_t, seed = case
    cfg = cases.sampler_cfg(make_config, case)
    cfg.device = "cuda"
    S = cfg.data.S
    m = _head_product_model(cfg, S, D, seed, fix)
    with torch.no_grad():
        xz, tz = torch.zeros((2, D), dtype=torch.long, device="cuda"), torch.ones(2, device="cuda")
        dense = m(xz, tz)                               # a foreign caller gets the reference's (B, D, S) logits tensor
        assert isinstance(dense, torch.Tensor) and dense.shape == (2, D, S)
        with ops.fused_head():                          # the samplers switch the un-materialised head on
            head = m(xz, tz)
            assert isinstance(head, ops.LogisticHead)
            assert torch.equal(head.logits(S), dense)
    fp = oracle_forward(fwd)

    def tables(ts, device):
        t32 = torch.tensor([float(t) for t in ts], dtype=torch.float64).to(torch.float32)
        Q = fp.transition(t32)
        return Q.to(device).contiguous(), Q.transpose(1, 2).contiguous().to(device), [float(b) for b in fp.beta(t32)]

    m.qt0_tables = tables
    cfg.sampler.name = cls
    sampler = sampling_utils.get_sampler(cfg)
    sampler.seed = seed
    args = ()
    if "condition_dim" in over:
        g = np.random.Generator(np.random.PCG64(seed))
        args = (torch.from_numpy(g.integers(0, S, (N, over["condition_dim"]))),)
    res = sampler.sample(m, N, *args)
    res = res if isinstance(res, tuple) else (res,)
    want = golden["head"][f"{name}/x"]
    assert res[0].shape == want.shape
    assert mismatch_fraction(res[0], want) <= 0.011


@pytest.mark.parametrize("S", [256, 32, 5])
def test_head_backward_kernel_matches_fp64_autograd(S):
    """d loss / d mu and d loss / d log_scale through the CUDA backward kernel against fp64 autograd of the reference
    formulas (oracle), for a random upstream gradient; the reference's own fp32 autograd is the noise yardstick."""
    from ctdd_b200 import ops
    N, D = 6, 50
    mu, ls = ho.head_inputs(N * D, 21 + S, -2.5, 2.5)
    g = torch.randn((N, D, S), generator=torch.Generator().manual_seed(5))
    for fix in (False, True):
        m64 = mu.double().view(N, D).requires_grad_(True)
        l64 = ls.double().view(N, D).requires_grad_(True)
        gm64, gl64 = torch.autograd.grad((ho.truncated_logistic_logits(m64, l64, S, fix) * g.double()).sum(), (m64, l64))
        m32 = mu.view(N, D).clone().requires_grad_(True)
        l32 = ls.view(N, D).clone().requires_grad_(True)
        gm32, gl32 = torch.autograd.grad((ho.truncated_logistic_logits(m32, l32, S, fix) * g).sum(), (m32, l32))
        mc = mu.view(N, D).cuda().requires_grad_(True)
        lc = ls.view(N, D).cuda().requires_grad_(True)
        out = ops.logistic_logits_autograd(mc, lc, S, fix)
        assert out.shape == (N, D, S)
        gm, gl = torch.autograd.grad((out * g.cuda()).sum(), (mc, lc))
        for got, want, ref32 in ((gm, gm64, gm32), (gl, gl64, gl32)):
            scale = float(want.abs().max())
            ref_err = float((ref32.double() - want).abs().max())
            err = float((got.cpu().double() - want).abs().max())
            assert err <= max(1e-4 * scale, 2.0 * ref_err), (S, fix, err, ref_err, scale)


def test_head_training_path_is_differentiable_and_matches_kernel():
    """With gradients enabled the head is one autograd.Function (CUDA forward + backward); its values agree with the
    no-grad kernel and the reference formulas, and sample_logistic keeps the reference's signature / layout."""
    from ctdd_b200.lib.models import models
    S = 256
    mu, ls = ho.head_inputs(2 * 3 * 4 * 4, 4, -2.0, 2.0)
    mu = mu.view(2, 3, 4, 4).cuda().requires_grad_(True)
    ls = ls.view(2, 3, 4, 4).cuda().requires_grad_(True)
    for fix in (False, True):
        lg = models.sample_logistic((mu, ls), 2, 3, 48, S, fix, "cuda")
        assert lg.shape == (2, 3, 4, 4, S) and lg.requires_grad
        torch.softmax(lg, -1)[..., 100].sum().backward()
        assert torch.isfinite(mu.grad).all() and torch.isfinite(ls.grad).all()
        with torch.no_grad():
            fast = models.sample_logistic((mu, ls), 2, 3, 48, S, fix, "cuda")
        assert fast.shape == lg.shape
        p_a, p_b = torch.softmax(lg.detach(), -1), torch.softmax(fast, -1)
        assert float((p_a - p_b).abs().max()) <= 2e-6
        l64 = ho.truncated_logistic_logits(mu.detach().cpu().double(), ls.detach().cpu().double(), S, fix)
        assert float((torch.softmax(lg.detach().cpu().double(), -1) - torch.softmax(l64, -1)).abs().max()) <= 2e-6
