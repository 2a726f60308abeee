"""CPU: the drop-in boundary — registry names, call conventions, and INTEGRATION.md option (ii) against the
real reference when it is mounted (build container only; the GPU box does not have /root/reference)."""
import inspect
import os

import pytest


def test_registries_hold_the_reference_names():
    from ctdd_b200.lib.sampling import sampling_utils as su
    from ctdd_b200.lib.losses import losses_utils as lu
    import ctdd_b200.lib.sampling.sampling  # noqa: F401
    import ctdd_b200.lib.losses.losses  # noqa: F401
    for name in ("TauL", "LBJF", "MidPointTauL", "PCTauL", "ConditionalTauLeaping", "ConditionalPCTauLeaping", "ExactSampling"):
        assert name in su._SAMPLERS, name
    for name in ("CTElbo", "NLL", "CTElboLambda", "CondCTElbo", "CatRM", "CatRMNLL", "ScoreElbo", "SDDMElbo", "NLLOriginal"):
        assert name in lu._LOSSES, name
    with pytest.raises(ValueError):                 # duplicate registration is an error, as in the reference
        su.register_sampler(su._SAMPLERS["TauL"])
    with pytest.raises(ValueError):
        lu.register_loss(lu._LOSSES["CTElbo"])


def test_sampler_and_loss_signatures_match_the_reference_conventions():
    from ctdd_b200.lib.sampling import sampling as s
    from ctdd_b200.lib.losses import losses as l
    assert list(inspect.signature(s.TauL.sample).parameters) == ["self", "model", "N"]
    assert list(inspect.signature(s.ConditionalTauLeaping.sample).parameters) == ["self", "model", "N", "conditioner"]
    assert list(inspect.signature(s.get_initial_samples).parameters)[:6] == ["N", "D", "device", "S", "initial_dist", "initial_dist_std"]
    for cls in (l.CTElbo, l.CatRM, l.SDDMElbo, l.ScoreElbo, l.CatRMNLL, l.NLLOriginal):
        assert len(inspect.signature(cls.calc_loss).parameters) >= 3     # (self, a, b[, label/writer]): both orders accepted


def test_no_cpu_fallback_is_offered():
    """Host tensors are refused instead of being routed to a CPU path."""
    import torch
    from ctdd_b200 import _native as nat
    with pytest.raises(RuntimeError):
        nat.ptr(torch.zeros(4))


@pytest.mark.skipif(not os.path.isdir("/root/reference/TAUnSDDM"), reason="reference not mounted")
def test_install_into_reference_overwrites_the_reference_registries():
    from oracle import ref_harness as rh
    ref = rh.import_reference()
    import ctdd_b200
    before = ref.su._SAMPLERS["TauL"]
    samplers, losses = ctdd_b200.install_into_reference()
    from ctdd_b200.lib.sampling import sampling as ours
    assert ref.su._SAMPLERS["TauL"] is ours.TauL and ref.su._SAMPLERS["TauL"] is not before
    assert "CTElbo" in losses and ref.lu._LOSSES["CTElbo"].__module__.startswith("ctdd_b200")
    # put the reference classes back so that other tests that drive the reference see the original registry
    import lib.sampling.sampling as rs
    import lib.losses.losses as rl
    for name in list(ref.su._SAMPLERS):
        if hasattr(rs, name):
            ref.su._SAMPLERS[name] = getattr(rs, name)
    for name in list(ref.lu._LOSSES):
        if hasattr(rl, name):
            ref.lu._LOSSES[name] = getattr(rl, name)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) writes exactly one JSON line to stdout,
    carrying the metric / unit / config of our arm plus cpu_baseline and a zero-copy e2e block."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reverse_step_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0


def test_staged_reference_imports_without_the_reference_checkout():
    """bench.py --impl reference runs on the GPU box, where only baseline/_ref/ (staged by tools/stage_reference.py) exists:
    the staged files must be importable on their own (lib/losses/losses.py pulls lib.d3pm at module level)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(root, "baseline", "_ref", "TAUnSDDM")
    if not os.path.isdir(os.path.join(staged, "lib", "sampling")):
        import pytest
        pytest.skip("no staged reference (baseline/_ref is created by __graft_entry__.build() where /root/reference exists)")
    code = ("import sys; sys.path.insert(0, %r); from oracle import ref_harness as rh; rh.REF_ROOT = %r; "
            "ref = rh.import_reference(); assert ref.ss.__file__.startswith(%r) and ref.ll.__file__.startswith(%r); print('ok')"
            % (root, staged, staged, staged))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
