"""CPU: the C-ABI library loads and exports every symbol include/ctdd.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctdd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctdd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ctdd_b200 import _native as nat
    L = nat.lib()
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ctdd.h but not exported"
    assert L.ctdd_version() == 1


def test_struct_layout_matches_header():
    """ctypes mirror of ctdd_step_params has the size the C compiler gives it."""
    import subprocess, tempfile
    from ctdd_b200 import _native as nat
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include "ctdd.h"\nint main(){printf("%zu %zu\\n", sizeof(ctdd_step_params), sizeof(ctdd_loss_params));return 0;}\n')
        exe = os.path.join(d, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(nat.StepParams) == a
    assert ctypes.sizeof(nat.LossParams) == b


def test_no_cpu_fallback():
    """Host tensors are refused loudly; registries and config keys behave like the reference's."""
    from ctdd_b200 import _native as nat, make_config
    from ctdd_b200.lib.models import forward_model as fm
    from ctdd_b200.lib.sampling import sampling_utils
    import ctdd_b200.lib.sampling.sampling  # noqa: F401  (registers)
    with pytest.raises(RuntimeError):
        nat.ptr(torch.zeros(3))
    cfg = make_config(data=dict(S=8), model=dict(rate_sigma=2.0, Q_sigma=5.0, time_exp=10.0, time_base=1.0), device="cpu")
    m = fm.GaussianTargetRate(cfg, "cpu")
    assert m.base_rate.shape == (8, 8)
    with pytest.raises(RuntimeError):
        m.transition(torch.tensor([0.5]))
    for name in ("TauL", "LBJF", "MidPointTauL", "PCTauL", "ConditionalTauLeaping", "ConditionalPCTauLeaping"):
        assert name in sampling_utils._SAMPLERS
    with pytest.raises(ValueError):
        sampling_utils.register_sampler(sampling_utils._SAMPLERS["TauL"])
    with pytest.raises(KeyError):
        sampling_utils.get_sampler(make_config(sampler=dict(name="NoSuchSampler")))


def test_base_rate_matches_reference_bits(golden):
    """Vectorised Gaussian base-rate construction equals the reference's nested-loop result bit for bit."""
    from ctdd_b200 import make_config
    from ctdd_b200.lib.models import forward_model as fm
    from oracle import cases
    for name in ("gauss256", "gauss32"):
        f = cases.FORWARD[name]
        cfg = make_config(data=dict(S=f["S"]), model=f["model"], device="cpu")
        m = fm.GaussianTargetRate(cfg, "cpu")
        np.testing.assert_array_equal(m.base_rate.numpy(), golden["forward"][f"{name}/base_rate"])
