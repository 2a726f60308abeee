"""GPU (needs 2 devices; skipped otherwise): the kernels launch on the CURRENT device's stream and keep per-device state,
so a model on cuda:1 while cuda:0 is current must either work (the classes and ops wrappers switch the device) or fail
loudly (raw _native.ptr) - never touch cuda:0's stream with cuda:1's pointers."""
import numpy as np
import pytest
import torch

from oracle import cases
from helpers import product_model

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
def test_model_on_second_device_while_first_is_current():
    from ctdd_b200 import _native as nat, make_config, ops
    from ctdd_b200.lib.sampling import sampling_utils
    import ctdd_b200.lib.sampling.sampling  # noqa: F401
    torch.cuda.set_device(0)
    case = cases.SAMPLERS[0]                       # TauL, gauss256
    name, cls, fwd, N, D, loss_name, logit_type, stub, over, max_t, seed = case
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        cfg = cases.sampler_cfg(make_config, case)
        cfg.device = dev
        m = product_model(fwd, cfg, D, seed, stub[0], stub[1], device=dev)
        sampler = sampling_utils.get_sampler(cfg)
        sampler.seed = seed
        assert torch.cuda.current_device() == 0
        outs.append(np.asarray(sampler.sample(m, N)[0]))
        assert torch.cuda.current_device() == 0    # the guard restores the caller's device
    assert np.array_equal(outs[0], outs[1])        # Philox + kernels are device independent
    # ops wrappers switch to the device of their tensors; the raw pointer helper refuses the mismatch
    x = torch.zeros((4, 256), device="cuda:1")
    assert ops.bgemm256(x.view(1, 4, 256), torch.eye(256, device="cuda:1").view(1, 256, 256)).device.index == 1
    with pytest.raises(RuntimeError):
        nat.ptr(x)
