"""ctdd_b200 — B200-native (sm_100a) reverse-CTMC hot path for continuous-time discrete diffusion.

Drop-in for the hot path of paulffm/Continuous-Time-Diffusion-Models-for-Discrete-Data (TAUnSDDM):
`lib.sampling.sampling`, `lib.losses.losses`, `lib.models.forward_model`, `lib.models.model_utils` and the three
registries keep the reference's names/signatures; the arithmetic runs in hand-written CUDA behind the C ABI of
include/ctdd.h (libctdd_b200.so).  See DESIGN.md / INTEGRATION.md.
"""
from . import _native  # noqa: F401
from .config import ConfigDict, make_config  # noqa: F401

__all__ = ["ConfigDict", "make_config", "install_into_reference", "lib"]
__version__ = "0.1.0"


def install_into_reference():
    """Overwrite the reference's registries with the ctdd_b200 classes (option ii).

    Call after the reference's `lib.sampling.sampling` / `lib.losses.losses` were imported by a train script:
    `_SAMPLERS[name]` / `_LOSSES[name]` then resolve to the CUDA-backed classes, and `get_sampler(cfg)` /
    `get_loss(cfg)` in the unmodified scripts pick them up."""
    import importlib
    import sys

    from .lib.sampling import sampling as _s, sampling_utils as _su
    from .lib.losses import losses as _l, losses_utils as _lu
    ref_su = sys.modules.get("lib.sampling.sampling_utils") or importlib.import_module("lib.sampling.sampling_utils")
    ref_lu = sys.modules.get("lib.losses.losses_utils") or importlib.import_module("lib.losses.losses_utils")
    if ref_su is not _su:
        ref_su._SAMPLERS.update(_su._SAMPLERS)
    if ref_lu is not _lu:
        ref_lu._LOSSES.update(_lu._LOSSES)
    return sorted(_su._SAMPLERS), sorted(_lu._LOSSES)
