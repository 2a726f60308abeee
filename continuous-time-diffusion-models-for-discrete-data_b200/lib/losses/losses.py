"""Training losses — placeholder until the fused loss kernels land (see ctdd.h ctdd_loss_*)."""
from . import losses_utils  # noqa: F401
