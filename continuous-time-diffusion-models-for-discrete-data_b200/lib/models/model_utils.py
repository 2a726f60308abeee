"""Model registry — same contract as the reference's lib/models/model_utils.py:5-27.

The reference's get_logprob_with_logits (model_utils.py:30-60: log-probabilities of the three logit types) has no
stand-alone counterpart here: it is evaluated inside the fused kernels (csrc/ctdd_step_*.cu for the samplers,
csrc/ctdd_loss.cu for the losses, all three `loss.logit_type` values), which never materialise its (B,D,S[,S]) tensors."""

_MODELS = {}


def register_model(cls):
    name = cls.__name__
    if name in _MODELS:
        raise ValueError(f"{name} is already registered!")
    _MODELS[name] = cls
    return cls


def get_model(name):
    return _MODELS[name]


def create_model(cfg, device, encoding=None, rank=None):
    if encoding is None:
        model = get_model(cfg.model.name)(cfg, device, rank)
    else:
        model = get_model(cfg.model.name)(cfg, device, encoding, rank)
    return model.to(device)
