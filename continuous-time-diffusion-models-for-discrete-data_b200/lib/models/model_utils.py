"""Model registry + get_logprob_with_logits — same contract as the reference's lib/models/model_utils.py:5-60."""
import torch
import torch.nn.functional as F

from ..utils import utils

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


def get_logprob_with_logits(cfg, model, xt, t, logits, xt_target=None):
    """(log_prob (B,D,S), log_xt (B,D)) by loss.logit_type (reference model_utils.py:30-60).

    Differentiable torch composition kept for API compatibility (callers outside the fused kernels); the samplers
    and loss classes of this package do not go through it — they use the fused CUDA kernels."""
    if xt_target is None:
        xt_target = xt
    xt_onehot = F.one_hot(xt_target.long(), cfg.data.S)
    if cfg.loss.logit_type == "direct":
        log_prob = F.log_softmax(logits, dim=-1)
    else:
        qt0 = model.transition(t)
        if cfg.loss.logit_type == "reverse_prob":
            p0t = F.softmax(logits, dim=-1)
            qt0 = utils.expand_dims(qt0, axis=list(range(1, xt.dim() - 1)))
            log_prob = torch.log(p0t @ qt0 + 1e-35)
        elif cfg.loss.logit_type == "reverse_logscale":
            log_p0t = F.log_softmax(logits, dim=-1)
            log_qt0 = torch.where(qt0 <= 1e-35, -1e9, torch.log(qt0))
            log_qt0 = utils.expand_dims(log_qt0, axis=list(range(1, xt.dim())))
            log_prob = torch.logsumexp(log_p0t.unsqueeze(-1) + log_qt0, dim=-2)
        else:
            raise ValueError("Unknown logit_type: %s" % cfg.loss.logit_type)
    log_xt = torch.sum(log_prob * xt_onehot, dim=-1)
    return log_prob, log_xt
