"""The `Standard` train step (reference lib/training/training.py:8-40): zero_grad, loss, backward, clip, warm-up,
optimiser step, EMA.  Same class name, constructor keys and return values.

What differs from the reference is only where the time goes: the loss is the fused-kernel loss of this package, the
NaN/inf guard is ONE device->host read (`isfinite`) instead of two (`isnan().any()`, `isinf().any()`), and the EMA is
one kernel launch (`EMA.update_ema` -> `ctdd_ema_update`) instead of three per parameter tensor.

The reference's scripts call the step two ways — `step(state, loss, minibatch, label)` (train_image.py:106) and
`step(state, minibatch, loss)` (train_maze.py:113, train_synthetic.py:103); both are accepted (the loss object is
the argument that has `calc_loss`)."""
import numpy as np
import torch

from . import training_utils


@training_utils.register_train_step
class Standard:
    def __init__(self, cfg):
        self.do_ema = "ema_decay" in cfg.model
        self.clip_grad = cfg.training.clip_grad
        self.grad_norm = cfg.training.grad_norm
        self.warmup = cfg.training.warmup
        self.lr = cfg.optimizer.lr
        self.device = cfg.device

    def step(self, state, loss, minibatch, label=None):
        if hasattr(minibatch, "calc_loss") and not hasattr(loss, "calc_loss"):
            loss, minibatch = minibatch, loss
        state["optimizer"].zero_grad()
        l = loss.calc_loss(state, minibatch, label)
        if not bool(torch.isfinite(l).all()):
            print("Loss is nan or inf")
            return torch.tensor(1e9, device=self.device)
        l.backward()
        if self.clip_grad:
            torch.nn.utils.clip_grad_norm_(state["model"].parameters(), self.grad_norm)
        if self.warmup > 0:
            for g in state["optimizer"].param_groups:
                g["lr"] = self.lr * np.minimum(state["n_iter"] / self.warmup, 1.0)
        state["optimizer"].step()
        if self.do_ema:
            state["model"].update_ema()
        return l.detach()
