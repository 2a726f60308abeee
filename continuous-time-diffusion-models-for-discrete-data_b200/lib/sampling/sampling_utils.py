"""Sampler registry — same contract as the reference's lib/sampling/sampling_utils.py:1-11."""
_SAMPLERS = {}


def register_sampler(cls):
    name = cls.__name__
    if name in _SAMPLERS:
        raise ValueError(f'{name} is already registered!')
    _SAMPLERS[name] = cls
    return cls


def register_alias(name, cls):
    """Legacy sampler names that configs use but the reference never registered."""
    _SAMPLERS.setdefault(name, cls)


def get_sampler(cfg):
    return _SAMPLERS[cfg.sampler.name](cfg)
