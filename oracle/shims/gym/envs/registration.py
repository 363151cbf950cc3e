import importlib

_REGISTRY = {}


class EnvSpec:
    def __init__(self, id, entry_point=None, reward_threshold=None, kwargs=None, **_):
        self.id, self.entry_point, self.reward_threshold = id, entry_point, reward_threshold
        self.kwargs = dict(kwargs or {})


def register(id, **kwargs):
    _REGISTRY[id] = EnvSpec(id, **kwargs)


def make(id, **kwargs):
    spec = _REGISTRY[id]
    entry = spec.entry_point
    if callable(entry):
        cls = entry
    else:
        mod_name, attr = entry.split(":")
        cls = getattr(importlib.import_module(mod_name), attr)
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    env = cls(**kw)
    env.spec = spec
    return env
