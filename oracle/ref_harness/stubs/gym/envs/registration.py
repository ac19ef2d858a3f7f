import importlib
from gym import error

_REGISTRY = {}


def register(id, entry_point=None, kwargs=None, max_episode_steps=None, **_):
    _REGISTRY[id] = dict(entry_point=entry_point, kwargs=kwargs or {}, max_episode_steps=max_episode_steps)


def make(id, **kw):
    if id not in _REGISTRY:
        raise error.Error(f"No registered env with id: {id}")
    spec = _REGISTRY[id]
    mod_name, cls_name = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    env = cls(**{**spec["kwargs"], **kw})
    if spec["max_episode_steps"] is not None:
        from gym.wrappers.time_limit import TimeLimit
        env = TimeLimit(env, max_episode_steps=spec["max_episode_steps"])
    return env
