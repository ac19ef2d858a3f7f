"""Harness that imports the UNMODIFIED reference (kevslinger/DTQN at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/gen_golden.py`` to produce the committed golden
fixtures and by the ``not gpu`` tests (when /root/reference is present) to re-validate the oracle port.
/root/reference does not exist on the GPU box: nothing on the ``-m gpu`` / smoke / bench path imports this.

Recipe:
  * stub ``gym`` 0.18 + ``pyglet`` packages (``stubs/``) go first on sys.path;
  * /root/reference goes on sys.path (read-only, PYTHONDONTWRITEBYTECODE so no .pyc is written);
  * shim 1: ``utils.random.set_global_seed`` uses ``random.randint(1, 1e6)`` which raises on
    Python >= 3.12 (utils/random.py:21-23) -> ``set_global_seed`` below restates it with 10**6;
  * shim 2: ``ReplayBuffer.episode_lengths`` is uint8 and ``uint8 - int`` wraps under numpy >= 2
    (replay_buffer.py:152) -> widen the array to int64 after constructing the agent.
"""
import os
import random
import sys

REFERENCE_ROOT = os.environ.get("DTQN_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dtqn", "networks", "dtqn.py"))


def activate() -> None:
    """Put the stubs and the reference on sys.path and register the reference's gym ids."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    os.environ.setdefault("WANDB_MODE", "disabled")
    for p in (REFERENCE_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import envs  # noqa: F401  (reference package: registers Memory-5-v0 / DiscreteCarFlag-v0)


def set_global_seed(seed: int, *envs_) -> None:
    """Restatement of utils/random.py:13-31 with the Python-3.12 fix (10**6 instead of 1e6)."""
    import numpy as np
    import torch
    from utils.random import RNG

    random.seed(seed)
    tseed = random.randint(1, 10**6)
    npseed = random.randint(1, 10**6)
    ospyseed = random.randint(1, 10**6)
    torch.manual_seed(tseed)
    np.random.seed(npseed)
    for env in envs_:
        env.seed(seed=seed)
        env.observation_space.seed(seed=seed)
        env.action_space.seed(seed=seed)
    os.environ["PYTHONHASHSEED"] = str(ospyseed)
    RNG.rng = np.random.Generator(np.random.PCG64(seed=seed))


def widen_episode_lengths(agent) -> None:
    """Shim 2 (numpy >= 2): restore numpy-1.x arithmetic in ReplayBuffer.sample without editing it."""
    import numpy as np

    agent.replay_buffer.episode_lengths = agent.replay_buffer.episode_lengths.astype(np.int64)
