"""Host-side glue mirroring the reference's utils/ package for the hot path:
``LinearAnneal`` (utils/epsilon_anneal.py:18-34), ``set_global_seed`` (utils/random.py:13-31) and ``get_agent``
(utils/agent_utils.py:36-168)."""
import os
import random
from typing import Sequence

import numpy as np
import torch

from dtqn_b200.agents import RNG, DtqnAgent
from dtqn_b200.envs import ENV_SPECS, BatchedEnv
from dtqn_b200.networks import DTQN


class Constant:
    def __init__(self, start):
        self.val = start

    def anneal(self):
        pass


class LinearAnneal:
    """Named "linear" in the reference but geometric toward ``end`` (utils/epsilon_anneal.py:33-34, SURVEY A-Q6)."""

    def __init__(self, start: float, end: float, duration: int):
        self.val, self.min, self.duration = start, end, duration

    def anneal(self):
        self.val = max(self.min, self.val - (self.val - self.min) / self.duration)


def set_global_seed(seed: int, *envs) -> None:
    """utils/random.py:13-31 (with ``10**6`` instead of the ``1e6`` that raises on Python >= 3.12, SURVEY A-Q8).
    Device envs are seeded at construction (seed + i per instance); host gym-style envs get ``env.seed(seed)``."""
    random.seed(seed)
    tseed, npseed, ospyseed = random.randint(1, 10**6), random.randint(1, 10**6), random.randint(1, 10**6)
    torch.manual_seed(tseed)
    np.random.seed(npseed)
    for env in envs:
        if hasattr(env, "seed"):
            env.seed(seed=seed)
    os.environ["PYTHONHASHSEED"] = str(ospyseed)
    RNG.rng = np.random.Generator(np.random.PCG64(seed=seed))


def get_agent(model_str: str, envs: Sequence, embed_per_obs_dim: int, action_dim: int, inner_embed: int,
              buffer_size: int, device, learning_rate: float, batch_size: int, context_len: int, max_env_steps: int,
              history: int, target_update_frequency: int, gamma: float, num_heads: int = 1, num_layers: int = 1,
              dropout: float = 0.0, identity: bool = False, gate: str = "res", pos: str = "learned", bag_size: int = 0,
              n_envs: int = None, **agent_kwargs) -> DtqnAgent:
    """utils/agent_utils.py:36-168 for the DTQN model family.  ``envs[0]`` is a ``BatchedEnv`` (or anything exposing
    obs_dim / num_actions / obs_mask / discrete / max_episode_steps)."""
    if model_str not in ("DTQN",):
        raise NotImplementedError(f"{model_str}: only DTQN is on the B200 hot path (baselines: SURVEY.md section 2 #21-22)")
    env = envs[0]
    env_obs_length, env_obs_mask = env.obs_dim, env.obs_mask
    if max_env_steps <= 0:
        max_env_steps = max(e.max_episode_steps for e in envs)
    obs_vocab_size = int(env_obs_mask) + 1                                   # agent_utils.py:92-95
    is_discrete_env = bool(env.discrete)
    if history < 1 or history > context_len:                                  # agent_utils.py:101-105
        print(f"History must be 1 < history <= context_len, but history is {history} and context len is {context_len}. "
              f"Clipping history to {np.clip(history, 1, context_len)}...")
        history = int(np.clip(history, 1, context_len))
    num_actions = env.num_actions
    n_envs = n_envs if n_envs is not None else getattr(env, "n_envs", 1)

    def network_factory():
        return DTQN(env_obs_length, num_actions, embed_per_obs_dim, action_dim, inner_embed, num_heads, num_layers,
                    context_len, dropout=dropout, gate=gate, identity=identity, pos=pos, discrete=is_discrete_env,
                    vocab_sizes=obs_vocab_size, target_update_frequency=target_update_frequency, bag_size=bag_size,
                    device=device)

    return DtqnAgent(network_factory, buffer_size, device, env_obs_length, max_env_steps, env_obs_mask, num_actions,
                     is_discrete_env, learning_rate=learning_rate, batch_size=batch_size, gamma=gamma,
                     context_len=context_len, embed_size=inner_embed, history=history,
                     target_update_frequency=target_update_frequency, bag_size=bag_size, n_envs=n_envs, **agent_kwargs)
