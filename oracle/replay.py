"""CPU oracle restatement of the episode-major replay buffer and the acting context (test infrastructure only).

Follows dtqn/buffers/replay_buffer.py:19-168 and utils/context.py:8-111.
Intended (numpy-1.x) integer semantics are used for the start draw: max(0, int(eplen) - ctx)
(replay_buffer.py:149-156 wraps under numpy >= 2 because episode_lengths is uint8 -- SURVEY.md A-Q1).
"""
import random

import numpy as np


class ReplayOracle:
    def __init__(self, buffer_size, obs_dim, obs_mask, max_episode_steps, context_len):
        self.max_size = buffer_size // max_episode_steps          # :27
        self.context_len = context_len
        self.obs_dim = obs_dim
        self.max_episode_steps = max_episode_steps
        self.obs_mask = obs_mask
        self.pos = [0, 0]
        S, E = self.max_size, max_episode_steps
        self.obss = np.full([S, E + 1, obs_dim], obs_mask, dtype=np.float32)   # :46-54
        self.actions = np.zeros([S, E + 1, 1], dtype=np.uint8)                 # :57-60
        self.rewards = np.zeros([S, E, 1], dtype=np.float32)                   # :61-64
        self.dones = np.ones([S, E, 1], dtype=np.bool_)                        # :65-68
        self.episode_lengths = np.zeros([S], dtype=np.int64)                   # :69 (uint8 in the reference)

    def store(self, obs, action, reward, done, episode_length=0):              # :71-86
        e, t = self.pos[0] % self.max_size, self.pos[1]
        self.obss[e, t + 1] = obs
        self.actions[e, t] = action
        self.rewards[e, t] = reward
        self.dones[e, t] = done
        self.episode_lengths[e] = episode_length
        self.pos = [self.pos[0], t + 1]

    def store_obs(self, obs):                                                  # :88-92 (+ cleanse :100-135)
        e = self.pos[0] % self.max_size
        self.obss[e] = self.obs_mask
        self.actions[e] = 0
        self.rewards[e] = 0
        self.dones[e] = True
        self.episode_lengths[e] = 0
        self.obss[e, 0] = obs

    def can_sample(self, batch_size):                                          # :94-95
        return batch_size < self.pos[0]

    def flush(self):                                                           # :97-98
        self.pos = [self.pos[0] + 1, 0]

    def draw_indices(self, batch_size):
        """Index draws of sample() (:141-156) with CPython ``random`` (MT19937)."""
        valid = [i for i in range(min(self.pos[0], self.max_size)) if i != self.pos[0] % self.max_size]
        eps = np.array([random.choice(valid) for _ in range(batch_size)], dtype=np.int64)
        starts = np.array([random.randint(0, max(0, int(self.episode_lengths[e]) - self.context_len)) for e in eps],
                          dtype=np.int64)
        return eps, starts

    def gather(self, eps, starts):
        """The payload gather of sample() (:157-168) for given (episode, start) pairs."""
        eps = np.asarray(eps, dtype=np.int64)[:, None]
        tr = np.asarray(starts, dtype=np.int64)[:, None] + np.arange(self.context_len)[None, :]
        return (self.obss[eps, tr], self.actions[eps, tr], self.rewards[eps, tr],
                self.obss[eps, 1 + tr], self.actions[eps, 1 + tr], self.dones[eps, tr],
                np.clip(self.episode_lengths[eps], 0, self.context_len))

    def sample(self, batch_size):
        return self.gather(*self.draw_indices(batch_size))


class ContextOracle:
    """utils/context.py Context.  ``obs`` inherits the dtype of np.full(obs_mask) = int64 for both envs
    (:46), so CarFlag observations are truncated toward zero when written."""

    def __init__(self, context_length, obs_mask, num_actions, obs_dim, rng):
        self.max_length, self.obs_mask, self.num_actions, self.obs_dim = context_length, obs_mask, num_actions, obs_dim
        self.rng = rng            # the global RNG.rng stand-in: an oracle.pcg64.PCG64
        self.timestep = 0

    def reset(self, obs):                                                       # :36-54
        self.obs = np.full([self.max_length, self.obs_dim], self.obs_mask)
        self.obs[0] = obs
        self.action = np.array([[self.rng.integers(self.num_actions)] for _ in range(self.max_length)], dtype=np.int64)
        self.timestep = 0

    def add_transition(self, o, a):                                             # :56-80
        self.timestep += 1
        if self.timestep >= self.max_length:                                    # roll :94-96
            self.obs = np.roll(self.obs, -1, axis=0)
            self.action = np.roll(self.action, -1, axis=0)
        t = min(self.timestep, self.max_length - 1)
        self.obs[t] = o
        self.action[t] = a

    def window(self):
        """The slice fed to the network by get_action (dtqn/agents/dtqn.py:81-92)."""
        n = min(self.max_length, self.timestep + 1)
        return self.obs[:n], self.action[:n]
