"""Batched lockstep POMDP environments on the device -- host-side mirror of envs/car_flag.py + envs/memory_cards.py
behind gym's TimeLimit (ids registered at envs/__init__.py:31-48).

``BatchedEnv`` owns the SoA state tensors of ``n_envs`` independent instances; instance i is seeded exactly like the
reference seeds one env -- ``Generator(PCG64(SeedSequence(seed_i)))`` (car_flag.py:70-74) -- with the seeding done on
the host BY numpy and only the raw PCG64 state uploaded.  A second per-instance stream plays the role of the
reference's global ``RNG.rng`` (utils/random.py:31: ``PCG64(seed)``, identical initial state) for the agent-side
draws (random / eps-greedy actions, Context.reset padding).
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from dtqn_b200 import _lib

ENV_SPECS = {
    # id: (kind, obs_dim, num_actions, max_episode_steps, obs_mask, discrete_obs)
    "DiscreteCarFlag-v0": (_lib.ENV_CARFLAG, 3, 3, 200, -5.0, False),
    "Memory-5-v0": (_lib.ENV_MEMORY, 10, 10, 50, 8.0, True),
}


def pcg64_states(seeds) -> tuple:
    """numpy does the SeedSequence hashing; returns ([4, n] uint64 words, [2, n] uint32 buffer) as numpy arrays."""
    n = len(seeds)
    words = np.zeros((4, n), dtype=np.uint64)
    buf = np.zeros((2, n), dtype=np.uint32)
    mask = (1 << 64) - 1
    for i, s in enumerate(seeds):
        st = np.random.PCG64(np.random.SeedSequence(int(s))).state
        words[0, i], words[1, i] = st["state"]["state"] >> 64, st["state"]["state"] & mask
        words[2, i], words[3, i] = st["state"]["inc"] >> 64, st["state"]["inc"] & mask
        buf[0, i], buf[1, i] = st["has_uint32"], st["uinteger"]
    return words, buf


def _u64(a: np.ndarray, dev) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int64).copy()).to(dev)


class BatchedEnv:
    """n_envs lockstep instances of one registered env id, stepped by the sm_100a kernels in csrc/env.cu."""

    def __init__(self, env_id: str, n_envs: int, seed: int = 1, device=None, seeds=None,
                 max_episode_steps: Optional[int] = None, agent_stream_of: Optional["BatchedEnv"] = None):
        """``agent_stream_of``: share another instance's agent-side streams (the reference has ONE global RNG.rng, drawn by
        the training loop and by run.evaluate alike -- utils/random.py:31, agents/dtqn.py:78, utils/context.py:50)."""
        if env_id not in ENV_SPECS:
            raise ValueError(f"Environment with id {env_id} not found (hot path covers {sorted(ENV_SPECS)})")
        dev = _lib.require_cuda(device)
        kind, O, A, max_steps, obs_mask, discrete = ENV_SPECS[env_id]
        self.env_id, self.kind, self.n_envs, self.device = env_id, kind, int(n_envs), dev
        self.obs_dim, self.num_actions, self.obs_mask, self.discrete = O, A, obs_mask, discrete
        self.max_episode_steps = int(max_episode_steps or max_steps)
        self.seeds = np.asarray(seeds if seeds is not None else np.arange(seed, seed + n_envs), dtype=np.int64)
        assert len(self.seeds) == n_envs
        words, buf = pcg64_states(self.seeds)
        n = self.n_envs
        self.rng, self.rng_buf = _u64(words, dev), torch.from_numpy(buf.view(np.int32).copy()).to(dev)
        if agent_stream_of is not None:
            assert agent_stream_of.n_envs == self.n_envs
            self.arng, self.arng_buf = agent_stream_of.arng, agent_stream_of.arng_buf
        else:
            self.arng, self.arng_buf = self.rng.clone(), self.rng_buf.clone()
        self.pos = torch.zeros(n, dtype=torch.float64, device=dev)
        self.vel = torch.zeros(n, dtype=torch.float64, device=dev)
        self.heaven = torch.ones(n, dtype=torch.int8, device=dev)
        self.cards = torch.zeros(n, dtype=torch.int64, device=dev)
        self.shown = torch.zeros(n, dtype=torch.int64, device=dev)
        self.cur = torch.zeros(n, dtype=torch.int32, device=dev)
        self.elapsed = torch.zeros(n, dtype=torch.int32, device=dev)
        self.done_flag = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.block_counts = torch.zeros((n + 255) // 256, dtype=torch.int32, device=dev)
        self.ep_stats = torch.zeros(4, dtype=torch.int64, device=dev)
        self.ep_return = torch.zeros(n, dtype=torch.int32, device=dev)
        self.env_acc = torch.zeros((n, 4), dtype=torch.int32, device=dev)   # per env: episodes, return, length, successes
        # step outputs (reused every step)
        self.actions = torch.zeros(n, dtype=torch.int32, device=dev)
        self.obs_out = torch.zeros((n, O), dtype=torch.float32, device=dev)
        self.reward_out = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done_out = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.truncated_out = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.success_out = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.struct = _lib.EnvStruct(
            kind=kind, n_envs=n, obs_dim=O, num_actions=A, max_episode_steps=self.max_episode_steps,
            stat_episodes_per_env=0,
            rng=_lib.ptr(self.rng), rng_buf=_lib.ptr(self.rng_buf), arng=_lib.ptr(self.arng),
            arng_buf=_lib.ptr(self.arng_buf), pos=_lib.ptr(self.pos), vel=_lib.ptr(self.vel),
            heaven=_lib.ptr(self.heaven), cards=_lib.ptr(self.cards), shown=_lib.ptr(self.shown),
            cur=_lib.ptr(self.cur), elapsed=_lib.ptr(self.elapsed), done_flag=_lib.ptr(self.done_flag),
            block_counts=_lib.ptr(self.block_counts), ep_stats=_lib.ptr(self.ep_stats),
            ep_return=_lib.ptr(self.ep_return), env_acc=_lib.ptr(self.env_acc))
        self.replay = None
        self.context = None

    def attach(self, replay=None, context=None) -> None:
        """Fuse ``agent.observe`` into the step: replay = dtqn_b200.buffers.ReplayBuffer, context = ContextWindow."""
        self.replay, self.context = replay, context

    def _rb(self):
        return C.byref(self.replay.struct) if self.replay is not None else None

    def _cx(self):
        return C.byref(self.context.struct) if self.context is not None else None

    def count_episodes_per_env(self, k: int) -> None:
        """k > 0: only each env's first k finished episodes (since reset_all) enter ep_stats / env_acc; 0: all of them."""
        self.struct.stat_episodes_per_env = int(k)

    def reset_all(self) -> None:
        _lib.check(_lib.lib.dtqn_env_reset_all(C.byref(self.struct), self._rb(), self._cx(), _lib.stream_ptr()),
                   "dtqn_env_reset_all")

    def step(self, actions: Optional[torch.Tensor] = None, mode: Optional[int] = None, epsilon: float = 0.0,
             q_last: Optional[torch.Tensor] = None, record: bool = True,
             epsilon_dev: Optional[torch.Tensor] = None) -> None:
        """One lockstep step of every instance (run.py:356-377 fused); results land in obs_out / reward_out /
        done_out / truncated_out / success_out / actions.  ``record=False`` skips the replay (evaluation)."""
        if mode is None:
            mode = _lib.ACT_GIVEN if actions is not None else _lib.ACT_RANDOM
        if actions is not None:
            self.actions.copy_(actions.to(torch.int32), non_blocking=True)
        io = _lib.StepIO(action_mode=mode, epsilon=float(epsilon), actions=_lib.ptr(self.actions),
                         q_last=_lib.ptr(q_last) if q_last is not None else None,
                         obs_out=_lib.ptr(self.obs_out), reward_out=_lib.ptr(self.reward_out),
                         done_out=_lib.ptr(self.done_out), truncated_out=_lib.ptr(self.truncated_out),
                         success_out=_lib.ptr(self.success_out),
                         epsilon_dev=_lib.ptr(epsilon_dev) if epsilon_dev is not None else None)
        _lib.check(_lib.lib.dtqn_env_step(C.byref(self.struct), self._rb() if record else None, self._cx(),
                                          C.byref(io), _lib.stream_ptr()), "dtqn_env_step")

    def current_obs(self) -> torch.Tensor:
        """The observation the next action conditions on, [n, O] float64 (what env.reset()/env.step() last returned)."""
        if self.kind == _lib.ENV_CARFLAG:
            p = self.pos
            d = torch.where((p >= 0.5 - 0.2) & (p <= 0.5 + 0.2), self.heaven.double(), torch.zeros_like(p))
            return torch.stack([p, self.vel, d], dim=1)
        sh = torch.arange(10, device=self.device, dtype=torch.int64) * 4
        return ((self.shown[:, None] >> sh[None, :]) & 0xF).double()

    def rng_state(self) -> np.ndarray:
        """[n, 6] uint64 (state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger) of the env streams."""
        w = self.rng.cpu().numpy().view(np.uint64)
        b = self.rng_buf.cpu().numpy().view(np.uint32).astype(np.uint64)
        return np.concatenate([w, b], axis=0).T.copy()


class ContextWindow:
    """Device mirror of utils/context.py Context for n_envs instances: obs ring [n, ctx, O] + timestep [n]."""

    def __init__(self, context_len: int, obs_mask: float, num_actions: int, env_obs_length: int, n_envs: int = 1,
                 device=None, trunc_obs: bool = True):
        dev = _lib.require_cuda(device)
        self.max_length, self.obs_mask, self.num_actions = int(context_len), float(obs_mask), int(num_actions)
        self.env_obs_length, self.n_envs, self.device = int(env_obs_length), int(n_envs), dev
        self.obs = torch.full((n_envs, context_len, env_obs_length), float(obs_mask), dtype=torch.float32, device=dev)
        self.timestep_t = torch.zeros(n_envs, dtype=torch.int32, device=dev)
        self.action = torch.zeros((n_envs, context_len), dtype=torch.uint8, device=dev)     # Context.action ring (a_embed > 0)
        self.trunc_obs = bool(trunc_obs)
        self.struct = _lib.ContextStruct(context_len=self.max_length, obs_dim=self.env_obs_length,
                                         trunc_obs=int(self.trunc_obs), obs_mask=self.obs_mask,
                                         obs=_lib.ptr(self.obs), timestep=_lib.ptr(self.timestep_t),
                                         action=_lib.ptr(self.action))

    @property
    def timestep(self) -> int:
        """Context.timestep of env 0 (run.py:231 reads it after an evaluation episode)."""
        return int(self.timestep_t[0].item())

    def windows(self):
        """Dense [n, ctx, O] windows in temporal order + valid lengths [n] (host-side helper for tests)."""
        t = self.timestep_t.long()
        n = torch.clamp(t + 1, max=self.max_length)
        j = torch.arange(self.max_length, device=self.device)[None, :]
        idx = (t[:, None] + 1 - n[:, None] + j) % self.max_length
        return torch.gather(self.obs, 1, idx[..., None].expand(-1, -1, self.env_obs_length)), n
