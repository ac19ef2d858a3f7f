"""Device-resident replay buffer -- host-side mirror of ``dtqn.buffers.replay_buffer.ReplayBuffer``.

Same constructor, attributes (``obss/actions/rewards/dones/episode_lengths/max_size/context_len``) and methods
(``store/store_obs/flush/can_sample/sample``) as dtqn/buffers/replay_buffer.py:8-168, but every array is a CUDA
tensor, the hot store path is fused into the env-step kernels (``dtqn_env_step``) for ``n_envs`` lockstep
environments, and ``sample`` is the ``dtqn_replay_sample_indices`` + ``dtqn_replay_gather`` kernels.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from dtqn_b200 import _lib


class ReplayBuffer:
    def __init__(self, buffer_size: int, env_obs_length: int, obs_mask: float, max_episode_steps: int,
                 context_len: Optional[int] = 1, n_envs: int = 1, device=None, sample_seed: int = 0, record_every: int = 1):
        dev = _lib.require_cuda(device)
        if isinstance(env_obs_length, tuple):
            raise NotImplementedError("image observations are outside the hot path (SURVEY.md section 2 #9)")
        self.max_size = buffer_size // max_episode_steps                      # replay_buffer.py:27
        if self.max_size < n_envs:
            raise ValueError(f"replay ring of {self.max_size} episode slots is smaller than n_envs={n_envs}")
        self.context_len = int(context_len)
        self.env_obs_length = int(env_obs_length)
        self.max_episode_steps = int(max_episode_steps)
        self.obs_mask = float(obs_mask)
        self.n_envs = int(n_envs)
        self.device = dev
        S, E, O = self.max_size, self.max_episode_steps, self.env_obs_length
        self.obss = torch.full((S, E + 1, O), self.obs_mask, dtype=torch.float32, device=dev)   # :46-54
        self.actions = torch.zeros((S, E + 1, 1), dtype=torch.uint8, device=dev)               # :57-60
        self.rewards = torch.zeros((S, E, 1), dtype=torch.float32, device=dev)                 # :61-64
        self.dones = torch.ones((S, E, 1), dtype=torch.uint8, device=dev)                      # :65-68 (bool)
        self.episode_lengths = torch.zeros((S,), dtype=torch.int32, device=dev)                # :69
        self.slot_open = torch.zeros((S,), dtype=torch.uint8, device=dev)
        self.counters = torch.zeros((4,), dtype=torch.int64, device=dev)   # started(cur), started(next), completed, dropped
        self.env_slot = torch.full((n_envs,), -1, dtype=torch.int32, device=dev)
        self.env_prev_len = torch.zeros((n_envs,), dtype=torch.int32, device=dev)
        self.draw_counter = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.sample_seed = int(sample_seed)
        self.record_every = int(record_every)       # K > 1: each env stores every K-th of its episodes (longer replay horizon)
        self._host_pos = [0, 0]            # used only by the single-env host-call API (store / flush)
        self._completed_seen = 0
        self.struct = _lib.ReplayStruct(
            n_slots=S, max_episode_steps=E, obs_dim=O, context_len=self.context_len, obs_mask=self.obs_mask,
            record_every=int(record_every),
            obss=_lib.ptr(self.obss), actions=_lib.ptr(self.actions), rewards=_lib.ptr(self.rewards),
            dones=_lib.ptr(self.dones), episode_lengths=_lib.ptr(self.episode_lengths),
            slot_open=_lib.ptr(self.slot_open), counters=_lib.ptr(self.counters),
            env_slot=_lib.ptr(self.env_slot), env_prev_len=_lib.ptr(self.env_prev_len))

    # ---- reference-compatible host-call API (one env, one call per transition; plumbing, not the hot path) ------
    @property
    def pos(self):
        """[episodes completed, transitions in the open episode] like ReplayBuffer.pos when driven by the host API;
        for the fused multi-env path use ``num_completed()``."""
        return list(self._host_pos)

    def store(self, obs, action, reward, done, episode_length: Optional[int] = 0) -> None:        # :71-86
        e, t = self._host_pos[0] % self.max_size, self._host_pos[1]
        self.obss[e, t + 1] = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
        self.actions[e, t] = int(action)
        self.rewards[e, t] = float(reward)
        self.dones[e, t] = int(bool(done))
        self.episode_lengths[e] = int(episode_length)
        self._host_pos = [self._host_pos[0], t + 1]

    def store_obs(self, obs) -> None:                                                             # :88-92,100-135
        e = self._host_pos[0] % self.max_size
        self.obss[e] = self.obs_mask
        self.actions[e] = 0
        self.rewards[e] = 0
        self.dones[e] = 1
        self.episode_lengths[e] = 0
        self.slot_open[e] = 1
        self.obss[e, 0] = torch.as_tensor(obs, dtype=torch.float32, device=self.device)

    def flush(self) -> None:                                                                      # :97-98
        e = self._host_pos[0] % self.max_size
        self.slot_open[e] = 0
        self._host_pos = [self._host_pos[0] + 1, 0]
        n = self._host_pos[0]
        self.slot_open[n % self.max_size] = 1      # the slot the next episode will use is excluded (:141-145)
        self.counters[0] = n + 1
        self.counters[1] = n + 1          # slots [0, n] are in use: n completed + the one about to be opened
        self.counters[2] = n

    # ---- shared -----------------------------------------------------------------------------------------------
    def num_completed(self) -> int:
        """Completed episodes (host sync).  Monotone, so once ``can_sample`` holds it is cached."""
        self._completed_seen = max(self._completed_seen, int(self.counters[2].item()))
        return self._completed_seen

    def can_sample(self, batch_size: int) -> bool:                                                # :94-95
        if batch_size < self._completed_seen:
            return True
        return batch_size < self.num_completed()

    def draw_indices(self, batch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        eps = torch.empty((batch_size,), dtype=torch.int32, device=self.device)
        starts = torch.empty((batch_size,), dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib.dtqn_replay_sample_indices(C.byref(self.struct), batch_size, self.sample_seed, 0,
                                                       _lib.ptr(self.draw_counter), _lib.ptr(eps), _lib.ptr(starts),
                                                       _lib.stream_ptr()), "dtqn_replay_sample_indices")
        return eps, starts

    def gather_windows(self, episodes: torch.Tensor, starts: torch.Tensor, out=None):
        """(L+1)-row windows for given indices: obs_win [B,L+1,O] f32, act_win [B,L+1] u8, rew [B,L] f32,
        done [B,L] u8, eplen [B] i32."""
        B, L, O = int(episodes.shape[0]), self.context_len, self.env_obs_length
        episodes = episodes.to(device=self.device, dtype=torch.int32).contiguous()
        starts = starts.to(device=self.device, dtype=torch.int32).contiguous()
        if out is None:
            out = (torch.empty((B, L + 1, O), dtype=torch.float32, device=self.device),
                   torch.empty((B, L + 1), dtype=torch.uint8, device=self.device),
                   torch.empty((B, L), dtype=torch.float32, device=self.device),
                   torch.empty((B, L), dtype=torch.uint8, device=self.device),
                   torch.empty((B,), dtype=torch.int32, device=self.device))
        _lib.check(_lib.lib.dtqn_replay_gather(C.byref(self.struct), B, _lib.ptr(episodes), _lib.ptr(starts),
                                               *[_lib.ptr(t) for t in out], _lib.stream_ptr()), "dtqn_replay_gather")
        return out

    def sample(self, batch_size: int, indices=None):
        """ReplayBuffer.sample (:137-168): (obss, actions, rewards, next_obss, next_actions, dones, episode_lengths)
        with the reference's shapes -- views of the gathered (L+1)-row windows, on the device."""
        eps, starts = indices if indices is not None else self.draw_indices(batch_size)
        obs_win, act_win, rew, done, eplen = self.gather_windows(eps, starts)
        return (obs_win[:, :-1], act_win[:, :-1, None], rew[..., None], obs_win[:, 1:], act_win[:, 1:, None],
                done[..., None].bool(), eplen[:, None])
