"""DtqnAgent -- host-side mirror of ``dtqn.agents.dtqn.DtqnAgent`` / ``dtqn.agents.dqn.DqnAgent``
(dtqn/agents/dtqn.py:15-269, dtqn/agents/dqn.py:24-140,208-210).

Same constructor and the surface ``run.py`` uses (``policy_network``, ``context_reset``, ``get_action``, ``observe``,
``replay_buffer.flush``, ``train``, ``eval_on/off``, ``context.timestep``, ``num_train_steps``, the eight
``RunningAverage`` attributes, ``target_update``), plus the batched B200 path: ``n_envs`` lockstep device environments
whose acting forward, replay append, window gather, 3 forwards + TD loss + backward, gradient allreduce, clip and Adam
all run as sm_100a kernels with no per-step host synchronisation.
"""
import ctypes as C
from enum import Enum
from typing import Callable, Optional, Union

import numpy as np
import torch
import torch.distributed as dist

from dtqn_b200 import _lib, checkpoint
from dtqn_b200.buffers import ReplayBuffer
from dtqn_b200.envs import ContextWindow
from dtqn_b200.networks import DTQN, NetCfg, ObsSrc, forward_groups
from dtqn_b200.parallel import allreduce_gradients

_l = _lib.lib
_l.dtqn_td_scratch_floats.argtypes = [C.POINTER(NetCfg), C.c_int32, C.c_int32]
_l.dtqn_td_scratch_floats.restype = C.c_int64
_l.dtqn_td_backward.argtypes = [C.POINTER(NetCfg), C.c_void_p, C.POINTER(ObsSrc), C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
_l.dtqn_td_backward.restype = C.c_int
_l.dtqn_clip_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                              C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int32, C.c_void_p]
_l.dtqn_clip_adam.restype = C.c_int
_l.dtqn_allreduce_clip_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
_l.dtqn_allreduce_clip_adam.restype = C.c_int

STAT_NAMES = ("td_errors", "qvalue_max", "qvalue_mean", "qvalue_min", "target_max", "target_mean", "target_min", "grad_norms")
RING = 100   # RunningAverage(100), dtqn/agents/dqn.py:82-89


class TrainMode(Enum):
    TRAIN = 1
    EVAL = 2


class RNG:
    """The reference's global generator (utils/random.py:9-10) for the single-env host-call API."""
    rng: np.random.Generator = None


class DeviceRunningAverage:
    """RunningAverage(100) (utils/logging_utils.py:10-24) whose samples live in a device ring written by the optimiser
    kernel; ``mean()`` is the only place that synchronises."""

    def __init__(self, agent, column: int):
        self._agent, self._col = agent, column

    def mean(self) -> float:
        n = min(self._agent.num_train_steps, RING)
        if n == 0:
            return 0.0
        self._agent.check_finite()
        return float(self._agent.stats_ring[:n, self._col].double().mean().item())


class DtqnAgent:
    def __init__(self, network_factory: Callable[[], torch.nn.Module], buffer_size: int, device, env_obs_length: int,
                 max_env_steps: int, obs_mask: Union[int, float], num_actions: int, is_discrete_env: bool,
                 learning_rate: float = 0.0003, batch_size: int = 32, context_len: int = 50, gamma: float = 0.99,
                 grad_norm_clip: float = 1.0, target_update_frequency: int = 10_000, history: int = 50,
                 bag_size: int = 0, n_envs: int = 1, trunc_context_obs: bool = True, sample_seed: int = 0,
                 record_every: int = 1, **kwargs):
        if bag_size:
            raise NotImplementedError("DTQN-bag is outside the hot path (SURVEY.md section 2 #23)")
        self.device = _lib.require_cuda(device)
        self.context_len, self.env_obs_length, self.n_envs = int(context_len), int(env_obs_length), int(n_envs)
        self.policy_network: DTQN = network_factory()
        self.target_network: DTQN = network_factory()
        self.target_update()                                                     # dqn.py:46-50
        self.target_network.eval()
        self.obs_tensor_type = torch.long if is_discrete_env else torch.float32  # dqn.py:54-59
        n = self.policy_network.n_flat
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.opt_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.learning_rate, self.betas, self.adam_eps = float(learning_rate), (0.9, 0.999), 1e-8   # dqn.py:64
        self.replay_buffer = ReplayBuffer(buffer_size, env_obs_length=env_obs_length, obs_mask=obs_mask,
                                          max_episode_steps=max_env_steps, context_len=context_len, n_envs=n_envs,
                                          device=self.device, sample_seed=sample_seed, record_every=record_every)
        self.batch_size, self.gamma, self.grad_norm_clip = int(batch_size), float(gamma), float(grad_norm_clip)
        self.target_update_frequency = int(target_update_frequency)
        self.history = int(history)
        self.num_train_steps = 0
        self.stats = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.stats_ring = torch.zeros((RING, 8), dtype=torch.float32, device=self.device)
        self.flags = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.opt_scratch = torch.zeros(1024, dtype=torch.float32, device=self.device)
        for col, name in enumerate(STAT_NAMES):
            setattr(self, name, DeviceRunningAverage(self, col))
        self.num_actions, self.obs_mask = int(num_actions), obs_mask
        self.train_mode = TrainMode.TRAIN
        self.train_context = ContextWindow(context_len, obs_mask, num_actions, env_obs_length, n_envs, self.device, trunc_context_obs)
        self.eval_context = ContextWindow(context_len, obs_mask, num_actions, env_obs_length, n_envs, self.device, trunc_context_obs)
        B, L, O = self.batch_size, self.context_len, self.env_obs_length
        self._win = (torch.empty((B, L + 1, O), dtype=torch.float32, device=self.device),
                     torch.empty((B, L + 1), dtype=torch.uint8, device=self.device),
                     torch.empty((B, L), dtype=torch.float32, device=self.device),
                     torch.empty((B, L), dtype=torch.uint8, device=self.device),
                     torch.empty((B,), dtype=torch.int32, device=self.device))
        self._q_all = torch.empty((3, B, L, self.num_actions), dtype=torch.float32, device=self.device)
        self._td_scratch = torch.zeros(int(_l.dtqn_td_scratch_floats(C.byref(self.policy_network.cfg), B, L)),
                                       dtype=torch.float32, device=self.device)
        self._q_last = torch.zeros((self.n_envs, self.num_actions), dtype=torch.float32, device=self.device)
        self._host_ctx = None          # single-env host-call API state
        self.strict_finite = False
        self.exchange = None           # parallel.PeerExchange when the gradient exchange runs over NVLink peer memory

    def use_peer_exchange(self, exchange) -> None:
        """Route the gradient collective through ``exchange`` (parallel.PeerExchange): the backward writes the local
        gradient straight into the IPC-exported buffer and the update becomes dtqn_allreduce_clip_adam (2 kernels,
        no NCCL call, graph-capturable).  Must be called before a CUDA graph is captured."""
        assert exchange.n == self.policy_network.n_flat
        self.exchange = exchange
        self.grads = exchange.grads

    # ---- mode / context (dqn.py:102-115) -------------------------------------------------------------------------------
    @property
    def context(self) -> ContextWindow:
        return self.train_context if self.train_mode == TrainMode.TRAIN else self.eval_context

    def eval_on(self) -> None:
        self.train_mode = TrainMode.EVAL
        self.policy_network.eval()

    def eval_off(self) -> None:
        self.train_mode = TrainMode.TRAIN
        self.policy_network.train()

    def target_update(self) -> None:
        """Hard update (dqn.py:208-210): one flat device-to-device copy."""
        self.target_network.flat.copy_(self.policy_network.flat)
        # the weight images (tcgen05 operands + the k-major copies the training forward streams) must follow even when the
        # loop runs as graph replays, where forward_groups' host-side staleness check never executes
        if self.policy_network.packed_stale:
            self.policy_network.repack()
        self.target_network.packed.copy_(self.policy_network.packed)
        self.target_network.packed_stale = False

    # ---- checkpoint / resume (dqn.py:212-327; file format in dtqn_b200/checkpoint.py) ----------------------------------------
    def save_mini_checkpoint(self, checkpoint_dir: str, wandb_id: Optional[str]) -> None:
        checkpoint.save_mini_checkpoint(self, checkpoint_dir, wandb_id)

    @staticmethod
    def load_mini_checkpoint(checkpoint_dir: str) -> dict:
        return checkpoint.load_mini_checkpoint(checkpoint_dir)

    def save_checkpoint(self, checkpoint_dir: str, wandb_id: Optional[str], episode_successes, episode_rewards,
                        episode_lengths, eps) -> None:
        checkpoint.save_agent(self, checkpoint_dir, wandb_id, episode_successes, episode_rewards, episode_lengths, eps)

    def load_checkpoint(self, checkpoint_dir: str):
        """-> (wandb_id, episode_successes, episode_rewards, episode_lengths, epsilon) like dqn.py:281-327."""
        return checkpoint.load_agent(self, checkpoint_dir)[:5]

    def check_finite(self) -> None:
        if self.exchange is not None and self.exchange.error():
            raise RuntimeError("gradient exchange timed out waiting for a peer rank (dtqn_p2p_error)")
        if int(self.flags.item()) != 0:
            raise RuntimeError("The total norm for gradients is non-finite, so it cannot be clipped.")  # agents/dtqn.py:257-261

    # ---- batched acting: get_action for every lockstep env (agents/dtqn.py:76-107) ------------------------------------------
    def q_last_batched(self) -> torch.Tensor:
        """Q of the last context position of every env ([n_envs, A]) from the device context ring."""
        cx, net = self.context, self.policy_network
        src = ObsSrc(obs=cx.obs.data_ptr(), seq_stride=cx.max_length * cx.env_obs_length,
                     timestep=cx.timestep_t.data_ptr(), ring_len=cx.max_length, obs_mask=cx.obs_mask,
                     actions=cx.action.data_ptr(), act_stride=cx.max_length, train_mode=int(net.training))
        forward_groups(net, [net], [src], self.n_envs, self.context_len, q_mode=1, save=0, q_out=self._q_last)
        return self._q_last

    def act_and_step(self, env, epsilon: float, record: Optional[bool] = None, epsilon_dev=None) -> None:
        """run.step (run.py:356-377) for all envs: acting forward -> eps-greedy -> env.step -> observe."""
        q = self.q_last_batched()
        rec = (self.train_mode == TrainMode.TRAIN) if record is None else record
        env.step(mode=_lib.ACT_EPS_GREEDY, epsilon=epsilon, q_last=q, record=rec, epsilon_dev=epsilon_dev)

    # ---- training step (agents/dtqn.py:162-269) ----------------------------------------------------------------------------
    def train_on_windows(self, obs_win, act_win, rew, done) -> None:
        """3 forwards + TD loss + backward + (allreduce) + clip + Adam on gathered (L+1)-row windows."""
        self.forward_backward(obs_win, act_win, rew, done)
        self.reduce_and_step()
        self.finish_step()

    def forward_backward(self, obs_win, act_win, rew, done) -> None:
        net, tgt = self.policy_network, self.target_network
        B, L, O = obs_win.shape[0], self.context_len, self.env_obs_length
        stride = (L + 1) * O
        # policy(obss, actions) and policy(next_obss, next_actions) in train mode, target(next_obss, next_actions) in eval
        # mode (agents/dtqn.py:215-233); actions / next_actions are columns [0, L) / [1, L+1) of the (L+1)-column window
        mk = lambda shift, train: ObsSrc(obs=obs_win.data_ptr() + 4 * O * shift, seq_stride=stride, timestep=None, ring_len=0,
                                         obs_mask=float(self.obs_mask), actions=act_win.data_ptr() + shift, act_stride=L + 1,
                                         train_mode=train)
        s_obs, s_next, s_tgt = mk(0, 1), mk(1, 1), mk(1, 0)
        ws = forward_groups(net, [net, net, tgt], [s_obs, s_next, s_tgt], B, L, q_mode=0, save=1,
                            q_out=self._q_all)
        st = _lib.stream_ptr()
        _lib.check(_l.dtqn_td_backward(C.byref(net.cfg), net.flat.data_ptr(), C.byref(s_obs), self._q_all.data_ptr(),
                                       act_win.data_ptr(), rew.data_ptr(), done.data_ptr(), B, L, self.history,
                                       self.gamma, ws.data_ptr(), ws.numel(), self._td_scratch.data_ptr(),
                                       self.grads.data_ptr(), self.stats.data_ptr(), st), "dtqn_td_backward")

    def reduce_and_step(self) -> None:
        """Gradient allreduce (the one collective, only when world > 1) -> /world -> global-norm clip -> Adam, identical
        on every rank."""
        net, st = self.policy_network, _lib.stream_ptr()
        ex = self.exchange
        if ex is not None:                                   # fused: barrier + NVLink reads + norm, then clip + Adam
            _lib.check(_l.dtqn_allreduce_clip_adam(net.flat.data_ptr(), ex.bases, ex.rank, ex.world, net.n_flat,
                                                   ex.reduced.data_ptr(), self.exp_avg.data_ptr(),
                                                   self.exp_avg_sq.data_ptr(), self.grad_norm_clip, self.learning_rate,
                                                   self.betas[0], self.betas[1], self.adam_eps, self.opt_step.data_ptr(),
                                                   self.opt_scratch.data_ptr(), self.stats.data_ptr(),
                                                   self.flags.data_ptr(), self.stats_ring.data_ptr(), RING, st),
                       "dtqn_allreduce_clip_adam")
            net.repack()
            return
        scale = allreduce_gradients(self.grads)              # sum of per-rank mean-MSE gradients; scale = 1/world
        _lib.check(_l.dtqn_clip_adam(net.flat.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), net.n_flat, scale, self.grad_norm_clip,
                                     self.learning_rate, self.betas[0], self.betas[1], self.adam_eps,
                                     self.opt_step.data_ptr(), self.opt_scratch.data_ptr(), self.stats.data_ptr(),
                                     self.flags.data_ptr(), self.stats_ring.data_ptr(), RING, st), "dtqn_clip_adam")
        net.repack()                                         # refresh the tensor-core weight image (1 launch)

    def finish_step(self) -> None:
        """Host-side bookkeeping of one update (agents/dtqn.py:266-269)."""
        self.num_train_steps += 1
        if self.strict_finite:
            self.check_finite()
        if self.num_train_steps % self.target_update_frequency == 0:             # agents/dtqn.py:268-269
            self.target_update()

    def train(self, indices=None) -> None:
        if not self.replay_buffer.can_sample(self.batch_size):                   # agents/dtqn.py:163-164
            return
        self.eval_off()
        rb = self.replay_buffer
        eps, starts = indices if indices is not None else rb.draw_indices(self.batch_size)
        rb.gather_windows(eps, starts, out=self._win)
        self.train_on_windows(*self._win[:4])

    # ---- reference-compatible single-env host-call API (plumbing for run.py; the hot path is the batched one) ---------------
    def context_reset(self, obs: np.ndarray) -> None:                           # agents/dtqn.py:109-114
        cx = self.context
        assert cx.n_envs == 1, "the host-call API drives one env; use BatchedEnv for n_envs > 1"
        o = torch.as_tensor(np.asarray(obs, dtype=np.float64), device=self.device)
        o = torch.trunc(o) if cx.trunc_obs else o
        cx.obs[0, 0] = o.float()
        cx.timestep_t.zero_()
        self._host_t = 0
        if RNG.rng is not None:                                                  # Context.reset random action padding
            cx.action[0, 0] = int(RNG.rng.integers(self.num_actions, size=(cx.max_length, 1))[0, 0])
        if self.train_mode == TrainMode.TRAIN:
            self.replay_buffer.store_obs(obs)

    def observe(self, obs: np.ndarray, action: int, reward: float, done: bool) -> None:   # agents/dtqn.py:116-160
        cx = self.context
        self._host_t += 1
        o = torch.as_tensor(np.asarray(obs, dtype=np.float64), device=self.device)
        o = torch.trunc(o) if cx.trunc_obs else o
        cx.obs[0, self._host_t % cx.max_length] = o.float()
        cx.action[0, self._host_t % cx.max_length] = int(action)                 # utils/context.py:77
        cx.timestep_t.fill_(self._host_t)
        if self.train_mode == TrainMode.TRAIN:
            self.replay_buffer.store(obs, action, reward, done, self._host_t)

    @torch.no_grad()
    def get_action(self, epsilon: float = 0.0) -> int:                          # agents/dtqn.py:76-107
        if RNG.rng is None:
            RNG.rng = np.random.Generator(np.random.PCG64(seed=0))
        if RNG.rng.random() < epsilon:
            return int(RNG.rng.integers(self.num_actions))
        return int(torch.argmax(self.q_last_batched()[0]).item())
