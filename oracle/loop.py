"""CPU oracle restatement of the reference's single-env training loop (test infrastructure / reported CPU baseline).

Follows run.py:246-405 (train / step / prepopulate) with the oracle env, context, replay buffer and trainer:
one epsilon-greedy action (full-context forward, batch 1) -> env.step -> buffer store -> one DtqnAgent.train()
(batch 32 x ctx windows: 3 forwards + backward + clip + Adam) per iteration.  bench.py times this on the GPU box's
host cores as the ``--impl reference`` arm and as ``cpu_baseline`` (kind "port": /root/reference cannot travel).
"""
import random

import numpy as np
import torch

from oracle import envs as oenvs
from oracle import network as onet
from oracle.agent import TrainerOracle
from oracle.pcg64 import PCG64
from oracle.replay import ContextOracle, ReplayOracle


class ReferenceLoop:
    def __init__(self, env_id="DiscreteCarFlag-v0", seed=1, inner_embed=64, heads=8, layers=2, context=50, batch=32,
                 buf_size=500_000, lr=3e-4, tuf=10_000, gamma=0.99, num_steps=2_000_000, embed_per_obs=8,
                 engine="module"):
        """engine "module": the network is ``oracle.network.ModuleNet`` (the torch.nn blocks the reference instantiates) with
        torch.optim.Adam / clip_grad_norm_ / F.mse_loss, i.e. the reference's own library calls op for op -- the timed CPU
        baseline.  engine "closed": the closed-form oracle + explicit Adam of ``oracle.agent`` (the numerical checker)."""
        random.seed(seed)
        torch.manual_seed(seed)
        self.env = oenvs.make(env_id, seed)
        self.rng = PCG64.from_seed(seed)                        # RNG.rng (utils/random.py:31)
        e = self.env
        self.discrete = env_id.startswith("Memory")
        self.ctx_len, self.batch, self.heads = context, batch, heads
        self.env.reset(); self.env.reset()                      # get_agent's hidden resets (env_processing.py:67)
        sd = onet.init_state_dict(e.obs_dim, e.num_actions, embed_per_obs, inner_embed, heads, layers, context,
                                  discrete=self.discrete, vocab_size=e.obs_mask + 1 if self.discrete else None)
        self.engine, self.gamma, self.tuf = engine, gamma, tuf
        if engine == "module":
            self.policy_net, self.target_net = onet.ModuleNet(sd, heads), onet.ModuleNet(sd, heads)
            self.target_net.eval()                                                   # dqn.py:46-50
            self.optimizer = torch.optim.Adam(self.policy_net.parameters(), lr=lr)   # dqn.py:64
            self.stats = []
        else:
            self.trainer = TrainerOracle(sd, heads, lr=lr, gamma=gamma, target_update_frequency=tuf)
        self.buffer = ReplayOracle(buf_size, e.obs_dim, e.obs_mask, e.max_episode_steps, context)
        self.context = ContextOracle(context, e.obs_mask, e.num_actions, e.obs_dim, self.rng)
        self.eps, self.eps_min, self.eps_dur = 1.0, 0.1, max(1, num_steps // 10)     # run.py:420
        self.env_steps = 0
        self.grad_steps = 0
        self._begin_episode()

    def _begin_episode(self):
        o = self.env.reset()
        self.context.reset(o)
        self.buffer.store_obs(o)

    def _observe(self, o, a, r, done, info):
        bd = False if info.get("TimeLimit.truncated", False) else done          # run.py:370-374
        self.context.add_transition(o, a)
        self.buffer.store(o, a, r, bd, self.context.timestep)

    def prepopulate(self, steps):                                                # run.py:380-405
        t = 0
        first = True
        while t < steps:
            if not first:
                self._begin_episode()
            first = False
            done = False
            while not done:
                a = self.rng.integers(self.env.num_actions)
                o, r, done, info = self.env.step(a)
                self._observe(o, a, r, done, info)
                t += 1
            self.buffer.flush()
        self._begin_episode()

    def get_action(self, epsilon):                                               # agents/dtqn.py:76-107
        if self.rng.random() < epsilon:
            return self.rng.integers(self.env.num_actions)
        obs, _ = self.context.window()
        x = torch.as_tensor(obs, dtype=torch.long if self.discrete else torch.float32).unsqueeze(0)
        with torch.no_grad():
            q = self.policy_net(x) if self.engine == "module" else onet.forward(self.trainer.policy, x, self.heads)
        return int(torch.argmax(q[:, -1, :]).item())

    def train(self):                                                             # agents/dtqn.py:162-269
        if not self.buffer.can_sample(self.batch):
            return None
        o, a, r, no, na, d, _ = self.buffer.sample(self.batch)
        conv = (lambda x: torch.as_tensor(x).long()) if self.discrete else (lambda x: torch.as_tensor(x, dtype=torch.float32))
        batch = (conv(o), torch.as_tensor(a.astype(np.int64)), torch.as_tensor(r), conv(no),
                 torch.as_tensor(na.astype(np.int64)), torch.as_tensor(d))
        if self.engine == "module":
            stats = self._train_module(batch)
        else:
            stats, _ = self.trainer.train_on_batch(batch)
        self.grad_steps += 1
        return stats

    def _train_module(self, batch):
        """agents/dtqn.py:215-269 with the reference's own torch calls, including its eight per-step ``.item()`` reads."""
        obss, actions, rewards, next_obss, _, dones = batch
        net, tgt = self.policy_net, self.target_net
        q = net(obss).gather(2, actions).squeeze()                                   # :215-219
        with torch.no_grad():
            a_star = torch.argmax(net(next_obss), dim=2).unsqueeze(-1)               # :226-229
            next_q = tgt(next_obss).gather(2, a_star).squeeze()                      # :230-233
            targets = rewards.squeeze() + (1 - dones.long().squeeze()) * (next_q * self.gamma)   # :236-238
        loss = torch.nn.functional.mse_loss(q, targets)                              # :243
        st = [loss.item(), q.max().item(), q.mean().item(), q.min().item(), targets.max().item(), targets.mean().item(),
              targets.min().item()]                                                  # :245-253
        self.optimizer.zero_grad(set_to_none=True)                                   # :255
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0, error_if_nonfinite=True)   # :257-261
        st.append(norm.item())
        self.optimizer.step()                                                        # :265
        if (self.grad_steps + 1) % self.tuf == 0:                                    # :268-269
            self.target_net.load_state_dict(net.state_dict())
        return dict(zip(("loss", "q_max", "q_mean", "q_min", "t_max", "t_mean", "t_min", "grad_norm"), st))

    def set_epsilon(self, v):
        self.eps = float(v)

    def step_only(self):
        """run.step + the episode roll of the loop body (run.py:291-296), no training."""
        a = self.get_action(self.eps)
        o, r, done, info = self.env.step(a)
        self._observe(o, a, r, done, info)
        self.env_steps += 1
        if done:
            self.buffer.flush()
            self._begin_episode()

    def train_only(self):
        return self.train()

    def iteration(self):
        """Loop body of run.train (run.py:290-298)."""
        self.step_only()
        stats = self.train()
        self.eps = max(self.eps_min, self.eps - (self.eps - self.eps_min) / self.eps_dur)   # epsilon_anneal.py:33-34
        return stats
