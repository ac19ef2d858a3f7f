"""The UNMODIFIED reference's own training loop behind the interface of ``oracle.loop.ReferenceLoop`` (test
infrastructure / reported CPU baseline, kind "reference").

Only usable where /root/reference exists (the build container): everything goes through ``oracle.ref_harness`` (stub
gym / pyglet + the two compatibility shims), and the functions timed are the reference's own ``run.prepopulate`` /
``run.step`` (run.py:356-405), ``DtqnAgent.train`` (dtqn/agents/dtqn.py:162-269) and the loop body of ``run.train``
(run.py:290-298).  On the GPU box the reference tree is absent and bench.py falls back to the oracle port (kind "port").
"""
import sys

import torch

from oracle import ref_harness as rh


def available() -> bool:
    return rh.available()


class RealReferenceLoop:
    def __init__(self, env_id="DiscreteCarFlag-v0", seed=1, inner_embed=64, heads=8, layers=2, context=50, batch=32,
                 buf_size=500_000, lr=3e-4, tuf=10_000, gamma=0.99, num_steps=2_000_000, embed_per_obs=8):
        rh.activate()
        import gym  # stub
        argv, sys.argv = sys.argv, ["run.py"]
        try:
            import run as ref_run
        finally:
            sys.argv = argv
        from utils import agent_utils, epsilon_anneal
        self._run = ref_run
        self.env = gym.make(env_id)
        rh.set_global_seed(seed, self.env)
        self.agent = agent_utils.get_agent("DTQN", [self.env], embed_per_obs, 0, inner_embed, buf_size, torch.device("cpu"),
                                           lr, batch, context, -1, context, tuf, gamma, num_heads=heads, num_layers=layers,
                                           dropout=0.0, identity=False, gate="res", pos="learned", bag_size=0)
        rh.widen_episode_lengths(self.agent)
        self.eps = epsilon_anneal.LinearAnneal(1.0, 0.1, max(1, num_steps // 10))      # run.py:420
        self.batch = batch
        self.env_steps = 0
        self.grad_steps = 0

    def prepopulate(self, steps):
        self._run.prepopulate(self.agent, steps, [self.env])                           # run.py:495
        self.agent.eval_off()
        self.agent.context_reset(self.env.reset())                                     # run.py:287-288

    def set_epsilon(self, v):
        self.eps.val = float(v)

    def step_only(self):
        """run.step + the episode roll of the loop body (run.py:291-296), no training."""
        done = self._run.step(self.agent, self.env, self.eps)
        self.env_steps += 1
        if done:
            self.agent.replay_buffer.flush()
            self.agent.context_reset(self.env.reset())

    def train_only(self):
        self.agent.train()                                                             # run.py:297
        self.grad_steps += 1

    def iteration(self):
        self.step_only()
        self.train_only()
        self.eps.anneal()                                                              # run.py:298
