"""Batched training loop -- the shape of run.train / run.step / run.prepopulate / run.evaluate (run.py:187-405) for
n_envs lockstep device environments per GPU, data-parallel over ranks (one process per GPU): envs and replay are
sharded per rank, parameters replicated, one gradient allreduce per update (agents.DtqnAgent.train_on_windows)."""
import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from dtqn_b200 import _lib, checkpoint
from dtqn_b200.envs import BatchedEnv
from dtqn_b200.parallel import broadcast_parameters, peer_exchange_wanted, rank_world, shard_seed, try_peer_exchange
from dtqn_b200.utils import LinearAnneal, get_agent



class BatchedTrainer:
    def __init__(self, env_id: str = "DiscreteCarFlag-v0", n_envs: int = 4096, seed: int = 1, device=None,
                 inner_embed: int = 64, heads: int = 8, layers: int = 2, context: int = 50, batch: int = 32,
                 buf_size: Optional[int] = None, lr: float = 3e-4, tuf: int = 10_000, gamma: float = 0.99,
                 history: Optional[int] = None, num_steps: int = 2_000_000, obs_embed: int = 8,
                 trunc_context_obs: bool = True, pos: str = "learned", max_episode_steps: Optional[int] = None,
                 a_embed: int = 0, dropout: float = 0.0, identity: bool = False, gate: str = "res", record_every: int = 1,
                 updates_per_step: int = 1):
        rank, world = rank_world()
        self.rank, self.world = rank, world
        self.device = _lib.require_cuda(device)
        self.env = BatchedEnv(env_id, n_envs, seed=shard_seed(seed, rank, n_envs), device=self.device,
                              max_episode_steps=max_episode_steps if max_episode_steps and max_episode_steps > 0 else None)
        self.eval_env = None
        E = self.env.max_episode_steps
        if buf_size is None:
            buf_size = max(500_000, 8 * n_envs * E)           # ring >= 8 slots per env so it never wraps onto open episodes
        torch.manual_seed(seed)                              # identical initial parameters on every rank
        self.agent = get_agent("DTQN", [self.env], obs_embed, a_embed, inner_embed, buf_size, self.device, lr, batch, context,
                               E, history or context, tuf, gamma, num_heads=heads, num_layers=layers, dropout=dropout,
                               identity=identity, gate=gate, pos=pos, n_envs=n_envs,
                               trunc_context_obs=trunc_context_obs, sample_seed=seed * 7919 + rank, record_every=record_every)
        if world > 1:
            broadcast_parameters(self.agent.policy_network.flat, src=0)
            self.agent.policy_network.packed_stale = True
            self.agent.target_update()
        # the one collective of the update: fused NVLink peer-memory exchange when the node allows it, else NCCL
        self.allreduce, self.allreduce_note = ("none" if world == 1 else "nccl"), ""
        if world > 1 and peer_exchange_wanted(world):
            ex, why = try_peer_exchange(self.agent.policy_network.n_flat, self.device)
            if ex is not None:
                self.agent.use_peer_exchange(ex)
                self.allreduce = "p2p-fused"
            else:
                self.allreduce_note = why
                if rank == 0:
                    print(f"[dtqn_b200] peer-memory gradient exchange unavailable ({why}); using the NCCL allreduce", flush=True)
        self._update_in_graph = world == 1 or self.allreduce == "p2p-fused"
        self.env.attach(self.agent.replay_buffer, self.agent.train_context)
        self.eps = LinearAnneal(1.0, 0.1, max(1, num_steps // 10))     # run.py:420
        self.env.reset_all()                                  # run.py:287-288
        self.n_envs = n_envs
        # the reference does 1 update per env step of ONE env (run.py:297); with thousands of lockstep envs per update the
        # data : update ratio is n_envs x larger -- K > 1 performs K sample / train rounds per lockstep step
        self.updates_per_step = int(updates_per_step)
        assert self.updates_per_step >= 1
        self.iterations = 0
        self._graph = None
        # exploration schedule mirrored on the device: a replayed graph reads / anneals it without any host write
        self._eps_state = torch.zeros(3, dtype=torch.float64, device=self.device)
        self._eps_dev = torch.zeros(1, dtype=torch.float64, device=self.device)

    def prepopulate(self, lockstep_steps: int) -> None:
        """run.prepopulate (run.py:380-405): uniform-random actions from the agent-side stream."""
        for _ in range(lockstep_steps):
            self.env.step(mode=_lib.ACT_RANDOM)

    def _device_iteration(self) -> None:
        """Everything of one loop iteration that runs on the device before the gradient collective."""
        agent, rb = self.agent, self.agent.replay_buffer
        self._device_epsilon()
        agent.act_and_step(self.env, 0.0, epsilon_dev=self._eps_dev)
        for k in range(self.updates_per_step):
            eps, starts = rb.draw_indices(agent.batch_size)
            rb.gather_windows(eps, starts, out=agent._win)
            agent.forward_backward(*agent._win[:4])
            if self._update_in_graph:
                agent.reduce_and_step()
            else:
                assert self.updates_per_step == 1, "updates_per_step > 1 needs the update inside the graph (1 GPU or the fused NVLink exchange)"

    def _sync_epsilon_to_device(self) -> None:
        self._eps_state.copy_(torch.tensor([self.eps.val, getattr(self.eps, "min", self.eps.val),
                                            getattr(self.eps, "duration", 1)], dtype=torch.float64))

    def _device_epsilon(self) -> None:
        _lib.check(_lib.lib.dtqn_eps_anneal(self._eps_state.data_ptr(), self._eps_dev.data_ptr(), _lib.stream_ptr()),
                   "dtqn_eps_anneal")

    def enable_graphs(self) -> None:
        """Capture one loop iteration (acting forward, env step + roll, sample, gather, 3 forwards, TD, backward and --
        single GPU -- clip + Adam) into a CUDA graph: ~60 launches replayed with one host call.  The per-step scalars
        (epsilon, Adam step, sampler draw counter) live in device memory so the replay needs no re-capture."""
        assert self.agent.replay_buffer.can_sample(self.agent.batch_size), "prepopulate before capturing"
        self.agent.eval_off()
        self._sync_epsilon_to_device()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up on a side stream (allocations, lazy module loads)
            self._device_iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._device_iteration()
        self._graph = g
        if not self._update_in_graph:
            self.agent.reduce_and_step()
        for _ in range(self.updates_per_step):
            self.agent.finish_step()
        self.eps.anneal()
        self.iterations += 1

    def disable_graphs(self) -> None:
        self._graph = None

    def capture(self, fn) -> torch.cuda.CUDAGraph:
        """Capture ``fn()`` (kernels only, no host sync) into a CUDA graph after one warm-up call on a side stream."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def act_only(self) -> None:
        """Acting half of an iteration: acting forward + eps-greedy + env step / append / roll."""
        self._device_epsilon()
        self.agent.act_and_step(self.env, 0.0, epsilon_dev=self._eps_dev)

    def train_only(self) -> None:
        """Training half of an iteration (single GPU): sample + gather + 3 forwards + TD + backward + clip + Adam."""
        agent, rb = self.agent, self.agent.replay_buffer
        eps, starts = rb.draw_indices(agent.batch_size)
        rb.gather_windows(eps, starts, out=agent._win)
        agent.forward_backward(*agent._win[:4])
        agent.reduce_and_step()

    def train_iteration(self) -> None:
        """One iteration of the run.train loop body (run.py:290-298) for all envs: step, train, anneal."""
        if self._graph is not None:
            self._graph.replay()
            if not self._update_in_graph:
                self.agent.reduce_and_step()
            for _ in range(self.updates_per_step):
                self.agent.finish_step()
        else:
            self.agent.act_and_step(self.env, self.eps.val)
            for _ in range(self.updates_per_step):
                self.agent.train()
        self.eps.anneal()
        self.iterations += 1

    # ---- host-policy loop: the user's policy code runs on the HOST between device calls ------------------------------------------
    def enable_host_loop(self) -> None:
        """Prepare ``host_q / host_step / host_train``: pinned host buffers plus three CUDA graphs (acting forward; env step
        + replay / context append + roll for host-chosen actions; sample + gather + 3 forwards + TD + backward + update),
        so each call is one graph replay and the copies it documents.  Needs a sampleable replay (prepopulate first).
        Capturing runs each piece once for real (one acting forward, one env step with the current ``env.actions``, one
        update)."""
        assert self.agent.replay_buffer.can_sample(self.agent.batch_size), "prepopulate before enable_host_loop()"
        env, agent, N = self.env, self.agent, self.n_envs
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        self.h_q, self.h_act = pin((N, env.num_actions), torch.float32), pin((N,), torch.int32)
        self.h_obs, self.h_rew = pin((N, env.obs_dim), torch.float32), pin((N,), torch.float32)
        self.h_done, self.h_stats = pin((N,), torch.uint8), pin((8,), torch.float32)
        agent.eval_off()
        self._hg_q = self.capture(agent.q_last_batched)
        self._hg_env = self.capture(lambda: env.step(mode=_lib.ACT_GIVEN))        # reads env.actions
        if self._update_in_graph:
            self._hg_train = self.capture(self.train_only)
        else:                                                                      # NCCL allreduce stays outside the graph
            def fb():
                rb = agent.replay_buffer
                eps, starts = rb.draw_indices(agent.batch_size)
                rb.gather_windows(eps, starts, out=agent._win)
                agent.forward_backward(*agent._win[:4])
            self._hg_train = self.capture(fb)
            agent.reduce_and_step()
        agent.finish_step()                                   # the capture warm-up performed one real update

    def host_q(self) -> torch.Tensor:
        """Q of every env's last context position as a pinned HOST tensor [n_envs, A] (agents/dtqn.py:81-107 batched)."""
        self._hg_q.replay()
        self.h_q.copy_(self.agent._q_last, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_q

    def host_step(self, actions: torch.Tensor):
        """env.step + agent.observe for host-chosen actions (int32 [n_envs], ideally pinned).  Returns pinned host
        (obs, reward, done); they are complete after the next synchronising call (``host_train`` / ``host_q``)."""
        self.env.actions.copy_(actions, non_blocking=True)
        self._hg_env.replay()
        self.h_obs.copy_(self.env.obs_out, non_blocking=True)
        self.h_rew.copy_(self.env.reward_out, non_blocking=True)
        self.h_done.copy_(self.env.done_out, non_blocking=True)
        return self.h_obs, self.h_rew, self.h_done

    def host_train(self) -> torch.Tensor:
        """agent.train() (agents/dtqn.py:162-269); returns the step's 8 statistics (loss first) as a pinned host tensor."""
        self._hg_train.replay()
        if not self._update_in_graph:
            self.agent.reduce_and_step()
        self.agent.finish_step()
        self.h_stats.copy_(self.agent.stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_stats

    def save_checkpoint(self, checkpoint_dir: str, wandb_id: Optional[str] = None, episode_successes=None,
                        episode_rewards=None, episode_lengths=None) -> None:
        """agent.save_checkpoint (dqn.py:222-279) + env streams, contexts and loop counters; one set of files per rank
        (pass a per-rank prefix under torchrun)."""
        checkpoint.save_trainer(self, checkpoint_dir, wandb_id, episode_successes, episode_rewards, episode_lengths)

    def load_checkpoint(self, checkpoint_dir: str):
        """-> (wandb_id, episode_successes, episode_rewards, episode_lengths); the loop continues bit-exactly."""
        return checkpoint.load_trainer(self, checkpoint_dir)

    def make_eval_env(self) -> BatchedEnv:
        if self.eval_env is None:
            # same seeds as the train envs (utils/random.py:26-29) and the SAME agent-side streams: the reference's
            # evaluation draws from the one global RNG.rng the training loop uses
            self.eval_env = BatchedEnv(self.env.env_id, self.n_envs, seeds=self.env.seeds, device=self.device,
                                       max_episode_steps=self.env.max_episode_steps, agent_stream_of=self.env)
            self.eval_env.attach(None, self.agent.eval_context)
        return self.eval_env

    @torch.no_grad()
    def evaluate(self, eval_episodes_per_env: int = 1, max_steps: Optional[int] = None, reduce_ranks: bool = True):
        """run.evaluate (run.py:187-243): greedy policy on a separate set of envs seeded like the train envs
        (utils/random.py:26-29), no replay writes, EXACTLY ``eval_episodes_per_env`` episodes per env: each env's first k
        finished episodes are counted (per-env counter in the step kernel), later ones are ignored, so short episodes are
        not over-represented.  Sums are reduced over ranks.  Returns (success_rate, mean_return, mean_episode_length);
        ``self.last_eval_per_env`` keeps the per-env [n, 4] sums (episodes, return, length, successes) of this rank."""
        agent = self.agent
        ev = self.make_eval_env()
        k = int(eval_episodes_per_env)
        agent.eval_on()
        ev.count_episodes_per_env(k)
        ev.reset_all()
        steps = max_steps or k * ev.max_episode_steps            # every episode ends within max_episode_steps (TimeLimit)
        for t in range(steps):
            agent.act_and_step(ev, 0.0, record=False)
            if t % 25 == 24 and bool((ev.env_acc[:, 0] >= k).all().item()):
                break
        agent.eval_off()
        self.last_eval_per_env = ev.env_acc.clone()
        tot = ev.env_acc.sum(dim=0, dtype=torch.int64)
        if reduce_ranks and self.world > 1:
            dist.all_reduce(tot)
        n, ret, length, succ = [int(v) for v in tot.tolist()]
        n = max(n, 1)
        return succ / n, ret / n, length / n
