"""Generate the golden fixtures by EXECUTING the unmodified reference (kevslinger/DTQN) in the build container.

    python tests/golden/gen_golden.py          # needs /root/reference; writes tests/golden/*.npz

The reference has no tests / golden vectors of its own, so these fixtures are what pins the oracle
port (``oracle/``) and, through it, the CUDA path.  They travel to the GPU box; /root/reference does not.
Everything here goes through ``oracle.ref_harness`` (stub gym/pyglet + two compatibility shims, no reference file
is modified).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh  # noqa: E402

rh.activate()
import gym  # noqa: E402  (stub)

ENV_SHAPES = {"DiscreteCarFlag-v0": (3, 3), "Memory-5-v0": (10, 10)}


def action_tapes(env_id, n_seeds, T, rng):
    A = ENV_SHAPES[env_id][1]
    tapes = rng.integers(0, A, size=(n_seeds, T)).astype(np.uint8)
    if env_id == "DiscreteCarFlag-v0":
        tapes[0, :] = 2          # drive right: terminates at +1 (heaven or hell)
        tapes[1, :] = 1          # coast: 200-step TimeLimit truncation
        tapes[2, :] = 0          # drive left
        tapes[3, : T // 2] = 2   # mixed
    else:
        tapes[0, :] = 0          # repeatedly pick card 0
        tapes[1, :] = np.arange(T) % A
    return tapes


def gen_env(env_id, n_seeds, T, first_seed=1):
    rng = np.random.default_rng(1234)
    O, _ = ENV_SHAPES[env_id]
    tapes = action_tapes(env_id, n_seeds, T, rng)
    seeds = np.arange(first_seed, first_seed + n_seeds, dtype=np.int64)
    out = dict(seeds=seeds, actions=tapes,
               initial_obs=np.zeros((n_seeds, O)), obs=np.zeros((n_seeds, T, O)), reward=np.zeros((n_seeds, T)),
               done=np.zeros((n_seeds, T), bool), truncated=np.zeros((n_seeds, T), bool),
               success=np.zeros((n_seeds, T), bool),
               reset_obs=np.full((n_seeds, T, O), np.nan), rng_state=np.zeros((n_seeds, 6), np.uint64))
    for i, seed in enumerate(seeds):
        env = gym.make(env_id)
        env.seed(int(seed))
        out["initial_obs"][i] = env.reset()
        for t in range(T):
            o, r, d, info = env.step(int(tapes[i, t]))
            out["obs"][i, t], out["reward"][i, t], out["done"][i, t] = o, r, d
            out["truncated"][i, t] = bool(info.get("TimeLimit.truncated", False))
            out["success"][i, t] = bool(info.get("is_success", False))
            if d:
                out["reset_obs"][i, t] = env.reset()
        st = env.unwrapped.np_random.bit_generator.state
        M = (1 << 64) - 1
        out["rng_state"][i] = [st["state"]["state"] >> 64, st["state"]["state"] & M, st["state"]["inc"] >> 64,
                               st["state"]["inc"] & M, st["has_uint32"], st["uinteger"]]
    return out


def perturb(net, gen):
    """Make every parameter non-trivial (the init has zero biases / zero position table / unit LayerNorm)."""
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("attn_mask"):
                continue
            if name.endswith("bias") or "position_encoding" in name:
                p.add_(torch.empty_like(p).normal_(0, 0.05, generator=gen))
            elif "layernorm" in name and name.endswith("weight"):
                p.add_(torch.empty_like(p).normal_(0, 0.1, generator=gen))
            else:
                p.mul_(3.0)   # larger weights -> O(1) Q values, harder numerics than the 0.02 init


def make_agent(env_id, inner_embed, layers, context, batch, seed=1, heads=8, history=None, buf=50_000):
    from utils import agent_utils
    env = gym.make(env_id)
    rh.set_global_seed(seed, env)
    agent = agent_utils.get_agent("DTQN", [env], 8, 0, inner_embed, buf, torch.device("cpu"), 3e-4, batch, context,
                                  -1, history or context, 10_000, 0.99, num_heads=heads, num_layers=layers,
                                  dropout=0.0, identity=False, gate="res", pos="learned", bag_size=0)
    rh.widen_episode_lengths(agent)
    return agent, env


def sd_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def gen_train(env_id, inner_embed, layers, context, batch, prepop, n_steps, tag):
    """Reference DtqnAgent.train() on injected batches: Q, loss, grads (step 1) and post-Adam parameters."""
    sys.argv = ["run.py"]
    import run as ref_run
    import random

    agent, env = make_agent(env_id, inner_embed, layers, context, batch)
    gen = torch.Generator().manual_seed(7)
    perturb(agent.policy_network, gen)
    agent.target_update()
    with torch.no_grad():   # make the target differ from the policy so a* and Q_tgt are distinguishable
        for name, p in agent.target_network.named_parameters():
            if not name.endswith("attn_mask"):
                p.add_(torch.empty_like(p).normal_(0, 0.01, generator=gen))
    ref_run.prepopulate(agent, prepop, [env])
    out = {}
    out.update(sd_np(agent.policy_network.state_dict(), "policy0/"))
    out.update(sd_np(agent.target_network.state_dict(), "target0/"))
    rb = agent.replay_buffer
    # the replay arrays after prepopulate (only the filled part) pin the store/flush layout
    n_ep = min(rb.pos[0], rb.max_size)
    out["replay/obss"], out["replay/actions"] = rb.obss[:n_ep].copy(), rb.actions[:n_ep].copy()
    out["replay/rewards"], out["replay/dones"] = rb.rewards[:n_ep].copy(), rb.dones[:n_ep].copy()
    out["replay/episode_lengths"] = rb.episode_lengths[:n_ep].copy()
    out["replay/pos"] = np.array(rb.pos)
    random.seed(99)
    real_sample = rb.sample
    for s in range(n_steps):
        state = random.getstate()
        batch_np = real_sample(batch)
        # recover the (episode, start) indices of this draw by replaying the same `random` stream
        random.setstate(state)
        valid = [i for i in range(min(rb.pos[0], rb.max_size)) if i != rb.pos[0] % rb.max_size]
        eps = np.array([random.choice(valid) for _ in range(batch)])
        starts = np.array([random.randint(0, max(0, int(rb.episode_lengths[e]) - rb.context_len)) for e in eps])
        out[f"step{s}/episodes"], out[f"step{s}/starts"] = eps, starts
        for name, arr in zip(("obss", "actions", "rewards", "next_obss", "next_actions", "dones", "eplens"), batch_np):
            out[f"step{s}/{name}"] = np.asarray(arr).copy()
        rb.sample = lambda bs, b=batch_np: b
        if s == 0:
            with torch.no_grad():
                ot = torch.as_tensor(batch_np[0], dtype=agent.obs_tensor_type)
                nt = torch.as_tensor(batch_np[3], dtype=agent.obs_tensor_type)
                out["step0/q_policy_obs"] = agent.policy_network(ot).numpy().copy()
                out["step0/q_policy_next"] = agent.policy_network(nt).numpy().copy()
                out["step0/q_target_next"] = agent.target_network(nt).numpy().copy()
        # capture the raw (unclipped) gradients: clip_grad_norm_ scales .grad in place, so hook backward
        raw = {}
        hooks = [p.register_hook(lambda g, n=n: raw.__setitem__(n, g.detach().clone()))
                 for n, p in agent.policy_network.named_parameters() if p.requires_grad]
        agent.train()
        for h in hooks:
            h.remove()
        if s == 0:
            for n, g in raw.items():
                out["step0/grad/" + n] = g.numpy().copy()
        rb.sample = real_sample
    stats = {}
    for nm in ("td_errors", "grad_norms", "qvalue_max", "qvalue_mean", "qvalue_min", "target_max", "target_mean", "target_min"):
        ra = getattr(agent, nm)
        vals = list(getattr(ra, "record", getattr(ra, "q", getattr(ra, "values", []))))
        stats[nm] = np.array(vals, dtype=np.float64)
    for k, v in stats.items():
        out["stats/" + k] = v
    out.update(sd_np(agent.policy_network.state_dict(), f"policy{n_steps}/"))
    out["meta"] = np.array([inner_embed, layers, context, batch, 8, n_steps])
    np.savez_compressed(os.path.join(HERE, f"train_{tag}.npz"), **out)
    print("wrote", f"train_{tag}.npz", {k: v.shape for k, v in stats.items()})


def gen_forward_short(env_id, inner_embed, layers, context, tag):
    """DTQN.forward on sequences shorter than the context (acting path, dtqn/agents/dtqn.py:81-107)."""
    agent, env = make_agent(env_id, inner_embed, layers, context, 32)
    gen = torch.Generator().manual_seed(11)
    perturb(agent.policy_network, gen)
    O, A = ENV_SHAPES[env_id]
    out = sd_np(agent.policy_network.state_dict(), "policy/")
    for L in (1, 7, context):
        if env_id == "Memory-5-v0":
            x = torch.randint(0, 9, (5, L, O), generator=gen)
        else:
            x = torch.empty(5, L, O).uniform_(-1.1, 1.1, generator=gen)
            x[:, :, 2] = torch.randint(-1, 2, (5, L), generator=gen).float()
        with torch.no_grad():
            q = agent.policy_network(x)
        out[f"L{L}/obss"], out[f"L{L}/q"] = x.numpy(), q.numpy()
    out["meta"] = np.array([inner_embed, layers, context, 8])
    np.savez_compressed(os.path.join(HERE, f"forward_{tag}.npz"), **out)
    print("wrote", f"forward_{tag}.npz")


def gen_acting(env_id, inner_embed, context, steps, tag):
    """The reference acting loop (run.step, run.py:356-377) with epsilon = 0.3: records the env/obs/action stream,
    the Context window fed to the network and the greedy Q of every step -> pins Context + get_action semantics."""
    sys.argv = ["run.py"]
    import run as ref_run
    from utils.random import RNG
    from utils import epsilon_anneal

    agent, env = make_agent(env_id, inner_embed, 2, context, 32)
    gen = torch.Generator().manual_seed(5)
    perturb(agent.policy_network, gen)
    eps = epsilon_anneal.Constant(0.3)
    agent.eval_off()
    agent.context_reset(env.reset())
    O, A = ENV_SHAPES[env_id]
    rec = dict(ctx_obs=[], ctx_len=[], action=[], obs=[], reward=[], done=[], buffer_done=[], greedy_q=[])
    for t in range(steps):
        n = min(agent.context.max_length, agent.context.timestep + 1)
        win = np.full((context, O), np.nan)
        win[:n] = agent.context.obs[:n]
        rec["ctx_obs"].append(win)
        rec["ctx_len"].append(n)
        with torch.no_grad():
            q = agent.policy_network(torch.as_tensor(agent.context.obs[:n], dtype=agent.obs_tensor_type).unsqueeze(0))
        rec["greedy_q"].append(q[0, -1].numpy().copy())
        action = agent.get_action(epsilon=eps.val)
        next_obs, reward, done, info = env.step(action)
        buffer_done = False if info.get("TimeLimit.truncated", False) else done
        agent.observe(next_obs, action, reward, buffer_done)
        rec["action"].append(action); rec["obs"].append(np.array(next_obs, dtype=np.float64))
        rec["reward"].append(reward); rec["done"].append(done); rec["buffer_done"].append(buffer_done)
        if done:
            agent.replay_buffer.flush()
            agent.context_reset(env.reset())
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(sd_np(agent.policy_network.state_dict(), "policy/"))
    out["meta"] = np.array([inner_embed, 2, context, 8])
    np.savez_compressed(os.path.join(HERE, f"acting_{tag}.npz"), **out)
    print("wrote", f"acting_{tag}.npz")


ABLATIONS = {                # name -> DTQN kwargs (run.py flags --identity / --gate gru / --a-embed 8 and their combination)
    "identity": dict(identity=True, gate="res", action_dim=0),
    "gru": dict(identity=False, gate="gru", action_dim=0),
    "aembed": dict(identity=False, gate="res", action_dim=8),
    "gtrxl_aembed": dict(identity=True, gate="gru", action_dim=8),
}


def gen_ablations():
    """Reference DTQN built directly with the ablation flags: forward Q for L = 1, 9 and the
    gradient of sum(Q^2) w.r.t. every trainable parameter -> pins the oracle's identity / GRU-gate / action-embedding
    restatement (the CUDA kernels for these flags follow the oracle)."""
    from dtqn.networks.dtqn import DTQN
    out = {}
    O, A, d, ctx, H = 3, 3, 32, 12, 4
    for name, kw in ABLATIONS.items():
        gen = torch.Generator().manual_seed(21)
        torch.manual_seed(3)
        net = DTQN(O, A, 8, kw["action_dim"], d, H, 2, ctx, dropout=0.0, gate=kw["gate"], identity=kw["identity"],
                   pos="learned", discrete=False)
        perturb(net, gen)
        out.update(sd_np(net.state_dict(), f"{name}/sd/"))
        for L in (1, 9):
            x = torch.empty(4, L, O).uniform_(-1.1, 1.1, generator=gen)
            a = torch.randint(0, A, (4, L, 1), generator=gen)
            q = net(x, a) if kw["action_dim"] else net(x)
            out[f"{name}/L{L}/obss"], out[f"{name}/L{L}/actions"] = x.numpy(), a.numpy()
            out[f"{name}/L{L}/q"] = q.detach().numpy().copy()
            if L == 9:
                net.zero_grad()
                (q ** 2).sum().backward()
                for n, p_ in net.named_parameters():
                    if p_.grad is not None:
                        out[f"{name}/grad/{n}"] = p_.grad.detach().numpy().copy()
    out["meta"] = np.array([d, 2, ctx, H])
    np.savez_compressed(os.path.join(HERE, "forward_ablations.npz"), **out)
    print("wrote forward_ablations.npz")


def gen_checkpoint():
    """A checkpoint WRITTEN BY THE REFERENCE (DqnAgent.save_checkpoint, dqn.py:222-279: `_checkpoint.pt` with pickled
    utils.logging_utils.RunningAverage objects + the five joblib `.sav` buffers + `_mini_checkpoint.pt`) after a short run,
    followed by one more reference train() on a recorded batch -> the parameters a correctly resumed agent must reach."""
    sys.argv = ["run.py"]
    import run as ref_run
    import random
    from utils import epsilon_anneal
    from utils.logging_utils import RunningAverage
    out_dir = os.path.join(HERE, "refckpt")
    os.makedirs(out_dir, exist_ok=True)
    batch, ctx = 4, 50
    agent, env = make_agent("DiscreteCarFlag-v0", 64, 2, ctx, batch, buf=4000)
    ref_run.prepopulate(agent, 1500, [env])
    agent.eval_off()
    agent.context_reset(env.reset())
    eps = epsilon_anneal.LinearAnneal(1.0, 0.1, 300)
    random.seed(5)
    for _ in range(40):                                     # the loop body of run.train (run.py:290-298)
        if ref_run.step(agent, env, eps):
            agent.replay_buffer.flush()
            agent.context_reset(env.reset())
        agent.train()
        eps.anneal()
    succ, rew, length = RunningAverage(10), RunningAverage(10), RunningAverage(10)
    for v in (0.0, 0.5, 1.0):
        succ.add(v); rew.add(-v); length.add(100 * v + 7)
    agent.save_checkpoint(os.path.join(out_dir, "ref"), None, succ, rew, length, eps)
    rb = agent.replay_buffer
    out = {"meta": np.array([64, 2, ctx, batch, 8, agent.num_train_steps]), "epsilon": np.array([eps.val]),
           "replay_pos": np.array(rb.pos), "td_errors_mean": np.array([agent.td_errors.mean()]),
           "succ_mean": np.array([succ.mean()])}
    state = random.getstate()
    batch_np = rb.sample(batch)
    random.setstate(state)
    valid = [i for i in range(min(rb.pos[0], rb.max_size)) if i != rb.pos[0] % rb.max_size]
    eps_idx = np.array([random.choice(valid) for _ in range(batch)])
    starts = np.array([random.randint(0, max(0, int(rb.episode_lengths[e]) - rb.context_len)) for e in eps_idx])
    out["episodes"], out["starts"] = eps_idx, starts
    for name, arr in zip(("obss", "actions", "rewards", "next_obss", "next_actions", "dones", "eplens"), batch_np):
        out["batch/" + name] = np.asarray(arr).copy()
    rb.sample = lambda bs, b=batch_np: b
    agent.train()
    out.update(sd_np(agent.policy_network.state_dict(), "after/"))
    out["after_grad_norm"] = np.array([list(agent.grad_norms.q)[-1]])
    np.savez_compressed(os.path.join(out_dir, "expect.npz"), **out)
    print("wrote refckpt/", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ablations":
        gen_ablations()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        gen_checkpoint()
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, "env_carflag.npz"), **gen_env("DiscreteCarFlag-v0", 24, 700))
    np.savez_compressed(os.path.join(HERE, "env_memory.npz"), **gen_env("Memory-5-v0", 24, 400))
    print("wrote env fixtures")
    gen_forward_short("DiscreteCarFlag-v0", 64, 2, 50, "carflag")
    gen_forward_short("Memory-5-v0", 128, 2, 50, "memory")
    gen_train("DiscreteCarFlag-v0", 64, 2, 50, 32, 12000, 3, "carflag")
    gen_train("Memory-5-v0", 128, 1, 50, 8, 1500, 2, "memory")
    gen_acting("DiscreteCarFlag-v0", 64, 50, 260, "carflag")
    gen_ablations()
    gen_checkpoint()
