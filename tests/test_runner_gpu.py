"""GPU checks of the batched training loop: eager vs CUDA-graph replay agree, learning statistics are finite, evaluation
runs without touching the replay, the host-call (reference-style) single-env API drives the same kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(**kw):
    from dtqn_b200.runner import BatchedTrainer
    args = dict(env_id="DiscreteCarFlag-v0", n_envs=512, seed=3, device="cuda", batch=16, num_steps=10_000)
    args.update(kw)
    return BatchedTrainer(**args)


@pytest.mark.parametrize("tuf", [10_000, 4])
def test_graph_replay_matches_eager(tuf):
    """tuf = 4: three hard target updates inside the window -- the target's weight images (which the graph's training forward
    streams) must follow the host-side target_update between replays."""
    a, b = _trainer(tuf=tuf), _trainer(tuf=tuf)
    for t in (a, b):
        t.prepopulate(230)
        assert t.agent.replay_buffer.can_sample(16)
    b.enable_graphs()
    a.train_iteration()            # the capture warm-up of b already performed one iteration
    for _ in range(12):
        a.train_iteration(); b.train_iteration()
    torch.cuda.synchronize()
    assert a.agent.num_train_steps == b.agent.num_train_steps == 13
    pa, pb = a.agent.policy_network.flat, b.agent.policy_network.flat
    assert torch.isfinite(pa).all() and torch.isfinite(pb).all()
    # identical algorithm and random streams, and every gradient sum has a fixed order (no fp32 atomics): bit-identical
    assert torch.equal(pa, pb) and torch.equal(a.agent.exp_avg_sq, b.agent.exp_avg_sq)
    assert torch.equal(a.agent.target_network.flat, b.agent.target_network.flat)
    if tuf == 4:      # the image the graph streams for the target network == a fresh pack of its current parameters
        img = b.agent.target_network.packed.clone()
        b.agent.target_network.repack(); torch.cuda.synchronize()
        assert torch.equal(img, b.agent.target_network.packed)
    assert np.array_equal(a.env.rng_state(), b.env.rng_state())
    assert a.agent.td_errors.mean() == b.agent.td_errors.mean()
    assert int(a.agent.opt_step.item()) == 13 and int(b.agent.opt_step.item()) == 13


def test_training_reduces_td_error_and_eval_runs():
    t = _trainer(n_envs=1024, batch=32, lr=1e-3)
    t.prepopulate(230)
    t.enable_graphs()
    first = None
    for i in range(300):
        t.train_iteration()
        if i == 20:
            first = t.agent.td_errors.mean()
    last = t.agent.td_errors.mean()
    assert np.isfinite(last) and last < first
    before = t.agent.replay_buffer.counters.clone()
    sr, ret, length = t.evaluate(1)
    assert 0.0 <= sr <= 1.0 and -1.0 <= ret <= 1.0 and 1 <= length <= 200
    assert torch.equal(before, t.agent.replay_buffer.counters)       # evaluation never writes the replay (agents/dtqn.py:159)


def test_memory_env_trains():
    t = _trainer(env_id="Memory-5-v0", n_envs=512, batch=16, inner_embed=128)
    t.prepopulate(60)
    for _ in range(5):
        t.train_iteration()
    t.agent.check_finite()
    assert np.isfinite(t.agent.td_errors.mean())


def test_host_call_api_single_env():
    """run.py-style use: host env loop calling context_reset / get_action / observe / replay_buffer.flush / train."""
    from dtqn_b200.utils import get_agent, set_global_seed
    from dtqn_b200.envs import BatchedEnv
    from oracle import envs as oenvs
    spec = BatchedEnv("DiscreteCarFlag-v0", 1, device="cuda")           # only used for its space metadata
    set_global_seed(1)
    agent = get_agent("DTQN", [spec], 8, 0, 64, 20_000, "cuda", 3e-4, 4, 50, -1, 50, 10_000, 0.99, num_heads=8,
                      num_layers=2, n_envs=1)
    env = oenvs.make("DiscreteCarFlag-v0", 1)
    agent.eval_off()
    agent.context_reset(env.reset())
    episodes = 0
    for step in range(1400):
        a = agent.get_action(epsilon=0.5)
        o, r, done, info = env.step(a)
        agent.observe(o, a, r, False if info.get("TimeLimit.truncated", False) else done)
        if done:
            agent.replay_buffer.flush()
            agent.context_reset(env.reset())
            episodes += 1
        agent.train()
    assert episodes >= 6 and agent.num_train_steps > 0
    agent.check_finite()
    assert agent.replay_buffer.pos[0] == episodes


def test_host_loop_graphs_match_eager_api():
    """BatchedTrainer.host_q / host_step / host_train (three CUDA graphs + pinned host buffers) against the same calls made
    eagerly through the agent / env API with the same host-chosen actions."""
    from dtqn_b200 import _lib
    a, b = _trainer(n_envs=256, batch=8), _trainer(n_envs=256, batch=8)
    for t in (a, b):
        t.prepopulate(230)
    b.enable_host_loop()
    # the three capture warm-ups of b executed one acting forward, one env step (with b.env.actions) and one train step
    a.agent.q_last_batched()
    a.env.step(actions=b.env.actions.clone(), mode=_lib.ACT_GIVEN)
    a.agent.train()
    rng = np.random.default_rng(0)
    for it in range(6):
        qa = a.agent.q_last_batched().cpu().numpy()
        qb = b.host_q().numpy().copy()
        assert np.abs(qa - qb).max() <= 2e-5 * max(1.0, np.abs(qa).max()), it
        act = torch.from_numpy(np.where(rng.random(256) < 0.3, rng.integers(0, 3, 256), qb.argmax(1)).astype(np.int32))
        a.env.step(actions=act.cuda(), mode=_lib.ACT_GIVEN)
        obs, rew, done = b.host_step(act.pin_memory())
        a.agent.train()
        stats = b.host_train().numpy().copy()
        assert np.array_equal(obs.numpy(), a.env.obs_out.cpu().numpy()) and np.array_equal(done.numpy(), a.env.done_out.cpu().numpy())
        assert np.array_equal(rew.numpy(), a.env.reward_out.cpu().numpy())
        assert np.allclose(stats, a.agent.stats.cpu().numpy(), rtol=1e-3, atol=1e-5), (it, stats, a.agent.stats)
    assert np.array_equal(a.env.rng_state(), b.env.rng_state())
    assert a.agent.num_train_steps == b.agent.num_train_steps
    assert torch.equal(a.agent.policy_network.flat, b.agent.policy_network.flat)


def test_evaluate_matches_oracle_greedy_rollouts(golden_dir):
    """run.evaluate (run.py:187-243) on the device: greedy policy, exactly k episodes per eval env, same seeds as the train
    envs.  Every env's (return, length, successes) over its k episodes must equal a CPU-oracle greedy rollout of the same
    seed with the same weights (oracle env + Context + closed-form network), and a second evaluate() continues the eval
    envs' streams like the reference's persistent eval env does."""
    import os
    from oracle import envs as oenvs, network as onet
    from oracle.pcg64 import PCG64
    from oracle.replay import ContextOracle
    z = np.load(os.path.join(golden_dir, "acting_carflag.npz"))
    sd = {k[len("policy/"):]: torch.from_numpy(z[k]).float() for k in z.files if k.startswith("policy/")}
    N, K = 24, 2
    t = _trainer(n_envs=N, seed=11, batch=8)
    t.agent.policy_network.load_state_dict(sd)
    first = t.evaluate(K)
    acc1 = t.last_eval_per_env.cpu().numpy()
    second = t.evaluate(K)
    acc2 = t.last_eval_per_env.cpu().numpy()
    assert (acc1[:, 0] == K).all() and (acc2[:, 0] == K).all()
    for got, acc in ((first, acc1), (second, acc2)):
        tot = acc.sum(0)
        assert got == (tot[3] / tot[0], tot[1] / tot[0], tot[2] / tot[0])
    mism = 0
    for i in range(N):
        env = oenvs.make("DiscreteCarFlag-v0", int(t.env.seeds[i]))
        cx = ContextOracle(50, -5, 3, 3, PCG64.from_seed(0))
        for acc in (acc1, acc2):
            ret = length = succ = 0
            for _ in range(K):
                cx.reset(env.reset())
                done, ep_r = False, 0.0
                while not done:
                    obs, _ = cx.window()
                    with torch.no_grad():
                        q = onet.forward(sd, torch.as_tensor(obs, dtype=torch.float32).unsqueeze(0), 8)
                    a = int(torch.argmax(q[0, -1]).item())
                    o, r, done, info = env.step(a)
                    cx.add_transition(o, a)
                    ep_r += r
                ret += ep_r; length += cx.timestep
                succ += int(info.get("is_success", False) or ep_r > 0)
            mism += (int(acc[i, 1]), int(acc[i, 2]), int(acc[i, 3])) != (int(ret), int(length), int(succ))
    assert mism == 0, f"{mism} of {2 * N} per-env evaluation records differ from the oracle rollout"


def test_updates_per_step_and_record_every():
    """New scale knobs (no reference analogue: it has one env): K gradient updates per lockstep step, and every K-th episode
    of each env stored in the replay.  Eager and graph replay stay bit-identical; the optimiser has taken K steps per
    iteration; with record_every = 4 a quarter of the finished episodes took a replay slot and every closed slot is a whole,
    valid episode."""
    a, b = _trainer(updates_per_step=3, record_every=4), _trainer(updates_per_step=3, record_every=4)
    for t in (a, b):
        t.prepopulate(400)
        assert t.agent.replay_buffer.can_sample(16)
    b.enable_graphs()
    a.train_iteration()
    for _ in range(5):
        a.train_iteration(); b.train_iteration()
    torch.cuda.synchronize()
    assert a.agent.num_train_steps == b.agent.num_train_steps == 18 and int(b.agent.opt_step.item()) == 18
    assert torch.equal(a.agent.policy_network.flat, b.agent.policy_network.flat)
    rb, env = a.agent.replay_buffer, a.env
    c = rb.counters.cpu().numpy()
    finished = int(env.ep_stats[3].item())
    per_env = env.env_acc[:, 0].cpu().numpy()
    recorded = int((per_env // 4 + 1).sum())                   # episodes 0, 4, 8, ... of each env (incl. the running one)
    assert c[1] == recorded and c[3] == 0 and finished == per_env.sum()
    used = min(int(c[1]), rb.max_size)
    lens = rb.episode_lengths[:used].cpu().numpy()
    open_ = rb.slot_open[:used].cpu().numpy().astype(bool)
    obss = rb.obss[:used].cpu().numpy()
    for s in np.where(~open_)[0][:300]:
        L = lens[s]
        assert 1 <= L <= 200 and np.all(obss[s, : L + 1, 0] != -5.0) and np.all(obss[s, L + 1:] == -5.0)
        assert np.all(np.abs(np.diff(obss[s, : L + 1, 0])) <= 0.0700001)        # one trajectory: |dp| <= max velocity
    slots = rb.env_slot.cpu().numpy()
    assert ((slots >= 0) == (per_env % 4 == 0)).all()          # envs whose running episode is not stored hold slot -1
