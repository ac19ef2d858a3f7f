"""GPU parity: batched env kernels vs the golden transitions recorded from the reference, vs the oracle at scale,
and the fused replay append / history-window gather vs the reference's buffer after prepopulate."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ENV_FILES = [("DiscreteCarFlag-v0", "env_carflag.npz"), ("Memory-5-v0", "env_memory.npz")]


def _cur_first(env):
    return env.current_obs()[:, 0].cpu().numpy()


@pytest.mark.parametrize("env_id,fname", ENV_FILES)
def test_env_step_bit_exact_vs_reference_golden(golden_dir, env_id, fname):
    from dtqn_b200.envs import BatchedEnv
    z = np.load(os.path.join(golden_dir, fname))
    seeds, tapes = z["seeds"], z["actions"]
    S, T = tapes.shape
    env = BatchedEnv(env_id, S, seeds=seeds, device="cuda")
    env.reset_all()
    assert np.array_equal(env.current_obs().cpu().numpy(), z["initial_obs"])          # f64 bit-exact
    for t in range(T):
        env.step(torch.from_numpy(tapes[:, t].astype(np.int32)).cuda())
        assert np.array_equal(env.obs_out.cpu().numpy(), z["obs"][:, t].astype(np.float32)), t
        assert np.array_equal(env.reward_out.cpu().numpy(), z["reward"][:, t].astype(np.float32)), t
        assert np.array_equal(env.done_out.cpu().numpy().astype(bool), z["done"][:, t]), t
        assert np.array_equal(env.truncated_out.cpu().numpy().astype(bool), z["truncated"][:, t]), t
        assert np.array_equal(env.success_out.cpu().numpy().astype(bool), z["success"][:, t]), t
        # full-precision state: the f64 observation after the step, or the reset observation where the episode ended
        want = np.where(z["done"][:, t, None], z["reset_obs"][:, t], z["obs"][:, t])
        assert np.array_equal(env.current_obs().cpu().numpy(), want), t
    assert np.array_equal(env.rng_state(), z["rng_state"])                            # PCG64 state after the tape


@pytest.mark.parametrize("env_id", ["DiscreteCarFlag-v0", "Memory-5-v0"])
def test_env_at_scale_vs_oracle(env_id):
    """4096 envs x 450 steps (crosses resets / truncation), a subset replayed through the CPU oracle."""
    from dtqn_b200.envs import BatchedEnv
    from oracle import envs as oenvs
    N, T = 4096, 450
    env = BatchedEnv(env_id, N, seed=1000, device="cuda")
    env.reset_all()
    g = torch.Generator(device="cuda").manual_seed(3)
    acts, obs, rew, done = [], [], [], []
    for t in range(T):
        a = torch.randint(0, env.num_actions, (N,), generator=g, device="cuda", dtype=torch.int32)
        env.step(a)
        acts.append(a.cpu().numpy()); obs.append(env.obs_out.cpu().numpy()); rew.append(env.reward_out.cpu().numpy())
        done.append(env.done_out.cpu().numpy())
    acts, obs, rew, done = map(np.stack, (acts, obs, rew, done))
    rs = env.rng_state()
    for i in list(range(0, N, 173)) + [N - 1]:
        r = oenvs.rollout(env_id, int(env.seeds[i]), acts[:, i])
        assert np.array_equal(obs[:, i], r["obs"].astype(np.float32)), i
        assert np.array_equal(rew[:, i], r["reward"].astype(np.float32)), i
        assert np.array_equal(done[:, i].astype(bool), r["done"]), i
        assert np.array_equal(rs[i], r["rng_state"]), i
    st = env.ep_stats.cpu().numpy()
    assert st[3] == done.sum() and st[1] > 0


@pytest.mark.parametrize("tag,env_id,prepop", [("carflag", "DiscreteCarFlag-v0", 12000), ("memory", "Memory-5-v0", 1500)])
def test_prepopulate_replay_identical_to_reference(golden_dir, tag, env_id, prepop):
    """N = 1: device env + device agent stream + fused replay append reproduce run.prepopulate (run.py:380-405)
    byte-for-byte: same random actions (global RNG.rng), same transitions, same buffer arrays."""
    from dtqn_b200.envs import BatchedEnv, ContextWindow
    from dtqn_b200.buffers import ReplayBuffer
    z = np.load(os.path.join(golden_dir, f"train_{tag}.npz"))
    ctx = int(z["meta"][2])
    env = BatchedEnv(env_id, 1, seed=1, device="cuda")
    env.reset_all(); env.reset_all()                 # get_agent's two hidden env.reset() (env_processing.py:67)
    rb = ReplayBuffer(50_000, env.obs_dim, env.obs_mask, env.max_episode_steps, ctx, n_envs=1, device="cuda")
    cx = ContextWindow(ctx, env.obs_mask, env.num_actions, env.obs_dim, n_envs=1, device="cuda")
    env.attach(rb, cx)
    env.reset_all()
    n_ep = int(z["replay/pos"][0])
    steps = 0
    # prepopulate finishes the episode in flight when the budget is reached -> run until n_ep episodes are complete
    while True:
        for _ in range(64):
            env.step()                                # DTQN_ACT_RANDOM
            steps += 1
        if int(rb.counters[2].item()) >= n_ep:
            break
        assert steps < 4 * prepop
    lens = rb.episode_lengths[:n_ep].cpu().numpy()
    assert np.array_equal(lens, z["replay/episode_lengths"])
    assert np.array_equal(rb.obss[:n_ep].cpu().numpy(), z["replay/obss"])
    assert np.array_equal(rb.actions[:n_ep].cpu().numpy(), z["replay/actions"])
    assert np.array_equal(rb.rewards[:n_ep].cpu().numpy(), z["replay/rewards"])
    assert np.array_equal(rb.dones[:n_ep].cpu().numpy().astype(bool), z["replay/dones"])
    # history-window gather == ReplayBuffer.sample for the reference's own index draw
    eps = torch.from_numpy(z["step0/episodes"]).int().cuda()
    starts = torch.from_numpy(z["step0/starts"]).int().cuda()
    out = rb.sample(len(eps), indices=(eps, starts))
    for got, name in zip(out, ("obss", "actions", "rewards", "next_obss", "next_actions", "dones", "eplens")):
        assert np.array_equal(got.cpu().numpy(), z["step0/" + name]), name


def test_replay_ring_multi_env_invariants():
    """4096 envs sharing one ring: slots handed out in env-index order, closed slots byte-identical to what a full
    cleanse would leave (fill values beyond the episode), open slots excluded from sampling."""
    from dtqn_b200.envs import BatchedEnv, ContextWindow
    from dtqn_b200.buffers import ReplayBuffer
    N, ctx = 4096, 50
    env = BatchedEnv("DiscreteCarFlag-v0", N, seed=7, device="cuda")
    rb = ReplayBuffer(200 * 8 * N, 3, -5.0, 200, ctx, n_envs=N, device="cuda", sample_seed=5)
    cx = ContextWindow(ctx, -5.0, 3, 3, n_envs=N, device="cuda")
    env.attach(rb, cx)
    env.reset_all()
    assert np.array_equal(rb.env_slot.cpu().numpy(), np.arange(N))
    total_done = 0
    for t in range(1300):
        env.step()
        total_done += int(env.done_out.sum().item()) if t % 50 == 0 else 0
    c = rb.counters.cpu().numpy()
    assert c[3] == 0 and c[1] == N + c[2] and c[2] == env.ep_stats[3].item()
    S, E = rb.max_size, 200
    used = min(int(c[1]), S)
    lens = rb.episode_lengths[:used].cpu().numpy()
    open_ = rb.slot_open[:used].cpu().numpy().astype(bool)
    assert open_.sum() == N
    closed = np.where(~open_)[0]
    obss = rb.obss[:used].cpu().numpy(); dones = rb.dones[:used].cpu().numpy()[..., 0]
    acts = rb.actions[:used].cpu().numpy()[..., 0]; rews = rb.rewards[:used].cpu().numpy()[..., 0]
    for s in closed[:: max(1, len(closed) // 400)]:
        L = lens[s]
        assert 1 <= L <= E
        assert np.all(obss[s, L + 1:] == -5.0) and np.all(obss[s, : L + 1, 0] != -5.0)
        assert np.all(dones[s, L:] == 1) and np.all(acts[s, L:] == 0) and np.all(rews[s, L:] == 0)
        assert np.all(dones[s, : L - 1] == 0)
    eps, starts = rb.draw_indices(8192)
    eps, starts = eps.cpu().numpy(), starts.cpu().numpy()
    assert not open_[eps].any() and np.all(lens[eps] > 0)
    assert np.all(starts >= 0) and np.all(starts <= np.maximum(0, lens[eps] - ctx))
    # context ring == last min(ctx, t+1) observations of the open episode (truncated toward zero, A-Q2)
    win, n = cx.windows()
    win, n = win.cpu().numpy(), n.cpu().numpy()
    slots = rb.env_slot.cpu().numpy(); el = env.elapsed.cpu().numpy()
    for i in range(0, N, 97):
        t = el[i]
        assert n[i] == min(ctx, t + 1)
        want = np.trunc(obss[slots[i], t + 1 - n[i]: t + 1])
        assert np.array_equal(win[i, : n[i]], want), i


@pytest.mark.parametrize("eps_on_device", [False, True])
def test_eps_greedy_stream_matches_reference(golden_dir, eps_on_device):
    """the N = 1 device loop at epsilon = 0.3 reproduces the reference's recorded acting loop draw for draw
    (agents/dtqn.py:76-107 through run.step, run.py:356-377): RNG.rng.random() < eps as an f64 comparison, integers(A)
    when exploring, argmax of the network's last-position Q otherwise; the Context window the network sees, the action,
    the observation / reward / done the env returns and the buffer_done the replay stores are compared every step."""
    from dtqn_b200.agents import DtqnAgent
    from dtqn_b200.envs import BatchedEnv
    from dtqn_b200.networks import DTQN
    z = np.load(os.path.join(golden_dir, "acting_carflag.npz"))
    d, layers, ctx, heads = [int(v) for v in z["meta"]]

    def mk():
        net = DTQN(3, 3, 8, 0, d, heads, layers, ctx, pos="learned", device="cuda")
        net.load_state_dict({k[len("policy/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("policy/")})
        return net
    agent = DtqnAgent(mk, 50_000, "cuda", 3, 200, -5, 3, False, context_len=ctx, n_envs=1)
    env = BatchedEnv("DiscreteCarFlag-v0", 1, seed=1, device="cuda")
    env.reset_all(); env.reset_all()                 # get_agent's two hidden env.reset() (env_processing.py:67)
    rb = agent.replay_buffer
    env.attach(rb, agent.train_context)
    env.reset_all()                                  # agent.context_reset(env.reset())
    eps_dev = torch.tensor([0.3], dtype=torch.float64, device="cuda") if eps_on_device else None
    T = len(z["action"])
    n_explore = 0
    for t in range(T):
        win, n = agent.context.windows()
        n = int(n[0].item())
        assert n == int(z["ctx_len"][t]), t
        assert np.array_equal(win[0, :n].cpu().numpy(), z["ctx_obs"][t][:n].astype(np.float32)), t
        slot, t_ep = int(rb.env_slot[0].item()), int(env.elapsed[0].item())
        agent.act_and_step(env, 0.0 if eps_on_device else 0.3, epsilon_dev=eps_dev)
        a = int(env.actions[0].item())
        assert a == int(z["action"][t]), (t, a, int(z["action"][t]), z["greedy_q"][t])
        n_explore += a != int(np.argmax(z["greedy_q"][t]))
        assert np.array_equal(env.obs_out[0].cpu().numpy(), z["obs"][t].astype(np.float32)), t
        assert float(env.reward_out[0].item()) == float(z["reward"][t]), t
        assert bool(env.done_out[0].item()) == bool(z["done"][t]), t
        assert bool(rb.dones[slot, t_ep, 0].item()) == bool(z["buffer_done"][t]), t
        assert int(rb.actions[slot, t_ep, 0].item()) == a and float(rb.rewards[slot, t_ep, 0].item()) == float(z["reward"][t])
    assert n_explore > 20 and z["done"].sum() >= 1        # the tape really exercises both branches and an episode roll


def test_eps_anneal_on_a_side_stream_tracks_the_host_schedule():
    """dtqn_eps_anneal takes a 64-bit stream handle (ctypes argtypes declared): run it on a non-default stream and compare
    the device schedule with LinearAnneal (utils/epsilon_anneal.py:33-34) bit for bit in f64."""
    from dtqn_b200 import _lib
    from dtqn_b200.utils import LinearAnneal
    eps = LinearAnneal(1.0, 0.1, 37)
    state = torch.tensor([eps.val, eps.min, eps.duration], dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(60):
            _lib.check(_lib.lib.dtqn_eps_anneal(state.data_ptr(), out.data_ptr(), _lib.stream_ptr()), "dtqn_eps_anneal")
            side.synchronize()
            assert float(out.item()) == eps.val
            eps.anneal()
    assert float(state[0].item()) == eps.val
