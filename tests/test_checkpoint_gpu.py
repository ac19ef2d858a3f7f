"""Checkpoint / resume of the batched loop in the reference's file format (dtqn/agents/dqn.py:212-327, run.py:469-499):
a resumed trainer continues exactly like the one that kept running (env streams, replay, sampler, Adam, epsilon)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REFERENCE_KEYS = {"step", "wandb_id", "replay_buffer_pos", "policy_net_state_dict", "target_net_state_dict",
                  "optimizer_state_dict", "epsilon", "episode_successes", "episode_rewards", "episode_lengths",
                  "td_errors", "grad_norms", "qvalue_max", "qvalue_mean", "qvalue_min", "target_max", "target_mean",
                  "target_min", "random_rng_state", "rng_bit_generator_state", "numpy_rng_state", "torch_rng_state",
                  "torch_cuda_rng_state"}                                           # dqn.py:236-270


def _trainer(seed, **kw):
    from dtqn_b200.runner import BatchedTrainer
    args = dict(env_id="DiscreteCarFlag-v0", n_envs=256, seed=seed, device="cuda", batch=16, num_steps=2_000, tuf=4)
    args.update(kw)
    return BatchedTrainer(**args)


@pytest.mark.parametrize("graphs", [False, True])
def test_resume_continues_like_the_uninterrupted_run(tmp_path, graphs):
    prefix = str(tmp_path / "model=DTQN_seed=3")
    a = _trainer(3)
    a.prepopulate(230)
    if graphs:
        a.enable_graphs()
    for _ in range(5):
        a.train_iteration()
    a.evaluate(1)
    a.save_checkpoint(prefix, wandb_id="abc")
    for suffix in ("_mini_checkpoint.pt", "_checkpoint.pt", "buffer_obss.sav", "buffer_actions.sav", "buffer_rewards.sav",
                   "buffer_dones.sav", "buffer_eplens.sav"):
        assert os.path.exists(prefix + suffix), suffix
    ck = torch.load(prefix + "_checkpoint.pt", weights_only=False)
    assert REFERENCE_KEYS <= set(ck) and ck["wandb_id"] == "abc" and ck["step"] == a.agent.num_train_steps
    assert a.agent.load_mini_checkpoint(prefix) == {"step": a.agent.num_train_steps, "wandb_id": "abc"}
    import joblib
    rb = a.agent.replay_buffer
    obss, dones, lens = (joblib.load(prefix + f"buffer_{k}.sav") for k in ("obss", "dones", "eplens"))
    assert obss.shape == (rb.max_size, 201, 3) and obss.dtype == np.float32                 # replay_buffer.py:46-54
    assert dones.dtype == np.bool_ and dones.shape == (rb.max_size, 200, 1) and lens.dtype == np.uint8
    opt = ck["optimizer_state_dict"]
    names = [n for n, _ in a.agent.policy_network.named_parameters()]
    assert opt["param_groups"][0]["params"] == list(range(len(names))) and opt["param_groups"][0]["lr"] == 3e-4
    assert all(float(st["step"]) == ck["step"] for st in opt["state"].values())
    assert not any(names[i].endswith("attn_mask") for i in opt["state"])

    def finish(t):
        for _ in range(6):
            t.train_iteration()
        ev = t.evaluate(1)
        torch.cuda.synchronize()
        ag = t.agent
        return dict(flat=ag.policy_network.flat.clone(), tgt=ag.target_network.flat.clone(), rng=t.env.rng_state(),
                    counters=ag.replay_buffer.counters.clone(), draws=int(ag.replay_buffer.draw_counter.item()),
                    opt=int(ag.opt_step.item()), steps=ag.num_train_steps, eps=t.eps.val, ev=ev, it=t.iterations,
                    lens=ag.replay_buffer.episode_lengths.clone(), obss=ag.replay_buffer.obss.clone(),
                    td=ag.td_errors.mean(), ctx=ag.train_context.obs.clone())

    ra = finish(a)
    b = _trainer(11)                      # different seeds: everything that matters must come from the files
    b.prepopulate(3)
    b.load_checkpoint(prefix)
    if graphs:
        b.enable_graphs()                 # performs the next iteration while capturing
        b_extra = 5
    rb_ = None
    if graphs:
        for _ in range(b_extra):
            b.train_iteration()
        ev = b.evaluate(1)
        torch.cuda.synchronize()
        ag = b.agent
        rb_ = dict(flat=ag.policy_network.flat.clone(), tgt=ag.target_network.flat.clone(), rng=b.env.rng_state(),
                   counters=ag.replay_buffer.counters.clone(), draws=int(ag.replay_buffer.draw_counter.item()),
                   opt=int(ag.opt_step.item()), steps=ag.num_train_steps, eps=b.eps.val, ev=ev, it=b.iterations,
                   lens=ag.replay_buffer.episode_lengths.clone(), obss=ag.replay_buffer.obss.clone(),
                   td=ag.td_errors.mean(), ctx=ag.train_context.obs.clone())
    else:
        rb_ = finish(b)
    if graphs:                            # the device-side exploration schedule tracks the host LinearAnneal bit for bit
        assert float(a._eps_state[0].item()) == a.eps.val and float(b._eps_state[0].item()) == b.eps.val
    for k in ("counters", "lens", "obss", "ctx"):
        assert torch.equal(ra[k], rb_[k]), k
    assert np.array_equal(ra["rng"], rb_["rng"])
    for k in ("draws", "opt", "steps", "eps", "it", "ev"):
        assert ra[k] == rb_[k], (k, ra[k], rb_[k])
    # parameters: same update sequence, fixed-order gradient sums -> the resumed run is bit-identical
    assert torch.equal(ra["flat"], rb_["flat"]) and torch.equal(ra["tgt"], rb_["tgt"])
    assert ra["td"] == rb_["td"]


def test_reference_style_checkpoint_loads_into_single_env_agent(tmp_path):
    """A file without the "b200" extras (what the reference itself writes) restores networks, optimiser and replay."""
    prefix = str(tmp_path / "ref")
    a = _trainer(5, n_envs=1, buf_size=200 * 64)
    a.prepopulate(4000)
    for _ in range(3):
        a.train_iteration()
    a.save_checkpoint(prefix)
    ck = torch.load(prefix + "_checkpoint.pt", weights_only=False)
    del ck["b200"]
    torch.save(ck, prefix + "_checkpoint.pt")
    b = _trainer(6, n_envs=1, buf_size=200 * 64)
    wid, succ, rew, length, eps = b.agent.load_checkpoint(prefix)
    assert wid is None and eps == a.eps.val and b.agent.num_train_steps == 3 and int(b.agent.opt_step.item()) == 3
    assert torch.equal(a.agent.policy_network.flat, b.agent.policy_network.flat)
    assert torch.equal(a.agent.exp_avg, b.agent.exp_avg) and torch.equal(a.agent.exp_avg_sq, b.agent.exp_avg_sq)
    assert torch.equal(a.agent.replay_buffer.obss, b.agent.replay_buffer.obss)
    x = torch.randn(2, 50, 3, device="cuda")
    assert torch.equal(a.agent.policy_network(x), b.agent.policy_network(x))
    assert abs(a.agent.td_errors.mean() - b.agent.td_errors.mean()) < 1e-7


def test_reference_written_checkpoint_resumes_like_the_reference(golden_dir):
    """a `_checkpoint.pt` + five `.sav` buffers WRITTEN BY THE REFERENCE (tests/golden/refckpt, made by
    gen_golden.py through DqnAgent.save_checkpoint, dqn.py:222-279) load into the CUDA agent: networks, Adam moments + step,
    replay arrays, statistics windows -- and the next train() on the reference's recorded batch lands on the parameters the
    reference itself reached after resuming (dqn.py:281-327 -> agents/dtqn.py:162-269)."""
    from dtqn_b200.agents import DtqnAgent
    from dtqn_b200.networks import DTQN
    d = os.path.join(golden_dir, "refckpt")
    z = np.load(os.path.join(d, "expect.npz"))
    dm, layers, ctx, B, heads, steps = [int(v) for v in z["meta"]]
    mk = lambda: DTQN(3, 3, 8, 0, dm, heads, layers, ctx, pos="learned", device="cuda")
    agent = DtqnAgent(mk, 4000, "cuda", 3, 200, -5, 3, False, batch_size=B, context_len=ctx)
    wandb_id, succ, rew, length, eps = agent.load_checkpoint(os.path.join(d, "ref"))
    assert wandb_id is None and agent.num_train_steps == steps and int(agent.opt_step.item()) == steps
    assert eps == float(z["epsilon"][0]) and abs(succ.mean() - float(z["succ_mean"][0])) < 1e-12
    assert abs(agent.td_errors.mean() - float(z["td_errors_mean"][0])) < 1e-6
    rb = agent.replay_buffer
    assert rb.pos == [int(z["replay_pos"][0]), 0] and rb.can_sample(B)
    out = rb.sample(B, indices=(torch.from_numpy(z["episodes"]).int().cuda(), torch.from_numpy(z["starts"]).int().cuda()))
    for got, name in zip(out, ("obss", "actions", "rewards", "next_obss", "next_actions", "dones", "eplens")):
        assert np.array_equal(got.cpu().numpy(), z["batch/" + name]), name
    agent.train(indices=(torch.from_numpy(z["episodes"]).int().cuda(), torch.from_numpy(z["starts"]).int().cuda()))
    agent.check_finite()
    assert abs(float(agent.stats[7]) - float(z["after_grad_norm"][0])) <= 1e-4 * max(1.0, float(z["after_grad_norm"][0]))
    for k, p in agent.policy_network.state_dict().items():
        if not k.endswith("attn_mask"):
            assert np.abs(p.cpu().numpy() - z["after/" + k]).max() < 2e-5, k
