"""GPU parity of the ablation-flag networks (run.py:98-103,151-167): --identity
(transformer.py:81-101), --gate gru (gates.py:5-31, gates shared by all layers, dtqn.py:107-131), --a-embed (dtqn.py:184-192,
utils/context.py:50,77) and --dropout.  Forward Q against the reference module's recorded outputs
(tests/golden/forward_ablations.npz, Q within 1e-3 rel as the north-star states; 1e-5 asserted), the training step (loss,
statistics, raw gradients, post-Adam parameters) against the CPU oracle, whose variants are pinned to the reference's autograd
by tests/test_oracle_vs_golden.py::test_oracle_ablation_variants_match_reference."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

VARIANTS = {"identity": dict(identity=True, gate="res", action_dim=0), "gru": dict(identity=False, gate="gru", action_dim=0),
            "aembed": dict(identity=False, gate="res", action_dim=8), "gtrxl_aembed": dict(identity=True, gate="gru", action_dim=8)}


def rel_err(got, ref):
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12))


def _sd(z, name):
    pre = f"{name}/sd/"
    return {k[len(pre):]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_forward_matches_reference_module(golden_dir, name):
    from dtqn_b200.networks import DTQN
    z = np.load(os.path.join(golden_dir, "forward_ablations.npz"))
    d, layers, ctx, H = [int(v) for v in z["meta"]]
    kw = VARIANTS[name]
    net = DTQN(3, 3, 8, kw["action_dim"], d, H, layers, ctx, gate=kw["gate"], identity=kw["identity"], pos="learned", device="cuda")
    sd = _sd(z, name)
    assert set(net.state_dict()) == set(sd)                       # the reference's key names, shared gates repeated per layer
    net.load_state_dict(sd)
    n_ref = sum(v.numel() for k, v in sd.items() if not k.startswith("transformer_layers.1.attn_gate")
                and not k.startswith("transformer_layers.1.mlp_gate"))
    assert sum(p.numel() for p in net.parameters()) == n_ref      # shared gates counted once, like nn.Module.parameters()
    for L in (1, 9):
        x = torch.from_numpy(z[f"{name}/L{L}/obss"])
        a = torch.from_numpy(z[f"{name}/L{L}/actions"])
        q = net(x, a).cpu().numpy()
        assert rel_err(q, z[f"{name}/L{L}/q"]) < 1e-5, (name, L)


@pytest.mark.parametrize("name,d,H,ctx,B", [("identity", 64, 8, 50, 6), ("gru", 64, 8, 50, 5), ("aembed", 64, 8, 50, 6),
                                            ("gtrxl_aembed", 32, 4, 12, 7), ("gtrxl_aembed", 128, 8, 20, 4)])
def test_variant_train_step_vs_oracle(name, d, H, ctx, B):
    from dtqn_b200.agents import DtqnAgent
    from dtqn_b200.networks import DTQN
    from oracle import network as onet, agent as oagent
    kw = VARIANTS[name]
    g = torch.Generator().manual_seed(5)
    O, A = 3, 3
    mk = lambda: DTQN(O, A, 8, kw["action_dim"], d, H, 2, ctx, gate=kw["gate"], identity=kw["identity"], pos="learned", device="cuda")
    agent = DtqnAgent(mk, 4000, "cuda", O, 200, -5, A, False, batch_size=B, context_len=ctx, history=ctx)
    with torch.no_grad():                                         # non-trivial parameters (the init has zero biases / position table)
        for n_, p in agent.policy_network.named_parameters():
            if n_.endswith("attn_mask"):
                continue
            noise = torch.empty(p.shape).normal_(0, 0.05, generator=g).cuda()
            p.copy_(p * 3.0 + noise if p.dim() > 1 else p + noise)
    agent.policy_network.packed_stale = True
    agent.target_update()
    with torch.no_grad():
        agent.target_network.flat.add_(torch.empty(agent.target_network.flat.shape).normal_(0, 0.01, generator=g).cuda())
    sd = {k: v.detach().cpu().clone() for k, v in agent.policy_network.state_dict().items()}
    tsd = {k: v.detach().cpu().clone() for k, v in agent.target_network.state_dict().items()}
    tr = oagent.TrainerOracle(sd, H, identity=kw["identity"])
    tr.target = tsd
    for step in range(2):
        win = torch.empty(B, ctx + 1, O).uniform_(-1.1, 1.1, generator=g)
        act = torch.randint(0, A, (B, ctx + 1), generator=g).to(torch.uint8)
        rew = torch.randint(-1, 2, (B, ctx), generator=g).float()
        done = (torch.rand(B, ctx, generator=g) < 0.1).to(torch.uint8)
        agent.train_on_windows(win.cuda(), act.cuda(), rew.cuda(), done.cuda())
        agent.check_finite()
        batch = (win[:, :-1], act[:, :-1, None].long(), rew[..., None], win[:, 1:], act[:, 1:, None].long(), done[..., None].bool())
        stats, grads = tr.train_on_batch(batch)
        st = agent.stats.cpu().numpy()
        ref = [stats[k] for k in ("loss", "q_max", "q_mean", "q_min", "t_max", "t_mean", "t_min", "grad_norm")]
        for k, (a_, r_) in enumerate(zip(st, ref)):
            assert abs(a_ - r_) <= 1e-4 * max(1.0, abs(r_)), (name, step, k, a_, r_)
        if step == 0:
            got = agent.policy_network.unflatten(agent.grads)
            gmax = max(float(v.abs().max()) for v in grads.values())
            assert set(got) == set(grads)
            for k, gr in grads.items():
                err = float((got[k].cpu() - gr).abs().max())
                assert err <= 1e-3 * max(float(gr.abs().max()), 1e-3 * gmax), (name, k, err, float(gr.abs().max()))
    for k, p in agent.policy_network.state_dict().items():
        if not k.endswith("attn_mask"):
            assert float((p.cpu() - tr.policy[k]).abs().max()) < 5e-5, (name, k)   # two Adam steps of 3e-4 each


def test_action_embedding_acting_context_matches_oracle():
    """--a-embed in the acting loop: the device Context keeps the action ring (utils/context.py:50,77: action[t] = the action
    that led to obs[t], slot 0 = the first random padding draw) and get_action feeds it as-is (agents/dtqn.py:87-103)."""
    from dtqn_b200.agents import DtqnAgent
    from dtqn_b200.envs import BatchedEnv
    from dtqn_b200.networks import DTQN
    from oracle import network as onet
    N, ctx, d, H = 6, 10, 32, 4
    mk = lambda: DTQN(3, 3, 8, 8, d, H, 2, ctx, pos="learned", device="cuda")
    agent = DtqnAgent(mk, 200 * 8 * N, "cuda", 3, 200, -5, 3, False, context_len=ctx, n_envs=N)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n_, p in agent.policy_network.named_parameters():
            if not n_.endswith("attn_mask"):
                p.copy_(p * 3.0 + torch.empty(p.shape).normal_(0, 0.05, generator=g).cuda())
    sd = {k: v.detach().cpu().clone() for k, v in agent.policy_network.state_dict().items()}
    env = BatchedEnv("DiscreteCarFlag-v0", N, seed=21, device="cuda")
    env.attach(agent.replay_buffer, agent.train_context)
    env.reset_all()
    hist = [[] for _ in range(N)]                                  # per env: actions of the running episode
    pad0 = agent.context.action[:, 0].cpu().numpy().copy()
    for t in range(37):
        win, n = agent.context.windows()
        ts = agent.context.timestep_t.cpu().numpy()
        ring_a = agent.context.action.cpu().numpy()
        q = agent.q_last_batched().cpu().numpy()
        for i in range(N):
            ni = int(n[i])
            # the reference's Context.action[:n]: slot 0 the padding draw (until evicted), then the actions taken so far
            full = [int(pad0[i])] + hist[i]
            want_a = np.array(full[-ni:] if len(full) > ctx else full[:ni])
            got_a = np.array([ring_a[i, (int(ts[i]) + 1 - ni + j) % ctx] for j in range(ni)])
            assert np.array_equal(got_a, want_a), (t, i)
            with torch.no_grad():
                ref = onet.forward(sd, win[i:i + 1, :ni].cpu(), H, actions=torch.from_numpy(want_a)[None, :, None])[0, -1].numpy()
            assert np.abs(q[i] - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (t, i)
        agent.act_and_step(env, 0.4)
        acts, done = env.actions.cpu().numpy(), env.done_out.cpu().numpy()
        for i in range(N):
            if done[i]:
                hist[i] = []
                pad0[i] = int(agent.context.action[i, 0].item())
            else:
                hist[i].append(int(acts[i]))


def test_dropout_masks_and_gradients():
    """--dropout p: (i) the mask generator keeps with probability 1 - p and scales by 1 / (1 - p); (ii) eval-mode forward ==
    the p = 0 network; (iii) train-mode forward changes from call to call but the backward regenerates the forward's masks:
    the analytic TD gradient matches central finite differences of the loss with the mask stream frozen."""
    from dtqn_b200 import _lib
    from dtqn_b200.agents import DtqnAgent
    from dtqn_b200.networks import DTQN
    lib = _lib.lib
    lib.dtqn_dropout_scales.argtypes = [C.c_uint64, C.c_uint32, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
    p = 0.25
    out = torch.zeros(1 << 20, device="cuda")
    for counter, site in ((0, 1), (7, 2), (7, 3)):
        assert lib.dtqn_dropout_scales(counter, site, p, out.numel(), out.data_ptr(), _lib.stream_ptr()) == 0
        keep = float((out > 0).float().mean())
        assert abs(keep - (1 - p)) < 3e-3 and abs(float(out.mean()) - 1.0) < 5e-3
        assert set(np.unique(out.cpu().numpy()).tolist()) == {0.0, np.float32(1 / (1 - p))}
    a = out.clone()
    lib.dtqn_dropout_scales(8, 3, p, out.numel(), out.data_ptr(), _lib.stream_ptr())
    assert 0.3 < float((a != out).float().mean()) < 0.45            # a new counter value draws fresh masks: 2 p (1 - p) differ
    O, A, d, H, ctx, B = 3, 3, 64, 8, 8, 3          # d = 64 so that the p = 0 twin is a default-architecture network
    mk = lambda pp: DTQN(O, A, 8, 0, d, H, 2, ctx, dropout=pp, pos="learned", device="cuda")
    agent = DtqnAgent(lambda: mk(p), 4000, "cuda", O, 200, -5, A, False, batch_size=B, context_len=ctx, history=ctx)
    g = torch.Generator().manual_seed(1)
    net = agent.policy_network
    with torch.no_grad():
        for n_, q in net.named_parameters():
            if not n_.endswith("attn_mask"):
                q.copy_(q * 4.0 + torch.empty(q.shape).normal_(0, 0.05, generator=g).cuda())
    agent.target_update()
    plain = mk(0.0)
    plain.load_state_dict(net.state_dict())
    x = torch.empty(5, ctx, O).uniform_(-1, 1, generator=g)
    net.eval()
    # eval mode: dropout is the identity (the p = 0 twin runs the fused default-architecture kernels: same values, other order)
    assert float((net(x) - plain(x)).abs().max()) <= 1e-5 * float(plain(x).abs().max())
    net.train()
    q1, q2 = net(x), net(x)
    assert not torch.equal(q1, q2) and not torch.equal(q1, plain(x))
    win = torch.empty(B, ctx + 1, O).uniform_(-1.1, 1.1, generator=g).cuda()
    act = torch.randint(0, A, (B, ctx + 1), generator=g).to(torch.uint8).cuda()
    rew = torch.randint(-1, 2, (B, ctx), generator=g).float().cuda()
    done = (torch.rand(B, ctx, generator=g) < 0.1).to(torch.uint8).cuda()

    def loss_at(counter):
        net.dropout_state.fill_(counter)
        agent.forward_backward(win, act, rew, done)                # TD loss + backward; advances the mask stream by one
        return float(agent.stats[0].item())
    loss_at(11)
    grads = agent.grads.clone()
    assert int(net.dropout_state.item()) == 12
    # the target y also moves with the policy parameters through a* = argmax policy(next): keep eps small, pick big gradients
    idx = torch.topk(grads.abs(), 6).indices.tolist()
    for i in idx:
        eps = 2e-3
        old = float(net.flat[i].item())
        net.flat[i] = old + eps; lp = loss_at(11)
        net.flat[i] = old - eps; lm = loss_at(11)
        net.flat[i] = old
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - float(grads[i])) <= 0.05 * abs(float(grads[i])) + 1e-4, (i, fd, float(grads[i]))
