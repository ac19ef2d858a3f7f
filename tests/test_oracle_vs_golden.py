"""Pins the CPU oracle port (oracle/) against the golden fixtures produced by executing the reference
(tests/golden/gen_golden.py).  CPU-only; this is what lets the GPU parity tests trust the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import envs as oenvs
from oracle import network as onet
from oracle import agent as oagent
from oracle.replay import ReplayOracle, ContextOracle
from oracle.pcg64 import PCG64


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def sd_from(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


@pytest.mark.parametrize("env_id,fname", [("DiscreteCarFlag-v0", "env_carflag.npz"), ("Memory-5-v0", "env_memory.npz")])
def test_env_rollouts_bit_exact(golden_dir, env_id, fname):
    z = load(golden_dir, fname)
    for i, seed in enumerate(z["seeds"]):
        r = oenvs.rollout(env_id, int(seed), z["actions"][i])
        assert np.array_equal(r["initial_obs"], z["initial_obs"][i])
        assert np.array_equal(r["obs"], z["obs"][i])                      # f64 bit-exact
        assert np.array_equal(r["reward"], z["reward"][i])
        assert np.array_equal(r["done"], z["done"][i])
        assert np.array_equal(r["truncated"], z["truncated"][i])
        assert np.array_equal(r["reset_obs"], z["reset_obs"][i], equal_nan=True)
        assert np.array_equal(r["rng_state"], z["rng_state"][i])         # PCG64 state after the whole tape


def test_known_answers_survey():
    """Known answers recorded in SURVEY.md section 8c (numpy 2.3.5, seed 1)."""
    env = oenvs.make("DiscreteCarFlag-v0", 1)
    o = env.reset()
    assert float(o[0]).hex() == "0x1.7105158c043bcp-3" and env.heaven == 1.0
    ps = []
    for a in [2, 2, 2, 0, 1]:
        o, *_ = env.step(a)
        ps.append((float(o[0]).hex(), float(o[1]).hex()))
    assert ps[0] == ("0x1.74178423918bap-3", "0x1.89374bc6a7efap-10")
    assert ps[3] == ("0x1.89988a486ebacp-3", "0x1.89374bc6a7efbp-9")
    m = oenvs.make("Memory-5-v0", 1)
    m.reset()
    assert m.cards == [5, 3, 4, 1, 1, 2, 3, 5, 4, 2] and m.cur == 2
    p = PCG64.from_seed(1)
    assert p.state == 0x9C5B484BFEDB756C2A6E7D6F320FBC7E and p.inc == 0x922AF2DA2645F895A19857B95740937B


@pytest.mark.parametrize("fname,heads", [("forward_carflag.npz", 8), ("forward_memory.npz", 8)])
def test_forward_matches_reference(golden_dir, fname, heads):
    z = load(golden_dir, fname)
    sd = sd_from(z, "policy/")
    ctx = int(z["meta"][2])
    for L in (1, 7, ctx):
        x = torch.from_numpy(z[f"L{L}/obss"])
        q = onet.forward(sd, x, heads).numpy()
        ref = z[f"L{L}/q"]
        assert np.abs(q - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("tag,obs_mask", [("carflag", -5), ("memory", 8)])
def test_train_step_matches_reference(golden_dir, tag, obs_mask):
    z = load(golden_dir, f"train_{tag}.npz")
    d, layers, ctx, B, heads, n_steps = [int(v) for v in z["meta"]]
    tr = oagent.TrainerOracle(sd_from(z, "policy0/"), heads)
    tr.target = sd_from(z, "target0/")
    is_disc = tag == "memory"
    for s in range(n_steps):
        conv = (lambda a: torch.from_numpy(a).long()) if is_disc else (lambda a: torch.from_numpy(a).float())
        batch = (conv(z[f"step{s}/obss"]), torch.from_numpy(z[f"step{s}/actions"].astype(np.int64)),
                 torch.from_numpy(z[f"step{s}/rewards"]), conv(z[f"step{s}/next_obss"]),
                 torch.from_numpy(z[f"step{s}/next_actions"].astype(np.int64)), torch.from_numpy(z[f"step{s}/dones"]))
        if s == 0:
            with torch.no_grad():
                q = onet.forward(tr.policy, batch[0], heads).numpy()
            assert np.abs(q - z["step0/q_policy_obs"]).max() < 5e-6
        stats, grads = tr.train_on_batch(batch)
        assert abs(stats["loss"] - z["stats/td_errors"][s]) <= 1e-5 * max(1, abs(z["stats/td_errors"][s]))
        assert abs(stats["grad_norm"] - z["stats/grad_norms"][s]) <= 2e-5 * max(1, z["stats/grad_norms"][s])
        for nm, key in (("q_max", "qvalue_max"), ("q_mean", "qvalue_mean"), ("q_min", "qvalue_min"),
                        ("t_max", "target_max"), ("t_mean", "target_mean"), ("t_min", "target_min")):
            assert abs(stats[nm] - z["stats/" + key][s]) < 1e-5
        if s == 0:
            for k, g in grads.items():
                ref = z["step0/grad/" + k]
                assert np.abs(g.numpy() - ref).max() <= 1e-5 * max(1e-3, np.abs(ref).max()) + 1e-8, k
    for k in tr.keys:
        ref = z[f"policy{n_steps}/" + k]
        assert np.abs(tr.policy[k].numpy() - ref).max() < 2e-6, k


@pytest.mark.parametrize("tag,env_id", [("carflag", "DiscreteCarFlag-v0"), ("memory", "Memory-5-v0")])
def test_replay_store_and_gather(golden_dir, tag, env_id):
    """Re-run prepopulate (run.py:380-405) through the oracle env + oracle buffer and compare the raw arrays and
    the sampled windows with the reference's."""
    z = load(golden_dir, f"train_{tag}.npz")
    d, layers, ctx, B, heads, n_steps = [int(v) for v in z["meta"]]
    prepop = {"carflag": 12000, "memory": 1500}[tag]
    env = oenvs.make(env_id, 1)
    rng = PCG64.from_seed(1)              # the global RNG.rng (utils/random.py:31); PCG64(seed) == PCG64(SeedSequence(seed))
    env.reset(); env.reset()              # get_agent's two hidden resets (env_processing.py:67 via :86,107)
    buf = ReplayOracle(50_000, env.obs_dim, env.obs_mask, env.max_episode_steps, ctx)
    cx = ContextOracle(ctx, env.obs_mask, env.num_actions, env.obs_dim, rng)
    t = 0
    while t < prepop:
        o = env.reset(); cx.reset(o); buf.store_obs(o)
        done = False
        while not done:
            a = rng.integers(env.num_actions)
            o, r, done, info = env.step(a)
            bd = False if info.get("TimeLimit.truncated", False) else done
            cx.add_transition(o, a); buf.store(o, a, r, bd, cx.timestep)
            t += 1
        buf.flush()
    n_ep = int(z["replay/pos"][0])
    assert buf.pos == list(z["replay/pos"])
    assert np.array_equal(buf.obss[:n_ep], z["replay/obss"])
    assert np.array_equal(buf.actions[:n_ep], z["replay/actions"])
    assert np.array_equal(buf.rewards[:n_ep], z["replay/rewards"])
    assert np.array_equal(buf.dones[:n_ep], z["replay/dones"])
    assert np.array_equal(buf.episode_lengths[:n_ep], z["replay/episode_lengths"])
    out = buf.gather(z["step0/episodes"], z["step0/starts"])
    for got, name in zip(out, ("obss", "actions", "rewards", "next_obss", "next_actions", "dones", "eplens")):
        assert np.array_equal(got, z["step0/" + name]), name


def test_acting_context_and_greedy_q(golden_dir):
    """Context window contents (int64 truncation quirk) and the greedy Q of the acting forward."""
    z = load(golden_dir, "acting_carflag.npz")
    sd = sd_from(z, "policy/")
    d, layers, ctx, heads = [int(v) for v in z["meta"]]
    T = len(z["action"])
    for t in range(0, T, 7):
        n = int(z["ctx_len"][t])
        win = z["ctx_obs"][t][:n]
        assert np.array_equal(win, np.trunc(win))     # the reference's acting context holds truncated observations
        with torch.no_grad():
            q = onet.forward(sd, torch.from_numpy(win).float()[None], heads)[0, -1].numpy()
        assert np.abs(q - z["greedy_q"][t]).max() < 5e-6


@pytest.mark.parametrize("name", ["identity", "gru", "aembed", "gtrxl_aembed"])
def test_oracle_ablation_variants_match_reference(golden_dir, name):
    """--identity (transformer.py:86-101), --gate gru (gates.py:5-31, shared across layers), --a-embed (dtqn.py:184-192):
    the oracle's restatement against the reference module's Q values and autograd gradients."""
    import torch
    from oracle import network as onet
    z = np.load(os.path.join(golden_dir, "forward_ablations.npz"))
    d, layers, ctx, H = [int(v) for v in z["meta"]]
    pre = f"{name}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith(pre)}
    identity = name in ("identity", "gtrxl_aembed")
    has_a = "action_embedding.embedding.0.weight" in sd
    assert ("transformer_layers.0.attn_gate.w_r.weight" in sd) == (name in ("gru", "gtrxl_aembed"))
    # shared gate modules: one leaf per distinct tensor, so autograd accumulates both layers' contributions like the reference
    leaves = {}
    for k in list(sd):
        if k.endswith("attn_mask"):
            continue
        base = k.replace("transformer_layers.1.attn_gate", "transformer_layers.0.attn_gate").replace(
            "transformer_layers.1.mlp_gate", "transformer_layers.0.mlp_gate")
        if base != k:
            assert torch.equal(sd[k], sd[base])
            sd[k] = leaves.setdefault(base, sd[base].requires_grad_(True))
        else:
            sd[k] = leaves.setdefault(k, sd[k].requires_grad_(True))
    for L in (1, 9):
        x = torch.from_numpy(z[f"{name}/L{L}/obss"])
        a = torch.from_numpy(z[f"{name}/L{L}/actions"])
        q = onet.forward(sd, x, H, actions=a if has_a else None, identity=identity)
        ref = z[f"{name}/L{L}/q"]
        assert np.abs(q.detach().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (name, L)
        if L == 9:
            (q ** 2).sum().backward()
            n_checked = 0
            for k in z.files:
                if not k.startswith(f"{name}/grad/"):
                    continue
                key = k[len(f"{name}/grad/"):]
                g, gr = leaves[key].grad.numpy(), z[k]
                assert np.abs(g - gr).max() <= 1e-4 * max(np.abs(gr).max(), 1e-6), (name, key)
                n_checked += 1
            assert n_checked >= 30


@pytest.mark.parametrize("fname", ["forward_carflag.npz", "forward_memory.npz"])
def test_module_form_matches_reference(golden_dir, fname):
    """oracle.network.ModuleNet (the torch.nn blocks the reference instantiates; the engine of the timed CPU baseline loop)
    loads the reference's state_dict unchanged and reproduces its recorded Q values."""
    z = load(golden_dir, fname)
    net = onet.ModuleNet(sd_from(z, "policy/"), 8)
    for L in (1, 7, int(z["meta"][2])):
        with torch.no_grad():
            q = net(torch.from_numpy(z[f"L{L}/obss"])).numpy()
        assert np.abs(q - z[f"L{L}/q"]).max() <= 2e-6 * max(1.0, np.abs(z[f"L{L}/q"]).max())


def test_baseline_loop_train_step_matches_reference(golden_dir):
    """The CPU baseline loop's train() (module engine: torch Adam / clip_grad_norm_ / mse_loss) on the reference's recorded
    batches reproduces its logged statistics and post-Adam parameters."""
    from oracle.loop import ReferenceLoop
    z = load(golden_dir, "train_carflag.npz")
    d, layers, ctx, B, heads, n_steps = [int(v) for v in z["meta"]]
    lp = ReferenceLoop("DiscreteCarFlag-v0", seed=1, inner_embed=d, heads=heads, layers=layers, context=ctx, batch=B)
    lp.policy_net.load_state_dict(sd_from(z, "policy0/")); lp.target_net.load_state_dict(sd_from(z, "target0/"))
    for s in range(n_steps):
        batch = (torch.from_numpy(z[f"step{s}/obss"]).float(), torch.from_numpy(z[f"step{s}/actions"].astype(np.int64)),
                 torch.from_numpy(z[f"step{s}/rewards"]), torch.from_numpy(z[f"step{s}/next_obss"]).float(),
                 torch.from_numpy(z[f"step{s}/next_actions"].astype(np.int64)), torch.from_numpy(z[f"step{s}/dones"]))
        st = lp._train_module(batch)
        lp.grad_steps += 1
        assert abs(st["loss"] - z["stats/td_errors"][s]) <= 1e-5 * max(1.0, abs(z["stats/td_errors"][s]))
        assert abs(st["grad_norm"] - z["stats/grad_norms"][s]) <= 1e-4 * max(1.0, abs(z["stats/grad_norms"][s]))
    for k, p in lp.policy_net.state_dict().items():
        if k.endswith("attn_mask"):                      # -inf entries
            assert np.array_equal(p.numpy(), z[f"policy{n_steps}/" + k])
        else:
            assert np.abs(p.numpy() - z[f"policy{n_steps}/" + k]).max() < 1e-5, k
