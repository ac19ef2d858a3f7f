"""Checkpoint / resume in the reference's on-disk format (dtqn/agents/dqn.py:212-327, run.py:469-499).

Files written for a path prefix ``P`` (the reference's ``policy_path``):

    P_mini_checkpoint.pt   {"step", "wandb_id"}                                   dqn.py:212-216
    P_checkpoint.pt        the reference's dict (same keys: step, wandb_id, replay_buffer_pos, policy/target
                           state_dicts with the reference's key names, a torch.optim.Adam state_dict, epsilon,
                           running averages, host RNG states) + one extra key "b200" holding what the reference
                           omits and the batched device loop needs for a bit-exact resume: env PCG64 streams and
                           physical state, acting contexts, replay slot bookkeeping, sampler draw counter, device
                           Adam step, statistics ring                                dqn.py:232-271
    Pbuffer_{obss,actions,rewards,dones,eplens}.sav   joblib dumps of the replay arrays with the reference's
                           shapes and dtypes ([S, max_ep+1, O] f32, u8, f32, bool, u8)  dqn.py:272-279

The pure functions (``adam_state_dict`` / ``load_adam_state_dict`` / ``running_average_*``) have no CUDA dependency
and are covered by the CPU tests; ``save_agent`` / ``load_agent`` / ``save_trainer`` / ``load_trainer`` move device
tensors through host memory (plumbing, never on the hot path).
"""
import pickle
import random
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

ADAM_DEFAULTS = dict(weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                     differentiable=False, fused=None)


class RunningAverage:
    """utils/logging_utils.py:10-24 (window mean of the last ``size`` samples) for the host-side result averages
    (success rate / return / episode length of the last 10 evaluations, run.py:491-493)."""

    def __init__(self, size: int, values: Iterable[float] = ()):
        self.size = int(size)
        self.values = [float(v) for v in values][-self.size:]

    def add(self, val: float) -> None:
        self.values.append(float(val))
        if len(self.values) > self.size:
            del self.values[0]

    def mean(self) -> float:
        return sum(self.values) / max(len(self.values), 1)

    def state(self) -> dict:
        return {"size": self.size, "values": list(self.values)}

    @staticmethod
    def from_state(st) -> "RunningAverage":
        if isinstance(st, RunningAverage):
            return st
        if isinstance(st, dict):
            return RunningAverage(st["size"], st["values"])
        return RunningAverage(getattr(st, "size", 10), list(getattr(st, "q", [])))     # a reference object, if unpickled


# ---- reading checkpoint pickles ------------------------------------------------------------------------------------------
class _RefRunningAverage:
    """Stand-in for the reference's pickled ``utils.logging_utils.RunningAverage`` objects (dqn.py:245-258: size, q, sum):
    the reference package is not importable where this library runs, and nothing of it needs to be."""
    size, q, sum = 10, (), 0


class _CheckpointPickle:
    """``pickle_module`` for torch.load: resolves only the globals a DTQN checkpoint legitimately holds (torch tensors /
    storages, numpy arrays and RNG state tuples, OrderedDict / deque, this module's classes) and maps the reference's
    RunningAverage to a plain holder; any other global raises instead of being imported and executed."""
    __name__ = "dtqn_b200.checkpoint._CheckpointPickle"
    _ALLOWED_MODULES = ("torch", "numpy", "collections", "_codecs", "copyreg")
    _ALLOWED_BUILTINS = {"set", "frozenset", "list", "dict", "tuple", "complex", "bytearray", "slice", "range", "int",
                         "float", "bool", "str", "bytes", "getattr"}

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if (module, name) == ("utils.logging_utils", "RunningAverage"):
                return _RefRunningAverage
            if module == "dtqn_b200.checkpoint" and name in ("RunningAverage", "_RefRunningAverage"):
                return globals()[name]
            if module == "builtins" and name in _CheckpointPickle._ALLOWED_BUILTINS and name != "getattr":
                return super().find_class(module, name)
            if module.split(".")[0] in _CheckpointPickle._ALLOWED_MODULES:
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}, which a DTQN checkpoint never holds")

    @staticmethod
    def load(f, **kw):
        return _CheckpointPickle.Unpickler(f, **kw).load()


def read_checkpoint_file(path: str) -> dict:
    return torch.load(path, weights_only=False, map_location="cpu", pickle_module=_CheckpointPickle)


# ---- torch.optim.Adam state_dict <-> flat moment buffers ------------------------------------------------------------------
def adam_state_dict(param_names, trainable, moments: Dict[str, Tuple[torch.Tensor, torch.Tensor]], step: int, lr: float,
                    betas=(0.9, 0.999), eps: float = 1e-8) -> dict:
    """The ``optimizer.state_dict()`` torch.optim.Adam would hold over ``policy_network.parameters()`` (dqn.py:64):
    parameter i = i-th entry of ``param_names`` (registration order, non-trainable ``attn_mask`` included in the group
    but without state, exactly as Adam leaves parameters that never received a gradient)."""
    state = {}
    for i, name in enumerate(param_names):
        if name in trainable and step > 0:
            m, v = moments[name]
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().cpu().clone(),
                        "exp_avg_sq": v.detach().cpu().clone()}
    group = dict(lr=lr, betas=tuple(betas), eps=eps, params=list(range(len(param_names))), **ADAM_DEFAULTS)
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(sd: dict, param_names, moments: Dict[str, Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[int, dict]:
    """Inverse of ``adam_state_dict``: copies exp_avg / exp_avg_sq into the (flat-buffer) views and returns
    (step, hyper-parameters).  Parameters without state get zero moments."""
    steps = set()
    for i, name in enumerate(param_names):
        if name not in moments:
            continue
        m, v = moments[name]
        st = sd["state"].get(i)
        if st is None:
            m.zero_(); v.zero_()
            continue
        m.copy_(st["exp_avg"].to(m.device).view_as(m))
        v.copy_(st["exp_avg_sq"].to(v.device).view_as(v))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"optimizer state holds different step counts {sorted(steps)}; the fused Adam keeps one")
    g = sd["param_groups"][0]
    return (steps.pop() if steps else 0), dict(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"])


# ---- agent ---------------------------------------------------------------------------------------------------------------------
def _param_names(net):
    return [n for n, _ in net.named_parameters()]


def _moments(agent):
    net = agent.policy_network
    m, v = net.unflatten(agent.exp_avg), net.unflatten(agent.exp_avg_sq)
    return {k: (m[k], v[k]) for k in m}


def _stat_state(agent) -> dict:
    n = min(agent.num_train_steps, agent.stats_ring.shape[0])
    return {"ring": agent.stats_ring.cpu(), "count": n}


def save_mini_checkpoint(agent, checkpoint_dir: str, wandb_id: Optional[str]) -> None:         # dqn.py:212-216
    torch.save({"step": agent.num_train_steps, "wandb_id": wandb_id}, checkpoint_dir + "_mini_checkpoint.pt")


def load_mini_checkpoint(checkpoint_dir: str) -> dict:                                         # dqn.py:218-220
    return read_checkpoint_file(checkpoint_dir + "_mini_checkpoint.pt")


def replay_arrays(rb) -> dict:
    """Host copies of the replay arrays in the reference's dtypes (replay_buffer.py:46-69)."""
    lens = rb.episode_lengths.cpu().numpy()
    return {"obss": rb.obss.cpu().numpy(), "actions": rb.actions.cpu().numpy(), "rewards": rb.rewards.cpu().numpy(),
            "dones": rb.dones.cpu().numpy().astype(np.bool_),
            "eplens": lens.astype(np.uint8) if rb.max_episode_steps <= 255 else lens}


def save_agent(agent, checkpoint_dir: str, wandb_id: Optional[str], episode_successes, episode_rewards, episode_lengths,
               eps, extra: Optional[dict] = None) -> None:
    """DqnAgent.save_checkpoint (dqn.py:222-279)."""
    import joblib
    from dtqn_b200.agents import RNG, STAT_NAMES
    torch.cuda.synchronize(agent.device)
    agent.check_finite()
    save_mini_checkpoint(agent, checkpoint_dir, wandb_id)
    rb, net = agent.replay_buffer, agent.policy_network
    names = _param_names(net)
    trainable = {n for n, p in net.named_parameters() if p.requires_grad}
    cpu_sd = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    ring = _stat_state(agent)
    ck = {
        "step": agent.num_train_steps,
        "wandb_id": wandb_id,
        # episodes started so far: slot pos[0] % S is the next one to be opened / cleansed (replay_buffer.py:79,90-91,97-98)
        "replay_buffer_pos": [rb.pos[0] if rb.pos[0] > 0 else int(rb.counters[0].item()), 0],
        "policy_net_state_dict": cpu_sd(net),
        "target_net_state_dict": cpu_sd(agent.target_network),
        "optimizer_state_dict": adam_state_dict(names, trainable, _moments(agent), int(agent.opt_step.item()),
                                                agent.learning_rate, agent.betas, agent.adam_eps),
        "epsilon": eps.val,
        "episode_successes": RunningAverage.from_state(episode_successes).state(),
        "episode_rewards": RunningAverage.from_state(episode_rewards).state(),
        "episode_lengths": RunningAverage.from_state(episode_lengths).state(),
        "random_rng_state": random.getstate(),
        "rng_bit_generator_state": RNG.rng.bit_generator.state if RNG.rng is not None else None,
        "numpy_rng_state": np.random.get_state(),
        "torch_rng_state": torch.get_rng_state(),
        "torch_cuda_rng_state": torch.cuda.get_rng_state(agent.device),
    }
    for col, name in enumerate(STAT_NAMES):           # td_errors ... target_min: the last <=100 samples, oldest first
        n, k = ring["count"], agent.num_train_steps
        rows = [(k - n + j) % ring["ring"].shape[0] for j in range(n)]
        ck[name] = {"size": ring["ring"].shape[0], "values": [float(ring["ring"][r, col]) for r in rows]}
    b200 = {
        "format": 1,
        "stats_ring": ring["ring"], "opt_step": int(agent.opt_step.item()),
        "replay": {k: getattr(rb, k).cpu() for k in ("slot_open", "counters", "env_slot", "env_prev_len", "draw_counter")},
        "replay_host_pos": list(rb._host_pos), "sample_seed": rb.sample_seed,
        "contexts": {k: {"obs": c.obs.cpu(), "timestep": c.timestep_t.cpu()}
                     for k, c in (("train", agent.train_context), ("eval", agent.eval_context))},
    }
    b200.update(extra or {})
    ck["b200"] = b200
    torch.save(ck, checkpoint_dir + "_checkpoint.pt")
    arrs = replay_arrays(rb)
    for key in ("obss", "actions", "rewards", "dones", "eplens"):
        joblib.dump(arrs[key], checkpoint_dir + f"buffer_{key}.sav")


def load_agent(agent, checkpoint_dir: str, validate=None):
    """DqnAgent.load_checkpoint (dqn.py:281-327).  Returns (wandb_id, successes, rewards, lengths, epsilon, b200-extra).
    Everything is read and checked against this agent (buffer shapes, parameter shapes, ``validate(ck)`` of the caller)
    BEFORE any of its state is overwritten, so a mismatched checkpoint raises and leaves the agent untouched."""
    import joblib
    from dtqn_b200.agents import RNG
    ck = read_checkpoint_file(checkpoint_dir + "_checkpoint.pt")
    rb, net, dev = agent.replay_buffer, agent.policy_network, agent.device
    if validate is not None:
        validate(ck)
    arrays = []
    for dst, key, dt in ((rb.obss, "obss", torch.float32), (rb.actions, "actions", torch.uint8),
                         (rb.rewards, "rewards", torch.float32), (rb.dones, "dones", torch.uint8),
                         (rb.episode_lengths, "eplens", torch.int32)):
        arr = np.asarray(joblib.load(checkpoint_dir + f"buffer_{key}.sav"))
        if tuple(arr.shape) != tuple(dst.shape):
            raise ValueError(f"buffer_{key}.sav has shape {arr.shape}, this replay buffer expects {tuple(dst.shape)}")
        arrays.append((dst, arr, dt))
    want = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    for which in ("policy_net_state_dict", "target_net_state_dict"):
        got = {k: tuple(v.shape) for k, v in ck[which].items()}
        if got != want:
            diff = sorted(set(got.items()) ^ set(want.items()))[:4]
            raise ValueError(f"{which} does not match this network (first differences: {diff})")
    agent.num_train_steps = int(ck["step"])
    for dst, arr, dt in arrays:
        dst.copy_(torch.from_numpy(arr.astype(np.uint8) if arr.dtype == np.bool_ else arr).to(dt))
    net.load_state_dict(ck["policy_net_state_dict"])
    agent.target_network.load_state_dict(ck["target_net_state_dict"])
    step, hyper = load_adam_state_dict(ck["optimizer_state_dict"], _param_names(net), _moments(agent))
    agent.opt_step.fill_(step)
    agent.learning_rate, agent.betas, agent.adam_eps = hyper["lr"], hyper["betas"], hyper["eps"]
    random.setstate(ck["random_rng_state"])
    if ck.get("rng_bit_generator_state") is not None:
        if RNG.rng is None:
            RNG.rng = np.random.Generator(np.random.PCG64(seed=0))
        RNG.rng.bit_generator.state = ck["rng_bit_generator_state"]
    np.random.set_state(ck["numpy_rng_state"])
    torch.set_rng_state(ck["torch_rng_state"])
    cuda_state = ck["torch_cuda_rng_state"]
    if cuda_state.numel() == torch.cuda.get_rng_state(dev).numel():
        torch.cuda.set_rng_state(cuda_state, dev)
    # else: written on a CUDA-less host, where the reference stores the CPU generator state under this key (dqn.py:266-268);
    # nothing on this path draws from torch's CUDA generator, so there is nothing to restore
    b = ck.get("b200")
    if b is not None:
        agent.stats_ring.copy_(b["stats_ring"])
        agent.opt_step.fill_(int(b["opt_step"]))
        for k, v in b["replay"].items():
            getattr(rb, k).copy_(v)
        rb._host_pos = list(b["replay_host_pos"])
        rb.sample_seed = int(b["sample_seed"])
        for k, c in (("train", agent.train_context), ("eval", agent.eval_context)):
            c.obs.copy_(b["contexts"][k]["obs"]); c.timestep_t.copy_(b["contexts"][k]["timestep"])
    else:
        # a checkpoint written by the reference: single env, host-call API bookkeeping only (dqn.py:287)
        n = int(ck["replay_buffer_pos"][0])
        rb._host_pos = [n, 0]
        rb.slot_open.zero_()
        rb.slot_open[n % rb.max_size] = 1
        rb.counters.copy_(torch.tensor([n + 1, n + 1, n, 0]))
        ring = agent.stats_ring
        from dtqn_b200.agents import STAT_NAMES
        for col, name in enumerate(STAT_NAMES):
            vals = list(RunningAverage.from_state(ck[name]).values)[-ring.shape[0]:]
            k = agent.num_train_steps
            for j, val in enumerate(vals):
                ring[(k - len(vals) + j) % ring.shape[0], col] = val
    rb._completed_seen = 0
    agent.flags.zero_()
    torch.cuda.synchronize(dev)
    ra = RunningAverage.from_state
    return (ck["wandb_id"], ra(ck["episode_successes"]), ra(ck["episode_rewards"]), ra(ck["episode_lengths"]),
            ck["epsilon"], b)


# ---- batched trainer (env streams + loop counters on top of the agent) -------------------------------------------------------------
ENV_STATE = ("rng", "rng_buf", "arng", "arng_buf", "pos", "vel", "heaven", "cards", "shown", "cur", "elapsed",
             "done_flag", "ep_stats", "ep_return")


def save_trainer(tr, checkpoint_dir: str, wandb_id: Optional[str] = None, episode_successes=None, episode_rewards=None,
                 episode_lengths=None) -> None:
    ra = lambda x: x if x is not None else RunningAverage(10)
    extra = {"env": {k: getattr(tr.env, k).cpu() for k in ENV_STATE}, "env_id": tr.env.env_id, "n_envs": tr.n_envs,
             "iterations": tr.iterations, "world": tr.world, "rank": tr.rank}
    if tr.eval_env is not None:                # evaluation envs keep their streams across evaluations (run.py:413-417)
        extra["eval_env"] = {k: getattr(tr.eval_env, k).cpu() for k in ENV_STATE}
    save_agent(tr.agent, checkpoint_dir, wandb_id, ra(episode_successes), ra(episode_rewards), ra(episode_lengths),
               tr.eps, extra)


def load_trainer(tr, checkpoint_dir: str):
    """Restores agent + env streams + loop counters.  A captured CUDA graph is dropped (its device scalars were
    overwritten); call ``enable_graphs()`` again -- it performs the next loop iteration while it re-captures."""
    def validate(ck):
        b = ck.get("b200")
        if b is not None and "env" in b and (b["env_id"] != tr.env.env_id or int(b["n_envs"]) != tr.n_envs):
            raise ValueError(f"checkpoint holds {b['n_envs']} x {b['env_id']}, trainer runs {tr.n_envs} x {tr.env.env_id}")

    wandb_id, succ, rew, length, epsilon, b = load_agent(tr.agent, checkpoint_dir, validate)
    tr.disable_graphs()
    if b is not None and "env" in b:
        for k, v in b["env"].items():
            getattr(tr.env, k).copy_(v)
        tr.iterations = int(b["iterations"])
        if "eval_env" in b:
            ev = tr.make_eval_env()
            for k, v in b["eval_env"].items():
                getattr(ev, k).copy_(v)
    tr.eps.val = epsilon                                                           # run.py:489
    torch.cuda.synchronize(tr.device)
    return wandb_id, succ, rew, length
