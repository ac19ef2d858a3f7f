"""Checkpoint format helpers (dtqn_b200/checkpoint.py) against torch.optim.Adam itself -- no GPU needed.
Reference: dtqn/agents/dqn.py:232-271 stores ``self.optimizer.state_dict()`` of ``optim.Adam(policy.parameters())``."""
import copy

import torch

from dtqn_b200 import checkpoint as ck


def _toy():
    torch.manual_seed(3)
    names = ["emb.weight", "emb.bias", "layer.attn_mask", "head.weight"]
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(4)),
              torch.nn.Parameter(torch.zeros(2, 2), requires_grad=False), torch.nn.Parameter(torch.randn(2, 4))]
    return names, params


def _step(opt, params, seed):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g)
    opt.step()


def test_adam_state_dict_loads_into_torch_adam_and_continues_identically():
    names, params = _toy()
    opt = torch.optim.Adam(params, lr=3e-4)
    for s in range(2):
        _step(opt, params, s)
    ref_sd = opt.state_dict()
    moments = {n: (ref_sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg_sq"])
               for i, n in enumerate(names) if i in ref_sd["state"]}
    trainable = {n for n, p in zip(names, params) if p.requires_grad}
    ours = ck.adam_state_dict(names, trainable, moments, step=2, lr=3e-4)
    assert set(ours["state"]) == set(ref_sd["state"]) == {0, 1, 3}          # attn_mask: in the group, no state
    assert ours["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert ours["param_groups"][0][k] == ref_sd["param_groups"][0][k]
    params2 = [torch.nn.Parameter(p.detach().clone(), requires_grad=p.requires_grad) for p in params]
    opt2 = torch.optim.Adam(params2, lr=1.0)                                   # lr comes from the loaded state
    opt2.load_state_dict(copy.deepcopy(ours))
    _step(opt, params, 7)
    _step(opt2, params2, 7)
    for a, b in zip(params, params2):
        assert torch.equal(a, b)


def test_load_adam_state_dict_round_trip():
    names, params = _toy()
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.8, 0.9))
    for s in range(3):
        _step(opt, params, s)
    sd = opt.state_dict()
    flat_m = {n: (torch.full(p.shape, 9.0), torch.full(p.shape, 9.0)) for n, p in zip(names, params) if n != "layer.attn_mask"}
    step, hyper = ck.load_adam_state_dict(sd, names, flat_m)
    assert step == 3 and hyper == dict(lr=1e-3, betas=(0.8, 0.9), eps=1e-8)
    for i, n in enumerate(names):
        if n in flat_m:
            assert torch.equal(flat_m[n][0], sd["state"][i]["exp_avg"]) and torch.equal(flat_m[n][1], sd["state"][i]["exp_avg_sq"])
    empty = {"state": {}, "param_groups": sd["param_groups"]}
    step, _ = ck.load_adam_state_dict(empty, names, flat_m)
    assert step == 0 and all(float(m.abs().sum()) == 0 and float(v.abs().sum()) == 0 for m, v in flat_m.values())


def test_running_average_window():
    ra = ck.RunningAverage(3)
    assert ra.mean() == 0
    for v in (1, 2, 3, 4):
        ra.add(v)
    assert ra.mean() == 3.0 and ra.state() == {"size": 3, "values": [2.0, 3.0, 4.0]}
    assert ck.RunningAverage.from_state(ra.state()).mean() == 3.0


def test_reference_written_checkpoint_unpickles_without_the_reference_package():
    """tests/golden/refckpt/ref_checkpoint.pt was written by the reference's DqnAgent.save_checkpoint (dqn.py:222-279) and
    pickles ``utils.logging_utils.RunningAverage`` objects; the loader must read it where that package does not exist."""
    import os
    import sys
    import pickle
    import numpy as np
    import pytest
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refckpt")
    c = ck.read_checkpoint_file(os.path.join(d, "ref_checkpoint.pt"))
    z = np.load(os.path.join(d, "expect.npz"))
    # the reference's RunningAverage objects come back as the loader's plain holder, whether or not the reference package is importable
    assert type(c["td_errors"]) is ck._RefRunningAverage and type(c["episode_successes"]) is ck._RefRunningAverage
    assert c["step"] == int(z["meta"][5]) and c["replay_buffer_pos"] == [int(z["replay_pos"][0]), 0]
    assert abs(c["epsilon"] - float(z["epsilon"][0])) == 0.0
    assert abs(ck.RunningAverage.from_state(c["td_errors"]).mean() - float(z["td_errors_mean"][0])) < 1e-9
    assert abs(ck.RunningAverage.from_state(c["episode_successes"]).mean() - float(z["succ_mean"][0])) < 1e-12
    assert set(c["optimizer_state_dict"]) == {"state", "param_groups"}
    assert ck.load_mini_checkpoint(os.path.join(d, "ref"))["step"] == c["step"]

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    import io
    buf = io.BytesIO()
    torch.save({"x": Evil()}, buf)
    buf.seek(0)
    with pytest.raises(pickle.UnpicklingError):
        torch.load(buf, weights_only=False, pickle_module=ck._CheckpointPickle)
