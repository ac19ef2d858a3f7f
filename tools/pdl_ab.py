"""A/B of programmatic dependent launch (DTQN_B200_PDL, csrc/common.cuh launch_k) on the bench workload: with the knob off
and on, time the training half and the whole iteration (CUDA graphs, device events) and check that the parameters after the
same number of iterations are BITWISE identical -- the launch attribute may only move when CTAs are scheduled, never what
they compute.    python tools/pdl_ab.py [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dtqn_b200 import _lib
from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600


def run(pdl):
    _lib.lib.dtqn_set_pdl(C.c_int32(pdl))
    tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=4096, seed=1, device="cuda:0", batch=32)
    tr.prepopulate(260)
    g_train = tr.capture(tr.train_only)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        g_train.replay()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        g_train.replay()
    e1.record(); torch.cuda.synchronize()
    t_train = e0.elapsed_time(e1) / iters
    tr.enable_graphs()
    for _ in range(20):
        tr.train_iteration()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        tr.train_iteration()
    e1.record(); torch.cuda.synchronize()
    t_iter = e0.elapsed_time(e1) / iters
    return t_train, t_iter, tr.agent.policy_network.flat_params().clone() if hasattr(tr.agent.policy_network, "flat_params") \
        else torch.cat([p.detach().flatten() for p in tr.agent.policy_network.parameters()]).clone()


res = {}
for pdl in (0, 1, 0, 1):
    t_train, t_iter, params = run(pdl)
    print(f"pdl={pdl} train_half_ms={t_train:.4f} iteration_ms={t_iter:.4f}", flush=True)
    res.setdefault(pdl, []).append(params)
same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][0], res[0][1])
print("params bitwise identical across pdl off/on and repeats:", same)
assert same
