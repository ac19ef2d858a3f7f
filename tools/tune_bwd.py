"""Launch-shape sweep of the backward pass on the bench workload (CarFlag, 4096 envs, batch 32 x 50 tokens):
dgrad rows per CTA x LayerNorm-backward fusion x head-backward tokens per CTA (dtqn_set_dgrad_rows / dtqn_set_fuse_ln_bwd /
dtqn_set_head_bwd_tokens), optionally with programmatic dependent launch.  For each setting: parameters after 3 training steps
against the baseline setting (bitwise for row-tile changes, relative difference otherwise), then the training half and the
whole iteration as CUDA graphs, device-timed.    python tools/tune_bwd.py [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dtqn_b200 import _lib
from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L = _lib.lib


def run(rows, fuse, head, pdl, whole):
    assert L.dtqn_set_dgrad_rows(C.c_int32(rows)) == 0 and L.dtqn_set_head_bwd_tokens(C.c_int32(head)) == 0
    L.dtqn_set_fuse_ln_bwd(C.c_int32(fuse)); L.dtqn_set_pdl(C.c_int32(pdl))
    tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=4096, seed=1, device="cuda:0", batch=32)
    tr.prepopulate(260)
    for _ in range(3):
        tr.train_only()
    torch.cuda.synchronize()
    params = tr.agent.policy_network.flat.clone()
    g = tr.capture(tr.train_only)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    t_train = e0.elapsed_time(e1) / iters
    t_iter = float("nan")
    if whole:
        tr.enable_graphs()
        for _ in range(20):
            tr.train_iteration()
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters):
            tr.train_iteration()
        e1.record(); torch.cuda.synchronize()
        t_iter = e0.elapsed_time(e1) / iters
    return params, t_train, t_iter


base = None
best = None
configs = [(64, 0, 64, 0, 1), (32, 0, 64, 0, 0), (16, 0, 64, 0, 0), (64, 1, 64, 0, 0), (32, 1, 64, 0, 1), (16, 1, 64, 0, 0),
           (32, 1, 32, 0, 0), (32, 1, 16, 0, 0), (32, 1, 32, 1, 1), (16, 1, 32, 1, 1), (64, 0, 64, 0, 1)]
for rows, fuse, head, pdl, whole in configs:
    params, t_train, t_iter = run(rows, fuse, head, pdl, whole)
    if base is None:
        base = params
    rel = ((params - base).abs().max() / base.abs().max()).item()
    print(f"dgrad_rows={rows} fuse_ln_bwd={fuse} head_tok={head} pdl={pdl}: train_half_ms={t_train:.4f} iteration_ms={t_iter:.4f} "
          f"params_vs_base bitwise={torch.equal(params, base)} max_rel={rel:.2e}", flush=True)
    if pdl == 0 and rel < 1e-4 and (best is None or t_train < best[0]):
        best = (t_train, rows, fuse, head)
print("best (pdl off):", best)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/best_bwd.env", "w") as f:
    f.write(f"export DTQN_B200_DGRAD_ROWS={best[1]} DTQN_B200_FUSE_LN_BWD={best[2]} DTQN_B200_HEAD_BWD_TOKENS={best[3]}\n")
