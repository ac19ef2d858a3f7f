"""End-to-end learning check: train DTQN on 4096 lockstep CarFlag envs with the reference hyper-parameters (batch 32,
lr 3e-4, tuf 10 000, gamma 0.99, ctx 50, eps geometric 1.0 -> 0.1 over num_steps/10) and log the greedy evaluation success
rate (run.evaluate semantics) over training.  Usage: python tools/learn_carflag.py [iterations] [eval_every]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
trunc = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True       # 1 = the reference's integer acting context (A-Q2)
tuf = int(sys.argv[5]) if len(sys.argv) > 5 else 10_000             # --tuf: one Bellman backup per target sync
n_envs = int(sys.argv[6]) if len(sys.argv) > 6 else 4096
slots = int(sys.argv[7]) if len(sys.argv) > 7 else 8 * n_envs       # replay ring size in episodes
tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=n_envs, seed=1, buf_size=slots * 200, device="cuda", inner_embed=64, context=50, batch=batch,
                    num_steps=iters, trunc_context_obs=trunc, tuf=tuf)
tr.prepopulate(260)
tr.enable_graphs()
log = []
sr, ret, length = tr.evaluate(1)
log.append(dict(iteration=0, env_steps=0, success_rate=sr, mean_return=ret, episode_length=length))
print(json.dumps(log[-1]), flush=True)
t0 = time.time()
for it in range(1, iters + 1):
    tr.train_iteration()
    if it % every == 0:
        torch.cuda.synchronize()
        wall = time.time() - t0
        sr, ret, length = tr.evaluate(1)
        a = tr.agent
        log.append(dict(iteration=it, env_steps=it * n_envs, success_rate=sr, mean_return=ret, episode_length=length,
                        td_error=a.td_errors.mean(), q_mean=a.qvalue_mean.mean(), grad_norm=a.grad_norms.mean(),
                        epsilon=tr.eps.val, train_wall_s=wall))
        print(json.dumps(log[-1]), flush=True)
        t0 = time.time() - wall          # exclude evaluation time from the training clock
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(batch=batch, trunc_context_obs=trunc, tuf=tuf, n_envs=n_envs, replay_slots=slots, log=log), open(f"gpurun_out/learn_carflag_b{batch}_t{int(trunc)}_tuf{tuf}_n{n_envs}.json", "w"), indent=1)
