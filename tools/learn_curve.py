"""Learning-curve run of the device loop with the reference's hyper-parameters (run.py:16-184: batch 32, ctx 50, lr 3e-4,
tuf 10 000, gamma 0.99, eps geometric 1.0 -> 0.1 over num_steps/10, 50 000 prepopulate transitions) and its evaluation
(run.evaluate: greedy, fixed number of episodes).  With --n-envs 1 the device loop reproduces the reference's single-env
stream (same env / agent draws, tests/test_env_gpu.py), so this is the reference experiment itself on the GPU.
One JSON line per evaluation, appended to --out as it is produced.

    python tools/learn_curve.py --n-envs 1 --iters 2000000 --eval-every 50000 --eval-episodes 10 --out gpurun_out/learn_n1.jsonl
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dtqn_b200.runner import BatchedTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="DiscreteCarFlag-v0")
ap.add_argument("--n-envs", type=int, default=1)
ap.add_argument("--iters", type=int, default=2_000_000)
ap.add_argument("--num-steps", type=int, default=None, help="length of the eps schedule (default: --iters)")
ap.add_argument("--eval-every", type=int, default=50_000)
ap.add_argument("--eval-episodes", type=int, default=10, help="greedy episodes per evaluation (total over the eval envs)")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--tuf", type=int, default=10_000)
ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--in-embed", type=int, default=64)
ap.add_argument("--buf-size", type=int, default=500_000)
ap.add_argument("--prepopulate", type=int, default=50_000, help="random-policy transitions before training (run.py:495)")
ap.add_argument("--trunc", type=int, default=1, help="1 = the reference's integer acting context (SURVEY A-Q2)")
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--record-every", type=int, default=1, help="each env stores every K-th episode (replay horizon x K)")
ap.add_argument("--updates-per-step", type=int, default=1, help="gradient updates per lockstep env step")
ap.add_argument("--budget-s", type=float, default=1e9, help="stop (cleanly) after this many seconds of wall-clock")
ap.add_argument("--out", default="gpurun_out/learn_curve.jsonl")
a = ap.parse_args()

E = 200 if "CarFlag" in a.env else 50
tr = BatchedTrainer(a.env, n_envs=a.n_envs, seed=a.seed, buf_size=max(a.buf_size, 8 * a.n_envs * E), device="cuda",
                    inner_embed=a.in_embed, context=50, batch=a.batch, lr=a.lr, tuf=a.tuf, num_steps=a.num_steps or a.iters,
                    trunc_context_obs=bool(a.trunc), record_every=a.record_every, updates_per_step=a.updates_per_step)
tr.prepopulate(max(1, a.prepopulate // a.n_envs))
while not tr.agent.replay_buffer.can_sample(a.batch):
    tr.prepopulate(64)
tr.enable_graphs()
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
k = max(1, -(-a.eval_episodes // a.n_envs))
cfg = {k_: getattr(a, k_) for k_ in ("env", "n_envs", "iters", "batch", "tuf", "lr", "in_embed", "buf_size", "trunc", "seed", "record_every", "updates_per_step")}


def log(it, wall):
    sr, ret, length = tr.evaluate(k)
    ag = tr.agent
    rec = dict(cfg, iteration=it, env_steps=it * a.n_envs, success_rate=sr, mean_return=ret, episode_length=length,
               eval_episodes=k * a.n_envs, td_error=ag.td_errors.mean(), q_mean=ag.qvalue_mean.mean(),
               target_mean=ag.target_mean.mean(), grad_norm=ag.grad_norms.mean(), epsilon=tr.eps.val, train_wall_s=round(wall, 1))
    with open(a.out, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


t0 = time.time()
wall = 0.0
log(0, 0.0)
start = time.time()
for it in range(1, a.iters + 1):
    tr.train_iteration()
    if it % a.eval_every == 0:
        torch.cuda.synchronize()
        wall += time.time() - t0
        log(it, wall)
        t0 = time.time()
        if time.time() - start > a.budget_s:
            break
