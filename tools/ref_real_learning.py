"""CPU control for the learning-curve claim: the UNMODIFIED reference (through oracle/ref_harness) trained on
DiscreteCarFlag-v0 with its default hyper-parameters (run.py:16-184: batch 32, ctx 50, lr 3e-4, tuf 10 000, buffer 500 000,
50 000 prepopulate steps, eps 1.0 -> 0.1 over num_steps/10) and evaluated with its own ``run.evaluate`` (greedy, 10
episodes, run.py:187-243) every ``every`` steps.  One JSON line per evaluation.

usage: python tools/ref_real_learning.py [total_steps] [eval_every] [threads] [num_steps_for_eps_schedule] [out.jsonl]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.ref_loop import RealReferenceLoop

total = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sched = int(sys.argv[4]) if len(sys.argv) > 4 else 2_000_000
out = sys.argv[5] if len(sys.argv) > 5 else None
torch.set_num_threads(threads)

lp = RealReferenceLoop("DiscreteCarFlag-v0", seed=1, num_steps=sched)
import gym  # noqa: E402  (stub, activated by the harness)
eval_env = gym.make("DiscreteCarFlag-v0")
eval_env.seed(1)                                    # utils/random.py:26-29: eval envs share the seed
lp.prepopulate(50_000)
f = open(out, "a") if out else None
t0 = time.time()
for it in range(1, total + 1):
    lp.iteration()
    if it % every == 0:
        sr, ret, length = lp._run.evaluate(lp.agent, eval_env, 10)
        a = lp.agent
        line = json.dumps({"impl": "reference-cpu", "step": it, "success": sr, "return": ret, "length": length,
                           "eps": round(lp.eps.val, 4), "td_error": a.td_errors.mean(), "q_mean": a.qvalue_mean.mean(),
                           "target_mean": a.target_mean.mean(), "wall_s": round(time.time() - t0, 1)})
        print(line, flush=True)
        if f:
            f.write(line + "\n"); f.flush()
