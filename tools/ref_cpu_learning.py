"""CPU-only diagnostic: does the reference algorithm itself (oracle restatement of run.py's single-env loop) learn
CarFlag within a few 1e5 updates?  Prints one JSON line per evaluation (greedy, run.evaluate semantics).

usage: python tools/ref_cpu_learning.py [total_steps] [eval_every] [tuf] [threads] [env_id]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import envs as oenvs
from oracle import network as onet
from oracle.loop import ReferenceLoop
from oracle.pcg64 import PCG64
from oracle.replay import ContextOracle

total = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
tuf = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 4
env_id = sys.argv[5] if len(sys.argv) > 5 else "DiscreteCarFlag-v0"
torch.set_num_threads(threads)

loop = ReferenceLoop(env_id, seed=1, tuf=tuf, num_steps=total)
loop.prepopulate(50_000)
eval_env = oenvs.make(env_id, 1)
eval_ctx = ContextOracle(loop.ctx_len, eval_env.obs_mask, eval_env.num_actions, eval_env.obs_dim, loop.rng)


def evaluate(episodes=20):
    succ = ret = steps = 0
    for _ in range(episodes):
        eval_ctx.reset(eval_env.reset())
        done, ep_r = False, 0.0
        while not done:
            obs, _ = eval_ctx.window()
            x = torch.as_tensor(obs, dtype=torch.long if loop.discrete else torch.float32).unsqueeze(0)
            with torch.no_grad():
                q = loop.policy_net(x) if loop.engine == "module" else onet.forward(loop.trainer.policy, x, loop.heads)
                a = int(torch.argmax(q[:, -1, :]).item())
            o, r, done, info = eval_env.step(a)
            eval_ctx.add_transition(o, a)
            ep_r += r
        succ += int(info.get("is_success", False) or ep_r > 0)
        ret += ep_r
        steps += eval_ctx.timestep
    return succ / episodes, ret / episodes, steps / episodes


t0 = time.time()
losses = []
for it in range(1, total + 1):
    st = loop.iteration()
    if st is not None:
        losses.append(float(st["loss"]) if isinstance(st, dict) else float(st[0]))
    if it % every == 0:
        sr, rt, ln = evaluate()
        print(json.dumps({"step": it, "success": sr, "return": rt, "length": ln, "eps": round(loop.eps, 4),
                          "loss_mean": float(np.mean(losses[-every:])) if losses else None,
                          "wall_s": round(time.time() - t0, 1)}), flush=True)
