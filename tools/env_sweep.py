"""Env-step kernel throughput vs number of lockstep envs (report achieved GB/s and the N at which it
saturates).  Random-policy steps (agent-side PCG64 stream), fused replay + context append, CUDA events, graph replay."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dtqn_b200 import _lib
from dtqn_b200.envs import BatchedEnv, ContextWindow
from dtqn_b200.buffers import ReplayBuffer

out = []
for env_id, bytes_per_step in (("DiscreteCarFlag-v0", 55.0), ("Memory-5-v0", 140.0)):
    for n in (4096, 32768, 262144, 1048576):
        env = BatchedEnv(env_id, n, seeds=list(range(1, 4097)) * (n // 4096), device="cuda")   # seeds repeat: throughput only
        rb = ReplayBuffer(2 * n * env.max_episode_steps, env.obs_dim, env.obs_mask, env.max_episode_steps, 50, n_envs=n, device="cuda")
        cx = ContextWindow(50, env.obs_mask, env.num_actions, env.obs_dim, n_envs=n, device="cuda")
        env.attach(rb, cx)
        env.reset_all()
        for _ in range(5):
            env.step(mode=_lib.ACT_RANDOM)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            env.step(mode=_lib.ACT_RANDOM)
        for _ in range(5):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 100
        torch.cuda.synchronize(); a.record()
        for _ in range(steps):
            g.replay()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        rec = dict(env=env_id, n_envs=n, us_per_lockstep_step=1e3 * ms, env_steps_per_sec=n / (ms * 1e-3),
                   algorithmic_GBps=bytes_per_step * n / (ms * 1e-3) / 1e9, dropped=int(rb.counters[3].item()))
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del env, rb, cx, g
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/env_sweep.json", "w"), indent=1)
