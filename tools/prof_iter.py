"""Profiling target for ncu: a few EAGER iterations of the whole loop (acting forward, env step + roll, sample, gather,
3 forwards, TD loss, backward, clip + Adam) of the bench workload inside a cudaProfilerStart/Stop range.

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof python tools/prof_iter.py [iters] [env] [n_envs]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
env_id = sys.argv[2] if len(sys.argv) > 2 else "DiscreteCarFlag-v0"
n_envs = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
tr = BatchedTrainer(env_id, n_envs=n_envs, seed=1, device="cuda:0", batch=32, inner_embed=64 if "CarFlag" in env_id else 128)
tr.prepopulate(260 if "CarFlag" in env_id else 80)
while not tr.agent.replay_buffer.can_sample(32):
    tr.prepopulate(16)
for _ in range(3):
    tr.train_iteration()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(iters):
    tr.train_iteration()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
