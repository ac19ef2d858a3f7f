"""Training-half time (sample + gather + 3 forwards + TD + backward + clip + Adam, one CUDA graph) vs the split-K chunk of
the weight-gradient GEMMs and with / without the side stream.  Usage: python tools/tune_wgrad.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dtqn_b200 import _lib
from dtqn_b200.runner import BatchedTrainer

tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=4096, seed=1, device="cuda:0", batch=32)
tr.prepopulate(260)
lib = _lib.lib


def timed(g, n=300):
    for _ in range(10):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for par in (1, 0):
    lib.dtqn_set_parallel_wgrad(par)
    for chunk in (64, 128, 256, 512, 1600):
        assert lib.dtqn_set_wgrad_chunk(chunk) == 0
        g = tr.capture(tr.train_only)
        print(f"side_stream={par} wgrad_chunk={chunk:5d}  train half {timed(g):7.1f} us", flush=True)
