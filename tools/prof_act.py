"""Profiling target: a few eager acting iterations (acting forward + env step) of the bench workload, for
`ncu --set full -k regex:act_fused ...` (see /opt/skills/guides/B200_PROFILING.md).  Usage: python tools/prof_act.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=4096, seed=1, device="cuda:0", batch=32)
tr.prepopulate(260)
tr._sync_epsilon_to_device()
for _ in range(iters):
    tr.act_only()
torch.cuda.synchronize()
if "--timeline" in sys.argv:
    import ctypes as C
    from dtqn_b200 import _lib
    buf = torch.zeros(256, dtype=torch.int64, device="cuda:0")
    _lib.lib.dtqn_set_act_fused_timeline(C.c_void_p(buf.data_ptr()))
    tr.act_only()
    torch.cuda.synchronize()
    _lib.lib.dtqn_set_act_fused_timeline(C.c_void_p(0))
    t = buf.cpu().numpy().reshape(8, 32)
    names = {0: "start", 1: "embed", 2: "w:qkv", 3: "dump0", 4: "attn", 5: "w:out", 6: "ln1", 7: "w:f1_0", 8: "w:a2e0", 9: "hid0",
             10: "w:f1_1", 11: "w:a2e1", 12: "hid1", 13: "w:f1_2", 14: "w:a2e2", 15: "hid2", 16: "w:f1_3", 17: "w:a2e3", 18: "hid3",
             19: "w:f2", 20: "ln2", 21: "w:l1", 22: "dump1", 23: "lastrow"}
    for i in range(1, 6):
        row = t[i]
        print("tile", i, "total cycles", int(t[i + 1][0] - row[0]) if i < 7 else 0,
              " ".join(f"{names[k]}={int(row[k] - row[k - 1])}" for k in range(1, 24)))
        # issuer (one thread): wake-up after the workers' arrive, MMA issue + commit, accumulator-ready wake-up of worker 0
        print("   issuer: ax0->wake", int(row[24] - row[1]), "issue qkv", int(row[25] - row[24]), "commit->worker", int(row[2] - row[25]),
              "| ao->wake", int(row[26] - row[4]), "issue out", int(row[27] - row[26]), "commit->worker", int(row[5] - row[27]),
              "| ax1->wake", int(row[28] - row[6]), "| f2 commit->worker", int(row[19] - row[29]),
              "| ax2->wake", int(row[30] - row[20]), "issue l1", int(row[31] - row[30]), "commit->worker", int(row[21] - row[31]))
print("ok")
