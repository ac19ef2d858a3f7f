"""world_size = 2 over gloo on CPU: the N > 1 host logic of the data-parallel path -- env seed sharding and the one
gradient collective (sum -> /world -> global-norm clip -> identical Adam) reproduce the single-process step on the
concatenated batch.  The per-rank gradients come from the CPU oracle (no GPU here); the collective helpers are the
product's own (dtqn_b200.parallel)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _batch(z, s, rows):
    f = lambda a: torch.from_numpy(a[rows])
    return (f(z[f"step{s}/obss"]).float(), f(z[f"step{s}/actions"].astype(np.int64)), f(z[f"step{s}/rewards"]),
            f(z[f"step{s}/next_obss"]).float(), f(z[f"step{s}/next_actions"].astype(np.int64)), f(z[f"step{s}/dones"]))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dtqn_b200 import parallel
    from oracle import agent as oagent, network as onet
    assert parallel.rank_world() == (rank, world)
    z = np.load(os.path.join(GOLDEN, "train_carflag.npz"))
    sd = {k[len("policy0/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("policy0/")}
    tgt = {k[len("target0/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("target0/")}
    if rank != 0:                                    # parameters are broadcast from rank 0 at start-up
        sd = {k: torch.zeros_like(v) if not k.endswith("attn_mask") else v for k, v in sd.items()}
    keys = onet.trainable_keys(sd)
    flat = torch.cat([sd[k].reshape(-1) for k in keys])
    parallel.broadcast_parameters(flat, src=0)
    o = 0
    for k in keys:
        n = sd[k].numel(); sd[k] = flat[o:o + n].view_as(sd[k]).clone(); o += n
    B = 32
    rows = np.arange(rank * (B // world), (rank + 1) * (B // world))     # this rank's shard of the batch
    for k in keys:
        sd[k].requires_grad_(True)
    loss, _, _ = oagent.td_loss(sd, tgt, _batch(z, 0, rows), 8)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys])
    g = torch.cat([x.reshape(-1) for x in grads])
    scale = parallel.allreduce_gradients(g)                               # the one collective
    assert scale == 1.0 / world
    g = g * scale
    total = torch.linalg.vector_norm(g)
    coef = torch.clamp(1.0 / (total + 1e-6), max=1.0)
    np.save(os.path.join(out_dir, f"g{rank}.npy"), (g * coef).numpy())
    np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array([total.item()]))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1)                                        # identical update on every rank
    z = np.load(os.path.join(GOLDEN, "train_carflag.npz"))
    # single-process reference gradient on the concatenated batch = the reference's own step-0 gradient
    from oracle import network as onet
    sd = {k[len("policy0/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("policy0/")}
    ref = np.concatenate([z["step0/grad/" + k].reshape(-1) for k in onet.trainable_keys(sd)])
    total = np.linalg.norm(ref)
    ref_clipped = ref * min(1.0, 1.0 / (total + 1e-6))
    assert abs(np.load(tmp_path / "n0.npy")[0] - z["stats/grad_norms"][0]) < 2e-5
    assert np.abs(g0 - ref_clipped).max() < 2e-6


def test_seed_sharding_is_disjoint_and_contiguous():
    from dtqn_b200.parallel import shard_seed
    n = 4096
    starts = [shard_seed(1, r, n) for r in range(8)]
    assert starts == [1 + r * n for r in range(8)]
    allseeds = np.concatenate([np.arange(s, s + n) for s in starts])
    assert len(np.unique(allseeds)) == 8 * n
