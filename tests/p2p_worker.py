"""torchrun worker of tests/test_p2p_gpu.py::test_two_processes_over_nvlink (2 ranks, one per GPU)."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    from dtqn_b200.runner import BatchedTrainer

    def same_on_all_ranks(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return all(torch.equal(parts[0], p) for p in parts)

    tr = BatchedTrainer("DiscreteCarFlag-v0", 256, seed=3, device=dev, batch=16, num_steps=2000)
    assert tr.allreduce == "p2p-fused", (tr.allreduce, tr.allreduce_note)
    ex = tr.agent.exchange
    # 1. the exchange kernel against NCCL on rank-specific gradients (two ranks: a + b is order-independent)
    g = torch.randn(ex.n, device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank)) * 1e-4
    ex.grads.copy_(g)
    want = g.clone()
    dist.all_reduce(want)
    before = tr.agent.policy_network.flat.clone()
    tr.agent.reduce_and_step()
    torch.cuda.synchronize()
    assert torch.equal(ex.reduced, want), float((ex.reduced - want).abs().max())
    assert not torch.equal(before, tr.agent.policy_network.flat)
    assert same_on_all_ranks(tr.agent.policy_network.flat) and same_on_all_ranks(tr.agent.exp_avg_sq)
    # 2. the training loop, eager then as one CUDA graph per iteration (collective inside the graph)
    tr.prepopulate(230)
    for _ in range(3):
        tr.train_iteration()
    tr.enable_graphs()
    for _ in range(20):
        tr.train_iteration()
    torch.cuda.synchronize()
    tr.agent.check_finite()
    assert same_on_all_ranks(tr.agent.policy_network.flat), "replicas diverged"
    assert same_on_all_ranks(tr.agent.exp_avg)
    assert tr.agent.num_train_steps == 24 and int(tr.agent.opt_step.item()) == 25      # + the bare update of part 1
    rs = tr.env.rng_state()
    other = [None] * world
    dist.all_gather_object(other, rs[:4].tolist())
    assert other[0] != other[1], "ranks must own different env shards"
    dist.barrier()
    if rank == 0:
        print("P2P_WORKER_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except Exception:
        import traceback
        traceback.print_exc()
        sys.exit(1)
