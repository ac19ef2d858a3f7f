"""Data-parallel plumbing (one process per GPU): env/replay sharding by rank and the single gradient collective.

The reference has no distributed code; the hot path shards naturally: envs and replay episodes
are independent, only the parameters couple ranks, so each update needs exactly one allreduce(sum) of the flat gradient
buffer, after which every rank divides by the world size, clips by the GLOBAL norm and applies the identical Adam step
(equal per-rank batch => mean of per-rank mean-MSE gradients == gradient of the mean over the concatenated batch)."""
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_seed(seed: int, rank: int, n_envs: int) -> int:
    """Rank g owns env seeds [seed + g*n_envs, seed + (g+1)*n_envs)."""
    return seed + rank * n_envs


def allreduce_gradients(flat_grads: torch.Tensor) -> float:
    """In-place sum over ranks of the flat gradient buffer; returns the scale (1/world) the optimiser kernel applies
    before the global-norm clip.  No-op (scale 1) in a single-process run."""
    _, world = rank_world()
    if world > 1:
        dist.all_reduce(flat_grads)
    return 1.0 / world


def broadcast_parameters(flat_params: torch.Tensor, src: int = 0) -> None:
    _, world = rank_world()
    if world > 1:
        dist.broadcast(flat_params, src=src)
