"""Data-parallel plumbing (one process per GPU): env/replay sharding by rank and the single gradient collective.

The reference has no distributed code; the hot path shards naturally: envs and replay episodes
are independent, only the parameters couple ranks, so each update needs exactly one allreduce(sum) of the flat gradient
buffer, after which every rank divides by the world size, clips by the GLOBAL norm and applies the identical Adam step
(equal per-rank batch => mean of per-rank mean-MSE gradients == gradient of the mean over the concatenated batch)."""
import ctypes as C
import os

import torch
import torch.distributed as dist

P2P_MAX_RANKS = 8


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


# ---- gradient exchange fused with the optimiser over NVLink peer memory (csrc/p2p.cu) ---------------------------------------
class _DeviceArray:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned device memory."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class PeerExchange:
    """One IPC-exported exchange buffer per rank + the peers' buffers mapped into this process.

    ``grads`` aliases the gradient area of the local buffer: the backward kernels write straight into it, and
    ``dtqn_allreduce_clip_adam`` reads every rank's area through NVLink (no NCCL call, graph-capturable).
    ``handles`` (world x 64 bytes, rank order) come from any host channel; ``PeerExchange.create`` uses
    torch.distributed for it."""

    def __init__(self, n_floats: int, device, rank: int, world: int):
        from dtqn_b200 import _lib
        self._lib, self.rank, self.world, self.n = _lib.lib, rank, world, int(n_floats)
        self.device = torch.device(device)
        if not 1 <= world <= P2P_MAX_RANKS:
            raise ValueError(f"peer exchange supports 1..{P2P_MAX_RANKS} ranks on one node, got {world}")
        l = self._lib
        l.dtqn_p2p_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_char_p]
        l.dtqn_p2p_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        l.dtqn_p2p_close.argtypes = [C.c_void_p]
        l.dtqn_p2p_free.argtypes = [C.c_void_p]
        l.dtqn_p2p_error.argtypes = [C.c_void_p]
        base, g = C.c_void_p(), C.c_void_p()
        self._handle = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(l.dtqn_p2p_alloc(self.n, C.byref(base), C.byref(g), self._handle), "dtqn_p2p_alloc")
        self.base = base.value
        self.grads = torch.as_tensor(_DeviceArray(g.value, self.n, self), device=self.device)
        self.reduced = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.peer_bases = [None] * world
        self.peer_bases[rank] = self.base
        self._opened = []

    @property
    def handle(self) -> bytes:
        return self._handle.raw

    def open_peers(self, handles) -> None:
        from dtqn_b200 import _lib
        with torch.cuda.device(self.device):
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                _lib.check(self._lib.dtqn_p2p_open(bytes(h), C.byref(p)), f"dtqn_p2p_open(rank {r})")
                self.peer_bases[r] = p.value
                self._opened.append(p.value)
        self.bases = (C.c_void_p * self.world)(*self.peer_bases)

    @classmethod
    def create_single(cls, n_floats: int, device) -> "PeerExchange":
        """World of one (the same kernels with a self-signalling barrier): unit tests and single-GPU runs."""
        ex = cls(n_floats, device, 0, 1)
        ex.open_peers([ex.handle])
        return ex

    def error(self) -> bool:
        """True if a bounded cross-rank wait expired (a peer never arrived).  Synchronises."""
        return self._lib.dtqn_p2p_error(C.c_void_p(self.base)) != 0

    def close(self) -> None:
        for p in self._opened:
            self._lib.dtqn_p2p_close(C.c_void_p(p))
        self._opened = []


def peer_exchange_wanted(world: int) -> bool:
    """P2P exchange is the default for 2..8 ranks of one node; DTQN_B200_ALLREDUCE=nccl selects the NCCL allreduce."""
    mode = os.environ.get("DTQN_B200_ALLREDUCE", "p2p").lower()
    if mode not in ("p2p", "nccl"):
        raise ValueError("DTQN_B200_ALLREDUCE must be 'p2p' or 'nccl'")
    if mode == "nccl" or world < 2 or world > P2P_MAX_RANKS or dist.get_backend() != "nccl":
        return False
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    return local_world == world and torch.cuda.device_count() >= world


def try_peer_exchange(n_floats: int, device):
    """(PeerExchange, "") over the default group, or (None, reason) when the ranks cannot map each other's memory.
    Every step is collective and exception-free between collectives: if ANY rank fails, all fall back to NCCL."""
    rank, world = rank_world()
    dev = torch.device(device)

    def all_ok(ok: bool) -> bool:
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    why = ""
    try:
        ok = all(torch.cuda.can_device_access_peer(dev.index, j) for j in range(world) if j != dev.index)
        why = "" if ok else "no peer access between the node's GPUs"
    except Exception as e:                                   # pragma: no cover
        ok, why = False, repr(e)
    if not all_ok(ok):
        return None, why or "a peer rank reported no peer access"
    ex = None
    try:
        ex = PeerExchange(n_floats, dev, rank, world)
    except Exception as e:
        why = repr(e)
    mine = torch.frombuffer(bytearray(ex.handle if ex is not None else bytes(64)), dtype=torch.uint8).to(dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    if not all_ok(ex is not None):
        return None, why or "a peer rank could not allocate its exchange buffer"
    try:
        ex.open_peers([h.cpu().numpy().tobytes() for h in allh])
        ok = True
    except Exception as e:
        ok, why = False, repr(e)
    if not all_ok(ok):                                       # also the barrier: nobody launches before all have mapped
        ex.close()
        return None, why or "a peer rank could not map the exchange buffers (cudaIpcOpenMemHandle)"
    return ex, ""
