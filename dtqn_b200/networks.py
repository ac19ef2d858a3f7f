"""DTQN Q-network -- host-side mirror of ``dtqn.networks.dtqn.DTQN`` (dtqn/networks/dtqn.py:15-218).

Same constructor signature, ``forward(obss, actions=None, bag_obss=None, bag_actions=None) -> [B, L, A]`` and
``state_dict`` key names / shapes as the reference, so reference checkpoints load both ways.
Every parameter is a view into ONE flat fp32 CUDA buffer (the layout ``dtqn_net_param_offsets`` reports); the forward,
backward and optimiser are the hand-written sm_100a kernels of ``libdtqn_b200.so``.  There is no eager fallback.
"""
import ctypes as C
from typing import Optional, Union

import numpy as np
import torch
import torch.nn as nn

from dtqn_b200 import _lib


class NetCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("obs_dim", "num_actions", "d_model", "n_heads", "n_layers", "context_len",
                                         "discrete", "vocab", "embed_per_obs", "pos_trainable",
                                         "action_dim", "identity", "gate_gru")] + [("dropout", C.c_float),
                                                                                    ("dropout_state", C.c_void_p)]


class ObsSrc(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("seq_stride", C.c_int64), ("timestep", C.c_void_p), ("ring_len", C.c_int32),
                ("obs_mask", C.c_float), ("actions", C.c_void_p), ("act_stride", C.c_int64), ("train_mode", C.c_int32),
                ("_pad", C.c_int32)]


_l = _lib.lib
_l.dtqn_net_param_count.argtypes = [C.POINTER(NetCfg)]
_l.dtqn_net_param_count.restype = C.c_int64
_l.dtqn_net_param_offsets.argtypes = [C.POINTER(NetCfg), C.POINTER(C.c_int64), C.c_int32]
_l.dtqn_net_param_offsets.restype = C.c_int
_l.dtqn_net_workspace_floats.argtypes = [C.POINTER(NetCfg), C.c_int64, C.c_int32]
_l.dtqn_net_workspace_floats.restype = C.c_int64
_l.dtqn_forward.argtypes = [C.POINTER(NetCfg), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(ObsSrc),
                            C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
_l.dtqn_forward.restype = C.c_int
_l.dtqn_packed_bytes.argtypes = [C.POINTER(NetCfg)]
_l.dtqn_packed_bytes.restype = C.c_int64
_l.dtqn_pack_weights.argtypes = [C.POINTER(NetCfg), C.c_void_p, C.c_void_p, C.c_void_p]
_l.dtqn_pack_weights.restype = C.c_int
_l.dtqn_set_tc_min_tokens.argtypes = [C.c_int32]
_l.dtqn_tc_error.restype = C.c_int


class _Holder(nn.Module):
    """Parameter container used to reproduce the reference's module tree (and therefore its state_dict keys)."""


def _sinusoid(context_len, d):
    # position_encodings.py:22-35
    position = torch.arange(context_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2) * (-np.log(10000.0) / d))
    pe = torch.zeros(1, context_len, d)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


class DTQN(nn.Module):
    def __init__(self, obs_dim: int, num_actions: int, embed_per_obs_dim: int, action_dim: int, inner_embed_size: int,
                 num_heads: int, num_layers: int, history_len: int, dropout: float = 0.0, gate: str = "res",
                 identity: bool = False, pos: Union[str, int] = "learned", discrete: bool = False,
                 vocab_sizes: Optional[Union[np.ndarray, int]] = None, bag_size: int = 0, device=None, **kwargs):
        super().__init__()
        if isinstance(obs_dim, tuple):
            raise NotImplementedError("image observations are outside the B200 hot path (SURVEY.md section 2 #9)")
        if gate not in ("res", "gru"):
            raise ValueError("Gate must be one of `gru`, `res`")            # dtqn.py:113-114
        if pos not in ("learned", "sin", "none"):
            raise ValueError(f"{pos!r} is not a valid PosEnum")             # dtqn.py:101 (PosEnum(pos))
        if bag_size:
            raise NotImplementedError("DTQN-bag (--bag-size > 0) is outside the hot path (SURVEY.md section 2 #23)")
        dev = _lib.require_cuda(device)
        self.obs_dim, self.num_actions, self.discrete = int(obs_dim), int(num_actions), bool(discrete)
        self.history_len, self.num_heads, self.num_layers = int(history_len), int(num_heads), int(num_layers)
        self.inner_embed_size, self.pos_kind, self.bag_size = int(inner_embed_size), pos, 0
        self.action_dim, self.identity, self.gate_kind, self.dropout_p = int(action_dim), bool(identity), gate, float(dropout)
        # mask-stream counter of the dropout kernels (device scalar: CUDA-graph replays advance it without host writes)
        self.dropout_state = torch.zeros(1, dtype=torch.int64, device=dev)
        vocab = int(np.max(vocab_sizes)) if (discrete and vocab_sizes is not None) else 0
        if discrete:
            assert vocab > 0, "Discrete environments need to have a vocab size for the token embeddings"
            assert embed_per_obs_dim > 1, "Each observation feature needs at least 1 embed dim"
        self.cfg = NetCfg(obs_dim=self.obs_dim, num_actions=self.num_actions, d_model=self.inner_embed_size,
                          n_heads=self.num_heads, n_layers=self.num_layers, context_len=self.history_len,
                          discrete=int(self.discrete), vocab=vocab, embed_per_obs=int(embed_per_obs_dim) if discrete else 0,
                          pos_trainable=int(pos == "learned"), action_dim=self.action_dim, identity=int(self.identity),
                          gate_gru=int(gate == "gru"), dropout=self.dropout_p,
                          dropout_state=self.dropout_state.data_ptr() if self.dropout_p > 0 else None)
        n = _l.dtqn_net_param_count(C.byref(self.cfg))
        if n < 0:
            raise ValueError(f"unsupported DTQN configuration for the sm_100a kernels (code {n})")
        self.n_flat = int(n)
        self.flat = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
        offs = (C.c_int64 * 256)()
        cnt = _l.dtqn_net_param_offsets(C.byref(self.cfg), offs, 256)
        assert cnt > 0
        self._offs = list(offs[:cnt])
        self._build_tree(dev)
        self._init_weights()
        self._ws = {}
        # tcgen05 operand image of the GEMM weights (bf16 hi/lo, K-major tiles); refreshed by repack()
        self.packed = torch.zeros(int(_l.dtqn_packed_bytes(C.byref(self.cfg))), dtype=torch.uint8, device=dev)
        self.packed_stale = True

    def repack(self) -> None:
        """Refresh the tensor-core weight image from the fp32 parameters (after an optimiser step / load_state_dict)."""
        _lib.check(_l.dtqn_pack_weights(C.byref(self.cfg), self.flat.data_ptr(), self.packed.data_ptr(),
                                        _lib.stream_ptr()), "dtqn_pack_weights")
        self.packed_stale = False

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.packed_stale = True
        return out

    # ---- parameter tree -----------------------------------------------------------------------------------------------
    def _view(self, shape, trainable=True):
        o = self._offs[self._k]; self._k += 1
        n = int(np.prod(shape))
        self._slices.append((o, n, tuple(shape)))
        return nn.Parameter(self.flat[o:o + n].view(*shape), requires_grad=trainable)

    def unflatten(self, flat: torch.Tensor) -> dict:
        """Views of another flat buffer with this layout (gradients, Adam moments) keyed by state_dict name."""
        names = [n for n, _ in self.named_parameters() if not n.endswith("attn_mask")]
        assert len(names) == len(self._slices)
        return {n: flat[o:o + k].view(*shape) for n, (o, k, shape) in zip(names, self._slices)}

    def _build_tree(self, dev):
        d, A, ctx, O = self.inner_embed_size, self.num_actions, self.history_len, self.obs_dim
        self._k = 0
        self._slices = []
        self.obs_embedding = _Holder()
        if self.discrete:
            E, V = self.cfg.embed_per_obs, self.cfg.vocab
            seq = _Holder()
            e0, e2 = _Holder(), _Holder()
            e0.weight = self._view((V, E))
            e2.weight = self._view((d - self.action_dim, O * E)); e2.bias = self._view((d - self.action_dim,))
            seq.add_module("0", e0); seq.add_module("2", e2)
            self.obs_embedding.observation_embedding = seq
        else:
            lin = _Holder()
            lin.weight = self._view((d - self.action_dim, O)); lin.bias = self._view((d - self.action_dim,))
            self.obs_embedding.observation_embedding = lin
        if self.action_dim > 0:                                                # representations.py:146-155
            self.action_embedding = _Holder()
            seq = _Holder(); a0 = _Holder()
            a0.weight = self._view((A, self.action_dim))
            seq.add_module("0", a0)
            self.action_embedding.embedding = seq
        self.position_embedding = _Holder()
        self.position_embedding.position_encoding = self._view((1, ctx, d), trainable=self.pos_kind == "learned")
        self.transformer_layers = _Holder()
        mask = torch.triu(torch.ones(ctx, ctx, device=dev), diagonal=1)
        mask[mask.bool()] = -float("inf")                                     # transformer.py:49-53
        for i in range(self.num_layers):
            blk = _Holder()
            blk.attn_mask = nn.Parameter(mask.clone(), requires_grad=False)
            blk.layernorm1, blk.layernorm2, blk.attention, blk.ffn = _Holder(), _Holder(), _Holder(), _Holder()
            blk.layernorm1.weight = self._view((d,)); blk.layernorm1.bias = self._view((d,))
            blk.layernorm2.weight = self._view((d,)); blk.layernorm2.bias = self._view((d,))
            blk.attention.in_proj_weight = self._view((3 * d, d)); blk.attention.in_proj_bias = self._view((3 * d,))
            blk.attention.out_proj = _Holder()
            blk.attention.out_proj.weight = self._view((d, d)); blk.attention.out_proj.bias = self._view((d,))
            f0, f2 = _Holder(), _Holder()
            f0.weight = self._view((4 * d, d)); f0.bias = self._view((4 * d,))
            f2.weight = self._view((d, 4 * d)); f2.bias = self._view((d,))
            blk.ffn.add_module("0", f0); blk.ffn.add_module("2", f2)
            if self.gate_kind == "gru":          # ONE attention gate and ONE mlp gate shared by every layer (dtqn.py:107-131)
                if i == 0:
                    gates = []
                    for _ in range(2):           # gates.py:13-18; w_z.bias = -2 (:22-24) is set by _init_weights
                        g = _Holder()
                        for nm in ("w_r", "u_r", "w_z", "u_z", "w_g", "u_g"):
                            setattr(g, nm, _Holder())
                        g.w_r.weight = self._view((d, d)); g.u_r.weight = self._view((d, d))
                        g.w_z.weight = self._view((d, d)); g.w_z.bias = self._view((d,))
                        g.u_z.weight = self._view((d, d)); g.w_g.weight = self._view((d, d)); g.u_g.weight = self._view((d, d))
                        gates.append(g)
                blk.attn_gate, blk.mlp_gate = gates
            self.transformer_layers.add_module(str(i), blk)
        self.ffn = _Holder()
        h0, h2 = _Holder(), _Holder()
        h0.weight = self._view((d, d)); h0.bias = self._view((d,))
        h2.weight = self._view((A, d)); h2.bias = self._view((A,))
        self.ffn.add_module("0", h0); self.ffn.add_module("2", h2)
        assert self._k == len(self._offs)

    @torch.no_grad()
    def _init_weights(self):
        """utils/torch_utils.py:4-15: N(0, 0.02) for Linear / Embedding / in_proj / out_proj weights, zero biases,
        LayerNorm (1, 0); learned position table zeros (position_encodings.py:41-43)."""
        for name, p in self.named_parameters():
            if name.endswith("attn_mask"):
                continue
            if "position_encoding" in name:
                p.copy_(_sinusoid(self.history_len, self.inner_embed_size) if self.pos_kind == "sin" else torch.zeros_like(p))
            elif "layernorm" in name:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            elif name.endswith("w_z.bias"):
                p.fill_(-2.0)                                                  # GTrXL gate bias (gates.py:22-24)
            elif name.endswith("bias"):
                p.zero_()
            else:
                p.copy_(torch.empty(p.shape).normal_(mean=0.0, std=0.02))

    def _apply(self, fn, recurse=True):
        # parameters are views of one flat CUDA buffer; moving them individually would break the kernels' layout
        probe = fn(torch.empty(0, device=self.flat.device))
        if probe.device != self.flat.device or probe.dtype != torch.float32:
            raise RuntimeError("dtqn_b200.DTQN lives on its CUDA device in fp32; construct it with device=...")
        return self

    # ---- kernels ------------------------------------------------------------------------------------------------------------
    def workspace(self, n_tokens: int, save: int) -> torch.Tensor:
        key = (int(n_tokens), int(save))
        ws = self._ws.get(key)
        if ws is None:
            n = _l.dtqn_net_workspace_floats(C.byref(self.cfg), n_tokens, save)
            ws = torch.empty(int(n), dtype=torch.float32, device=self.flat.device)
            self._ws[key] = ws
        return ws

    def forward(self, obss: torch.Tensor, actions: Optional[torch.Tensor] = None, bag_obss=None, bag_actions=None):
        """obss [B, L, O] (float, or integer ids for discrete envs) -> Q [B, L, A] (dtqn.py:158-218)."""
        assert obss.dim() == 3, "obss is batch x seq_len x obs_dim"
        B, L, O = obss.shape
        if B == 0 or L == 0:
            return torch.empty((B, L, self.num_actions), dtype=torch.float32, device=self.flat.device)
        assert L <= self.history_len, "Cannot forward, history is longer than expected."           # dtqn.py:171-173
        assert O == self.obs_dim, f"Obs dim is incorrect. Expected {self.obs_dim} got {O}"          # dtqn.py:177-179
        x = obss.to(device=self.flat.device, dtype=torch.float32).contiguous()
        q = torch.empty((B, L, self.num_actions), dtype=torch.float32, device=self.flat.device)
        act = None
        if self.action_dim > 0:                                                # dtqn.py:184-192
            assert actions is not None, "a network built with action_dim > 0 needs the actions"
            act = actions.to(device=self.flat.device).reshape(B, L).to(torch.uint8).contiguous()
        src = ObsSrc(obs=x.data_ptr(), seq_stride=L * O, timestep=None, ring_len=0, obs_mask=0.0,
                     actions=act.data_ptr() if act is not None else None, act_stride=L, train_mode=int(self.training))
        forward_groups(self, [self], [src], B, L, q_mode=0, save=0, q_out=q)
        return q


def forward_groups(net: DTQN, nets, srcs, n_seq: int, L: int, q_mode: int, save: int, q_out: torch.Tensor,
                   workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dtqn_forward over len(nets) groups (group g uses nets[g]'s parameters) sharing one launch sequence.
    Returns the workspace used."""
    G = len(nets)
    ws = workspace if workspace is not None else net.workspace(G * n_seq * L, save)
    for m in nets:
        if m.packed_stale:
            m.repack()
    pp = (C.c_void_p * G)(*[m.flat.data_ptr() for m in nets])
    pk = (C.c_void_p * G)(*[m.packed.data_ptr() for m in nets])
    ss = (ObsSrc * G)(*srcs)
    _lib.check(_l.dtqn_forward(C.byref(net.cfg), G, pp, pk, ss, n_seq, L, q_mode, save, ws.data_ptr(), ws.numel(),
                               q_out.data_ptr(), _lib.stream_ptr()), "dtqn_forward")
    return ws


def set_tc_min_tokens(n: int) -> None:
    """Groups with >= n tokens run their GEMMs on the tcgen05 path (default 4096)."""
    _l.dtqn_set_tc_min_tokens(int(n))


def tc_error() -> bool:
    return bool(_l.dtqn_tc_error())
