"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/dtqn_b200.h declares
(no compute calls -- there is no GPU here) and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dtqn_b200.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(dtqn_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "dtqn_b200", "libdtqn_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 5
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dtqn_b200.h but not exported"
    lib.dtqn_version.restype = ctypes.c_int
    header_ver = int(re.search(r"#define DTQN_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "dtqn_b200.h")).read()).group(1))
    assert lib.dtqn_version() == header_ver


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dtqn_b200 import _lib
    from dtqn_b200.envs import BatchedEnv
    with pytest.raises(_lib.DtqnLibError):
        BatchedEnv("DiscreteCarFlag-v0", 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dtqn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


class _NetCfg(ctypes.Structure):      # dtqn_net_cfg (include/dtqn_b200.h)
    _fields_ = [(n, ctypes.c_int32) for n in ("obs_dim", "num_actions", "d_model", "n_heads", "n_layers", "context_len",
                                              "discrete", "vocab", "embed_per_obs", "pos_trainable", "action_dim", "identity",
                                              "gate_gru")] + [("dropout", ctypes.c_float), ("dropout_state", ctypes.c_void_p)]


def test_layout_queries_run_on_the_host():
    """The layout functions are pure host arithmetic: parameter count == the reference's trainable parameters + attn_mask-free
    padding rules, and the weight-image size == tcgen05 tiles + the fused acting kernel's 240 KB image + the k-major copies."""
    lib = ctypes.CDLL(os.path.join(ROOT, "dtqn_b200", "libdtqn_b200.so"))
    lib.dtqn_net_param_count.restype = ctypes.c_int64
    lib.dtqn_packed_bytes.restype = ctypes.c_int64
    lib.dtqn_net_workspace_floats.restype = ctypes.c_int64
    lib.dtqn_net_workspace_floats.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32]
    d, A, O, ctx, layers = 64, 3, 3, 50, 2
    cfg = _NetCfg(O, A, d, 8, layers, ctx, 0, 0, 0, 1)
    n = lib.dtqn_net_param_count(ctypes.byref(cfg))
    # 107 779 trainable parameters for CarFlag (112 779 with the two non-trainable 50 x 50 attn_mask tensors);
    # the flat buffer pads every tensor to 4 floats
    trainable = d * O + d + ctx * d + layers * (4 * d + 3 * d * d + 3 * d + d * d + d + 4 * d * d + 4 * d + 4 * d * d + d) + d * d + d + A * d + A
    assert trainable == 107_779 and trainable <= n <= trainable + 4 * 8
    gemm = layers * 12 * d * d + d * d                       # in/out/ffn.0/ffn.2 per layer + head ffn.0
    tc_tiles = 4 * (gemm + layers * 3 * d * d)              # bf16 hi + lo = 4 bytes per weight; + K|V and Q operands per layer
    act_img = 240 * 1024                                    # in_proj0, out_proj0, ffn.0, ffn.2, in_proj1
    wt = 4 * gemm                                           # fp32 transposes
    assert lib.dtqn_packed_bytes(ctypes.byref(cfg)) == tc_tiles + act_img + wt
    # Memory-5: d = 128 -> no fused acting image, no k-major copies
    cfg_m = _NetCfg(10, 10, 128, 8, 2, 50, 1, 9, 8, 1)
    dm = 128
    assert lib.dtqn_packed_bytes(ctypes.byref(cfg_m)) == 4 * (2 * 12 * dm * dm + dm * dm + 2 * 3 * dm * dm)
    assert lib.dtqn_net_param_count(ctypes.byref(cfg_m)) >= 431_186
    # invalid configurations are rejected with a negative status, never a crash
    bad = _NetCfg(O, A, 96, 8, layers, ctx, 0, 0, 0, 1)
    assert lib.dtqn_net_param_count(ctypes.byref(bad)) < 0
    assert lib.dtqn_net_workspace_floats(ctypes.byref(cfg), 0, 0) < 0
    # ablation flags (run.py:98-103,151-167): --a-embed 8 swaps 8 observation-embedding columns for an [A, 8] action table,
    # --gate gru adds ONE attention gate + ONE mlp gate (6 d^2 + d each) whatever the layer count, any d_model % 4 == 0 goes
    n_a = lib.dtqn_net_param_count(ctypes.byref(_NetCfg(O, A, d, 8, layers, ctx, 0, 0, 0, 1, 8)))
    assert n_a - n == A * 8 - 8 * O - 8
    n_g = lib.dtqn_net_param_count(ctypes.byref(_NetCfg(O, A, d, 8, layers, ctx, 0, 0, 0, 1, 0, 0, 1)))
    assert n_g - n == 2 * (6 * d * d + d)
    assert lib.dtqn_net_param_count(ctypes.byref(_NetCfg(O, A, 96, 8, layers, ctx, 0, 0, 0, 1, 0, 1))) > 0
    assert lib.dtqn_net_param_count(ctypes.byref(_NetCfg(O, A, d, 8, layers, ctx, 0, 0, 0, 1, 0, 0, 0, 0.5))) < 0   # dropout needs its state
    ws_act = lib.dtqn_net_workspace_floats(ctypes.byref(cfg), 4096 * ctx, 0)
    ws_train = lib.dtqn_net_workspace_floats(ctypes.byref(cfg), 3 * 32 * ctx, 1)
    assert ws_act > 4096 * ctx * d and ws_train > 3 * 32 * ctx * d * 2 * 10
