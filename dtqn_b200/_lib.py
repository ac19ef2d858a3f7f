"""ctypes binding of libdtqn_b200.so (the C ABI declared in include/dtqn_b200.h).

The product path has NO CPU fallback: importing this module without the built library, or calling into it
without a CUDA device, raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C dtqn_b200/csrc``.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtqn_b200.so")

ENV_CARFLAG, ENV_MEMORY = 0, 1
ACT_GIVEN, ACT_RANDOM, ACT_EPS_GREEDY = 0, 1, 2

_p = C.c_void_p


class EnvStruct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_envs", C.c_int32), ("obs_dim", C.c_int32), ("num_actions", C.c_int32),
                ("max_episode_steps", C.c_int32), ("stat_episodes_per_env", C.c_int32),
                ("rng", _p), ("rng_buf", _p), ("arng", _p), ("arng_buf", _p),
                ("pos", _p), ("vel", _p), ("heaven", _p), ("cards", _p), ("shown", _p), ("cur", _p),
                ("elapsed", _p), ("done_flag", _p), ("block_counts", _p), ("ep_stats", _p), ("ep_return", _p), ("env_acc", _p)]


class ReplayStruct(C.Structure):
    _fields_ = [("n_slots", C.c_int32), ("max_episode_steps", C.c_int32), ("obs_dim", C.c_int32),
                ("context_len", C.c_int32), ("obs_mask", C.c_float), ("record_every", C.c_int32),
                ("obss", _p), ("actions", _p), ("rewards", _p), ("dones", _p), ("episode_lengths", _p),
                ("slot_open", _p), ("counters", _p), ("env_slot", _p), ("env_prev_len", _p)]


class ContextStruct(C.Structure):
    _fields_ = [("context_len", C.c_int32), ("obs_dim", C.c_int32), ("trunc_obs", C.c_int32), ("obs_mask", C.c_float),
                ("obs", _p), ("timestep", _p), ("action", _p)]


class StepIO(C.Structure):
    _fields_ = [("action_mode", C.c_int32), ("_pad", C.c_int32), ("epsilon", C.c_double), ("actions", _p), ("q_last", _p),
                ("obs_out", _p), ("reward_out", _p), ("done_out", _p), ("truncated_out", _p), ("success_out", _p),
                ("epsilon_dev", _p)]


class DtqnLibError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise DtqnLibError(
            f"{LIB_PATH} is missing: build the sm_100a library first (make -C dtqn_b200/csrc). "
            "dtqn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.dtqn_version.restype = C.c_int
    sigs = {
        "dtqn_env_reset_all": [C.POINTER(EnvStruct), C.POINTER(ReplayStruct), C.POINTER(ContextStruct), _p],
        "dtqn_env_step": [C.POINTER(EnvStruct), C.POINTER(ReplayStruct), C.POINTER(ContextStruct), C.POINTER(StepIO), _p],
        "dtqn_replay_sample_indices": [C.POINTER(ReplayStruct), C.c_int32, C.c_uint64, C.c_uint64, _p, _p, _p, _p],
        "dtqn_replay_gather": [C.POINTER(ReplayStruct), C.c_int32, _p, _p, _p, _p, _p, _p, _p, _p],
        "dtqn_eps_anneal": [_p, _p, _p],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    # opt-in: programmatic dependent launch for the training chain (csrc/common.cuh launch_k); off by default
    lib.dtqn_set_pdl(C.c_int32(1 if os.environ.get("DTQN_B200_PDL", "0") == "1" else 0))
    # launch-shape overrides of the backward pass (defaults are compiled in; see include/dtqn_b200.h)
    for var, fn in (("DTQN_B200_DGRAD_ROWS", lib.dtqn_set_dgrad_rows), ("DTQN_B200_FUSE_LN_BWD", lib.dtqn_set_fuse_ln_bwd),
                    ("DTQN_B200_HEAD_BWD_TOKENS", lib.dtqn_set_head_bwd_tokens)):
        if var in os.environ and fn(C.c_int32(int(os.environ[var]))) != 0:
            raise DtqnLibError(f"{var}={os.environ[var]}: unsupported value")
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise ValueError(f"{what}: invalid argument (code {rc})")
    raise DtqnLibError(f"{what}: CUDA error {rc}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DtqnLibError("dtqn_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise DtqnLibError("dtqn_b200 requires a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")
