"""dtqn_b200 -- B200-native (sm_100a) data-parallel hot path of Deep Transformer Q-Networks.

Host side in Python over PyTorch-owned device tensors; all compute is hand-written CUDA behind the C ABI declared in
``include/dtqn_b200.h`` (``libdtqn_b200.so``, built in-tree).  No CPU fallback: importing without the built library
raises.
"""
from dtqn_b200 import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__all__ = ["_lib"]
