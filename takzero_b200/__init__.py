"""takzero_b200: B200-native batched self-play search (CUDA library + thin host mirrors).

The product is `libtakzero_b200.so` (C ABI in include/takzero_b200.h).  Nothing here computes
on the CPU; `capi.lib()` raises ImportError when the CUDA library has not been built."""
from . import capi  # noqa: F401
from .capi import BatchedMCTS, TakzeroError  # noqa: F401
