"""comet_b200 -- B200-native (sm_100a) device layer for wizenheimer/comet's vector search path.

The product is `libcomet_b200.so` (C ABI in include/comet_b200.h).  This package only loads it and
exposes a thin ctypes binding (`comet_b200.capi`) used by the tests, bench.py and the Python
rendering of comet's builder API (`comet_b200.api`).  There is no CPU fallback: importing is fine
without a GPU, every compute call fails loudly without one.
"""
from .capi import lib, CometError  # noqa: F401
