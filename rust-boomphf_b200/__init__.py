"""B200-native MPHF construction and batched lookup for u64 keys -- the hot path of
10XGenomics/rust-boomphf behind the reference's own interface (Mphf / BoomHashMap).

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/boomphf_cuda.h); this package is the thin host-side mirror used by tests and bench.
"""
from ._native import LIB_PATH, NONE, SYMBOLS, build_native, lib  # noqa: F401
from .mphf import Mphf, MphfPanic  # noqa: F401
from .keys import StreamKeys, encode_keys  # noqa: F401
from .hashmap import BoomHashMap, BoomHashMap2, BoomHashMapAny, NoKeyBoomHashMap, NoKeyBoomHashMap2  # noqa: F401
