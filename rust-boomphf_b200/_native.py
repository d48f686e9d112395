"""Loader for libboomphf_cuda.so (the C ABI declared in include/boomphf_cuda.h).

There is deliberately no CPU fallback here: if the CUDA library is missing, cannot be loaded
or finds no sm_100 device, every entry point raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# measurement aid: BPHF_LIB_VARIANT=<tag> builds / loads libboomphf_cuda_<tag>.so with the extra nvcc flags of
# BPHF_NVCC_EXTRA (e.g. -DBPHF_PASS_VARIANT=1), so that two kernel variants can be timed by the same bench command
_VARIANT = os.environ.get("BPHF_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "libboomphf_cuda%s.so" % ("_" + _VARIANT if _VARIANT else ""))
_REPO = os.path.dirname(_HERE)

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]

BPHF_OK, BPHF_ERR_ARG, BPHF_ERR_GAMMA, BPHF_ERR_MAX_ITERS, BPHF_ERR_CUDA, BPHF_ERR_NOMEM, BPHF_ERR_INTERNAL = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1
NONE = 0xFFFFFFFFFFFFFFFF
MAX_LEVELS = 101

# every symbol include/boomphf_cuda.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bphf_abi_version", "bphf_last_error", "bphf_device_count", "bphf_wy_swap_8byte_tail",
    "bphf_build_u64", "bphf_num_levels", "bphf_level_dims", "bphf_level_export", "bphf_export_all",
    "bphf_total_dims", "bphf_from_levels", "bphf_free", "bphf_hash_batch_u64", "bphf_map_permute_u64",
    "bphf_map_get_batch_u64", "bphf_get_build_stats", "bphf_set_profile", "bphf_set_build_mode", "bphf_set_query_mode", "bphf_set_query_partition",
    "bphf_query_partition_stats", "bphf_keygen_u64",
    "bphf_check_permutation", "bphf_checksum_u64", "bphf_host_alloc", "bphf_host_free",
    "bphf_comm_create", "bphf_comm_ipc_handle", "bphf_comm_connect_ipc", "bphf_comm_connect_local",
    "bphf_comm_destroy", "bphf_build_sharded_u64", "bphf_build_chunked_u64",
    "bphf_build_keys", "bphf_hash_batch_keys", "bphf_from_levels_keys", "bphf_key_kind",
    "bphf_build_stream", "bphf_hash_batch_stream", "bphf_set_ingest_mode", "bphf_set_xchg_bits", "bphf_l2_probe",
]


class BuildStats(C.Structure):
    _fields_ = [
        ("n_levels", C.c_uint32),
        ("tail_level", C.c_uint32),
        ("keys_in", C.c_uint64 * (MAX_LEVELS + 1)),
        ("bits", C.c_uint64 * (MAX_LEVELS + 1)),
        ("kernel_launches", C.c_uint32),
        ("host_syncs", C.c_uint32),
        ("pass_ms", C.c_float * (MAX_LEVELS + 1)),
        ("fin_ms", C.c_float * (MAX_LEVELS + 1)),
        ("tail_ms", C.c_float),
        ("ranks_ms", C.c_float),
        ("total_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("bucketed_levels", C.c_uint32),
        ("rebuilds", C.c_uint32),
        ("merge_ms", C.c_float * (MAX_LEVELS + 1)),
        ("cascade_ms", C.c_float),
        ("streamed_ingest", C.c_uint32),
        ("peak_device_bytes", C.c_uint64),
    ]


def _sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(_REPO, "include", "boomphf_cuda.h")]


def build_native(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> rust-boomphf_b200/libboomphf_cuda.so (in-tree)."""
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in _sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    # several ranks of one torchrun may get here at once: build under a lock, into a temporary, then rename
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and os.path.exists(LIB_PATH) and \
                    os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(p) for p in _sources()):
                return LIB_PATH  # another process built it while we waited
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            extra = os.environ.get("BPHF_NVCC_EXTRA", "").split() if _VARIANT else []
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [
                "-o", tmp, os.path.join(_CSRC, "boomphf_cuda.cu")]
            subprocess.check_call(cmd)
            os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C ABI.  Raises if the library is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libboomphf_cuda.so is not built (run __graft_entry__.build()); this package has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    u64, u32, i32, dbl, vp = C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_void_p
    pu64, pu32 = C.POINTER(u64), C.POINTER(u32)
    L.bphf_abi_version.restype = i32
    L.bphf_last_error.restype = C.c_char_p
    L.bphf_device_count.restype = i32
    L.bphf_device_count.argtypes = [C.POINTER(i32)]
    L.bphf_wy_swap_8byte_tail.restype = i32
    L.bphf_build_u64.restype = i32
    L.bphf_build_u64.argtypes = [vp, u64, dbl, i32, u64, i32, i32, vp, C.POINTER(vp)]
    L.bphf_num_levels.restype = i32
    L.bphf_num_levels.argtypes = [vp, pu32]
    L.bphf_level_dims.restype = i32
    L.bphf_level_dims.argtypes = [vp, u32, pu64, pu64, pu64]
    L.bphf_level_export.restype = i32
    L.bphf_level_export.argtypes = [vp, u32, vp, vp, i32]
    L.bphf_export_all.restype = i32
    L.bphf_export_all.argtypes = [vp, vp, vp, i32]
    L.bphf_total_dims.restype = i32
    L.bphf_total_dims.argtypes = [vp, pu64, pu64, pu64]
    L.bphf_from_levels.restype = i32
    L.bphf_from_levels.argtypes = [u32, vp, vp, vp, i32, C.POINTER(vp)]
    L.bphf_free.restype = None
    L.bphf_free.argtypes = [vp]
    L.bphf_hash_batch_u64.restype = i32
    L.bphf_hash_batch_u64.argtypes = [vp, vp, u64, vp, i32, vp]
    L.bphf_map_permute_u64.restype = i32
    L.bphf_map_permute_u64.argtypes = [vp, vp, vp, u64, vp, vp, i32, vp]
    L.bphf_map_get_batch_u64.restype = i32
    L.bphf_map_get_batch_u64.argtypes = [vp, vp, vp, u64, vp, u64, vp, vp, i32, vp]
    L.bphf_get_build_stats.restype = i32
    L.bphf_get_build_stats.argtypes = [vp, C.POINTER(BuildStats)]
    L.bphf_set_profile.restype = i32
    L.bphf_set_profile.argtypes = [i32]
    L.bphf_set_query_mode.restype = i32
    L.bphf_set_query_mode.argtypes = [i32]
    L.bphf_set_query_partition.restype = i32
    L.bphf_set_query_partition.argtypes = [i32, u32, u64, u64, dbl]
    L.bphf_query_partition_stats.restype = i32
    L.bphf_query_partition_stats.argtypes = [i32, pu64, pu64, pu32]
    L.bphf_set_build_mode.restype = i32
    L.bphf_set_build_mode.argtypes = [i32, u64]
    L.bphf_keygen_u64.restype = i32
    L.bphf_keygen_u64.argtypes = [vp, u64, u64, u64, i32, i32, vp]
    L.bphf_check_permutation.restype = i32
    L.bphf_check_permutation.argtypes = [vp, u64, i32, pu64]
    L.bphf_checksum_u64.restype = i32
    L.bphf_checksum_u64.argtypes = [vp, u64, i32, pu64, pu64]
    L.bphf_host_alloc.restype = vp
    L.bphf_host_alloc.argtypes = [C.c_size_t]
    L.bphf_host_free.restype = None
    L.bphf_host_free.argtypes = [vp]
    L.bphf_comm_create.restype = i32
    L.bphf_comm_create.argtypes = [i32, i32, i32, u64, dbl, C.POINTER(vp)]
    L.bphf_comm_ipc_handle.restype = i32
    L.bphf_comm_ipc_handle.argtypes = [vp, vp]
    L.bphf_comm_connect_ipc.restype = i32
    L.bphf_comm_connect_ipc.argtypes = [vp, vp]
    L.bphf_comm_connect_local.restype = i32
    L.bphf_comm_connect_local.argtypes = [vp, C.POINTER(vp)]
    L.bphf_comm_destroy.restype = None
    L.bphf_comm_destroy.argtypes = [vp]
    L.bphf_build_keys.restype = i32
    L.bphf_build_keys.argtypes = [vp, u64, i32, i32, dbl, i32, u64, i32, i32, C.POINTER(vp)]
    L.bphf_hash_batch_keys.restype = i32
    L.bphf_hash_batch_keys.argtypes = [vp, vp, u64, vp, i32, vp]
    L.bphf_from_levels_keys.restype = i32
    L.bphf_from_levels_keys.argtypes = [u32, vp, vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.bphf_set_ingest_mode.restype = i32
    L.bphf_set_ingest_mode.argtypes = [i32, u64]
    L.bphf_set_xchg_bits.restype = i32
    L.bphf_set_xchg_bits.argtypes = [dbl]
    L.bphf_l2_probe.restype = i32
    L.bphf_l2_probe.argtypes = [i32, u64, u64, i32, C.POINTER(C.c_float)]
    L.bphf_build_stream.restype = i32
    L.bphf_build_stream.argtypes = [vp, u64, vp, u64, vp, u64, dbl, i32, u64, i32, i32, C.POINTER(vp)]
    L.bphf_hash_batch_stream.restype = i32
    L.bphf_hash_batch_stream.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, i32, vp]
    L.bphf_key_kind.restype = i32
    L.bphf_key_kind.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.bphf_build_chunked_u64.restype = i32
    L.bphf_build_chunked_u64.argtypes = [vp, vp, u64, u64, dbl, i32, i32, u64, i32, i32, C.POINTER(vp)]
    L.bphf_build_sharded_u64.restype = i32
    L.bphf_build_sharded_u64.argtypes = [vp, vp, u64, u64, dbl, i32, u64, i32, C.POINTER(vp)]
    _lib = L
    return L


def last_error():
    return (lib().bphf_last_error() or b"").decode("utf-8", "replace")
