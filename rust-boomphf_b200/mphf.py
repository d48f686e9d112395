"""Host-side mirror of `boomphf::Mphf<u64>` over the C ABI (include/boomphf_cuda.h).

Same names, argument meaning and failure behaviour as the reference
(/root/reference/src/lib.rs): `Mphf.new` (:235), `Mphf.new_parallel` (:363), `hash` (:324),
`try_hash` (:340).  Where the reference panics, this raises `MphfPanic` with the same text.
Key sets are numpy uint64 arrays (host) or CUDA torch tensors of dtype int64/uint64 (device).
"""
import ctypes as C

import numpy as np

from . import _native as N
from .keys import StreamKeys, encode_keys, is_stream_like


class MphfPanic(RuntimeError):
    """What a Rust panic of the reference turns into on this side of the boundary."""

    def __init__(self, code, text):
        super().__init__(text)
        self.code = code


def _check(rc):
    if rc != N.BPHF_OK:
        raise MphfPanic(rc, N.last_error() or "boomphf_cuda error %d" % rc)


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and hasattr(x, "data_ptr") and bool(x.is_cuda)


KEY_U64, KEY_UINT, KEY_BYTES, KEY_STREAM = 0, 1, 2, 3


def _as_stream(objects, rust_type=None):
    """StreamKeys as they are; plain Python collections through the recorder stand-in (keys.encode_keys)"""
    if isinstance(objects, StreamKeys):
        return objects
    return encode_keys(objects, rust_type)


def _stream_args(sk):
    return (C.c_void_p(sk.data.ctypes.data), int(sk.data.size), C.c_void_p(sk.seg_ends.ctypes.data),
            int(sk.seg_ends.size), C.c_void_p(sk.key_segs.ctypes.data), len(sk))


def _kind_of(dtype_itemsize, ndim, shape, is_uint8):
    """(kind, width) by element size / shape: 8-byte ints -> u64; 4/2/1-byte ints -> UINT; (n, N<=8) uint8 -> [u8; N]"""
    if ndim == 2:
        if not is_uint8 or not 1 <= shape[1] <= 8:
            raise TypeError("byte-array keys must be uint8 of shape (n, N <= 8)")
        return KEY_BYTES, int(shape[1])
    if dtype_itemsize == 8:
        return KEY_U64, 8
    if dtype_itemsize in (1, 2, 4):
        return KEY_UINT, int(dtype_itemsize)
    raise TypeError("unsupported key type")


def _as_any_keys(keys):
    """keys of any supported kind -> (pointer, n, mem, device_or_None, keepalive, kind, width)"""
    if _is_torch_cuda(keys):
        k = keys.contiguous()
        kind, width = _kind_of(k.element_size(), k.dim(), tuple(k.shape), str(k.dtype) == "torch.uint8")
        n = k.shape[0] if k.dim() == 2 else k.numel()
        return C.c_void_p(k.data_ptr()), n, N.MEM_DEVICE, k.device.index or 0, k, kind, width
    if hasattr(keys, "is_cuda") and hasattr(keys, "numpy"):
        keys = keys.contiguous().numpy()
    a = keys if isinstance(keys, np.ndarray) else np.asarray(keys, dtype=np.uint64)
    if a.dtype.kind not in "iu":
        raise TypeError("keys must be integers")
    a = np.ascontiguousarray(a)
    kind, width = _kind_of(a.dtype.itemsize, a.ndim, a.shape, a.dtype == np.uint8)
    n = a.shape[0] if a.ndim == 2 else a.size
    return C.c_void_p(a.ctypes.data), n, N.MEM_HOST, None, a, kind, width


def _as_keys(keys):
    """-> (pointer, n, mem, device_or_None, keepalive)"""
    if _is_torch_cuda(keys):
        import torch
        if keys.dtype not in (torch.int64, torch.uint64):
            raise TypeError("device keys must be int64/uint64")
        k = keys.contiguous()
        return C.c_void_p(k.data_ptr()), k.numel(), N.MEM_DEVICE, k.device.index or 0, k
    if hasattr(keys, "is_cuda") and hasattr(keys, "numpy"):  # CPU torch tensor
        import torch
        if keys.dtype not in (torch.int64, torch.uint64):
            raise TypeError("keys must be int64/uint64")
        keys = keys.contiguous().view(torch.int64).numpy().view(np.uint64)
    a = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
    return C.c_void_p(a.ctypes.data), a.size, N.MEM_HOST, None, a


def _stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class Mphf:
    """A minimal perfect hash function over a set of u64 keys, resident on one B200."""

    def __init__(self, handle, device, rust_type=None):
        self._h = C.c_void_p(handle)
        self.device = device
        self.rust_type = rust_type  # stream keys only: the T of Mphf<T>, used to encode single items
        self._async_streams = []    # torch streams with possibly unfinished device-side lookups on this handle

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_h", None):
            # bphf_free must not overtake lookups still running on a caller's stream (include/boomphf_cuda.h)
            for st in getattr(self, "_async_streams", ()):
                try:
                    st.synchronize()
                except Exception:
                    pass
            self._async_streams = []
            N.lib().bphf_free(self._h)
            self._h = None

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def _build(cls, gamma, objects, starting_seed, device, stream, rust_type=None):
        if rust_type is not None or is_stream_like(objects):
            # Mphf<T> for any other T: Hash (Vec<u8>, String, Vec<String>, u128, tuples, ...): the recorded writes
            sk = _as_stream(objects, rust_type)
            out = C.c_void_p()
            _check(N.lib().bphf_build_stream(*_stream_args(sk), float(gamma), 0 if starting_seed is None else 1,
                                             0 if starting_seed is None else int(starting_seed), int(device),
                                             N.MEM_HOST, C.byref(out)))
            return cls(out.value, int(device), sk.rust_type)
        ptr, n, mem, dev, keep, kind, width = _as_any_keys(objects)
        if dev is not None:
            device = dev
        out = C.c_void_p()
        if kind == KEY_U64:
            rc = N.lib().bphf_build_u64(ptr, n, float(gamma), 0 if starting_seed is None else 1,
                                        0 if starting_seed is None else int(starting_seed), int(device), mem,
                                        _stream_ptr(stream), C.byref(out))
        else:  # Mphf<u32>, Mphf<[u8; N]>, ...: SURVEY 8f row 4
            rc = N.lib().bphf_build_keys(ptr, n, kind, width, float(gamma), 0 if starting_seed is None else 1,
                                         0 if starting_seed is None else int(starting_seed), int(device), mem,
                                         C.byref(out))
        del keep
        _check(rc)
        return cls(out.value, int(device))

    @classmethod
    def new(cls, gamma, objects, device=0, stream=None, rust_type=None):
        """Mphf::new(gamma, objects)  (src/lib.rs:235).  Same result as new_parallel.
        rust_type (e.g. "Vec<String>", "u128", "(u32, String)") or a `StreamKeys` selects the any-T path."""
        return cls._build(gamma, objects, None, device, stream, rust_type)

    @classmethod
    def new_parallel(cls, gamma, objects, starting_seed=None, device=0, stream=None, rust_type=None):
        """Mphf::new_parallel(gamma, objects, starting_seed)  (src/lib.rs:363)."""
        return cls._build(gamma, objects, starting_seed, device, stream, rust_type)

    @classmethod
    def _build_chunked(cls, gamma, objects, n, parallel, max_iters, device):
        parts = [_as_keys(c) for c in objects]
        mems = {p[2] for p in parts}
        if len(mems) > 1:
            raise TypeError("chunks must all live in host memory or all on the device")
        mem = mems.pop() if mems else N.MEM_HOST
        for p in parts:
            if p[3] is not None:
                device = p[3]
        ptrs = (C.c_void_p * max(len(parts), 1))(*[p[0].value for p in parts])
        lens = (C.c_uint64 * max(len(parts), 1))(*[p[1] for p in parts])
        out = C.c_void_p()
        rc = N.lib().bphf_build_chunked_u64(ptrs, lens, len(parts), int(n), float(gamma), 1 if parallel else 0,
                                            0 if max_iters is None else 1, 0 if max_iters is None else int(max_iters),
                                            int(device), mem, C.byref(out))
        del parts
        _check(rc)
        return cls(out.value, int(device))

    @classmethod
    def from_chunked_iterator(cls, gamma, objects, n, device=0):
        """Mphf::from_chunked_iterator(gamma, &objects, n)  (src/lib.rs:112): objects = iterable of key chunks"""
        return cls._build_chunked(gamma, list(objects), n, False, None, device)

    @classmethod
    def from_chunked_iterator_parallel(cls, gamma, objects, max_iters, n, num_threads=0, device=0):
        """Mphf::from_chunked_iterator_parallel(gamma, &objects, max_iters, n, num_threads)  (src/lib.rs:539).
        num_threads is accepted for signature parity; the GPU is the pool."""
        return cls._build_chunked(gamma, list(objects), n, True, max_iters, device)

    @classmethod
    def from_levels(cls, levels, device=0, key_kind=KEY_U64, key_width=8, rust_type=None):
        """Upload a deserialised Mphf<T>: levels = [(bits, words u64[], ranks u64[]), ...]; key_kind / key_width say
        what T is (default u64); rust_type names T for a stream-key Mphf (key_kind = KEY_STREAM)."""
        if rust_type is not None:
            key_kind, key_width = KEY_STREAM, 0
        n = len(levels)
        bits = np.array([l[0] for l in levels], dtype=np.uint64)
        words = [np.ascontiguousarray(l[1], dtype=np.uint64) for l in levels]
        ranks = [np.ascontiguousarray(l[2], dtype=np.uint64) for l in levels]
        for b, w, r in zip(bits, words, ranks):
            if len(w) != (int(b) + 63) // 64 or len(r) != (len(w) + 7) // 8:
                raise ValueError("level arrays do not match bits")
        wp = (C.c_void_p * max(n, 1))(*[w.ctypes.data for w in words])
        rp = (C.c_void_p * max(n, 1))(*[r.ctypes.data for r in ranks])
        out = C.c_void_p()
        _check(N.lib().bphf_from_levels_keys(n, C.c_void_p(bits.ctypes.data), wp, rp, int(key_kind), int(key_width),
                                             int(device), C.byref(out)))
        return cls(out.value, int(device), rust_type)

    @property
    def key_kind(self):
        k, w = C.c_int(), C.c_int()
        _check(N.lib().bphf_key_kind(self._h, C.byref(k), C.byref(w)))
        return int(k.value), int(w.value)

    # ---- data model -------------------------------------------------------------------------
    @property
    def num_levels(self):
        v = C.c_uint32()
        _check(N.lib().bphf_num_levels(self._h, C.byref(v)))
        return int(v.value)

    def level_dims(self, level):
        b, w, r = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = N.lib().bphf_level_dims(self._h, level, C.byref(b), C.byref(w), C.byref(r))
        if rc != N.BPHF_OK:
            raise MphfPanic(rc, "that level doesn't exist")  # src/lib.rs:302
        return int(b.value), int(w.value), int(r.value)

    def level(self, level):
        """(bits, words, ranks) of one level == (BitVector{bits, vector}, Box<[u64]>)"""
        bits, nw, nr = self.level_dims(level)
        words = np.empty(nw, dtype=np.uint64)
        ranks = np.empty(nr, dtype=np.uint64)
        _check(N.lib().bphf_level_export(self._h, level, C.c_void_p(words.ctypes.data),
                                         C.c_void_p(ranks.ctypes.data), N.MEM_HOST))
        return bits, words, ranks

    def levels(self):
        """All levels with two bulk copies (bphf_export_all)."""
        tw, tr, _ = self.total_dims()
        words = np.empty(tw, dtype=np.uint64)
        ranks = np.empty(tr, dtype=np.uint64)
        _check(N.lib().bphf_export_all(self._h, C.c_void_p(words.ctypes.data), C.c_void_p(ranks.ctypes.data),
                                       N.MEM_HOST))
        out, wo, ro = [], 0, 0
        for l in range(self.num_levels):
            bits, nw, nr = self.level_dims(l)
            out.append((bits, words[wo:wo + nw], ranks[ro:ro + nr]))
            wo += nw
            ro += nr
        return out

    def total_dims(self):
        w, r, k = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(N.lib().bphf_total_dims(self._h, C.byref(w), C.byref(r), C.byref(k)))
        return int(w.value), int(r.value), int(k.value)

    def build_stats_raw(self):
        """the bphf_build_stats struct itself (cheap: for callers that read a few fields inside a timed loop)"""
        st = N.BuildStats()
        _check(N.lib().bphf_get_build_stats(self._h, C.byref(st)))
        return st

    def build_stats(self):
        st = self.build_stats_raw()
        nl = int(st.n_levels)
        return {
            "n_levels": nl, "tail_level": int(st.tail_level),
            "keys_in": [int(st.keys_in[i]) for i in range(nl)], "bits": [int(st.bits[i]) for i in range(nl)],
            "kernel_launches": int(st.kernel_launches), "host_syncs": int(st.host_syncs),
            "pass_ms": [float(st.pass_ms[i]) for i in range(nl)], "fin_ms": [float(st.fin_ms[i]) for i in range(nl)],
            "tail_ms": float(st.tail_ms), "ranks_ms": float(st.ranks_ms), "total_ms": float(st.total_ms),
            "bucketed_levels": int(st.bucketed_levels), "rebuilds": int(st.rebuilds),
            "merge_ms": [float(st.merge_ms[i]) for i in range(nl)], "cascade_ms": float(st.cascade_ms),
            "streamed_ingest": int(st.streamed_ingest), "peak_device_bytes": int(st.peak_device_bytes),
        }

    # ---- lookups ----------------------------------------------------------------------------
    def hash_batch(self, keys, out=None, stream=None):
        """try_hash over a batch; entries equal to NONE (u64::MAX) are `None`.
        numpy in -> numpy out; CUDA tensor in -> CUDA tensor out (int64 view of the u64 values).
        A stream-key Mphf takes a `StreamKeys` or a collection of Python values of its rust_type."""
        if self.key_kind[0] == KEY_STREAM:
            sk = _as_stream(keys, self.rust_type)
            if out is None:
                out = np.empty(len(sk), dtype=np.uint64)
            _check(N.lib().bphf_hash_batch_stream(self._h, *_stream_args(sk), C.c_void_p(out.ctypes.data), N.MEM_HOST,
                                                  _stream_ptr(stream)))
            return out
        ptr, n, mem, dev, keep, kind, width = _as_any_keys(keys)
        if (kind, width) != self.key_kind:
            raise TypeError("keys are not of the type this Mphf was built for")
        fn = N.lib().bphf_hash_batch_u64 if kind == KEY_U64 else N.lib().bphf_hash_batch_keys
        if mem == N.MEM_DEVICE:
            import torch
            if dev != self.device:
                raise ValueError("keys live on another device than the Mphf")
            if out is None:
                out = torch.empty(n, dtype=torch.int64, device=keep.device)
            if stream is None:
                stream = torch.cuda.current_stream(keep.device)
            _check(fn(self._h, ptr, n, C.c_void_p(out.data_ptr()), mem, _stream_ptr(stream)))
            if hasattr(stream, "synchronize") and not any(st is stream or st == stream for st in self._async_streams):
                self._async_streams.append(stream)
            return out
        if out is None:
            out = np.empty(n, dtype=np.uint64)
        _check(fn(self._h, ptr, n, C.c_void_p(out.ctypes.data), mem, _stream_ptr(stream)))
        return out

    def try_hash(self, item):
        """Mphf::try_hash (src/lib.rs:340): Some(rank) or None"""
        kind, width = self.key_kind
        if kind == KEY_STREAM:
            q = [item]
        elif kind == KEY_BYTES:
            q = np.frombuffer(bytes(item), dtype=np.uint8).reshape(1, width)
        elif kind == KEY_UINT:
            q = np.array([int(item) & ((1 << (8 * width)) - 1)], dtype=np.dtype("u%d" % width))
        else:
            q = np.array([int(item) & (2 ** 64 - 1)], dtype=np.uint64)
        v = int(self.hash_batch(q)[0])
        return None if v == N.NONE else v

    def hash(self, item):
        """Mphf::hash (src/lib.rs:324); an item outside the construction set may panic"""
        v = self.try_hash(item)
        if v is None:
            raise MphfPanic(N.BPHF_ERR_ARG, "internal error: entered unreachable code: must find a hash value")
        return v
