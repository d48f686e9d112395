"""ctypes wrapper around oracle/libboomphf_oracle.so (CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package
(rust-boomphf_b200/).  See the header of boomphf_oracle.c for the parity status
("parity unpinned" at the wyhash boundary).
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libboomphf_oracle.so")
_lib = None

NONE = 0xFFFFFFFFFFFFFFFF  # what bo_try_hash returns where try_hash returns None
BO_OK, BO_ERR_ARG, BO_ERR_GAMMA, BO_ERR_MAX_ITERS, BO_ERR_NOMEM = 0, -1, -2, -3, -5


def build_lib(force=False):
    """Compile the oracle with the committed Makefile (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "boomphf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libboomphf_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build_lib()
    L = C.CDLL(_LIB_PATH)
    u64, u32, i32, dbl, vp = C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_void_p
    L.bo_wyhash_bytes.restype = u64
    L.bo_wyhash_bytes.argtypes = [vp, u64, u64]
    L.bo_wyrng.restype = u64
    L.bo_wyrng.argtypes = [C.POINTER(u64)]
    L.bo_set_key_kind.restype = i32
    L.bo_set_key_kind.argtypes = [i32, i32]
    L.bo_set_key_stream.restype = i32
    L.bo_set_key_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bo_hash_with_seed.restype = u64
    L.bo_hash_with_seed.argtypes = [u64, u64]
    L.bo_hashmod.restype = u64
    L.bo_hashmod.argtypes = [u64, u64, u64]
    L.bo_level_size.restype = u64
    L.bo_level_size.argtypes = [dbl, u64]
    L.bo_build_serial.restype = i32
    L.bo_build_serial.argtypes = [vp, u64, dbl, C.POINTER(vp)]
    L.bo_build_parallel.restype = i32
    L.bo_build_parallel.argtypes = [vp, u64, dbl, i32, u64, i32, C.POINTER(vp)]
    L.bo_build_chunked_serial.restype = i32
    L.bo_build_chunked_serial.argtypes = [vp, vp, u64, u64, dbl, C.POINTER(vp)]
    L.bo_build_chunked_parallel.restype = i32
    L.bo_build_chunked_parallel.argtypes = [vp, vp, u64, u64, dbl, i32, u64, i32, C.POINTER(vp)]
    L.bo_from_levels.restype = i32
    L.bo_from_levels.argtypes = [u32, vp, vp, vp, C.POINTER(vp)]
    L.bo_num_levels.restype = u32
    L.bo_num_levels.argtypes = [vp]
    L.bo_level_dims.restype = i32
    L.bo_level_dims.argtypes = [vp, u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.bo_level_words.restype = vp
    L.bo_level_words.argtypes = [vp, u32]
    L.bo_level_ranks.restype = vp
    L.bo_level_ranks.argtypes = [vp, u32]
    L.bo_try_hash.restype = u64
    L.bo_try_hash.argtypes = [vp, u64]
    L.bo_hash_batch.restype = None
    L.bo_hash_batch.argtypes = [vp, vp, u64, vp, i32]
    L.bo_create_map.restype = i32
    L.bo_create_map.argtypes = [vp, vp, vp, u64]
    L.bo_map_get_batch.restype = None
    L.bo_map_get_batch.argtypes = [vp, vp, vp, u64, vp, u64, vp, vp]
    L.bo_mark_level.restype = None
    L.bo_mark_level.argtypes = [vp, u64, u64, u64, vp, vp]
    L.bo_filter_level.restype = u64
    L.bo_filter_level.argtypes = [vp, u64, u64, u64, vp, vp]
    L.bo_keygen.restype = None
    L.bo_keygen.argtypes = [vp, u64, u64, u64, i32]
    L.bo_free.restype = None
    L.bo_free.argtypes = [vp]
    L.bo_max_threads.restype = i32
    L.bo_wy_swap_8byte_tail.restype = i32
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


KIND_U64, KIND_UINT, KIND_BYTES, KIND_STREAM = 0, 1, 2, 3


class Stream:
    """Stream keys (kind 3): the recorded Hasher::write calls of n keys -- `data` all writes back to back (uint8),
    `seg_ends[j]` the end offset of write j, `key_segs[i]..key_segs[i+1]` the writes of key i."""

    def __init__(self, data, seg_ends, key_segs):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.seg_ends = np.ascontiguousarray(seg_ends, dtype=np.uint64)
        self.key_segs = np.ascontiguousarray(key_segs, dtype=np.uint64)

    def __len__(self):
        return len(self.key_segs) - 1

    def use(self):
        lib().bo_set_key_stream(_ptr(self.data), _ptr(self.seg_ends), _ptr(self.key_segs))

    def indices(self):
        return np.arange(len(self), dtype=np.uint64)


def key_kind(keys):
    """(kind, width) of a key array by dtype / shape: u64-like -> (0, 8); u32/u16/u8 (and signed) -> (1, w);
    2-D uint8 (n, N <= 8) = [u8; N] keys -> (2, N)"""
    a = keys if isinstance(keys, np.ndarray) else None
    if a is None:
        return KIND_U64, 8
    if a.ndim == 2:
        if a.dtype != np.uint8 or not 1 <= a.shape[1] <= 8:
            raise TypeError("byte-array keys must be uint8 of shape (n, N <= 8)")
        return KIND_BYTES, int(a.shape[1])
    if a.dtype.itemsize == 8:
        return KIND_U64, 8
    if a.dtype.kind in "iu" and a.dtype.itemsize in (1, 2, 4):
        return KIND_UINT, int(a.dtype.itemsize)
    raise TypeError("unsupported key dtype %s" % a.dtype)


def _as_u64(keys):
    """keys as u64 values holding the key's bytes little-endian in their low bytes"""
    if isinstance(keys, np.ndarray) and keys.ndim == 2:
        out = np.zeros(keys.shape[0], dtype=np.uint64)
        for b in range(keys.shape[1]):
            out |= keys[:, b].astype(np.uint64) << np.uint64(8 * b)
        return out
    if isinstance(keys, np.ndarray) and keys.dtype.itemsize < 8:
        unsigned = keys.view(np.dtype("u%d" % keys.dtype.itemsize))
        return np.ascontiguousarray(unsigned.astype(np.uint64))
    if isinstance(keys, np.ndarray) and keys.dtype == np.int64:
        return np.ascontiguousarray(keys.view(np.uint64))
    a = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64))
    return a


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__({BO_ERR_ARG: "bad argument", BO_ERR_GAMMA: "assertion failed: gamma > 1.01",
                          BO_ERR_MAX_ITERS: "counldn't find unique hashes", BO_ERR_NOMEM: "out of memory"}.get(code, str(code)))
        self.code = code


class OracleMphf:
    """Mirror of `Mphf<u64>` (/root/reference/src/lib.rs:90-96) over the C restatement."""

    def __init__(self, handle, kind=(KIND_U64, 8)):
        self._h = C.c_void_p(handle)
        self.kind = kind

    def _use_kind(self):
        if self.kind[0] == KIND_STREAM:
            raise TypeError("stream-key Mphf: use hash_batch_stream")
        lib().bo_set_key_kind(*self.kind)

    @classmethod
    def new_stream(cls, gamma, stream, parallel=False, starting_seed=None, nthreads=0):
        """Mphf::new / new_parallel for keys given as their Hash byte streams (`Stream`)"""
        idx = stream.indices()
        stream.use()
        out = C.c_void_p()
        try:
            if parallel:
                rc = lib().bo_build_parallel(_ptr(idx), len(idx), float(gamma), 0 if starting_seed is None else 1,
                                             0 if starting_seed is None else int(starting_seed), int(nthreads), C.byref(out))
            else:
                rc = lib().bo_build_serial(_ptr(idx), len(idx), float(gamma), C.byref(out))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value, (KIND_STREAM, 0))

    def hash_batch_stream(self, stream, nthreads=1):
        """try_hash of every key of `stream` (NONE = no level matched)"""
        idx = stream.indices()
        out = np.empty(len(idx), dtype=np.uint64)
        stream.use()
        try:
            lib().bo_hash_batch(self._h, _ptr(idx), len(idx), _ptr(out), int(nthreads))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        return out

    def __del__(self):
        try:
            if self._h:
                lib().bo_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- constructors -------------------------------------------------------------------
    @classmethod
    def new(cls, gamma, keys):
        """Mphf::new (src/lib.rs:235)"""
        kind = key_kind(keys)
        keys = _as_u64(keys)
        lib().bo_set_key_kind(*kind)
        out = C.c_void_p()
        try:
            rc = lib().bo_build_serial(_ptr(keys), len(keys), float(gamma), C.byref(out))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value, kind)

    @classmethod
    def new_parallel(cls, gamma, keys, starting_seed=None, nthreads=0):
        """Mphf::new_parallel (src/lib.rs:363)"""
        kind = key_kind(keys)
        keys = _as_u64(keys)
        lib().bo_set_key_kind(*kind)
        out = C.c_void_p()
        try:
            rc = lib().bo_build_parallel(_ptr(keys), len(keys), float(gamma), 0 if starting_seed is None else 1,
                                         0 if starting_seed is None else int(starting_seed), int(nthreads), C.byref(out))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value, kind)

    @classmethod
    def from_chunked_iterator(cls, gamma, chunks, n):
        """Mphf::from_chunked_iterator(gamma, &chunks, n)  (src/lib.rs:112)"""
        arrs = [_as_u64(c) for c in chunks]
        k = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint64)
        lens = np.array([a.size for a in arrs], dtype=np.uint64)
        out = C.c_void_p()
        rc = lib().bo_build_chunked_serial(_ptr(k), _ptr(lens), lens.size, int(n), float(gamma), C.byref(out))
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value)

    @classmethod
    def from_chunked_iterator_parallel(cls, gamma, chunks, max_iters, n, num_threads=0):
        """Mphf::from_chunked_iterator_parallel(gamma, &chunks, max_iters, n, num_threads)  (src/lib.rs:539)"""
        arrs = [_as_u64(c) for c in chunks]
        k = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint64)
        lens = np.array([a.size for a in arrs], dtype=np.uint64)
        out = C.c_void_p()
        rc = lib().bo_build_chunked_parallel(_ptr(k), _ptr(lens), lens.size, int(n), float(gamma),
                                             0 if max_iters is None else 1, 0 if max_iters is None else int(max_iters),
                                             int(num_threads), C.byref(out))
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value)

    @classmethod
    def from_levels(cls, levels):
        """levels: list of (bits, words ndarray, ranks ndarray) -- the serde data model."""
        n = len(levels)
        bits = np.array([l[0] for l in levels], dtype=np.uint64)
        words = [np.ascontiguousarray(l[1], dtype=np.uint64) for l in levels]
        ranks = [np.ascontiguousarray(l[2], dtype=np.uint64) for l in levels]
        wp = (C.c_void_p * max(n, 1))(*[w.ctypes.data for w in words])
        rp = (C.c_void_p * max(n, 1))(*[r.ctypes.data for r in ranks])
        out = C.c_void_p()
        rc = lib().bo_from_levels(n, _ptr(bits), wp, rp, C.byref(out))
        if rc != BO_OK:
            raise OracleError(rc)
        return cls(out.value)

    # -- data model ---------------------------------------------------------------------
    @property
    def num_levels(self):
        return int(lib().bo_num_levels(self._h))

    def level(self, l):
        """-> (bits, words ndarray u64, ranks ndarray u64) (copies)"""
        b, w, r = C.c_uint64(), C.c_uint64(), C.c_uint64()
        rc = lib().bo_level_dims(self._h, l, C.byref(b), C.byref(w), C.byref(r))
        if rc != BO_OK:
            raise OracleError(rc)
        wp = lib().bo_level_words(self._h, l)
        rp = lib().bo_level_ranks(self._h, l)
        words = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_uint64)), shape=(w.value,)).copy() if w.value else np.zeros(0, np.uint64)
        ranks = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint64)), shape=(r.value,)).copy() if r.value else np.zeros(0, np.uint64)
        return int(b.value), words, ranks

    def levels(self):
        return [self.level(l) for l in range(self.num_levels)]

    # -- queries ------------------------------------------------------------------------
    def try_hash(self, key):
        """Mphf::try_hash (src/lib.rs:340); None when no level matched"""
        self._use_kind()
        try:
            v = int(lib().bo_try_hash(self._h, int(key)))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        return None if v == NONE else v

    def hash(self, key):
        """Mphf::hash (src/lib.rs:324); the reference hits unreachable!() on a miss"""
        v = self.try_hash(key)
        if v is None:
            raise RuntimeError("internal error: entered unreachable code: must find a hash value")
        return v

    def hash_batch(self, keys, nthreads=1):
        keys = _as_u64(keys)
        out = np.empty(len(keys), dtype=np.uint64)
        self._use_kind()
        try:
            lib().bo_hash_batch(self._h, _ptr(keys), len(keys), _ptr(out), int(nthreads))
        finally:
            lib().bo_set_key_kind(KIND_U64, 8)
        return out

    def create_map(self, keys, values):
        """BoomHashMap::create_map (src/hashmap.rs:27-40): returns permuted copies."""
        k = _as_u64(keys).copy()
        v = _as_u64(values).copy()
        rc = lib().bo_create_map(self._h, _ptr(k), _ptr(v), len(k))
        if rc != BO_OK:
            raise OracleError(rc)
        return k, v

    def map_get_batch(self, map_keys, map_values, queries):
        """BoomHashMap::get (src/hashmap.rs:49-66) over a batch -> (values, found)"""
        mk, mv, q = _as_u64(map_keys), _as_u64(map_values), _as_u64(queries)
        vals = np.empty(len(q), dtype=np.uint64)
        found = np.empty(len(q), dtype=np.uint8)
        lib().bo_map_get_batch(self._h, _ptr(mk), _ptr(mv), len(mk), _ptr(q), len(q), _ptr(vals), _ptr(found))
        return vals, found.astype(bool)


def fingerprint(levels):
    """sha256 over per-level little-endian bits || words || ranks."""
    h = hashlib.sha256()
    for bits, words, ranks in levels:
        h.update(np.uint64(bits).tobytes())
        h.update(np.ascontiguousarray(words, dtype="<u8").tobytes())
        h.update(np.ascontiguousarray(ranks, dtype="<u8").tobytes())
    return h.hexdigest()


def serialize_bincode(levels):
    """bincode (fixint, LE) image of the serde data model of Mphf<u64>
    (src/lib.rs:91-96, src/bitvector.rs:41-68): u64 L; per level u64 bits, u64 W, W*u64,
    u64 R, R*u64.  PhantomData serialises to nothing."""
    out = [np.uint64(len(levels)).tobytes()]
    for bits, words, ranks in levels:
        out.append(np.uint64(bits).tobytes())
        out.append(np.uint64(len(words)).tobytes())
        out.append(np.ascontiguousarray(words, dtype="<u8").tobytes())
        out.append(np.uint64(len(ranks)).tobytes())
        out.append(np.ascontiguousarray(ranks, dtype="<u8").tobytes())
    return b"".join(out)


def deserialize_bincode(buf):
    mv = memoryview(buf)
    off = 0

    def rd():
        nonlocal off
        v = int(np.frombuffer(mv[off:off + 8], dtype="<u8")[0])
        off += 8
        return v

    levels = []
    for _ in range(rd()):
        bits = rd()
        w = rd()
        words = np.frombuffer(mv[off:off + 8 * w], dtype="<u8").astype(np.uint64)
        off += 8 * w
        r = rd()
        ranks = np.frombuffer(mv[off:off + 8 * r], dtype="<u8").astype(np.uint64)
        off += 8 * r
        levels.append((bits, words, ranks))
    return levels


def keygen(n, seed=1, kind=0, start=0):
    """distinct-by-construction synthetic keys. kind 0 random u64, 1 62-bit 31-mers, 2 = 2*i"""
    out = np.empty(n, dtype=np.uint64)
    lib().bo_keygen(_ptr(out), int(start), int(n), int(seed), int(kind))
    return out


def hash_stream_with_seed(it, stream, i=0):
    """hash_with_seed(it, &key_i) for a stream key"""
    stream.use()
    try:
        return int(lib().bo_hash_with_seed(int(it), int(i)))
    finally:
        lib().bo_set_key_kind(KIND_U64, 8)


def hash_with_seed(it, key):
    return int(lib().bo_hash_with_seed(int(it), int(key)))


def hashmod(it, key, n):
    return int(lib().bo_hashmod(int(it), int(key), int(n)))


def hash_key_bytes(it, data, prefixed):
    """hash_with_seed(it, key) for a key that writes `data` (with an 8-byte usize length prefix first when
    `prefixed`), computed with the byte-string core alone: an independent cross-check of bo_set_key_kind"""
    h = 1 << ((2 * it) & 63)
    size = 0
    L = lib()
    L.bo_wyhash_core.restype = C.c_uint64
    L.bo_wyhash_core.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
    L.bo_wyhash_finish.restype = C.c_uint64
    L.bo_wyhash_finish.argtypes = [C.c_uint64, C.c_uint64]
    if prefixed:
        h = L.bo_wyhash_core(len(data).to_bytes(8, "little"), 8, h)
        size += 8
    h = L.bo_wyhash_core(bytes(data), len(data), h)
    size += len(data)
    return int(L.bo_wyhash_finish(size, h))


def wyhash_bytes(data, seed):
    b = bytes(data)
    buf = C.create_string_buffer(b, len(b))
    return int(lib().bo_wyhash_bytes(C.cast(buf, C.c_void_p), len(b), int(seed)))


def wyrng(seed):
    s = C.c_uint64(seed)
    return int(lib().bo_wyrng(C.byref(s))), int(s.value)


def level_size(gamma, k):
    return int(lib().bo_level_size(float(gamma), int(k)))


def max_threads():
    return int(lib().bo_max_threads())


def mark_level(keys, seed_index, bits, a, collide):
    """find_collisions of `keys` into the caller's word arrays (numpy uint64, modified in place)"""
    k = _as_u64(keys)
    lib().bo_mark_level(_ptr(k), k.size, int(seed_index), int(bits), _ptr(a), _ptr(collide))


def filter_level(keys, seed_index, bits, collide):
    """keys whose slot collided (the redo list of Context::filter), order preserved"""
    k = _as_u64(keys)
    out = np.empty(k.size, dtype=np.uint64)
    m = lib().bo_filter_level(_ptr(k), k.size, int(seed_index), int(bits), _ptr(collide), _ptr(out))
    return out[:m].copy()


def or_collide(a, c, a2, c2):
    """the merge monoid of SURVEY.md 8e: (a,c) (+) (a',c') = (a|a', c|c'|(a&a'))"""
    return a | a2, c | c2 | (a & a2)


def ranks_from_levels(level_words):
    """Mphf::compute_ranks (src/lib.rs:277-297): one running popcount over all levels, one entry per 8 words"""
    out, pop = [], 0
    for w in level_words:
        pc = np.array([bin(int(x)).count("1") for x in w], dtype=np.uint64) if len(w) < 4096 else \
            np.unpackbits(w.view(np.uint8)).reshape(len(w), 64).sum(axis=1).astype(np.uint64)
        cum = np.concatenate(([0], np.cumsum(pc)))[:-1] + pop
        out.append(cum[::8].astype(np.uint64))
        pop += int(pc.sum())
    return out
