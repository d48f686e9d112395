"""Host-side mirror of the reference's `hashmap` module for u64 keys and u64 values
(/root/reference/src/hashmap.rs): BoomHashMap (:16-132) and NoKeyBoomHashMap (:434-523).

`create_map`'s in-place cycle permutation (:27-40) becomes an out-of-place scatter on the GPU
(keys_out[hash(k)] = k); the resulting arrays are identical because the permutation is unique.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .mphf import Mphf, _check


def _u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


class BoomHashMap:
    """BoomHashMap<u64, u64>: keys and values stored densely in MPHF order."""

    GAMMA = 1.7  # hard-coded by the reference (src/hashmap.rs:44, :171)

    def __init__(self, mphf, keys, values):
        self.mphf, self.keys, self.values = mphf, keys, values

    @classmethod
    def _create_map(cls, keys, values, mphf):
        keys, values = _u64(keys), _u64(values)
        if len(keys) != len(values):
            raise ValueError("keys and values differ in length")
        ko, vo = np.empty_like(keys), np.empty_like(values)
        _check(N.lib().bphf_map_permute_u64(mphf._h, C.c_void_p(keys.ctypes.data), C.c_void_p(values.ctypes.data),
                                            len(keys), C.c_void_p(ko.ctypes.data), C.c_void_p(vo.ctypes.data),
                                            N.MEM_HOST, None))
        return cls(mphf, ko, vo)

    @classmethod
    def new(cls, keys, data, device=0):
        """BoomHashMap::new (src/hashmap.rs:43-46)"""
        return cls._create_map(keys, data, Mphf.new(cls.GAMMA, _u64(keys), device=device))

    @classmethod
    def new_parallel(cls, keys, data, device=0):
        """BoomHashMap::new_parallel (src/hashmap.rs:170-173)"""
        return cls._create_map(keys, data, Mphf.new_parallel(cls.GAMMA, _u64(keys), None, device=device))

    @classmethod
    def new_with_mphf(cls, mphf, keys, data):
        """the gamma sweep of BASELINE config 5 builds the Mphf directly and wraps it"""
        return cls._create_map(keys, data, mphf)

    def get_batch(self, queries):
        """BoomHashMap::get over a batch (src/hashmap.rs:49-66) -> (values, found mask)"""
        q = _u64(queries)
        vals = np.empty(len(q), dtype=np.uint64)
        found = np.empty(len(q), dtype=np.uint8)
        _check(N.lib().bphf_map_get_batch_u64(self.mphf._h, C.c_void_p(self.keys.ctypes.data),
                                              C.c_void_p(self.values.ctypes.data), len(self.keys),
                                              C.c_void_p(q.ctypes.data), len(q), C.c_void_p(vals.ctypes.data),
                                              C.c_void_p(found.ctypes.data), N.MEM_HOST, None))
        return vals, found.astype(bool)

    def get(self, key):
        vals, found = self.get_batch([key])
        return int(vals[0]) if found[0] else None

    def get_key_id(self, key):
        """src/hashmap.rs:88-105"""
        q = _u64([key])
        vals = np.empty(1, dtype=np.uint64)
        found = np.empty(1, dtype=np.uint8)
        _check(N.lib().bphf_map_get_batch_u64(self.mphf._h, C.c_void_p(self.keys.ctypes.data), None, len(self.keys),
                                              C.c_void_p(q.ctypes.data), 1, C.c_void_p(vals.ctypes.data),
                                              C.c_void_p(found.ctypes.data), N.MEM_HOST, None))
        return int(vals[0]) if found[0] else None

    def get_key(self, id):
        """src/hashmap.rs:117-124 (keeps the reference's `id > len` bound check)"""
        if id > len(self.keys):
            return None
        return int(self.keys[id])  # id == len raises IndexError like the reference panics

    def __len__(self):
        return len(self.keys)

    def is_empty(self):
        return len(self.keys) == 0

    def iter(self):
        return zip(self.keys.tolist(), self.values.tolist())


class NoKeyBoomHashMap:
    """NoKeyBoomHashMap<u64, u64> (src/hashmap.rs:434-523): values only, no key check on get."""

    def __init__(self, mphf, values):
        self.mphf, self.values = mphf, values

    @classmethod
    def new_with_mphf(cls, mphf, values):
        """src/hashmap.rs:495-497 -- values must already be in MPHF order"""
        return cls(mphf, _u64(values))

    @classmethod
    def new_parallel(cls, keys, data, device=0):
        """src/hashmap.rs:489-493"""
        m = BoomHashMap.new_parallel(keys, data, device=device)
        return cls(m.mphf, m.values)

    new = new_parallel

    def get_batch(self, queries):
        """src/hashmap.rs:500-510: Some(values[pos]) whenever try_hash is Some"""
        pos = self.mphf.hash_batch(_u64(queries))
        ok = pos < np.uint64(len(self.values))
        out = np.zeros(len(pos), dtype=np.uint64)
        out[ok] = self.values[pos[ok].astype(np.int64)]
        return out, ok


class BoomHashMap2:
    """BoomHashMap2<u64, u64, u64> (src/hashmap.rs:222-394): keys plus two value arrays in MPHF order.
    create_map (:277-302) swaps all three arrays along the same permutation; here two out-of-place scatters."""

    GAMMA = 1.7  # src/hashmap.rs:306, :425

    def __init__(self, mphf, keys, values, aux_values):
        self.mphf, self.keys, self.values, self.aux_values = mphf, keys, values, aux_values

    @classmethod
    def _create_map(cls, keys, values, aux_values, mphf):
        keys, values, aux = _u64(keys), _u64(values), _u64(aux_values)
        if not (len(keys) == len(values) == len(aux)):
            raise ValueError("keys, values and aux_values differ in length")
        a = BoomHashMap._create_map(keys, values, mphf)
        b = BoomHashMap._create_map(keys, aux, mphf)
        return cls(mphf, a.keys, a.values, b.values)

    @classmethod
    def new(cls, keys, values, aux_values, device=0):
        """BoomHashMap2::new (src/hashmap.rs:305-308)"""
        return cls._create_map(keys, values, aux_values, Mphf.new(cls.GAMMA, _u64(keys), device=device))

    @classmethod
    def new_parallel(cls, keys, values, aux_values, device=0):
        """BoomHashMap2::new_parallel (src/hashmap.rs:424-427)"""
        return cls._create_map(keys, values, aux_values, Mphf.new_parallel(cls.GAMMA, _u64(keys), None, device=device))

    def get_batch(self, queries):
        """BoomHashMap2::get over a batch (src/hashmap.rs:310-328) -> (values, aux_values, found mask)"""
        one = BoomHashMap(self.mphf, self.keys, self.values)
        two = BoomHashMap(self.mphf, self.keys, self.aux_values)
        v, f = one.get_batch(queries)
        a, _ = two.get_batch(queries)
        return v, a, f

    def get(self, key):
        v, a, f = self.get_batch([key])
        return (int(v[0]), int(a[0])) if f[0] else None

    def get_key_id(self, key):
        """src/hashmap.rs:351-368"""
        return BoomHashMap(self.mphf, self.keys, self.values).get_key_id(key)

    def __len__(self):
        return len(self.keys)

    def is_empty(self):
        return len(self.keys) == 0

    def iter(self):
        return zip(self.keys.tolist(), self.values.tolist(), self.aux_values.tolist())


class NoKeyBoomHashMap2:
    """NoKeyBoomHashMap2<u64, u64, u64> (src/hashmap.rs:530-641): two value arrays, no key check on get."""

    def __init__(self, mphf, values, aux_values):
        self.mphf, self.values, self.aux_values = mphf, values, aux_values

    @classmethod
    def new_parallel(cls, keys, values, aux_values, device=0):
        """src/hashmap.rs:600-605"""
        m = BoomHashMap2.new_parallel(keys, values, aux_values, device=device)
        return cls(m.mphf, m.values, m.aux_values)

    new = new_parallel

    @classmethod
    def new_with_mphf(cls, mphf, values, aux_values):
        """src/hashmap.rs:608-614 -- arrays must already be in MPHF order"""
        return cls(mphf, _u64(values), _u64(aux_values))

    def get_batch(self, queries):
        """src/hashmap.rs:617-627: Some((values[pos], aux_values[pos])) whenever try_hash is Some"""
        pos = self.mphf.hash_batch(_u64(queries))
        ok = pos < np.uint64(len(self.values))
        v = np.zeros(len(pos), dtype=np.uint64)
        a = np.zeros(len(pos), dtype=np.uint64)
        idx = pos[ok].astype(np.int64)
        v[ok] = self.values[idx]
        a[ok] = self.aux_values[idx]
        return v, a, ok


class BoomHashMapAny:
    """BoomHashMap<K, D> for any other K: Hash (String, Vec<u8>, Vec<String>, u128, tuples, ... -- stream keys,
    keys.py): the Mphf is built and queried on the GPU; keys and values are host-side Python lists held in MPHF order
    (src/hashmap.rs:16-132).  `get` checks the stored key like the reference (:49-66)."""

    GAMMA = 1.7  # src/hashmap.rs:44, :171

    def __init__(self, mphf, keys, values):
        self.mphf, self.keys, self.values = mphf, keys, values

    @classmethod
    def new(cls, keys, data, rust_type=None, device=0):
        """BoomHashMap::new (src/hashmap.rs:43-46) / new_parallel (:170-173) -- the same structure either way"""
        keys, data = list(keys), list(data)
        if len(keys) != len(data):
            raise ValueError("keys and values differ in length")
        from .keys import encode_keys
        sk = encode_keys(keys, rust_type)
        mphf = Mphf.new(cls.GAMMA, sk, device=device)
        pos = mphf.hash_batch(sk)
        ko, vo = [None] * len(keys), [None] * len(keys)
        for i, p in enumerate(pos.tolist()):  # create_map (:27-40): element i goes to slot hash(key i)
            ko[p], vo[p] = keys[i], data[i]
        return cls(mphf, ko, vo)

    new_parallel = new

    def get_key_id_batch(self, queries):
        """positions of the queries that are keys of the map, None for the others (src/hashmap.rs:88-105)"""
        queries = list(queries)
        if not queries:
            return []
        pos = self.mphf.hash_batch(queries).tolist()
        n = len(self.keys)
        return [p if p < n and self.keys[p] == q else None for p, q in zip(pos, queries)]

    def get_batch(self, queries):
        return [None if p is None else self.values[p] for p in self.get_key_id_batch(queries)]

    def get(self, key):
        return self.get_batch([key])[0]

    def get_key_id(self, key):
        return self.get_key_id_batch([key])[0]

    def get_key(self, id):
        """src/hashmap.rs:117-124"""
        if id > len(self.keys):
            return None
        return self.keys[id]

    def __len__(self):
        return len(self.keys)

    def is_empty(self):
        return len(self.keys) == 0

    def iter(self):
        return zip(self.keys, self.values)
