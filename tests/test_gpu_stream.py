"""GPU parity for stream keys (BPHF_KEY_STREAM: any T: Hash as its recorded Hasher::write calls), through the C
ABI against the oracle: bit-identical levels and ranks, identical hash values, the reference's remaining quickcheck
types (Vec<String>, Vec<u8>; /root/reference/src/lib.rs:851-879), the error paths, device-resident tables."""
import ctypes as C
import importlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from stream_cases import STREAM_TYPES, make_keys
from test_gpu_parity import assert_same_levels

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


@pytest.fixture(scope="module")
def K(pkg):
    return importlib.import_module(pkg.__name__ + ".keys")


def ostream(oracle, sk):
    return oracle.Stream(sk.data, sk.seg_ends, sk.key_segs)


@pytest.mark.parametrize("rust_type", STREAM_TYPES)
@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000, 5000, 20000])
def test_stream_build_and_hash_parity(pkg, oracle, K, rust_type, n):
    vals = make_keys(rust_type, n, seed=1)
    sk = K.encode_keys(vals, rust_type)
    m = pkg.Mphf.new(1.7, sk)
    ref = oracle.OracleMphf.new_stream(1.7, ostream(oracle, sk))
    assert m.key_kind == (3, 0)
    assert_same_levels(m.levels(), ref.levels())
    if n:
        h = m.hash_batch(sk)
        assert np.array_equal(h, ref.hash_batch_stream(ostream(oracle, sk)))
        assert np.array_equal(np.sort(h), np.arange(n, dtype=np.uint64))
    # absent keys: whatever the oracle says (None or an arbitrary index), the device says too
    other = K.encode_keys(make_keys(rust_type, 300, seed=77), rust_type)
    assert np.array_equal(m.hash_batch(other), ref.hash_batch_stream(ostream(oracle, other)))


@pytest.mark.parametrize("gamma", [1.02, 1.7, 2.0, 5.0])
@pytest.mark.parametrize("seed", [None, 0, 3, 40])
def test_stream_gamma_and_starting_seed(pkg, oracle, K, gamma, seed):
    vals = make_keys("Vec<String>", 8000, seed=2)
    sk = K.encode_keys(vals, "Vec<String>")
    m = pkg.Mphf.new_parallel(gamma, sk, seed)
    ref = oracle.OracleMphf.new_stream(gamma, ostream(oracle, sk), parallel=True, starting_seed=seed)
    assert_same_levels(m.levels(), ref.levels())


def test_stream_large_set(pkg, oracle, K):
    """300k byte-string keys: several full-size levels through the StreamCfg level kernels before the tail"""
    rng = np.random.default_rng(5)
    n = 300_000
    lens = rng.integers(0, 48, size=n)
    body = rng.integers(0, 256, size=int(lens.sum()), dtype=np.uint8).tobytes()
    vals, pos = [], 0
    for i, l in enumerate(lens.tolist()):
        vals.append(body[pos:pos + l] + i.to_bytes(4, "little"))  # distinct by construction
        pos += l
    sk = K.encode_keys(vals, "Vec<u8>")
    m = pkg.Mphf.new(1.7, sk)
    ref = oracle.OracleMphf.new_stream(1.7, ostream(oracle, sk), parallel=True)
    assert_same_levels(m.levels(), ref.levels())
    h = m.hash_batch(sk)
    assert np.array_equal(h, ref.hash_batch_stream(ostream(oracle, sk), nthreads=0))
    assert np.array_equal(np.sort(h), np.arange(n, dtype=np.uint64))


@pytest.mark.parametrize("rust_type,as_array", [
    ("u64", lambda v: np.array(v, dtype=np.uint64)),
    ("u32", lambda v: np.array(v, dtype=np.uint32)),
    ("[u8; 5]", lambda v: np.frombuffer(b"".join(v), dtype=np.uint8).reshape(-1, 5)),
])
def test_stream_and_word_kinds_agree_on_the_device(pkg, K, rust_type, as_array):
    vals = make_keys(rust_type, 30000, seed=5)
    a = pkg.Mphf.new(1.7, as_array(vals))
    b = pkg.Mphf.new(1.7, vals, rust_type=rust_type)
    assert_same_levels(a.levels(), b.levels())
    assert np.array_equal(a.hash_batch(as_array(vals)), b.hash_batch(vals))


def test_plain_python_collections_and_single_items(pkg, K):
    words = ["alpha", "beta", "gamma", "delta", "", "ε"]
    m = pkg.Mphf.new(1.7, words)                       # list of str -> Mphf<String>
    assert m.rust_type == "String"
    assert sorted(m.hash(w) for w in words) == list(range(len(words)))
    assert sorted(int(v) for v in m.hash_batch(words)) == list(range(len(words)))
    blobs = [b"", b"\x00", b"\x00\x00", b"0123456789abcdef0123456789abcdef!", b"k"]
    mb = pkg.Mphf.new_parallel(2.0, blobs, None)       # list of bytes -> Mphf<Vec<u8>>
    assert sorted(mb.hash(b) for b in blobs) == list(range(len(blobs)))
    rows = [["a", "b"], ["ab"], ["a", "", "b"], ["b", "a"]]
    mr = pkg.Mphf.new(1.7, rows)                       # list of lists of str -> Mphf<Vec<String>>
    assert sorted(mr.hash(r) for r in rows) == list(range(len(rows)))
    big = pkg.Mphf.new(1.7, [1, 2, 1 << 127], rust_type="u128")
    assert sorted(big.hash(v) for v in (1, 2, 1 << 127)) == [0, 1, 2]


def test_stream_levels_round_trip(pkg, oracle, K):
    vals = make_keys("(u32, String)", 4000, seed=8)
    m = pkg.Mphf.new(1.7, vals, rust_type="(u32, String)")
    blob = oracle.serialize_bincode(m.levels())
    m2 = pkg.Mphf.from_levels(oracle.deserialize_bincode(blob), rust_type="(u32, String)")
    assert m2.key_kind == (3, 0)
    assert np.array_equal(m.hash_batch(vals), m2.hash_batch(vals))
    o = oracle.OracleMphf.from_levels(m.levels())
    o.kind = (oracle.KIND_STREAM, 0)
    sk = K.encode_keys(vals, "(u32, String)")
    assert np.array_equal(o.hash_batch_stream(ostream(oracle, sk)), m.hash_batch(sk))


def test_stream_duplicates_and_bad_arguments(pkg, K):
    N = importlib.import_module(pkg.__name__ + "._native")
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new(1.7, ["same", "same", "other"])
    assert e.value.code == N.BPHF_ERR_MAX_ITERS and "counldn't find unique hashes" in str(e.value)
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new(1.0, ["a", "b"])
    assert e.value.code == N.BPHF_ERR_GAMMA
    good = K.encode_keys([b"ab", b"cde"], "Vec<u8>")
    for seg_ends, key_segs in (([8, 10, 18, 99], [0, 2, 4]),      # write past the byte array
                               ([8, 10, 9, 21], [0, 2, 4]),       # not monotonic
                               ([8, 10, 18, 21], [0, 3, 2]),      # key table not monotonic / does not end at n_segs
                               ([8, 10, 18, 21], [1, 2, 4])):     # does not start at 0
        bad = K.StreamKeys(good.data, seg_ends, key_segs)
        with pytest.raises(pkg.MphfPanic) as e:
            pkg.Mphf.new(1.7, bad)
        assert e.value.code == N.BPHF_ERR_ARG
    m = pkg.Mphf.new(1.7, good)
    with pytest.raises(pkg.MphfPanic):                             # word keys into a stream handle
        out = np.empty(2, dtype=np.uint64)
        rc = N.lib().bphf_hash_batch_u64(m._h, C.c_void_p(out.ctypes.data), 2, C.c_void_p(out.ctypes.data), N.MEM_HOST, None)
        if rc != N.BPHF_OK:
            raise pkg.MphfPanic(rc, N.last_error())
    mu = pkg.Mphf.new(1.7, np.arange(100, dtype=np.uint64))
    out = np.empty(2, dtype=np.uint64)
    rc = N.lib().bphf_hash_batch_stream(mu._h, C.c_void_p(good.data.ctypes.data), good.data.size,
                                        C.c_void_p(good.seg_ends.ctypes.data), good.seg_ends.size,
                                        C.c_void_p(good.key_segs.ctypes.data), 2, C.c_void_p(out.ctypes.data), N.MEM_HOST, None)
    assert rc == N.BPHF_ERR_ARG                                    # stream keys into a u64 handle


def test_stream_tables_resident_on_the_device(pkg, oracle, K):
    import torch
    N = importlib.import_module(pkg.__name__ + "._native")
    vals = make_keys("Vec<u8>", 10000, seed=4)
    sk = K.encode_keys(vals, "Vec<u8>")
    d = torch.from_numpy(sk.data.copy()).cuda()
    se = torch.from_numpy(sk.seg_ends.view(np.int64).copy()).cuda()
    ks = torch.from_numpy(sk.key_segs.view(np.int64).copy()).cuda()
    out = C.c_void_p()
    rc = N.lib().bphf_build_stream(C.c_void_p(d.data_ptr()), d.numel(), C.c_void_p(se.data_ptr()), se.numel(),
                                   C.c_void_p(ks.data_ptr()), len(sk), 1.7, 0, 0, 0, N.MEM_DEVICE, C.byref(out))
    assert rc == N.BPHF_OK, N.last_error()
    m = pkg.Mphf(out.value, 0, "Vec<u8>")
    ref = oracle.OracleMphf.new_stream(1.7, ostream(oracle, sk))
    assert_same_levels(m.levels(), ref.levels())
    res = torch.empty(len(sk), dtype=torch.int64, device="cuda")
    rc = N.lib().bphf_hash_batch_stream(m._h, C.c_void_p(d.data_ptr()), d.numel(), C.c_void_p(se.data_ptr()), se.numel(),
                                        C.c_void_p(ks.data_ptr()), len(sk), C.c_void_p(res.data_ptr()), N.MEM_DEVICE, None)
    assert rc == N.BPHF_OK, N.last_error()
    assert np.array_equal(res.cpu().numpy().view(np.uint64), ref.hash_batch_stream(ostream(oracle, sk)))


@settings(max_examples=25, deadline=None)
@given(st.sets(st.lists(st.text(max_size=6), max_size=4).map(tuple), max_size=60))
def test_quickcheck_check_string_gpu(pkg, oracle, K, v):
    """check_string (src/lib.rs:851): HashSet<Vec<String>> -> a permutation, and the oracle's levels"""
    vals = [list(x) for x in v]
    sk = K.encode_keys(vals, "Vec<String>")
    m = pkg.Mphf.new(1.7, sk)
    assert sorted(m.hash_batch(sk).tolist()) == list(range(len(vals)))
    assert_same_levels(m.levels(), oracle.OracleMphf.new_stream(1.7, ostream(oracle, sk)).levels())


@settings(max_examples=25, deadline=None)
@given(st.sets(st.binary(max_size=70), max_size=80))
def test_quickcheck_check_vec_u8_gpu(pkg, oracle, K, v):
    """check_vec_u8 (src/lib.rs:875): HashSet<Vec<u8>>"""
    vals = list(v)
    sk = K.encode_keys(vals, "Vec<u8>")
    m = pkg.Mphf.new_parallel(1.7, sk, None)
    assert sorted(m.hash_batch(sk).tolist()) == list(range(len(vals)))
    assert_same_levels(m.levels(), oracle.OracleMphf.new_stream(1.7, ostream(oracle, sk), parallel=True).levels())


@pytest.mark.parametrize("entry", GOLDEN["stream_keys"], ids=[e["rust_type"] for e in GOLDEN["stream_keys"]])
def test_golden_stream_fingerprints_gpu(pkg, oracle, K, entry):
    sk = K.encode_keys(make_keys(entry["rust_type"], entry["n"], seed=entry["keys_seed"]), entry["rust_type"])
    m = pkg.Mphf.new(entry["gamma"], sk)
    assert [l[0] for l in m.levels()] == entry["bits"] and oracle.fingerprint(m.levels()) == entry["sha256"]
    assert [int(v) for v in m.hash_batch(sk)[:4]] == entry["hash_first"]


def test_boomhashmap_with_string_keys(pkg):
    """BoomHashMap<String, _> (src/hashmap.rs:16-132) over stream keys: every key finds its value, absent keys None"""
    keys = make_keys("String", 5000, seed=12)
    vals = [("v", i) for i in range(len(keys))]
    m = pkg.BoomHashMapAny.new(keys, vals)
    assert len(m) == len(keys) and not m.is_empty()
    assert m.get_batch(keys) == vals
    assert sorted(m.get_key_id_batch(keys)) == list(range(len(keys)))
    absent = [k + "\x00absent" for k in keys[:200]]
    assert m.get_batch(absent) == [None] * 200
    assert m.get(keys[7]) == vals[7] and m.get("no such key, surely") is None
    assert m.get_key(m.get_key_id(keys[3])) == keys[3]
    assert dict(m.iter()) == dict(zip(keys, vals))
    mb = pkg.BoomHashMapAny.new([b"a", b"bb", b""], [1, 2, 3], rust_type="Vec<u8>")
    assert [mb.get(b"a"), mb.get(b"bb"), mb.get(b""), mb.get(b"c")] == [1, 2, 3, None]
