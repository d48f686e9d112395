"""Stream keys (any T: Hash as its recorded Hasher::write calls) -- CPU side: the recorder stand-in
(rust-boomphf_b200/keys.py) against hand-written byte streams, and the oracle's stream hash against its own
single-word key kinds, a pure-Python restatement of the wyhash v1 core, and the MPHF permutation property for the
reference's remaining quickcheck types (Vec<String>, Vec<u8>; /root/reference/src/lib.rs:851-879)."""
import importlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from stream_cases import STREAM_TYPES, make_keys

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


@pytest.fixture(scope="module")
def K():
    import __graft_entry__ as ge
    p = ge.load_package()
    return importlib.import_module(p.__name__ + ".keys")


def ostream(oracle, sk):
    return oracle.Stream(sk.data, sk.seg_ends, sk.key_segs)


# ---- the recorder stand-in ------------------------------------------------------------------------------
def test_encoder_spells_out_the_std_hash_impls(K):
    le8 = lambda v: int(v).to_bytes(8, "little")
    assert K.encode_keys([["ab", "c"], []], "Vec<String>").writes_of(0) == [le8(2), b"ab", b"\xff", b"c", b"\xff"]
    assert K.encode_keys([["ab", "c"], []], "Vec<String>").writes_of(1) == [le8(0)]
    assert K.encode_keys([b"xyz", b""]).writes_of(0) == [le8(3), b"xyz"]            # Vec<u8>: prefix + ONE write
    assert K.encode_keys([b"xyz", b""]).writes_of(1) == [le8(0), b""]              # ... even when it is empty
    assert K.encode_keys(["hé"]).writes_of(0) == ["hé".encode(), b"\xff"]          # String: bytes, 0xff
    assert K.encode_keys([1 << 100], "u128").writes_of(0) == [(1 << 100).to_bytes(16, "little")]
    assert K.encode_keys([-2], "i128").writes_of(0) == [b"\xfe" + b"\xff" * 15]
    assert K.encode_keys([[1, 2]], "[u32; 2]").writes_of(0) == [le8(2), b"\x01\0\0\0\x02\0\0\0"]   # array = slice
    assert K.encode_keys([[True, False]], "Vec<bool>").writes_of(0) == [le8(2), b"\x01", b"\x00"]  # per element
    assert K.encode_keys([(7, "k", -1)], "(u16, &str, i8)").writes_of(0) == [b"\x07\0", b"k", b"\xff", b"\xff"]
    assert K.encode_keys([("a",)], "(char,)").writes_of(0) == [b"a\0\0\0"]
    assert K.encode_keys([5], "&'a Box<usize>").writes_of(0) == [le8(5)]
    sk = K.encode_keys([b"ab", b"cde"], "Vec<u8>")
    assert list(sk.seg_ends) == [8, 10, 18, 21] and list(sk.key_segs) == [0, 2, 4] and len(sk) == 2
    with pytest.raises(TypeError):
        K.encode_keys([1.5], "f64")            # f64 is not Hash
    with pytest.raises(TypeError):
        K.encode_keys([{1: 2}])


# ---- the oracle's stream hash ---------------------------------------------------------------------------
P0, P1, P2, P3, P4, P5 = (0xa0761d6478bd642f, 0xe7037ed1a0b428db, 0x8ebc6af09c88c6e3, 0x589965cc75374cc3,
                          0x1d8e4e27c47d124f, 0xeb44accab455d165)
M64 = (1 << 64) - 1


def _mum(a, b):
    r = (a & M64) * (b & M64)
    return (r >> 64) ^ (r & M64)


def _rd(b):
    return int.from_bytes(b, "little")


def _swapped(b):
    return (_rd(b[:4]) << 32) | _rd(b[4:8])


def _rest(b):  # read_rest of the wyhash crate, 1..8 bytes
    n = len(b)
    if n in (1, 2, 4):
        return _rd(b)
    if n == 3:
        return (_rd(b[:2]) << 8) | b[2]
    if n == 5:
        return (_rd(b[:4]) << 8) | b[4]
    if n == 6:
        return (_rd(b[:4]) << 16) | _rd(b[4:6])
    if n == 7:
        return (_rd(b[:4]) << 24) | (_rd(b[4:6]) << 8) | b[6]
    return _swapped(b)


def py_core(b, seed):
    """pure-Python wyhash v1 core (one Hasher::write), written from the published algorithm"""
    i = 0
    while i + 32 <= len(b):
        c = b[i:i + 32]
        seed = _mum(seed ^ P0, _mum(_rd(c[0:8]) ^ P1, _rd(c[8:16]) ^ P2) ^ _mum(_rd(c[16:24]) ^ P3, _rd(c[24:32]) ^ P4))
        i += 32
    seed ^= P0
    t = b[i:]
    if t:
        q = (len(t) - 1) // 8
        if q == 0:
            seed = _mum(seed, _rest(t) ^ P1)
        elif q == 1:
            seed = _mum(_swapped(t[0:8]) ^ seed, _rest(t[8:]) ^ P2)
        elif q == 2:
            seed = _mum(_swapped(t[0:8]) ^ seed, _swapped(t[8:16]) ^ P2) ^ _mum(seed, _rest(t[16:]) ^ P3)
        else:
            seed = _mum(_swapped(t[0:8]) ^ seed, _swapped(t[8:16]) ^ P2) ^ _mum(_swapped(t[16:24]) ^ seed, _rest(t[24:]) ^ P4)
    return seed


def py_stream_hash(it, writes):
    h, size = 1 << ((2 * it) & 63), 0
    for w in writes:
        if len(w):                      # a zero-length write leaves the hasher alone
            h = py_core(w, h)
        size += len(w)
    return _mum(h, size ^ P5)


def test_py_core_reproduces_the_readme_known_answer(oracle):
    assert _mum(py_core(bytes([0, 1, 2]), 3), 3 ^ P5) == 0xb0f941520b1ad95d
    for n in list(range(0, 70)) + [95, 96, 97, 128, 1000]:
        data = bytes((i * 37 + n) & 0xff for i in range(n))
        assert _mum(py_core(data, 12345), n ^ P5) == oracle.wyhash_bytes(data, 12345), n


@pytest.mark.parametrize("rust_type", STREAM_TYPES)
def test_oracle_stream_hash_matches_the_python_restatement(oracle, K, rust_type):
    vals = make_keys(rust_type, 60, seed=3)
    sk = K.encode_keys(vals, rust_type)
    os_ = ostream(oracle, sk)
    for i in range(len(vals)):
        for it in (0, 1, 5, 31, 32, 40):
            assert oracle.hash_stream_with_seed(it, os_, i) == py_stream_hash(it, sk.writes_of(i)), (rust_type, i, it)


@pytest.mark.parametrize("rust_type,as_array", [
    ("u64", lambda v: np.array(v, dtype=np.uint64)),
    ("u32", lambda v: np.array(v, dtype=np.uint32)),
    ("[u8; 5]", lambda v: np.frombuffer(b"".join(v), dtype=np.uint8).reshape(-1, 5)),
])
def test_stream_and_word_kinds_agree(oracle, K, rust_type, as_array):
    """the same keys through the one-word kinds (0/1/2) and as recorded streams: identical MPHF"""
    vals = make_keys(rust_type, 3000, seed=5)
    a = oracle.OracleMphf.new(1.7, as_array(vals))
    b = oracle.OracleMphf.new_stream(1.7, ostream(oracle, K.encode_keys(vals, rust_type)))
    assert oracle.fingerprint(a.levels()) == oracle.fingerprint(b.levels())


@pytest.mark.parametrize("rust_type", STREAM_TYPES)
@pytest.mark.parametrize("n", [0, 1, 2, 500])
def test_oracle_stream_mphf_is_a_permutation(oracle, K, rust_type, n):
    vals = make_keys(rust_type, n, seed=9)
    s = ostream(oracle, K.encode_keys(vals, rust_type))
    m = oracle.OracleMphf.new_stream(1.7, s)
    mp = oracle.OracleMphf.new_stream(1.7, s, parallel=True)
    assert oracle.fingerprint(m.levels()) == oracle.fingerprint(mp.levels())  # new == new_parallel(None)
    assert sorted(m.hash_batch_stream(s).tolist()) == list(range(n))


@settings(max_examples=40, deadline=None)
@given(st.sets(st.lists(st.text(max_size=6), max_size=4).map(tuple), max_size=60))
def test_quickcheck_check_string(oracle, K, v):
    """check_string (src/lib.rs:851): HashSet<Vec<String>>"""
    vals = [list(x) for x in v]
    s = ostream(oracle, K.encode_keys(vals, "Vec<String>"))
    m = oracle.OracleMphf.new_stream(1.7, s)
    assert sorted(m.hash_batch_stream(s).tolist()) == list(range(len(vals)))


@settings(max_examples=40, deadline=None)
@given(st.sets(st.binary(max_size=70), max_size=80))
def test_quickcheck_check_vec_u8(oracle, K, v):
    """check_vec_u8 (src/lib.rs:875): HashSet<Vec<u8>>"""
    vals = list(v)
    s = ostream(oracle, K.encode_keys(vals, "Vec<u8>"))
    m = oracle.OracleMphf.new_stream(1.7, s)
    assert sorted(m.hash_batch_stream(s).tolist()) == list(range(len(vals)))


# ---- committed golden vectors (tests/golden/make_golden.py) ------------------------------------------------
def test_golden_stream_hashes(oracle):
    for e in GOLDEN["stream_hashes"]:
        w = [bytes.fromhex(x) for x in e["writes"]]
        ends = np.cumsum([len(x) for x in w]).astype(np.uint64)
        s = oracle.Stream(np.frombuffer(b"".join(w), dtype=np.uint8), ends, [0, len(w)])
        for it, h in e["h"].items():
            assert "%016x" % oracle.hash_stream_with_seed(int(it), s) == h
            assert "%016x" % py_stream_hash(int(it), w) == h


@pytest.mark.parametrize("entry", GOLDEN["stream_keys"], ids=[e["rust_type"] for e in GOLDEN["stream_keys"]])
def test_golden_stream_fingerprints(oracle, K, entry):
    sk = K.encode_keys(make_keys(entry["rust_type"], entry["n"], seed=entry["keys_seed"]), entry["rust_type"])
    s = ostream(oracle, sk)
    m = oracle.OracleMphf.new_stream(entry["gamma"], s, parallel=True)
    assert [l[0] for l in m.levels()] == entry["bits"] and oracle.fingerprint(m.levels()) == entry["sha256"]
    assert [int(v) for v in m.hash_batch_stream(s)[:4]] == entry["hash_first"]


@settings(max_examples=200, deadline=None)
@given(st.lists(st.binary(max_size=80), min_size=0, max_size=6), st.integers(min_value=0, max_value=70))
def test_oracle_stream_hash_on_arbitrary_write_sequences(oracle, writes, it):
    """any sequence of Hasher::write calls (empty writes included): C oracle == the Python restatement above"""
    ends = np.cumsum([len(w) for w in writes]).astype(np.uint64) if writes else np.zeros(0, np.uint64)
    data = np.frombuffer(b"".join(writes), dtype=np.uint8) if writes else np.zeros(0, np.uint8)
    s = oracle.Stream(data, ends, [0, len(writes)])
    assert oracle.hash_stream_with_seed(it, s) == py_stream_hash(it, writes)
