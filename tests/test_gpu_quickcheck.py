"""The reference's own property tests (quickcheck, /root/reference/src/lib.rs:702-879), restated with hypothesis and run
through the C ABI on the GPU: for random key sets, `Mphf::new`, `Mphf::new_parallel`, `from_chunked_iterator` and
`from_chunked_iterator_parallel` (2 threads) must map the keys onto exactly 0..n-1 -- and, beyond what the reference
asserts, agree with the CPU oracle bit for bit."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_gpu_parity import assert_same_levels

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))


def check_mphf(pkg, oracle, keys):
    """check_mphf (src/lib.rs:702-710): serial and parallel builds, hashes sorted == 0..n"""
    ref = oracle.OracleMphf.new(1.7, keys)
    for build in (lambda: pkg.Mphf.new(1.7, keys), lambda: pkg.Mphf.new_parallel(1.7, keys, None)):
        m = build()
        if len(keys):
            h = m.hash_batch(keys)
            assert np.array_equal(np.sort(h), np.arange(len(keys), dtype=np.uint64))
        assert_same_levels(m.levels(), ref.levels())


def check_chunked_mphf(pkg, oracle, chunks, n):
    """check_chunked_mphf (src/lib.rs:757-794): from_chunked_iterator and _parallel(.., 2 threads)"""
    allk = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint64)
    a = pkg.Mphf.from_chunked_iterator(1.7, chunks, n)
    b = pkg.Mphf.from_chunked_iterator_parallel(1.7, chunks, None, n, 2)
    assert_same_levels(a.levels(), oracle.OracleMphf.from_chunked_iterator(1.7, chunks, n).levels())
    assert_same_levels(b.levels(), oracle.OracleMphf.from_chunked_iterator_parallel(1.7, chunks, None, n).levels())
    if n:
        for m in (a, b):
            assert np.array_equal(np.sort(m.hash_batch(allk)), np.arange(n, dtype=np.uint64))


@settings(**SETTINGS)
@given(st.sets(st.integers(0, 2**64 - 1), max_size=3000))
def test_check_u64(pkg, oracle, v):       # src/lib.rs:869-873
    check_mphf(pkg, oracle, np.array(sorted(v), dtype=np.uint64))


@settings(**SETTINGS)
@given(st.sets(st.integers(0, 2**32 - 1), max_size=3000))
def test_check_u32(pkg, oracle, v):       # src/lib.rs:857-861
    check_mphf(pkg, oracle, np.array(sorted(v), dtype=np.uint32))


@settings(**SETTINGS)
@given(st.sets(st.integers(-2**63, 2**63 - 1), max_size=3000))
def test_check_isize(pkg, oracle, v):     # src/lib.rs:863-867
    check_mphf(pkg, oracle, np.array(sorted(v), dtype=np.int64))


@settings(**SETTINGS)
@given(st.integers(1, 8).flatmap(lambda w: st.sets(st.binary(min_size=w, max_size=w), max_size=1500)))
def test_check_vec_u8_fixed_length(pkg, oracle, v):   # src/lib.rs:875-879, restricted to equal-length keys of <= 8 bytes
    if not v:
        return
    keys = np.frombuffer(b"".join(sorted(v)), dtype=np.uint8).reshape(len(v), -1)
    check_mphf(pkg, oracle, keys)


@settings(**SETTINGS)
@given(st.sets(st.integers(0, 2**64 - 1), max_size=3000), st.lists(st.integers(0, 400), max_size=12))
def test_check_int_slices(pkg, oracle, v, lens):       # src/lib.rs:822-855: random chunk lengths
    keys = np.array(sorted(v), dtype=np.uint64)
    chunks, pos = [], 0
    for ln in lens:
        chunks.append(keys[pos:pos + ln])
        pos = min(len(keys), pos + ln)
    chunks.append(keys[pos:])
    check_chunked_mphf(pkg, oracle, chunks, len(keys))


def test_from_ints_serial(pkg, oracle):                # src/lib.rs:881-885: 1 M keys 2*i (i32 there; u64 and i32 here)
    for dtype in (np.uint64, np.int32):
        items = (np.arange(1_000_000) * 2).astype(dtype)
        m = pkg.Mphf.new(2.0, items)
        assert np.array_equal(np.sort(m.hash_batch(items)), np.arange(len(items), dtype=np.uint64))
        assert_same_levels(m.levels(), oracle.OracleMphf.new(2.0, items).levels())
