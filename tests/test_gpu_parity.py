"""Parity tests proper (run on a B200 with -m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs -- bit-exact per-level words, ranks and hash values -- plus the committed
golden fingerprints and size-independent properties at the BASELINE sizes."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
from test_oracle import _keys_for  # noqa: E402


def assert_same_levels(got, exp):
    assert len(got) == len(exp), ("level count", len(got), len(exp))
    for l, ((b0, w0, r0), (b1, w1, r1)) in enumerate(zip(got, exp)):
        assert b0 == b1, ("bits", l, b0, b1)
        assert np.array_equal(w0, w1), ("words differ at level", l)
        assert np.array_equal(r0, r1), ("ranks differ at level", l)


@pytest.mark.parametrize("entry", GOLDEN["mphf"], ids=[e["name"] for e in GOLDEN["mphf"]])
def test_golden_fingerprints(pkg, oracle, entry):
    keys = _keys_for(entry["name"], oracle)
    seed = entry["starting_seed"]
    m = pkg.Mphf.new_parallel(entry["gamma"], keys, seed)
    lv = m.levels()
    assert [l[0] for l in lv] == entry["bits"]
    assert oracle.fingerprint(lv) == entry["sha256"]
    if seed is None:
        m2 = pkg.Mphf.new(entry["gamma"], keys)
        assert oracle.fingerprint(m2.levels()) == entry["sha256"]
        if len(keys):
            assert [int(v) for v in m.hash_batch(keys[:6])] == entry["hash_first"]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 149, 150, 151, 1000, 4095, 4096, 4097, 8191, 8192, 8193,
                               20000, 65537, 300001, 1 << 20, 3_000_000])
def test_build_and_hash_parity_sizes(pkg, oracle, n):
    keys = oracle.keygen(n, seed=n + 1, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    if n:
        h = m.hash_batch(keys)
        assert np.array_equal(h, ref.hash_batch(keys, nthreads=0))
        assert np.array_equal(np.sort(h), np.arange(n, dtype=np.uint64))
    st = m.build_stats()
    assert st["n_levels"] == ref.num_levels and sum(st["keys_in"][:1]) == n


@pytest.mark.parametrize("gamma", [1.0100001, 1.02, 1.3, 1.7, 2.0, 3.0, 5.0, 17.5, 100.0])
def test_gamma_sweep_parity(pkg, oracle, gamma):
    keys = oracle.keygen(200_000, seed=11, kind=1)
    m = pkg.Mphf.new_parallel(gamma, keys, None)
    ref = oracle.OracleMphf.new_parallel(gamma, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


@pytest.mark.parametrize("seed", [0, 1, 5, 30, 31, 32, 40, 2**63, 2**64 - 3])
def test_starting_seed_parity(pkg, oracle, seed):
    # build seed index = starting_seed + level (src/lib.rs:370,385); crosses the 1<<(2*iter) wrap at 32
    keys = oracle.keygen(100_000, seed=1, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, seed)
    ref = oracle.OracleMphf.new_parallel(1.7, keys, seed)
    assert_same_levels(m.levels(), ref.levels())
    # lookups use seed index = level index on both sides, so even the inconsistent answers agree
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_many_levels_crosses_seed_wrap(pkg, oracle):
    # gamma just above the floor on a larger set needs > 32 levels: level 32 reuses level 0's seed
    keys = oracle.keygen(3_000_000, seed=3, kind=0)
    m = pkg.Mphf.new_parallel(1.0100001, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.0100001, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_level_size_at_and_above_2_pow_32_uses_modulo(pkg, oracle):
    # gamma * n >= 2^32 switches level 0 from fast-range to `%` (src/lib.rs:81); later levels switch back
    n = 4_200_000
    gamma = 1024.0
    assert oracle.level_size(gamma, n) >= 2**32
    keys = oracle.keygen(n, seed=21, kind=0)
    m = pkg.Mphf.new_parallel(gamma, keys, None)
    ref = oracle.OracleMphf.new_parallel(gamma, keys)
    got, exp = m.levels(), ref.levels()
    assert got[0][0] >= 2**32 and got[-1][0] < 2**32
    assert_same_levels(got, exp)
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_key_order_does_not_matter(pkg, oracle):
    keys = oracle.keygen(500_000, seed=2, kind=0)
    a = pkg.Mphf.new(1.7, keys)
    rng = np.random.default_rng(0)
    shuffled = keys.copy()
    rng.shuffle(shuffled)
    b = pkg.Mphf.new(1.7, shuffled)
    assert oracle.fingerprint(a.levels()) == oracle.fingerprint(b.levels())


def test_literal_bench_input(pkg, oracle):
    # benches/build.rs:11-12 -- keys 0,2,4,...,1999998, gamma 2.0; scan1_ser hashes them in order
    keys = oracle.keygen(1_000_000, 0, kind=2)
    m = pkg.Mphf.new(2.0, keys)
    ref = oracle.OracleMphf.new(2.0, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_absent_keys_match_oracle_including_none(pkg, oracle):
    keys = oracle.keygen(100_000, 1, 0)
    other = oracle.keygen(100_000, 99, 0)
    m = pkg.Mphf.new(1.7, keys)
    ref = oracle.OracleMphf.new(1.7, keys)
    out = m.hash_batch(other)
    assert np.array_equal(out, ref.hash_batch(other))
    assert (out == np.uint64(pkg.NONE)).any()
    assert m.try_hash(int(keys[5])) == ref.try_hash(int(keys[5]))
    missing = int(other[np.flatnonzero(out == np.uint64(pkg.NONE))[0]])
    assert m.try_hash(missing) is None
    with pytest.raises(pkg.MphfPanic):
        m.hash(missing)


@pytest.mark.parametrize("n", [1, 10, 1000, 20000])
def test_absent_keys_against_small_mphfs_are_none_not_garbage(pkg, oracle, n):
    # shallow MPHFs queried with keys outside the set: far more than 32 keys of a warp tile miss every level, all of
    # them must come back as None (try_hash, src/lib.rs:340-355), never as the raw key
    keys = oracle.keygen(n, seed=3 * n + 1, kind=0)
    absent = oracle.keygen(10_000, seed=777, kind=0)
    for gamma in (1.7, 5.0):
        m = pkg.Mphf.new(gamma, keys)
        ref = oracle.OracleMphf.new(gamma, keys)
        out = m.hash_batch(absent)
        assert np.array_equal(out, ref.hash_batch(absent))
        none = out == np.uint64(pkg.NONE)
        assert none.sum() > 0 and (out[~none] < n).all()
        mixed = np.concatenate([absent[:300], keys, absent[300:900]])
        assert np.array_equal(m.hash_batch(mixed), ref.hash_batch(mixed))


def test_duplicate_keys_report_max_iters(pkg):
    for n in (10, 100_000):
        keys = np.arange(n, dtype=np.uint64) * 7 + 1
        keys[-1] = keys[0]
        with pytest.raises(pkg.MphfPanic) as e:
            pkg.Mphf.new_parallel(1.7, keys, None)
        assert e.value.code == -3 and "counldn't find unique hashes" in str(e.value)


def test_heavily_duplicated_input_grows_buffers_then_reports_max_iters(pkg):
    # every key twice: no level shrinks, so the planned arena and redo lists overflow and must be grown
    base = np.arange(50_000, dtype=np.uint64) * 3 + 5
    keys = np.concatenate([base, base])
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new_parallel(1.7, keys, None)
    assert e.value.code == -3


def test_device_resident_keys_and_outputs(pkg, oracle):
    import torch
    n = 2_000_000
    keys = oracle.keygen(n, seed=8, kind=0)
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    m = pkg.Mphf.new_parallel(1.7, dk, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    out = m.hash_batch(dk)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), ref.hash_batch(keys, nthreads=0))
    # the same on a caller-provided stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        m2 = pkg.Mphf.new_parallel(1.7, dk, None, stream=s)
        out2 = m2.hash_batch(dk, stream=s)
    s.synchronize()
    assert torch.equal(out, out2)


def test_device_keygen_matches_oracle_keygen(pkg, oracle):
    import ctypes as C
    import torch
    for kind in (0, 1, 2):
        d = torch.empty(100_003, dtype=torch.int64, device="cuda")
        rc = pkg.lib().bphf_keygen_u64(C.c_void_p(d.data_ptr()), 17, d.numel(), 5, kind, 0, None)
        assert rc == 0
        assert np.array_equal(d.cpu().numpy().view(np.uint64), oracle.keygen(d.numel(), 5, kind, start=17))


def test_from_levels_upload_round_trip(pkg, oracle):
    keys = oracle.keygen(300_000, 4, 0)
    ref = oracle.OracleMphf.new(1.7, keys)
    blob = oracle.serialize_bincode(ref.levels())          # an Mphf serialised by the reference side
    m = pkg.Mphf.from_levels(oracle.deserialize_bincode(blob))
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))
    assert oracle.serialize_bincode(m.levels()) == blob
    # and the other direction: built on the GPU, serialised, read by the reference side
    g = pkg.Mphf.new(1.7, keys)
    assert oracle.serialize_bincode(g.levels()) == blob
    for l in range(g.num_levels):
        b, w, r = g.level(l)
        assert b == ref.level(l)[0] and np.array_equal(w, ref.level(l)[1]) and np.array_equal(r, ref.level(l)[2])
    with pytest.raises(pkg.MphfPanic):
        g.level_dims(g.num_levels)


def test_boomhashmap_parity(pkg, oracle):
    keys = oracle.keygen(400_000, 6, 1)
    vals = keys * np.uint64(3) + np.uint64(1)
    hm = pkg.BoomHashMap.new_parallel(keys, vals)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    mk, mv = ref.create_map(keys, vals)
    assert np.array_equal(hm.keys, mk) and np.array_equal(hm.values, mv)
    q = np.concatenate([keys[:1000], oracle.keygen(1000, 77, 0)])
    v, found = hm.get_batch(q)
    rv, rfound = ref.map_get_batch(mk, mv, q)
    assert np.array_equal(found, rfound) and np.array_equal(v, rv)
    assert hm.get(int(keys[3])) == int(vals[3]) and hm.get(12345) is None
    assert hm.get_key_id(int(keys[3])) == ref.try_hash(int(keys[3]))
    assert len(hm) == len(keys) and hm.get_key(7) == int(mk[7])
    nk = pkg.NoKeyBoomHashMap.new_parallel(keys, vals)
    out, ok = nk.get_batch(keys[:500])
    assert ok.all() and np.array_equal(out, vals[:500])


def test_boomhashmap_gamma_sweep_with_prebuilt_mphf(pkg, oracle):
    # BASELINE config 5 at a size the oracle finishes in seconds; gamma = 1.0 must be rejected
    keys = oracle.keygen(1_000_000, 5, 1)
    vals = np.arange(len(keys), dtype=np.uint64)
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new_parallel(1.0, keys, None)
    assert e.value.code == -2
    for gamma in (1.7, 2.0, 5.0):
        m = pkg.Mphf.new_parallel(gamma, keys, None)
        ref = oracle.OracleMphf.new_parallel(gamma, keys)
        assert_same_levels(m.levels(), ref.levels())
        hm = pkg.BoomHashMap.new_with_mphf(m, keys, vals)
        v, found = hm.get_batch(keys[::97])
        assert found.all() and np.array_equal(v, vals[::97])


def test_ten_million_full_oracle_comparison(pkg, oracle):
    n = 10_000_000
    keys = oracle.keygen(n, seed=1, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_baseline_size_100m_properties(pkg, oracle):
    """BASELINE configs[1]: 1e8 random u64 keys, gamma 1.7, one B200.  Size-independent checks:
    the outputs are a permutation of 0..n (the reference's own property), total popcount == n,
    ranks == cumulative popcounts, rebuilding gives the same structure (checksum of checksums),
    and a 1e6 sample of keys agrees with an oracle that was fed the exported levels."""
    import ctypes as C
    import torch
    n = 100_000_000
    dk = torch.empty(n, dtype=torch.int64, device="cuda")
    assert pkg.lib().bphf_keygen_u64(C.c_void_p(dk.data_ptr()), 0, n, 1, 0, 0, None) == 0
    m = pkg.Mphf.new_parallel(1.7, dk, None)
    out = m.hash_batch(dk)
    torch.cuda.synchronize()
    bad = C.c_uint64(123)
    assert pkg.lib().bphf_check_permutation(C.c_void_p(out.data_ptr()), n, 0, C.byref(bad)) == 0
    assert bad.value == 0
    lv = m.levels()
    pop = 0
    for bits, words, ranks in lv:
        pc = np.bitwise_count(words).astype(np.uint64)
        blocks = np.add.reduceat(pc, np.arange(0, len(pc), 8))
        excl = np.concatenate([[0], np.cumsum(blocks)[:-1]]).astype(np.uint64) + np.uint64(pop)
        assert np.array_equal(ranks, excl)
        pop += int(pc.sum())
    assert pop == n and m.total_dims()[2] == n
    fp = oracle.fingerprint(lv)
    m2 = pkg.Mphf.new_parallel(1.7, dk, None)
    assert oracle.fingerprint(m2.levels()) == fp
    sample = oracle.keygen(1_000_000, 1, 0, start=54_321_000)
    ref = oracle.OracleMphf.from_levels(lv)
    assert np.array_equal(m.hash_batch(sample), ref.hash_batch(sample, nthreads=0))
    # level sizes follow max(255, floor(gamma * k)) with k = keys left (src/lib.rs:369,384)
    st = m.build_stats()
    assert st["keys_in"][0] == n and [oracle.level_size(1.7, k) for k in st["keys_in"]] == st["bits"]


def test_baseline_size_100m_is_the_oracles_structure_word_for_word(pkg, oracle):
    """BASELINE configs[1] at full size against the ORACLE (not against itself): 1e8 keys, every level's bits, words and
    ranks equal what the restated new_parallel builds from the same keys (about 2 s of host time on 16 cores)."""
    import ctypes as C
    import torch
    n = 100_000_000
    dk = torch.empty(n, dtype=torch.int64, device="cuda")
    assert pkg.lib().bphf_keygen_u64(C.c_void_p(dk.data_ptr()), 0, n, 1, 0, 0, None) == 0
    m = pkg.Mphf.new_parallel(1.7, dk, None)
    keys = oracle.keygen(n, seed=1, kind=0)
    assert np.array_equal(dk[:1000].cpu().numpy().view(np.uint64), keys[:1000])   # same generator on both sides
    ref = oracle.OracleMphf.new_parallel(1.7, keys, None, nthreads=len(os.sched_getaffinity(0)))
    assert_same_levels(m.levels(), ref.levels())


def test_natural_policy_buckets_level0_at_3e8_keys_and_matches_the_oracle(pkg, oracle):
    """3e8 keys: level 0 has 5.1e8 bits > 4.5e8, so the DEFAULT policy builds it with the bucketed (shared-memory)
    kernels on its own -- no forced mode, no small thresholds -- followed by the plain striped-list levels; all words and
    ranks against the oracle's new_parallel (about 7 s of host time)."""
    import ctypes as C
    import torch
    n = 300_000_000
    dk = torch.empty(n, dtype=torch.int64, device="cuda")
    assert pkg.lib().bphf_keygen_u64(C.c_void_p(dk.data_ptr()), 0, n, 9, 0, 0, None) == 0
    m = pkg.Mphf.new_parallel(1.7, dk, None)
    st = m.build_stats()
    assert st["bucketed_levels"] >= 1 and st["rebuilds"] == 0
    keys = oracle.keygen(n, seed=9, kind=0)
    ref = oracle.OracleMphf.new_parallel(1.7, keys, None, nthreads=len(os.sched_getaffinity(0)))
    assert_same_levels(m.levels(), ref.levels())
    out = m.hash_batch(dk)
    bad = C.c_uint64(1)
    assert pkg.lib().bphf_check_permutation(C.c_void_p(out.data_ptr()), n, 0, C.byref(bad)) == 0 and bad.value == 0


def test_level0_beyond_2_pow_32_bits_refuses_buckets_and_matches_the_oracle(pkg, oracle):
    """gamma * n >= 2^32: level 0 takes the `%` reduction of the full 64-bit hash (src/lib.rs:81) and cannot be bucketed
    (32-bit slot arithmetic); gamma = 30 keeps the key set small (1.5e8 keys -> 4.5e9 bits) so the oracle finishes in
    seconds.  Words, ranks and hashes against the oracle."""
    import ctypes as C
    import torch
    n, gamma = 150_000_000, 30.0
    assert oracle.level_size(gamma, n) >= 2 ** 32
    dk = torch.empty(n, dtype=torch.int64, device="cuda")
    assert pkg.lib().bphf_keygen_u64(C.c_void_p(dk.data_ptr()), 0, n, 4, 0, 0, None) == 0
    m = pkg.Mphf.new_parallel(gamma, dk, None)
    st = m.build_stats()
    assert st["bits"][0] >= 2 ** 32 and st["bucketed_levels"] == 0
    keys = oracle.keygen(n, seed=4, kind=0)
    ref = oracle.OracleMphf.new_parallel(gamma, keys, None, nthreads=len(os.sched_getaffinity(0)))
    assert_same_levels(m.levels(), ref.levels())
    sample = keys[::997]
    assert np.array_equal(m.hash_batch(sample), ref.hash_batch(sample, nthreads=0))


# ---- construction strategies: bucketed shared-memory kernels vs plain L2-atomic kernels -----------------------
@pytest.fixture
def small_buckets(pkg):
    """force the bucketed kernels onto small levels so the unit-test sizes exercise them (and every transition)"""
    pkg.lib().bphf_set_build_mode(2, 3000)
    yield
    pkg.lib().bphf_set_build_mode(0, 0)


@pytest.fixture
def plain_only(pkg):
    pkg.lib().bphf_set_build_mode(1, 0)
    yield
    pkg.lib().bphf_set_build_mode(0, 0)


@pytest.mark.parametrize("n", [2999, 3000, 3001, 5000, 9000, 20_000, 65_537, 300_001, 1 << 20, 3_000_000])
def test_bucketed_build_parity_sizes(pkg, oracle, small_buckets, n):
    keys = oracle.keygen(n, seed=n + 3, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    st = m.build_stats()
    assert st["rebuilds"] == 0
    assert (st["bucketed_levels"] > 0) == (n > 4096)  # <= 4096 keys: the tail kernel owns every level
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


@pytest.mark.parametrize("gamma", [1.0100001, 1.3, 1.7, 2.0, 5.0, 17.5, 100.0, 1000.0])
def test_bucketed_build_parity_gammas(pkg, oracle, small_buckets, gamma):
    # large gamma: the level after a bucketed one is already tail sized (scatter -> plain list -> tail kernel)
    keys = oracle.keygen(150_000, seed=12, kind=1)
    m = pkg.Mphf.new_parallel(gamma, keys, None)
    ref = oracle.OracleMphf.new_parallel(gamma, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert m.build_stats()["bucketed_levels"] > 0
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


@pytest.mark.parametrize("seed", [5, 31, 2**64 - 3])
def test_bucketed_build_parity_seeds(pkg, oracle, small_buckets, seed):
    keys = oracle.keygen(100_000, seed=1, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, seed)
    assert_same_levels(m.levels(), oracle.OracleMphf.new_parallel(1.7, keys, seed).levels())


def test_host_keys_in_chunks_every_strategy(pkg, oracle):
    # 40 M host keys = 3 H2D chunks, each consumed by the level-0 entry kernel as it lands
    n = 40_000_000
    keys = oracle.keygen(n, seed=2, kind=0)
    exp = oracle.OracleMphf.new_parallel(1.7, keys).levels()
    try:
        for mode in (2, 1, 0):
            pkg.lib().bphf_set_build_mode(mode, 0)
            m = pkg.Mphf.new_parallel(1.7, keys, None)
            assert (m.build_stats()["bucketed_levels"] >= 4) == (mode == 2)
            assert_same_levels(m.levels(), exp)
    finally:
        pkg.lib().bphf_set_build_mode(0, 0)


def test_plain_and_bucketed_strategies_agree(pkg, oracle, plain_only):
    keys = oracle.keygen(3_000_000, seed=5, kind=0)
    a = pkg.Mphf.new_parallel(1.7, keys, None)
    assert a.build_stats()["bucketed_levels"] == 0
    pkg.lib().bphf_set_build_mode(2, 0)
    b = pkg.Mphf.new_parallel(1.7, keys, None)
    assert b.build_stats()["bucketed_levels"] > 0
    pkg.lib().bphf_set_build_mode(0, 0)
    c = pkg.Mphf.new_parallel(1.7, keys, None)   # default: persistent cascade kernel
    assert c.build_stats()["bucketed_levels"] == 0 and c.build_stats()["kernel_launches"] < a.build_stats()["kernel_launches"]
    assert oracle.fingerprint(a.levels()) == oracle.fingerprint(b.levels()) == oracle.fingerprint(c.levels())
    assert_same_levels(a.levels(), oracle.OracleMphf.new_parallel(1.7, keys).levels())


def test_bucket_overflow_falls_back_to_plain_kernels(pkg, small_buckets):
    # every key 40 times: all copies of a key land in one bucket, the region overflows, the build is redone
    # with the L2-atomic kernels and reports the reference's MAX_ITERS failure
    base = np.arange(5_000, dtype=np.uint64) * 3 + 5
    keys = np.tile(base, 40)
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new_parallel(1.7, keys, None)
    assert e.value.code == -3
    # a skewed but valid case: distinct keys plus one key repeated would be invalid input, so instead check that
    # a valid build right after the failed one is unaffected
    ok = pkg.Mphf.new_parallel(1.7, base, None)
    assert ok.total_dims()[2] == len(base)


# ---- chunked constructors ---------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 99, 100, 101, 5000, 300_001, 3_000_000])
@pytest.mark.parametrize("gamma", [1.7, 2.0, 5.0])
def test_from_chunked_iterator_parity(pkg, oracle, n, gamma):
    keys = oracle.keygen(n, seed=n + 17, kind=0)
    chunks = [keys[:0]] + list(np.array_split(keys, 4)) if n else [keys]
    m = pkg.Mphf.from_chunked_iterator(gamma, chunks, n)
    assert_same_levels(m.levels(), oracle.OracleMphf.from_chunked_iterator(gamma, chunks, n).levels())
    mp = pkg.Mphf.from_chunked_iterator_parallel(gamma, chunks, None, n, 4)
    ref = oracle.OracleMphf.from_chunked_iterator_parallel(gamma, chunks, None, n)
    assert_same_levels(mp.levels(), ref.levels())
    if n:
        h = mp.hash_batch(keys)
        assert np.array_equal(h, ref.hash_batch(keys, nthreads=0))
        assert np.array_equal(np.sort(h), np.arange(n, dtype=np.uint64))


def test_from_chunked_iterator_parallel_gamma_switch_and_limits(pkg, oracle):
    keys = oracle.keygen(1_000_000, seed=4, kind=1)
    chunks = np.array_split(keys, 9)
    for gamma in (1.02, 1.3, 17.5):
        m = pkg.Mphf.from_chunked_iterator_parallel(gamma, chunks, None, len(keys))
        assert_same_levels(m.levels(), oracle.OracleMphf.from_chunked_iterator_parallel(gamma, chunks, None, len(keys)).levels())
    # max_iters guards the chunked levels only
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.from_chunked_iterator_parallel(2.0, chunks, 2, len(keys))
    assert e.value.code == -3
    ok = pkg.Mphf.from_chunked_iterator_parallel(2.0, chunks, 8, len(keys))
    assert_same_levels(ok.levels(), oracle.OracleMphf.from_chunked_iterator_parallel(2.0, chunks, 8, len(keys)).levels())
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.from_chunked_iterator(1.7, chunks, len(keys) + 1)     # n does not match the chunks
    assert e.value.code == -1
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.from_chunked_iterator_parallel(1.0, chunks, None, len(keys))
    assert e.value.code == -2


def _chunked_parallel_both(pkg, oracle, gamma, chunks, max_iters, n):
    """(levels or 'panic') from the GPU path and from the oracle."""
    try:
        got = pkg.Mphf.from_chunked_iterator_parallel(gamma, chunks, max_iters, n).levels()
    except pkg.MphfPanic as e:
        assert e.code == -3
        got = "panic"
    try:
        exp = oracle.OracleMphf.from_chunked_iterator_parallel(gamma, chunks, max_iters, n).levels()
    except oracle.OracleError:
        exp = "panic"
    return got, exp


def test_from_chunked_iterator_parallel_max_iters_boundaries(pkg, oracle):
    # `iter > max_iters` sits at the top of the reference's loop (src/lib.rs:568-573): before the keys_remaining == 0
    # break, and never reached again once the buffering level has run (:645-648)
    keys = oracle.keygen(1_000_000, seed=4, kind=1)
    chunks = np.array_split(keys, 5)
    free = oracle.OracleMphf.from_chunked_iterator_parallel(2.0, chunks, None, len(keys))
    thr = int(np.float32(0.01) * np.float32(len(keys)))
    # the buffering level = first level entered with fewer than min(50 M, thr) keys
    sizes = [b for b, _, _ in free.levels()]
    buf_level = next(l for l, b in enumerate(sizes) if b < 2.0 * min(50_000_000, thr))
    assert 1 <= buf_level < len(sizes) - 1
    for mi in (buf_level - 1, buf_level, buf_level + 1):
        got, exp = _chunked_parallel_both(pkg, oracle, 2.0, chunks, mi, len(keys))
        assert (got == "panic") == (exp == "panic"), (mi, buf_level)
        if exp != "panic":
            assert_same_levels(got, exp)
    assert _chunked_parallel_both(pkg, oracle, 2.0, chunks, buf_level - 1, len(keys))[1] == "panic"
    assert _chunked_parallel_both(pkg, oracle, 2.0, chunks, buf_level, len(keys))[1] != "panic"
    # small n: no buffering at all (thr = 0); the check after the LAST level fires even though nothing is left
    small = oracle.keygen(60, seed=8, kind=0)
    full = oracle.OracleMphf.from_chunked_iterator_parallel(1.7, [small], None, len(small))
    last = full.num_levels - 1
    for mi in (last - 1, last, last + 1):
        got, exp = _chunked_parallel_both(pkg, oracle, 1.7, [small], mi, len(small))
        assert (got == "panic") == (exp == "panic"), (mi, last)
        if exp != "panic":
            assert_same_levels(got, exp)
    assert _chunked_parallel_both(pkg, oracle, 1.7, [small], last, len(small))[1] == "panic"
    assert _chunked_parallel_both(pkg, oracle, 1.7, [small], last + 1, len(small))[1] != "panic"


def test_from_chunked_iterator_device_chunks(pkg, oracle):
    import torch
    keys = oracle.keygen(2_000_000, seed=12, kind=0)
    dev = [torch.from_numpy(c.view(np.int64)).cuda() for c in np.array_split(keys, 3)]
    m = pkg.Mphf.from_chunked_iterator_parallel(2.0, dev, None, len(keys))
    assert_same_levels(m.levels(), oracle.OracleMphf.from_chunked_iterator_parallel(2.0, [keys], None, len(keys)).levels())
    one = pkg.Mphf.from_chunked_iterator(1.7, [torch.from_numpy(keys.view(np.int64)).cuda()], len(keys))
    assert_same_levels(one.levels(), oracle.OracleMphf.new(1.7, keys).levels())


# ---- key types other than u64 (the reference's check_u32 / check_isize / check_vec_u8) ------
def _distinct(rng, dtype, n):
    info = np.iinfo(dtype)
    span = int(info.max) - int(info.min) + 1
    n = min(n, span)
    if span <= 1 << 20:
        vals = rng.permutation(span)[:n] + int(info.min)
    else:
        vals = np.unique(rng.integers(info.min, info.max, size=2 * n, dtype=np.int64))[:n]
    return vals.astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.uint16, np.int16, np.uint8, np.int8, np.int64])
@pytest.mark.parametrize("n", [200, 5000, 300_000])
def test_integer_key_types_parity(pkg, oracle, dtype, n):
    keys = _distinct(np.random.default_rng(n), dtype, n)
    m = pkg.Mphf.new(1.7, keys)
    ref = oracle.OracleMphf.new(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    assert_same_levels(pkg.Mphf.new_parallel(1.7, keys, None).levels(), ref.levels())
    h = m.hash_batch(keys)
    assert np.array_equal(h, ref.hash_batch(keys))
    assert np.array_equal(np.sort(h), np.arange(len(keys), dtype=np.uint64))
    assert m.try_hash(int(keys[0])) == int(h[0])


@pytest.mark.parametrize("width", [1, 2, 3, 4, 5, 6, 7, 8])
def test_byte_array_keys_parity(pkg, oracle, width):
    rng = np.random.default_rng(width)
    keys = np.unique(rng.integers(0, 256, size=(60_000, width), dtype=np.uint8), axis=0)
    keys = keys[rng.permutation(len(keys))]
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    h = m.hash_batch(keys)
    assert np.array_equal(h, ref.hash_batch(keys))
    assert np.array_equal(np.sort(h), np.arange(len(keys), dtype=np.uint64))
    assert m.key_kind == (2, width)
    assert m.try_hash(bytes(keys[5])) == int(h[5])
    # upload of the exported levels keeps the key type; a u64 view of the same levels hashes differently
    m2 = pkg.Mphf.from_levels(m.levels(), key_kind=2, key_width=width)
    assert np.array_equal(m2.hash_batch(keys), h)
    with pytest.raises(TypeError):
        m.hash_batch(np.arange(10, dtype=np.uint64))


def test_u32_device_keys_and_absent_keys(pkg, oracle):
    import torch
    keys = _distinct(np.random.default_rng(7), np.uint32, 1_000_000)
    dk = torch.from_numpy(keys.view(np.int32)).cuda()
    m = pkg.Mphf.new_parallel(2.0, dk, None)
    ref = oracle.OracleMphf.new_parallel(2.0, keys)
    assert_same_levels(m.levels(), ref.levels())
    absent = _distinct(np.random.default_rng(8), np.uint32, 100_000)
    got = m.hash_batch(torch.from_numpy(absent.view(np.int32)).cuda()).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, ref.hash_batch(absent))


@pytest.mark.parametrize("entry", GOLDEN["key_kinds"], ids=[e["name"] for e in GOLDEN["key_kinds"]])
def test_key_kind_golden_fingerprints_gpu(pkg, oracle, entry):
    from test_oracle import _golden_kind_keys
    keys = _golden_kind_keys(entry["name"])
    m = pkg.Mphf.new_parallel(entry["gamma"], keys, None)
    assert [l[0] for l in m.levels()] == entry["bits"] and oracle.fingerprint(m.levels()) == entry["sha256"]
    assert [int(v) for v in m.hash_batch(keys[:4])] == entry["hash_first"]


def test_chunked_parallel_golden_fingerprint_gpu(pkg, oracle):
    e = GOLDEN["chunked_parallel"]
    keys = oracle.keygen(e["n"], e["keygen_seed"], 0)
    m = pkg.Mphf.from_chunked_iterator_parallel(e["gamma"], np.array_split(keys, 7), None, e["n"])
    assert [l[0] for l in m.levels()] == e["bits"] and oracle.fingerprint(m.levels()) == e["sha256"]


def test_boomhashmap2_and_nokey2_parity(pkg, oracle):
    # BoomHashMap2 / NoKeyBoomHashMap2 (src/hashmap.rs:222-427, 530-641): three arrays along one permutation
    keys = oracle.keygen(200_000, seed=41, kind=1)
    vals = np.arange(len(keys), dtype=np.uint64) * 3 + 1
    aux = np.arange(len(keys), dtype=np.uint64) ^ np.uint64(0xABCDEF)
    m = pkg.BoomHashMap2.new_parallel(keys, vals, aux)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    k1, v1 = ref.create_map(keys, vals)
    k2, a1 = ref.create_map(keys, aux)
    assert np.array_equal(m.keys, k1) and np.array_equal(m.values, v1) and np.array_equal(m.aux_values, a1)
    absent = oracle.keygen(1000, seed=42, kind=1)
    q = np.concatenate([keys[:1000], absent])
    v, a, f = m.get_batch(q)
    assert f[:1000].all() and np.array_equal(v[:1000], vals[:1000]) and np.array_equal(a[:1000], aux[:1000])
    assert not f[1000:][~np.isin(absent, keys)].any()
    assert m.get(int(keys[7])) == (int(vals[7]), int(aux[7])) and len(m) == len(keys)
    nk = pkg.NoKeyBoomHashMap2.new_parallel(keys, vals, aux)
    v, a, ok = nk.get_batch(keys[:500])
    assert ok.all() and np.array_equal(v, vals[:500]) and np.array_equal(a, aux[:500])


def test_concurrent_builds_and_lookups_from_host_threads(pkg, oracle):
    """Mphf is Send + Sync and `hash` takes &self (src/lib.rs:324): one handle answers lookups from many host threads
    at once, and independent builds may run concurrently (each thread gets its own streams inside the library)."""
    import threading
    shared_keys = oracle.keygen(400_000, seed=90, kind=0)
    shared = pkg.Mphf.new_parallel(1.7, shared_keys, None)
    want_shared = oracle.OracleMphf.new_parallel(1.7, shared_keys).hash_batch(shared_keys, nthreads=0)
    n_threads, errors = 8, []

    def worker(t):
        try:
            keys = oracle.keygen(150_000 + 1000 * t, seed=100 + t, kind=t % 2)
            ref = oracle.OracleMphf.new(1.7, keys)
            for rep in range(3):
                m = pkg.Mphf.new_parallel(1.7, keys, None)
                assert oracle.fingerprint(m.levels()) == oracle.fingerprint(ref.levels()), ("build", t, rep)
                lo = (t * 50_000) % len(shared_keys)
                got = shared.hash_batch(shared_keys[lo:lo + 50_000])
                assert np.array_equal(got, want_shared[lo:lo + 50_000]), ("shared lookup", t, rep)
                assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys)), ("own lookup", t, rep)
        except BaseException as e:  # noqa: BLE001 -- reported in the main thread
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not any(th.is_alive() for th in threads), "a worker thread hung"
    assert not errors, errors


# ---- streamed ingest of host key sets (out-of-core side of SURVEY 8f row 3) -----------------------------------
@pytest.fixture()
def streamed_ingest(pkg):
    """force every host key set through the device ring (3 slots of 10 000 keys) for the duration of a test"""
    pkg.lib().bphf_set_ingest_mode(1, 10_000)
    yield
    pkg.lib().bphf_set_ingest_mode(0, 0)


@pytest.mark.parametrize("n", [0, 1, 100, 4096, 9_999, 10_000, 10_001, 30_000, 123_457, 1_000_000, 3_000_000])
def test_streamed_ingest_parity_sizes(pkg, oracle, streamed_ingest, n):
    keys = oracle.keygen(n, seed=n + 7, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    assert_same_levels(m.levels(), ref.levels())
    if n:
        assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))
    # small sets (level 1 would already belong to the tail kernel) are handed to the in-core schedule
    assert m.build_stats()["streamed_ingest"] == (1 if n >= 9_999 else 0)


def test_streamed_ingest_keeps_the_keys_out_of_device_memory(pkg, oracle):
    n = 4_000_000
    keys = oracle.keygen(n, seed=77, kind=0)
    peak = {}
    for mode in (2, 1):
        pkg.lib().bphf_set_ingest_mode(mode, 100_000)
        pkg.lib().bphf_set_profile(2)  # the in-core build only keeps the memory statistic when profiling
        try:
            m = pkg.Mphf.new_parallel(1.7, keys, None)
        finally:
            pkg.lib().bphf_set_ingest_mode(0, 0)
            pkg.lib().bphf_set_profile(0)
        st = m.build_stats()
        assert st["streamed_ingest"] == (1 if mode == 1 else 0)
        peak[mode] = st["peak_device_bytes"]
        del m
    # in core: the 32 MB of keys + lists + arenas; streamed: 3 ring slots of 0.8 MB instead of the keys
    assert peak[2] >= 8 * n and peak[1] <= peak[2] - 8 * n * 0.9, peak


@pytest.mark.parametrize("gamma", [1.02, 2.0, 5.0, 100.0, 3000.0])
@pytest.mark.parametrize("seed", [None, 7])
def test_streamed_ingest_gamma_and_seed(pkg, oracle, streamed_ingest, gamma, seed):
    # gamma = 3000: almost nothing survives level 0 -> the streamed schedule hands over to the in-core one
    keys = oracle.keygen(250_000, seed=3, kind=1)
    m = pkg.Mphf.new_parallel(gamma, keys, seed)
    ref = oracle.OracleMphf.new_parallel(gamma, keys, seed)
    assert_same_levels(m.levels(), ref.levels())


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_streamed_ingest_is_independent_of_the_build_mode(pkg, oracle, streamed_ingest, mode):
    keys = oracle.keygen(2_000_000, seed=21, kind=0)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    try:
        pkg.lib().bphf_set_build_mode(mode, 0)
        m = pkg.Mphf.new_parallel(1.7, keys, None)
    finally:
        pkg.lib().bphf_set_build_mode(0, 0)
    assert_same_levels(m.levels(), ref.levels())
    st = m.build_stats()
    assert st["bucketed_levels"] == 0  # the streamed schedule marks through the L2 whatever the mode says


@pytest.mark.parametrize("parallel", [False, True])
def test_streamed_ingest_of_host_chunks(pkg, oracle, streamed_ingest, parallel):
    """from_chunked_iterator[_parallel] with chunks that never become device resident as a whole: ragged chunk
    lengths around the ring size, empty chunks, the gamma switch of the parallel variant"""
    keys = oracle.keygen(600_000, seed=31, kind=0)
    cuts = [0, 0, 5, 9_999, 20_000, 20_000, 45_001, 300_000, 600_000]
    chunks = [keys[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    if parallel:
        m = pkg.Mphf.from_chunked_iterator_parallel(2.5, chunks, None, len(keys))
        ref = oracle.OracleMphf.from_chunked_iterator_parallel(2.5, chunks, None, len(keys))
    else:
        m = pkg.Mphf.from_chunked_iterator(2.5, chunks, len(keys))
        ref = oracle.OracleMphf.from_chunked_iterator(2.5, chunks, len(keys))
    assert_same_levels(m.levels(), ref.levels())
    assert np.array_equal(m.hash_batch(keys), ref.hash_batch(keys, nthreads=0))


def test_streamed_ingest_duplicates_and_regrowth(pkg, oracle, streamed_ingest):
    keys = oracle.keygen(100_000, seed=5, kind=0)
    dup = np.concatenate([keys, keys[:10]])
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new_parallel(1.7, dup, None)
    assert "counldn't find unique hashes" in str(e.value)
    # heavily duplicated input makes level 1 much larger than planned: the lists / arena grow between the two passes
    many = np.concatenate([keys[:50_000]] * 3 + [keys])
    with pytest.raises(pkg.MphfPanic):
        pkg.Mphf.new_parallel(1.7, many, None)
