"""Sharded (multi-GPU) construction on the GPU: G shards -- here G logical shards on ONE device, one
thread each, talking through exactly the peer-memory kernels a multi-GPU run uses -- must produce, on every
shard, the structure an unsharded build of the union produces (bit-exact vs the CPU oracle)."""
import numpy as np
import pytest

from test_gpu_parity import assert_same_levels

pytestmark = pytest.mark.gpu


def _sharded(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".sharded")


def _split(keys, world, how):
    n = len(keys)
    if how == "even":
        b = _sharded_bounds(n, world)
    elif how == "first_empty":
        b = [0, 0] + _sharded_bounds(n, world - 1)[1:]
    elif how == "skewed":
        cut = sorted(set([0, n] + [min(n, (n * (2 ** r - 1)) // (2 ** world - 1)) for r in range(1, world)]))
        while len(cut) < world + 1:
            cut.append(n)
        b = sorted(cut)
    else:
        raise ValueError(how)
    return [keys[b[r]:b[r + 1]] for r in range(world)]


def _sharded_bounds(n, world):
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def _run(pkg, oracle, keys, world, gamma=1.7, seed=None, how="even", max_keys=None):
    sh = _sharded(pkg)
    comms = sh.ShardComm.local_group([0] * world, max_keys or max(len(keys), 1), gamma)
    try:
        shards = _split(keys, world, how)
        assert sum(len(s) for s in shards) == len(keys)
        ms = sh.build_in_threads(comms, gamma, shards, seed)
        ref = oracle.OracleMphf.new_parallel(gamma, keys, seed)
        exp = ref.levels()
        for m in ms:
            assert_same_levels(m.levels(), exp)
        if len(keys) and seed is None:
            h = ms[-1].hash_batch(keys)
            assert np.array_equal(h, ref.hash_batch(keys, nthreads=0))
            assert np.array_equal(np.sort(h), np.arange(len(keys), dtype=np.uint64))
        return ms
    finally:
        for c in comms:
            c.close()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("n", [0, 1, 100, 4096, 5000, 9001, 70_000, 1_000_003])
def test_sharded_build_parity(pkg, oracle, world, n):
    _run(pkg, oracle, oracle.keygen(n, seed=3 * n + world, kind=0), world)


@pytest.mark.parametrize("how", ["first_empty", "skewed"])
def test_unbalanced_shards(pkg, oracle, how):
    _run(pkg, oracle, oracle.keygen(300_001, seed=9, kind=0), 4, how=how)


@pytest.mark.parametrize("gamma", [1.02, 2.0, 5.0])
def test_sharded_gamma(pkg, oracle, gamma):
    _run(pkg, oracle, oracle.keygen(200_000, seed=21, kind=1), 3, gamma=gamma)


@pytest.mark.parametrize("seed", [5, 31, 2 ** 64 - 3])
def test_sharded_starting_seed(pkg, oracle, seed):
    _run(pkg, oracle, oracle.keygen(150_000, seed=4, kind=0), 2, seed=seed)


def test_comm_is_reusable_and_matches_unsharded_gpu_build(pkg, oracle):
    sh = _sharded(pkg)
    world = 4
    comms = sh.ShardComm.local_group([0] * world, 6_000_000, 2.0)
    try:
        for n, gamma in [(6_000_000, 1.7), (1_234_567, 2.0), (6_000_000, 1.7)]:
            keys = oracle.keygen(n, seed=n, kind=0)
            ms = sh.build_in_threads(comms, gamma, _split(keys, world, "even"))
            one = pkg.Mphf.new_parallel(gamma, keys, None)
            fp = oracle.fingerprint(one.levels())
            for m in ms:
                assert oracle.fingerprint(m.levels()) == fp
    finally:
        for c in comms:
            c.close()


def test_duplicates_across_shards_report_max_iters(pkg):
    sh = _sharded(pkg)
    keys = np.arange(1000, dtype=np.uint64)
    comms = sh.ShardComm.local_group([0, 0], 2000, 1.7)
    try:
        with pytest.raises(pkg.MphfPanic) as e:
            sh.build_in_threads(comms, 1.7, [keys, keys.copy()])  # every key appears in both shards
        assert "unique hashes" in str(e.value)
    finally:
        for c in comms:
            c.close()


def test_sharded_argument_errors(pkg):
    sh = _sharded(pkg)
    with pytest.raises(pkg.MphfPanic):
        sh.ShardComm.create(0, 0, 9, 1000)            # world > 8
    with pytest.raises(pkg.MphfPanic):
        sh.ShardComm.create(0, 2, 2, 1000)            # rank out of range
    c = sh.ShardComm.create(0, 0, 2, 1000)
    try:
        with pytest.raises(pkg.MphfPanic):            # not connected yet
            c.build(1.7, np.arange(10, dtype=np.uint64), 10)
    finally:
        c.close()


# ---- sharded build with the bucketed (shared-memory) marking kernels ---------------------------------------------
@pytest.fixture
def small_buckets(pkg):
    """force the bucketed kernels onto small levels so the unit-test sizes exercise them (and every transition)"""
    pkg.lib().bphf_set_build_mode(2, 3000)
    yield
    pkg.lib().bphf_set_build_mode(0, 0)


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("n", [2999, 5000, 9001, 70_000, 1_000_003, 3_000_000])
def test_sharded_bucketed_parity(pkg, oracle, small_buckets, world, n):
    ms = _run(pkg, oracle, oracle.keygen(n, seed=5 * n + world, kind=0), world)
    assert (ms[0].build_stats()["bucketed_levels"] > 0) == (n > 4096)


@pytest.mark.parametrize("gamma", [1.3, 2.0, 5.0, 17.5])
def test_sharded_bucketed_gammas(pkg, oracle, small_buckets, gamma):
    _run(pkg, oracle, oracle.keygen(300_001, seed=33, kind=1), 3, gamma=gamma)


def test_sharded_bucketed_unbalanced_shards_fail_cleanly_or_match(pkg, oracle, small_buckets):
    # all keys on the last shard: its bucket regions are sized for a 1/4 share -> overflow is reported as an error
    # on every shard (no hang, no corruption); the comm must then be discarded
    sh = _sharded(pkg)
    keys = oracle.keygen(200_000, seed=8, kind=0)
    comms = sh.ShardComm.local_group([0] * 4, len(keys), 1.7)
    try:
        shards = [keys[:0], keys[:0], keys[:0], keys]
        try:
            ms = sh.build_in_threads(comms, 1.7, shards)
        except pkg.MphfPanic as e:
            assert e.code in (-4, -5)
        else:
            assert_same_levels(ms[0].levels(), oracle.OracleMphf.new_parallel(1.7, keys).levels())
    finally:
        for c in comms:
            c.close()


def test_wrong_global_count_is_an_error_on_every_shard_not_a_hang(pkg, oracle):
    # the first barrier of level 0 carries each shard's key count: a sum that differs from n_global stops the build
    # with an error status on all shards (no timeout, no hang, no wrong structure)
    import threading
    import time
    sh = _sharded(pkg)
    keys = oracle.keygen(100_000, seed=77, kind=0)
    comms = sh.ShardComm.local_group([0, 0], 200_000, 1.7)
    errs = [None, None]

    def run(r):
        try:
            comms[r].build(1.7, keys[r * 50_000:(r + 1) * 50_000], 120_000)   # shards hold 100 000 keys in total
        except pkg.MphfPanic as e:
            errs[r] = e

    t0 = time.time()
    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    try:
        assert all(e is not None for e in errs) and time.time() - t0 < 3.0
        assert all(e.code in (-6, -4) for e in errs)
    finally:
        for c in comms:
            c.close()


# ---- sharded build with the exchange ("owner computes") kernels for the large levels ------------------------------------
@pytest.fixture
def small_xchg(pkg):
    """lower the exchange threshold so that unit-test sizes take the path 8 x 1e8 keys take (and every transition out of it)"""
    pkg.lib().bphf_set_xchg_bits(2.0e5)
    yield
    pkg.lib().bphf_set_xchg_bits(4.5e8)


def _xchg_levels(m):
    return m.build_stats()["bucketed_levels"] // 1000


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("n", [70_000, 300_001, 1_000_003, 3_000_000])
def test_sharded_exchange_parity(pkg, oracle, small_xchg, world, n):
    ms = _run(pkg, oracle, oracle.keygen(n, seed=7 * n + world, kind=0), world)
    # 1.7 n bits at level 0, 0.76 n at level 1, ...: levels above 2e5 bits are exchange levels
    want = sum(1 for k in ms[0].build_stats()["keys_in"][:4] if 1.7 * k > 2.0e5)
    assert _xchg_levels(ms[0]) == want and (want > 0) == (n > 120_000)


@pytest.mark.parametrize("how", ["first_empty", "skewed"])
def test_sharded_exchange_unbalanced_shards(pkg, oracle, small_xchg, how):
    ms = _run(pkg, oracle, oracle.keygen(900_001, seed=9, kind=0), 4, how=how)
    assert _xchg_levels(ms[0]) >= 2


@pytest.mark.parametrize("gamma", [1.02, 2.0, 5.0, 17.5])
def test_sharded_exchange_gammas(pkg, oracle, small_xchg, gamma):
    ms = _run(pkg, oracle, oracle.keygen(400_000, seed=21, kind=1), 3, gamma=gamma)
    # (gamma 17.5 leaves 2.8 % of the keys per level: two levels and the tail -- too few for an exchange level, which
    # needs two ordinary levels behind it)
    assert _xchg_levels(ms[0]) >= (1 if gamma <= 5.0 else 0)


@pytest.mark.parametrize("seed", [5, 31, 2 ** 64 - 3])
def test_sharded_exchange_starting_seed(pkg, oracle, small_xchg, seed):
    _run(pkg, oracle, oracle.keygen(500_000, seed=4, kind=0), 2, seed=seed)


def test_sharded_exchange_comm_is_reusable(pkg, oracle, small_xchg):
    sh = _sharded(pkg)
    world = 4
    comms = sh.ShardComm.local_group([0] * world, 2_000_000, 2.0)
    try:
        for n, gamma in [(2_000_000, 1.7), (654_321, 2.0), (2_000_000, 1.7), (5000, 1.7)]:
            keys = oracle.keygen(n, seed=n, kind=0)
            ms = sh.build_in_threads(comms, gamma, _split(keys, world, "even"))
            fp = oracle.fingerprint(oracle.OracleMphf.new_parallel(gamma, keys).levels())
            for m in ms:
                assert oracle.fingerprint(m.levels()) == fp
    finally:
        for c in comms:
            c.close()


def test_sharded_exchange_duplicates_across_shards_report_max_iters(pkg, small_xchg):
    sh = _sharded(pkg)
    keys = np.arange(400_000, dtype=np.uint64)
    comms = sh.ShardComm.local_group([0, 0], 800_000, 1.7)
    try:
        with pytest.raises(pkg.MphfPanic) as e:
            sh.build_in_threads(comms, 1.7, [keys, keys.copy()])  # every key appears in both shards
        assert "unique hashes" in str(e.value) or e.value.code in (-4, -5)   # or the shards' buffers give out first
    finally:
        for c in comms:
            c.close()


def test_sharded_exchange_all_keys_on_one_shard_fails_cleanly_or_matches(pkg, oracle, small_xchg):
    # the exchange blocks are sized for shards of up to twice the even share: one shard holding everything either still
    # fits (small sets) or is reported as an error on every shard -- no hang, no wrong structure
    sh = _sharded(pkg)
    keys = oracle.keygen(3_000_000, seed=8, kind=0)
    comms = sh.ShardComm.local_group([0] * 4, len(keys), 1.7)
    try:
        shards = [keys[:0], keys[:0], keys[:0], keys]
        try:
            ms = sh.build_in_threads(comms, 1.7, shards)
        except pkg.MphfPanic as e:
            assert e.code in (-4, -5)
        else:
            assert_same_levels(ms[0].levels(), oracle.OracleMphf.new_parallel(1.7, keys).levels())
    finally:
        for c in comms:
            c.close()


def test_sharded_exchange_wrong_global_count_is_an_error_not_a_hang(pkg, oracle, small_xchg):
    import threading
    import time
    sh = _sharded(pkg)
    keys = oracle.keygen(400_000, seed=77, kind=0)
    comms = sh.ShardComm.local_group([0, 0], 800_000, 1.7)
    errs = [None, None]

    def run(r):
        try:
            comms[r].build(1.7, keys[r * 200_000:(r + 1) * 200_000], 480_000)   # shards hold 400 000 keys in total
        except pkg.MphfPanic as e:
            errs[r] = e

    t0 = time.time()
    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    try:
        assert all(e is not None for e in errs) and time.time() - t0 < 5.0
        assert all(e.code in (-6, -4) for e in errs)
    finally:
        for c in comms:
            c.close()


def test_exchange_schedule_with_a_single_shard_matches_the_oracle(pkg, oracle, small_xchg, monkeypatch):
    # BPHF_XCHG_WORLD1: the one shard owns every slot and routes to itself -- the configuration the exchange kernels are
    # profiled in (profiles/r02_summary.md); same structure as any other build
    monkeypatch.setenv("BPHF_XCHG_WORLD1", "1")
    ms = _run(pkg, oracle, oracle.keygen(1_000_003, seed=12, kind=0), 1)
    assert _xchg_levels(ms[0]) >= 2
