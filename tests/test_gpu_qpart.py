"""Partitioned lookup (csrc/qpart_kernels.cuh: Mphf::hash / try_hash, src/lib.rs:324-355, for structures larger than the
L2) against the oracle and against the plain lookup kernel: the same ranks and the same Nones, whatever the slicing, the
number of stages, the batch size, and when a slice list overflows (the batch is then redone by the plain kernel)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stats(pkg, device=0):
    b, f, s = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    assert pkg.lib().bphf_query_partition_stats(device, C.byref(b), C.byref(f), C.byref(s)) == 0
    return b.value, f.value, s.value


@pytest.fixture
def part(pkg):
    """set(mode, slice_shift, walk_bytes, batch_keys, cap_scale); the defaults come back afterwards"""
    def set_(mode, slice_shift=0, walk_bytes=0, batch_keys=0, cap_scale=0.0):
        assert pkg.lib().bphf_set_query_partition(mode, slice_shift, walk_bytes, batch_keys, cap_scale) == 0
    yield set_
    assert pkg.lib().bphf_set_query_partition(0, 0, 0, 0, 0.0) == 0


def _queries(oracle, keys, n_absent, seed):
    rng = np.random.default_rng(seed)
    q = np.concatenate([keys, oracle.keygen(n_absent, seed=seed + 1000, kind=0)])
    rng.shuffle(q)
    return q


@pytest.mark.parametrize("n,slice_shift,walk_bytes,batch", [
    (5_000, 9, 64, 0),              # 17 slices wanted at level 0 -> a wider slice is taken (15 buckets at most)
    (300_000, 15, 4096, 0),         # 16 slices wanted -> 8 of twice the size, 4 stages, one batch
    (300_000, 17, 4096, 70_001),    # 4 buckets, batches that end inside a chunk
    (300_000, 18, 30_000, 2048),    # 2 buckets at level 0, 1 stage (level 1 is one slice), batch = one chunk
    (3_000_000, 19, 100_000, 1_000_000),
    (3_000_000, 22, 1 << 20, 0),    # 2 buckets, then the walk
])
def test_partitioned_lookup_equals_oracle_and_plain_kernel(pkg, oracle, part, n, slice_shift, walk_bytes, batch):
    keys = oracle.keygen(n, seed=n + 5, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    q = _queries(oracle, keys, n // 3 + 17, seed=n)
    expect = ref.hash_batch(q, nthreads=0)
    part(1)
    plain = m.hash_batch(q)
    assert np.array_equal(plain, expect)
    b0, f0, _ = _stats(pkg)
    part(2, slice_shift, walk_bytes, batch)
    got = m.hash_batch(q)
    b1, f1, stages = _stats(pkg)
    assert b1 > b0 and stages >= 1 and f1 == f0, "the partitioned kernels did not run"
    assert np.array_equal(got, expect)
    assert (got == np.uint64(pkg.NONE)).any()
    # in-set keys alone: a permutation of 0..n-1
    h = m.hash_batch(keys)
    assert np.array_equal(np.sort(h), np.arange(n, dtype=np.uint64))


def test_partitioned_lookup_of_absent_keys_only(pkg, oracle, part):
    keys = oracle.keygen(100_000, seed=3, kind=0)
    m = pkg.Mphf.new(1.7, keys)
    ref = oracle.OracleMphf.new(1.7, keys)
    absent = oracle.keygen(250_000, seed=4, kind=0)
    part(2, 14, 1024, 0)
    out = m.hash_batch(absent)
    assert np.array_equal(out, ref.hash_batch(absent, nthreads=0))
    none = out == np.uint64(pkg.NONE)
    assert none.sum() > 1000 and (out[~none] < 100_000).all()  # most absent keys land on a set bit of some level


@pytest.mark.parametrize("cap_scale", [0.05, 0.5])
def test_slice_list_overflow_falls_back_to_the_plain_walk_on_the_device(pkg, oracle, part, cap_scale):
    keys = oracle.keygen(400_000, seed=21, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    q = _queries(oracle, keys, 100_000, seed=9)
    expect = ref.hash_batch(q, nthreads=0)
    _, f0, _ = _stats(pkg)
    part(2, 16, 4096, 150_000, cap_scale)
    got = m.hash_batch(q)
    _, f1, _ = _stats(pkg)
    assert f1 > f0, "no batch overflowed: the test does not exercise the fallback"
    assert np.array_equal(got, expect)


def test_one_key_repeated_overflows_its_slice_and_is_still_answered(pkg, oracle, part):
    keys = oracle.keygen(200_000, seed=33, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    q = np.full(120_000, keys[77], dtype=np.uint64)
    q[::1000] = keys[:120]
    _, f0, _ = _stats(pkg)
    part(2, 15, 4096, 0)
    got = m.hash_batch(q)
    _, f1, _ = _stats(pkg)
    assert f1 > f0
    assert np.array_equal(got, ref.hash_batch(q, nthreads=0))


def test_partitioned_lookup_on_a_caller_stream_with_device_keys(pkg, oracle, part):
    import torch
    keys = oracle.keygen(1_000_003, seed=8, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    q = _queries(oracle, keys, 50_000, seed=2)
    dq = torch.from_numpy(q.view(np.int64)).cuda()
    part(2, 18, 8192, 300_000)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    outs = [m.hash_batch(dq, stream=st) for _ in range(3)]  # back to back on one stream: scratch is reused in order
    st.synchronize()
    expect = ref.hash_batch(q, nthreads=0)
    for o in outs:
        assert np.array_equal(o.cpu().numpy().view(np.uint64), expect)


def test_partitioned_lookup_of_u32_keys(pkg, oracle, part):
    rng = np.random.default_rng(5)
    keys = rng.choice(1 << 32, size=200_000, replace=False).astype(np.uint32)
    m = pkg.Mphf.new(1.7, keys)
    ref = oracle.OracleMphf.new(1.7, keys)
    q = np.concatenate([keys, rng.integers(0, 1 << 32, size=50_000, dtype=np.uint64).astype(np.uint32)])
    rng.shuffle(q)
    part(2, 15, 2048, 60_000)
    assert np.array_equal(m.hash_batch(q), ref.hash_batch(q))


def test_gamma_and_seed_do_not_matter_to_the_partition(pkg, oracle, part):
    keys = oracle.keygen(250_000, seed=12, kind=1)
    q = _queries(oracle, keys, 10_000, seed=13)
    for gamma, seed in ((1.0100001, None), (3.0, 5), (17.5, None)):
        m = pkg.Mphf.new_parallel(gamma, keys, seed)
        ref = oracle.OracleMphf.new_parallel(gamma, keys, seed)
        part(2, 15, 4096, 100_000)
        assert np.array_equal(m.hash_batch(q), ref.hash_batch(q, nthreads=0))


def test_automatic_mode_takes_the_partitioned_path_beyond_the_l2_and_agrees_with_the_plain_kernel(pkg, part):
    """natural scale: 4e8 keys -> 230 MB of packed lookup structure (the policy's threshold is 200 MB: below it the L2
    still serves the plain kernel better), 2^26 device-resident queries (in-set and absent); the policy picks the
    partitioned kernels on its own and the ranks equal the plain kernel's, value for value.  A 2e8-key structure
    (115 MB) must stay with the plain kernel."""
    import torch
    n, nq = 400_000_000, 1 << 26
    L = pkg.lib()
    keys = torch.empty(n, dtype=torch.int64, device="cuda")
    assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None) == 0
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    q = torch.empty(nq, dtype=torch.int64, device="cuda")
    half = nq // 2
    q[:half] = keys[:half]
    assert L.bphf_keygen_u64(C.c_void_p(q[half:].data_ptr()), n + 12345, nq - half, 1, 0, 0, None) == 0  # not in the set
    q = q[torch.randperm(nq, device="cuda")]
    del keys
    part(1)
    plain = m.hash_batch(q)
    torch.cuda.synchronize()
    b0, f0, _ = _stats(pkg)
    part(0)
    got = m.hash_batch(q)
    torch.cuda.synchronize()
    b1, f1, stages = _stats(pkg)
    assert b1 == b0 + 1 and f1 == f0 and stages >= 1, (b0, b1, f0, f1, stages, m.total_dims())
    assert torch.equal(got, plain)
    assert int((got == -1).sum()) > 1000  # (most absent keys land on a set bit of some level)
    del m, got, plain
    small_keys = torch.empty(200_000_000, dtype=torch.int64, device="cuda")
    assert L.bphf_keygen_u64(C.c_void_p(small_keys.data_ptr()), 0, 200_000_000, 1, 0, 0, None) == 0
    small = pkg.Mphf.new_parallel(1.7, small_keys, None)
    del small_keys
    b2, _, _ = _stats(pkg)
    small.hash_batch(q)
    torch.cuda.synchronize()
    assert _stats(pkg)[0] == b2, "a structure the L2 serves must stay with the plain kernel"


def test_in_place_lookup_keeps_to_the_plain_kernel_and_is_right(pkg, oracle, part):
    import torch
    keys = oracle.keygen(300_000, seed=41, kind=0)
    m = pkg.Mphf.new_parallel(1.7, keys, None)
    ref = oracle.OracleMphf.new_parallel(1.7, keys)
    q = _queries(oracle, keys, 20_000, seed=6)
    buf = torch.from_numpy(q.view(np.int64).copy()).cuda()
    part(2, 15, 4096, 0)
    b0, _, _ = _stats(pkg)
    m.hash_batch(buf, out=buf)          # out == keys
    torch.cuda.synchronize()
    b1, _, _ = _stats(pkg)
    assert b1 == b0, "an in-place batch must not go through the partitioned kernels (their fallback re-reads the keys)"
    assert np.array_equal(buf.cpu().numpy().view(np.uint64), ref.hash_batch(q, nthreads=0))
