"""BASELINE configs[3]: batched Mphf::hash of n in-set u64 keys (default 1e9) against a prebuilt gamma = 1.7 MPHF on one
GPU, keys and results device resident; warm build + timed lookup passes with the plain kernel (every probe of the big
levels is a random DRAM sector) and with the partitioned kernels (csrc/qpart_kernels.cuh); outputs checked to be equal
and a permutation of 0..n-1.  argv: n [slice_shift walk_bytes batch_keys]; per-kernel times with BPHF_Q_PROFILE=1."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
tune = [int(float(a)) for a in sys.argv[2:5]] + [0] * 3
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
m = pkg.Mphf.new_parallel(1.7, keys, None)
t0 = time.perf_counter(); m2 = pkg.Mphf.new_parallel(1.7, keys, None); tb = time.perf_counter() - t0
del m2
out = torch.empty_like(keys)
tw, tr, _ = m.total_dims()
st = m.build_stats()
print("config 4: %d keys, %d levels, structure %.0f MB (words + ranks), build %.1f ms (%.1f Gkeys/s)"
      % (n, st["n_levels"], 8 * (tw + tr) / 1e6, tb * 1e3, n / tb / 1e9))


def timed(label):
    m.hash_batch(keys, out=out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); m.hash_batch(keys, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    bad = C.c_uint64(1)
    assert L.bphf_check_permutation(C.c_void_p(out.data_ptr()), n, 0, C.byref(bad)) == 0
    print("  %-12s %.2f ms = %.1f Gkeys/s (%.0f GB/s of keys + results + structure), permutation: %s"
          % (label, best * 1e3, n / best / 1e9, (16.0 * n + 8 * (tw + tr)) / best / 1e9, bad.value == 0), flush=True)
    return best


assert L.bphf_set_query_partition(1, 0, 0, 0, 0.0) == 0
timed("plain")
plain = out.clone()
assert L.bphf_set_query_partition(0, tune[0], tune[1], tune[2], 0.0) == 0
timed("partitioned")
b, f, s = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
L.bphf_query_partition_stats(0, C.byref(b), C.byref(f), C.byref(s))
print("  partitioned batches %d, fallbacks %d, stages %d; equal to the plain kernel's output: %s"
      % (b.value, f.value, s.value, bool(torch.equal(out, plain))))
