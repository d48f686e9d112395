"""Profiling target for the partitioned lookup: an MPHF over n keys (default 1e9), then `reps` lookups of nq device-
resident in-set keys (default 2^27 = one batch).  Under ncu:
  ncu --kernel-name regex:'qsplit|qwalk|qrestore|qfallback|query_sync' --launch-skip 9 -c 9 --metrics ... python tools/prof_qpart.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 27
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tune = [int(float(a)) for a in sys.argv[4:7]] + [0] * 3
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
m = pkg.Mphf.new_parallel(1.7, keys, None)
q = keys[torch.randperm(n, device="cuda")[:nq]] if nq < n else keys
out = torch.empty_like(q)
assert L.bphf_set_query_partition(0, tune[0], tune[1], tune[2], 0.0) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record(); m.hash_batch(q, out=out); e1.record(); torch.cuda.synchronize()
    print("lookup %d: %.3f ms (%.1f Gkeys/s)" % (r, e0.elapsed_time(e1), nq / e0.elapsed_time(e1) / 1e6), flush=True)
