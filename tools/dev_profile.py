"""dev helper: per-kernel CUDA-event times of one build (bphf_set_profile(2)) + lookup timing."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
gamma = float(sys.argv[2]) if len(sys.argv) > 2 else 1.7
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
min_keys = int(float(sys.argv[5])) if len(sys.argv) > 5 else 0
pkg = ge.load_package()
L = pkg.lib()
L.bphf_set_build_mode(mode, min_keys)
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
torch.cuda.synchronize()
for prof in (0, 2):
    L.bphf_set_profile(prof)
    for r in range(reps):
        t0 = time.perf_counter()
        m = pkg.Mphf.new_parallel(gamma, keys, None)
        dt = time.perf_counter() - t0
        st = m.build_stats()
        print("profile", prof, "rep", r, "wall_ms %.3f total_ms %.3f launches %d syncs %d levels %d tail_level %d bucketed %d rebuilds %d" % (
            dt * 1e3, st["total_ms"], st["kernel_launches"], st["host_syncs"], st["n_levels"], st["tail_level"], st["bucketed_levels"], st["rebuilds"]))
st = m.build_stats()
print("pass_ms", ["%.3f" % x for x in st["pass_ms"]])
print("fin_ms ", ["%.3f" % x for x in st["fin_ms"]])
print("tail_ms %.3f ranks_ms %.3f cascade_ms %.3f sum %.3f" % (st["tail_ms"], st["ranks_ms"], st["cascade_ms"], sum(st["pass_ms"]) + sum(st["fin_ms"]) + st["tail_ms"] + st["ranks_ms"] + st["cascade_ms"]))
print("keys_in", st["keys_in"])
out = torch.empty_like(keys)
for r in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.hash_batch(keys, out=out)
    torch.cuda.synchronize()
    print("lookup ms %.3f" % ((time.perf_counter() - t0) * 1e3))
