"""Full-size property check: build over n keys on one GPU, hash all of them, verify the outputs are exactly 0..n-1
and that every level has max(255, floor(gamma * keys_left)) bits (src/lib.rs:369,384)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])); gamma = float(sys.argv[2]) if len(sys.argv) > 2 else 1.7
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
torch.cuda.synchronize()
t0 = time.perf_counter(); m = pkg.Mphf.new_parallel(gamma, keys, None); t1 = time.perf_counter()
out = torch.empty_like(keys)
m.hash_batch(keys, out=out); torch.cuda.synchronize(); t2 = time.perf_counter()
bad = C.c_uint64(1)
assert L.bphf_check_permutation(C.c_void_p(out.data_ptr()), n, 0, C.byref(bad)) == 0
st = m.build_stats()
sizes_ok = [max(255, int(gamma * float(k))) for k in st["keys_in"]] == st["bits"]  # src/lib.rs:369,384
print("n %d gamma %.2f: build %.1f ms, lookup %.1f ms, levels %d, bucketed %d, rebuilds %d, syncs %d, not-a-permutation count %d, level sizes ok %s"
      % (n, gamma, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["n_levels"], st["bucketed_levels"], st["rebuilds"], st["host_syncs"], bad.value, sizes_ok))
assert bad.value == 0 and sizes_ok and st["rebuilds"] == 0
