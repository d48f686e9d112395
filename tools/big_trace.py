import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = 1_000_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
L.bphf_set_profile(2)
for i in range(4):
    t0 = time.perf_counter(); m = pkg.Mphf.new_parallel(1.7, keys, None); dt = time.perf_counter() - t0
    st = m.build_stats()
    sys.stderr.write("build %d wall %.1f ms dev %.1f ms | pass %s | fin %s | cascade %.2f tail %.2f ranks %.2f launches %d syncs %d bucketed %d rebuilds %d\n" % (
        i, dt * 1e3, st["total_ms"], ["%.1f" % x for x in st["pass_ms"][:4]], ["%.1f" % x for x in st["fin_ms"][:4]], st["cascade_ms"], st["tail_ms"], st["ranks_ms"],
        st["kernel_launches"], st["host_syncs"], st["bucketed_levels"], st["rebuilds"]))
    if i == 1: del m
