import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])); mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L.bphf_set_build_mode(mode, 0)
k = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(k.data_ptr()), 0, n, 1, 0, 0, None)
torch.cuda.synchronize()
for _ in range(2):
    m = pkg.Mphf.new_parallel(1.7, k, None)
st = m.build_stats()
print("levels", st["n_levels"], "syncs", st["host_syncs"], "launches", st["kernel_launches"], "total_ms %.3f" % st["total_ms"])
print(st["keys_in"])
print([b // 64 for b in st["bits"]])
