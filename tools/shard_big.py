import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import __graft_entry__ as ge, torch, importlib
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
n = int(float(sys.argv[1])); world = 2
comms = sh.ShardComm.local_group([0] * world, n * world, 1.7)
shards = []
for r in range(world):
    k = torch.empty(n, dtype=torch.int64, device="cuda")
    L.bphf_keygen_u64(C.c_void_p(k.data_ptr()), r * n, n, 1, 0, 0, None)
    shards.append(k)
torch.cuda.synchronize()
try:
    ms = sh.build_in_threads(comms, 1.7, shards)
    st = ms[0].build_stats()
    print("ok levels", st["n_levels"], st["keys_in"][:6], "total_ms", st["total_ms"])
except Exception as e:
    print("ERR", e)
