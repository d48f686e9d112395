"""Leak check: free device memory must not drift over many build / lookup / map / sharded / chunked cycles."""
import os, sys, importlib
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
keys = np.unique(np.random.default_rng(7).integers(0, 2**63, 2_100_000, dtype=np.uint64))[:2_000_000]
k32 = np.unique(np.random.default_rng(1).integers(0, 2**32, 500_000, dtype=np.uint64)).astype(np.uint32)
def cycle():
    for mode in (0, 1, 2):
        L.bphf_set_build_mode(mode, 3000 if mode == 2 else 0)
        m = pkg.Mphf.new_parallel(1.7, keys, None); m.hash_batch(keys[:100000]); m.levels(); m.close()
    L.bphf_set_build_mode(0, 0)
    m = pkg.Mphf.from_chunked_iterator_parallel(2.0, np.array_split(keys, 3), None, len(keys)); m.close()
    m = pkg.Mphf.new(1.7, k32); m.hash_batch(k32[:1000]); m.close()
    bm = pkg.BoomHashMap.new_parallel(keys[:300000], keys[:300000]); bm.get_batch(keys[:1000]); bm.mphf.close()
    comms = sh.ShardComm.local_group([0, 0], len(keys), 1.7)
    ms = sh.build_in_threads(comms, 1.7, np.array_split(keys, 2))
    for x in ms: x.close()
    for c in comms: c.close()
    try:
        pkg.Mphf.new_parallel(1.7, np.tile(keys[:1000], 3), None)   # duplicates: error path
    except pkg.MphfPanic:
        pass
for _ in range(3): cycle()
torch.cuda.synchronize(); free0 = torch.cuda.mem_get_info()[0]
for _ in range(30): cycle()
torch.cuda.synchronize(); free1 = torch.cuda.mem_get_info()[0]
print("free before %.1f MB, after 30 more cycles %.1f MB, drift %.1f MB" % (free0 / 1e6, free1 / 1e6, (free0 - free1) / 1e6))
assert free0 - free1 < 64e6
