"""Quick smoke over every construction strategy (L2-atomic cascade, per-level, bucketed, chunked, sharded) against the oracle."""
import os, sys, importlib
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); L = pkg.lib()
def same(a, b):
    return len(a) == len(b) and all(x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) for x, y in zip(a, b))
for mode, mk in ((0, 0), (1, 0), (2, 3000)):
    L.bphf_set_build_mode(mode, mk)
    for n in (0, 1, 150, 5000, 70_000, 400_000):
        keys = orc.keygen(n, seed=n + 1)
        m = pkg.Mphf.new_parallel(1.7, keys, None)
        assert same(m.levels(), orc.OracleMphf.new_parallel(1.7, keys).levels()), (mode, n)
        if n:
            assert np.array_equal(m.hash_batch(keys), orc.OracleMphf.new_parallel(1.7, keys).hash_batch(keys))
    print("mode", mode, "ok", flush=True)
L.bphf_set_build_mode(0, 0)
keys = orc.keygen(300_000, seed=5)
m = pkg.Mphf.from_chunked_iterator_parallel(2.0, np.array_split(keys, 3), None, len(keys))
assert same(m.levels(), orc.OracleMphf.from_chunked_iterator_parallel(2.0, [keys], None, len(keys)).levels())
print("chunked ok", flush=True)
sh = importlib.import_module(pkg.__name__ + ".sharded")
for mode, mk in ((0, 0), (2, 3000)):
    L.bphf_set_build_mode(mode, mk)
    comms = sh.ShardComm.local_group([0, 0, 0], 200_000, 1.7)
    keys = orc.keygen(200_000, seed=9)
    ms = sh.build_in_threads(comms, 1.7, np.array_split(keys, 3))
    assert same(ms[1].levels(), orc.OracleMphf.new_parallel(1.7, keys).levels())
    for c in comms: c.close()
    print("sharded mode", mode, "ok", flush=True)
