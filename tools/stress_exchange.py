"""stress: the 8-logical-shard exchange build over 3e6 keys, repeated; prints the first error text"""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import numpy as np
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
L.bphf_set_xchg_bits(2.0e5)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
import torch
n = 3_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(keys.data_ptr(), 0, n, 21000008, 0, 0, None)
hk = keys.cpu().numpy().view(np.uint64)
one = pkg.Mphf.new_parallel(1.7, keys, None)
tw, tr, _ = one.total_dims()
ref = [np.concatenate([w for _, w, _ in one.levels()]), np.concatenate([r for _, _, r in one.levels()])]
b = [i * n // world for i in range(world + 1)]
bad = 0
for rep in range(reps):
    comms = sh.ShardComm.local_group([0] * world, n, 1.7)
    try:
        ms = sh.build_in_threads(comms, 1.7, [hk[b[r]:b[r + 1]] for r in range(world)])
        for r, m in enumerate(ms):
            lv = m.levels()
            w = np.concatenate([x for _, x, _ in lv]); rk = np.concatenate([x for _, _, x in lv])
            if len(w) != len(ref[0]) or not np.array_equal(w, ref[0]) or not np.array_equal(rk, ref[1]):
                bad += 1
                print("rep", rep, "shard", r, "DIFFERS: words equal", len(w) == len(ref[0]) and np.array_equal(w, ref[0]),
                      "ranks equal", len(rk) == len(ref[1]) and np.array_equal(rk, ref[1]), flush=True)
    except Exception as e:
        bad += 1
        print("rep", rep, "ERROR", repr(e), flush=True)
    finally:
        for c in comms:
            c.close()
print("done, bad =", bad)
