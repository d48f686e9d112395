"""ncu target: warm build, then one build + one lookup over n keys (short; profile with -k regex filters)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
gamma = float(sys.argv[2]) if len(sys.argv) > 2 else 1.7
pkg = ge.load_package()
L = pkg.lib()
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
out = torch.empty_like(keys)
for _ in range(2):
    m = pkg.Mphf.new_parallel(gamma, keys, None)
    m.hash_batch(keys, out=out)
torch.cuda.synchronize()
print("ok", m.num_levels)
