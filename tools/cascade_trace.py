"""Per-CTA phase timestamps of the persistent cascade kernel (library built with -DBPHF_TRACE):
    BPHF_LIB_VARIANT=trace BPHF_NVCC_EXTRA=-DBPHF_TRACE python tools/cascade_trace.py [n]
For every level: when the CTAs entered the filter pass, how long each CTA's pass took (min / median / max), how long
they then waited at the grid barrier, the finalize step and the second barrier."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import numpy as np
import torch

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
pkg = ge.load_package()
pkg.build_native()
L = pkg.lib()
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
for _ in range(3):
    m = pkg.Mphf.new_parallel(1.7, keys, None)
st = m.build_stats()
LEVELS, POINTS, CTAS = 16, 5, 1024
buf = np.zeros((LEVELS, POINTS, CTAS), dtype=np.uint64)
fn = L.bphf_debug_cascade_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_uint64]
assert fn(buf.ctypes.data, buf.nbytes) == 0
t = buf.astype(np.int64)
ctas = int((t[1, 0] != 0).sum())
print("CTAs", ctas, "cascade_ms %.3f" % st["cascade_ms"], "keys_in", st["keys_in"][:14])
t0 = t[1, 0, :ctas].min()
print("level  keys      start_us  pass(min/med/max)us      wait1(min/med/max)   finalize(med) wait2(med)  level_total")
for l in range(1, LEVELS):
    if t[l, 0, 0] == 0:
        break
    x = t[l, :, :ctas].astype(np.float64) / 1e3
    p = x[1] - x[0]
    w1 = x[2] - x[1]
    f = x[3] - x[2]
    w2 = x[4] - x[3]
    print("%3d %10d %9.1f   %7.1f %7.1f %7.1f     %7.1f %7.1f %7.1f   %7.1f %9.1f %10.1f" % (
        l, st["keys_in"][l - 1], (x[0].min() - t0 / 1e3), p.min(), np.median(p), p.max(), w1.min(), np.median(w1), w1.max(),
        np.median(f), np.median(w2), x[4].max() - x[0].min()))

