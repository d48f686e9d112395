"""ncu target for the exchange (owner computes) kernels on ONE GPU: a one-shard comm with BPHF_XCHG_WORLD1=1 runs the
sharded schedule with itself as the only owner -- per-GPU work of an N x 1e8-key build (1e8 keys routed, 1e8 offsets
marked and answered), stores local instead of NVLink.   BPHF_XCHG_WORLD1=1 BPHF_XCHG_BITS=1e8 python tools/prof_xchg_world1.py [n]"""
import ctypes as C, os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BPHF_XCHG_WORLD1", "1")
os.environ.setdefault("BPHF_XCHG_BITS", "1e8")
import __graft_entry__ as ge
import torch
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
comm = sh.ShardComm.create(0, 0, 1, n, 1.7)
L.bphf_set_profile(2)
for _ in range(2):
    m = comm.build(1.7, keys, n)
st = m.build_stats()
one = pkg.Mphf.new_parallel(1.7, keys, None)
same = all((a[0] == b[0]) and (a[1] == b[1]).all() and (a[2] == b[2]).all() for a, b in zip(m.levels(), one.levels()))
print("exchange levels %d, total %.3f ms, xscatter %s, owner phase %s, identical to the plain build: %s"
      % (st["bucketed_levels"] // 1000, st["total_ms"], ["%.3f" % x for x in st["pass_ms"][:3]], ["%.3f" % x for x in st["fin_ms"][:3]], same))
del m, one
comm.close()
