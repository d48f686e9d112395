"""BASELINE config 5: BoomHashMap over n 2-bit-packed 31-mer u64 keys (62-bit bijection), gamma sweep.
Build the Mphf with new_parallel(gamma), create_map as an out-of-place permutation (bphf_map_permute_u64), batched
get of all keys (bphf_map_get_batch_u64); everything device resident.  gamma = 1.0 must be rejected like the
reference's assert!(gamma > 1.01)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 3, 1, 0, None)
vals = torch.arange(n, dtype=torch.int64, device="cuda")
ko = torch.empty_like(keys); vo = torch.empty_like(keys); got = torch.empty_like(keys)
found = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
try:
    pkg.Mphf.new_parallel(1.0, keys, None)
    print("gamma 1.0: NOT rejected")
except pkg.MphfPanic as e:
    print("gamma 1.0: rejected (%s)" % e)
def timed(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, r
for gamma in (1.7, 2.0, 5.0):
    tb, m = timed(lambda: pkg.Mphf.new_parallel(gamma, keys, None))
    tw, tr, _ = m.total_dims()
    def permute():
        assert L.bphf_map_permute_u64(m._h, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), n,
                                      C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), 1, None) == 0
    def get():
        assert L.bphf_map_get_batch_u64(m._h, C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), n,
                                        C.c_void_p(keys.data_ptr()), n, C.c_void_p(got.data_ptr()),
                                        C.c_void_p(found.data_ptr()), 1, None) == 0
    tp, _ = timed(permute)
    tg, _ = timed(get)
    ok = bool(found.all().item()) and bool((got == vals).all().item())
    print("gamma %.1f: levels %d bits/key %.2f (with ranks %.2f) | build %.1f ms (%.1f Gkeys/s) | create_map %.1f ms (%.1f Gkeys/s) | get %.1f ms (%.1f Gkeys/s) | all found and values right: %s"
          % (gamma, m.num_levels, 64.0 * tw / n, 64.0 * (tw + tr) / n, tb * 1e3, n / tb / 1e9, tp * 1e3, n / tp / 1e9, tg * 1e3, n / tg / 1e9, ok))
    del m
