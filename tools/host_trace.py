"""Host-side timeline of back-to-back builds (BPHF_HOST_TRACE=1 makes the library print one line per build), with a
device-wide synchronize injected before some of them: what does the first build after torch.cuda.synchronize() pay?"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = 100_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
L.bphf_set_profile(1)
stream = torch.cuda.Stream()
for i in range(40):
    if i in (20, 30):
        sys.stderr.write("-- torch.cuda.synchronize() --\n"); torch.cuda.synchronize()
    if i == 35:
        sys.stderr.write("-- stream.synchronize() + sleep 5 ms --\n"); stream.synchronize(); time.sleep(0.005)
    t0 = time.perf_counter()
    m = pkg.Mphf.new_parallel(1.7, keys, None, stream=stream)
    t1 = time.perf_counter()
    if i >= 17:
        sys.stderr.write("python: call %.0f us\n" % ((t1 - t0) * 1e6))
