import time, sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch, pynvml as nv
pkg = ge.load_package(); L = pkg.lib()
n = 100_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
for _ in range(3): m = pkg.Mphf.new_parallel(1.7, keys, None)
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
def timed(label, f):
    # build loop with call f() injected between builds
    t0 = time.perf_counter()
    for i in range(20):
        m = pkg.Mphf.new_parallel(1.7, keys, None)
        if f: 
            c0 = time.perf_counter(); f(); c1 = time.perf_counter()
    dt = (time.perf_counter() - t0) / 20 * 1e3
    print("%-40s %.3f ms/build" % (label, dt), ("call %.3f ms" % ((c1 - c0) * 1e3)) if f else "")
timed("no nvml", None)
timed("clockinfo SM", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
timed("event reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h))
timed("power", lambda: nv.nvmlDeviceGetPowerUsage(h))
timed("no nvml", None)
