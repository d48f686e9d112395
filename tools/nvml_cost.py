"""What does an NVML clock query cost the build loop?  A sampler thread queries SM clock + event reasons every
`period` ms while the main thread builds back to back; prints the duration of every query and the step-time
distribution.  (bench.py must sample clocks during its timed region; this is how its sampling period was chosen.)"""
import time, sys, os, ctypes as C, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch, pynvml as nv
pkg = ge.load_package(); L = pkg.lib()
n = 100_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
for _ in range(5): m = pkg.Mphf.new_parallel(1.7, keys, None)
t0 = time.perf_counter(); nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0); print("nvmlInit + handle: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
def q():
    a = time.perf_counter(); nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM); b = time.perf_counter()
    nv.nvmlDeviceGetCurrentClocksEventReasons(h); c = time.perf_counter()
    return (b - a) * 1e3, (c - b) * 1e3
print("first queries (idle GPU): clock %.3f ms, reasons %.3f ms" % q())
print("second queries (idle GPU): clock %.3f ms, reasons %.3f ms" % q())
def loop(period_ms, steps=60):
    stop = threading.Event(); calls = []
    def run():
        while not stop.is_set():
            calls.append(q())
            stop.wait(period_ms / 1e3)
    th = threading.Thread(target=run, daemon=True) if period_ms else None
    ts = []
    if th: th.start()
    for i in range(steps):
        a = time.perf_counter(); m = pkg.Mphf.new_parallel(1.7, keys, None); ts.append((time.perf_counter() - a) * 1e3)
    stop.set()
    if th: th.join()
    ts_sorted = sorted(ts)
    print("period %4s ms: step mean %.3f p50 %.3f max %.3f ms | %d queries, clock mean %.3f max %.3f ms, reasons mean %.3f max %.3f ms" % (
        period_ms or "off", sum(ts) / len(ts), ts_sorted[len(ts) // 2], ts_sorted[-1], len(calls),
        sum(c[0] for c in calls) / max(1, len(calls)), max([c[0] for c in calls] or [0]),
        sum(c[1] for c in calls) / max(1, len(calls)), max([c[1] for c in calls] or [0])))
for p in (0, 500, 50, 10, 2, 0):
    loop(p)
