import ctypes as C, sys, os, time
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = 100_000_000
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, 0, None)
stream = torch.cuda.current_stream()
def build(tag):
    t0 = time.perf_counter(); m = pkg.Mphf.new_parallel(1.7, keys, None, stream=stream); dt = (time.perf_counter() - t0) * 1e3
    if dt > 3.2 or tag: sys.stderr.write("%-28s build %.2f ms\n" % (tag, dt))
    return m
mode = sys.argv[1]
for i in range(30): m = build("")
sys.stderr.write("-- phase 1: 30 plain builds done; now mode %s\n" % mode)
for i in range(30):
    m = build("")
    if mode == "tensor_item":
        x = torch.tensor([1], dtype=torch.int32, device="cuda"); x.item()
    elif mode == "tensor_only":
        x = torch.tensor([1], dtype=torch.int32, device="cuda")
    elif mode == "item_prealloc":
        if i == 0: y = torch.zeros(1, dtype=torch.int32, device="cuda")
        y.item()
for i in range(3): m = build("after-mode")
sys.stderr.write("-- torch.cuda.synchronize()\n"); torch.cuda.synchronize()
for i in range(3): m = build("after-sync-1")
sys.stderr.write("-- torch.cuda.synchronize()\n"); torch.cuda.synchronize()
for i in range(3): m = build("after-sync-2")
