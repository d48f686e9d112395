"""Streamed ingest at size: n u64 keys in PINNED HOST memory, built (a) in core -- keys copied whole -- and
(b) streamed through the device ring (bphf_set_ingest_mode(1)); prints wall time, the device-memory high-water mark of
each, checks both structures are identical and that lookups (streamed from the host too) are a permutation.
With a third argument G the tool finally occupies G GB of device memory and builds once more in the AUTOMATIC ingest mode:
the library must notice that an in-core build no longer fits and stream by itself.
usage: out_of_core.py [n] [ring_keys] [hog_gb]"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch, numpy as np
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ring = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
host = torch.empty(n, dtype=torch.int64, pin_memory=True)
piece = 1 << 27
tmp = torch.empty(min(piece, n), dtype=torch.int64, device="cuda")
for lo in range(0, n, piece):
    cnt = min(piece, n - lo)
    L.bphf_keygen_u64(C.c_void_p(tmp.data_ptr()), lo, cnt, 1, 0, 0, None)
    host[lo:lo + cnt].copy_(tmp[:cnt])
torch.cuda.synchronize(); del tmp; torch.cuda.empty_cache()
hk = host.numpy().view(np.uint64)
res = {}
for name, mode in (("in core", 2), ("streamed", 1)):
    L.bphf_set_ingest_mode(mode, ring)
    L.bphf_set_profile(2)  # keeps the memory high-water mark for the in-core build too
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        m = pkg.Mphf.new_parallel(1.7, hk, None)
        best = min(best, time.perf_counter() - t0)
        st = m.build_stats()
        if _ == 0:
            del m
    tw, tr, _k = m.total_dims()
    w = torch.empty(tw, dtype=torch.int64, device="cuda"); r = torch.empty(tr, dtype=torch.int64, device="cuda")
    assert L.bphf_export_all(m._h, C.c_void_p(w.data_ptr()), C.c_void_p(r.data_ptr()), 1) == 0
    res[name] = (best, st, w, r, m)
    print("%-9s: %8.1f ms (%.2f Gkeys/s), device memory high-water %.2f GB, streamed_ingest=%d, levels %d"
          % (name, best * 1e3, n / best / 1e9, st["peak_device_bytes"] / 1e9, st["streamed_ingest"], st["n_levels"]), flush=True)
L.bphf_set_ingest_mode(0, 0)
L.bphf_set_profile(0)
same = bool(torch.equal(res["in core"][2], res["streamed"][2])) and bool(torch.equal(res["in core"][3], res["streamed"][3]))
m = res["streamed"][4]
out = torch.empty(min(n, 200_000_000), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
t0 = time.perf_counter(); m.hash_batch(hk[:len(out)], out=out); tl = time.perf_counter() - t0
chk = torch.from_numpy(out.view(np.int64)).cuda()
distinct = int(torch.unique(chk).numel()) == len(out) and int(chk.max()) < n and int(chk.min()) >= 0
print("identical structures: %s | lookup of %d host keys %.1f ms, distinct and in range: %s" % (same, len(out), tl * 1e3, distinct))

if len(sys.argv) > 3:
    for k in list(res):
        del res[k]
    del m, chk
    torch.cuda.empty_cache()
    hog = torch.empty(int(float(sys.argv[3]) * 1e9), dtype=torch.uint8, device="cuda")
    free_b, total_b = torch.cuda.mem_get_info()
    t0 = time.perf_counter(); m = pkg.Mphf.new_parallel(1.7, hk, None); dt = time.perf_counter() - t0
    st = m.build_stats()
    print("automatic mode with %.0f GB of the device occupied (cudaMemGetInfo free %.1f GB): %.1f ms, streamed_ingest=%d, high-water %.2f GB"
          % (hog.numel() / 1e9, free_b / 1e9, dt * 1e3, st["streamed_ingest"], st["peak_device_bytes"] / 1e9))
