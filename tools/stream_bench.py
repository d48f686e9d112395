"""Stream keys at size: n distinct Vec<u8> keys of `width` bytes each (write_usize(width) + one write of
the bytes = 2 recorded writes per key), tables resident on the device; build and batched lookup through the C ABI,
checked to be a permutation of 0..n-1.  usage: stream_bench.py [n] [width]"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge, torch
pkg = ge.load_package(); L = pkg.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = "cuda"
# key i: 8-byte prefix (= width), then `width` bytes whose first 8 are a bijection of i (distinct by construction)
body = torch.randint(0, 256, (n, width), dtype=torch.uint8, device=dev)
ids = torch.empty(n, dtype=torch.int64, device=dev)
L.bphf_keygen_u64(C.c_void_p(ids.data_ptr()), 0, n, 7, 0, 0, None)
body[:, :8] = ids.view(torch.uint8).view(n, 8)
rec = torch.empty((n, 8 + width), dtype=torch.uint8, device=dev)
rec[:, :8] = torch.tensor([width], dtype=torch.int64, device=dev).view(torch.uint8).view(1, 8)
rec[:, 8:] = body
data = rec.reshape(-1).contiguous()
base = torch.arange(n, dtype=torch.int64, device=dev) * (8 + width)
seg_ends = torch.stack([base + 8, base + 8 + width], dim=1).reshape(-1).contiguous()
key_segs = (torch.arange(n + 1, dtype=torch.int64, device=dev) * 2).contiguous()
out = torch.empty(n, dtype=torch.int64, device=dev)
args = (C.c_void_p(data.data_ptr()), data.numel(), C.c_void_p(seg_ends.data_ptr()), seg_ends.numel(),
        C.c_void_p(key_segs.data_ptr()), n)
def build():
    h = C.c_void_p()
    rc = L.bphf_build_stream(*args, 1.7, 0, 0, 0, 1, C.byref(h))
    assert rc == 0, pkg._native.last_error() if hasattr(pkg, "_native") else rc
    return pkg.Mphf(h.value, 0, "Vec<u8>")
def timed(f, reps=3):
    best, r = 1e9, None
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, r
tb, m = timed(build)
def look():
    assert L.bphf_hash_batch_stream(m._h, *args, C.c_void_p(out.data_ptr()), 1, None) == 0
tl, _ = timed(look)
ok = bool((torch.sort(out).values == torch.arange(n, dtype=torch.int64, device=dev)).all().item())
st = m.build_stats()
print("stream keys: n %d Vec<u8> of %d bytes (%.0f MB of writes) | levels %d | build %.2f ms (%.2f Gkeys/s) | lookup %.2f ms (%.2f Gkeys/s) | permutation: %s"
      % (n, width, data.numel() / 1e6, m.num_levels, tb * 1e3, n / tb / 1e9, tl * 1e3, n / tl / 1e9, ok))
