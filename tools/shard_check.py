"""torchrun target -- BASELINE configs[2] as a correctness check: ONE Mphf over world * n keys built sharded across the
GPUs must be word for word the Mphf a single GPU builds from the same keys (a different code path: no merge, no
exchange), and hash the keys to a permutation of 0..world*n-1.  usage: shard_check.py <keys_per_gpu>
(1.25e8 on 8 GPUs = the 1e9-key target of north_star)"""
import ctypes as C, os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
n = int(float(sys.argv[1])); total = n * world
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), rank * n, n, 1, 0, lr, None)
comm = sh.ShardComm.from_process_group(lr, total, 1.7)
best = 1e9
for r in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = comm.build(1.7, keys, total)
    best = min(best, time.perf_counter() - t0)
st = m.build_stats()


def export(mphf):
    tw, tr, _ = mphf.total_dims()
    w = torch.empty(tw, dtype=torch.int64, device="cuda"); r = torch.empty(tr, dtype=torch.int64, device="cuda")
    assert L.bphf_export_all(mphf._h, C.c_void_p(w.data_ptr()), C.c_void_p(r.data_ptr()), 1) == 0
    return w, r


# every replica identical: order-free checksum of (words, ranks) must agree across ranks
w, r = export(m)
sig = torch.stack([w.sum(), (w ^ (w >> 7)).sum(), r.sum(), torch.tensor(w.numel(), device="cuda")])
sigs = [torch.empty_like(sig) for _ in range(world)]
dist.all_gather(sigs, sig)
replicas_equal = all(bool((s == sigs[0]).all().item()) for s in sigs)
if rank == 0:
    allk = torch.empty(total, dtype=torch.int64, device="cuda")
    L.bphf_keygen_u64(C.c_void_p(allk.data_ptr()), 0, total, 1, 0, lr, None)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one = pkg.Mphf.new_parallel(1.7, allk, None)
    t_one = time.perf_counter() - t0
    w1, r1 = export(one)
    same = w.numel() == w1.numel() and bool(torch.equal(w, w1)) and bool(torch.equal(r, r1))
    dims_same = [m.level_dims(l) for l in range(m.num_levels)] == [one.level_dims(l) for l in range(one.num_levels)]
    del w1, r1, one
    out = torch.empty(total, dtype=torch.int64, device="cuda")
    m.hash_batch(allk, out=out)
    bad = C.c_uint64(1)
    assert L.bphf_check_permutation(C.c_void_p(out.data_ptr()), total, lr, C.byref(bad)) == 0
    print("sharded build of %d keys on %d GPUs: %.2f ms wall (device %.2f ms, %.1f Gkeys/s), %d levels | single-GPU build of the same keys %.1f ms | "
          "identical words+ranks: %s, identical level dims: %s, replicas identical: %s, hash is a permutation of 0..n-1: %s"
          % (total, world, best * 1e3, st["total_ms"], total / best / 1e9, m.num_levels, t_one * 1e3, same, dims_same,
             replicas_equal, bad.value == 0), flush=True)
    ok = same and dims_same and replicas_equal and bad.value == 0
else:
    ok = True
del m
torch.cuda.synchronize(); dist.barrier(); comm.close(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
