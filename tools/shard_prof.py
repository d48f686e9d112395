"""torchrun target: sharded build timing with per-level breakdown (profile 2). usage: shard_prof.py <keys_per_gpu> [reps]"""
import ctypes as C, os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
n = int(float(sys.argv[1])); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), rank * n, n, 1, 0, lr, None)
comm = sh.ShardComm.from_process_group(lr, n * world, 1.7)
for prof in (0, 2):
    L.bphf_set_profile(prof)
    for r in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = comm.build(1.7, keys, n * world)
        dt = time.perf_counter() - t0
        if rank == 0:
            st = m.build_stats()
            print("profile", prof, "rep", r, "wall_ms %.3f total_ms %.3f launches %d levels %d" % (dt * 1e3, st["total_ms"], st["kernel_launches"], st["n_levels"]), flush=True)
if rank == 0:
    st = m.build_stats()
    print("pass_ms ", ["%.3f" % x for x in st["pass_ms"]])
    print("merge_ms", ["%.3f" % x for x in st["merge_ms"]])
    print("fin_ms  ", ["%.3f" % x for x in st["fin_ms"]])
    print("tail_ms %.3f ranks_ms %.3f" % (st["tail_ms"], st["ranks_ms"]))
del m
torch.cuda.synchronize(); dist.barrier(); comm.close(); dist.destroy_process_group()
