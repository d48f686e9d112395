"""torchrun target: back-to-back sharded builds WITHOUT a barrier in between (the bench pattern); prints wall and device time"""
import ctypes as C, os, sys, time, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = ge.load_package(); L = pkg.lib()
sh = importlib.import_module(pkg.__name__ + ".sharded")
n = int(float(sys.argv[1])); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
keep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
keys = torch.empty(n, dtype=torch.int64, device="cuda")
L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), rank * n, n, 1, 0, lr, None)
comm = sh.ShardComm.from_process_group(lr, n * world, 1.7)
L.bphf_set_profile(2)
m = None
out = []
for r in range(reps):
    t0 = time.perf_counter()
    mm = comm.build(1.7, keys, n * world)
    dt = time.perf_counter() - t0
    st = mm.build_stats()
    if keep:
        m = mm
    else:
        del mm
    out.append("rep %d wall %.2f dev %.2f | p0 %.2f f0 %.2f m0 %.2f p1 %.2f" % (r, dt * 1e3, st["total_ms"], st["pass_ms"][0], st["fin_ms"][0], st["merge_ms"][0], st["pass_ms"][1]))
if rank == 0:
    print("\n".join(out), flush=True)
m = None
torch.cuda.synchronize(); dist.barrier(); comm.close(); dist.destroy_process_group()
