#!/usr/bin/env python
"""bench.py -- MPHF build / batched lookup throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one complete MPHF construction (Mphf::new_parallel, gamma = 1.7) over the workload's
synthetic key set, keys resident in HBM when the timed region starts.  The same run also reports
 * lookup:  batched Mphf::hash of every key against the built structure (Mkeys/s),
 * e2e:     the same build through the C ABI with HOST (pinned) keys: H2D copy, build, and the D2H
            export of all level words + ranks inside the timed region,
 * roofline for the dominant kernel (CUDA events inside the library, on the launching stream),
 * cpu_baseline: the oracle's restatement of new_parallel on this box's host cores.
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
# (profiles/r02_summary.md; n = 1e8, gamma = 1.7, one GPU -- only quoted for that workload)
NCU_TRAFFIC_1E8 = {"mark_list_kernel": 0.8472e9 + 9.463e6, "cascade_kernel": 1.450e9 + 641.4e6}
NCU_TRAFFIC_SOURCE = "ncu --set full, profiles/r02_summary.md (capture of this code state)"
# the exchange-level kernels of a sharded build, per GPU at 1e8 keys per GPU: captured in one-shard mode
# (BPHF_XCHG_WORLD1: the same per-GPU work, stores local instead of NVLink), profiles/r02d_ncu_xchg_table.md
NCU_TRAFFIC_XCHG_1E8 = {"xscatter_kernel": 0.800e9 + 1.149e9,
                        "xmark + xfinalize + xbits": (442.8e6 + 4.1e6) + 42.5e6 + (421.3e6 + 15.9e6)}
NCU_TRAFFIC_XCHG_SOURCE = ("ncu --set full in one-shard exchange mode (same per-GPU work, local stores instead of NVLink), "
                           "profiles/r02d_ncu_xchg_table.md")

METRIC = "mphf_build_throughput"
UNIT = "Mkeys/s"
GAMMA = 1.7


def host_threads():
    """cores this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arms pass
    this count to the oracle explicitly instead of letting OpenMP read the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML (an nvidia-smi
    subprocess per sample stalls the driver for tens of ms and would perturb a 3 ms step)."""

    def __init__(self, torch_index, enabled=True, period=0.05):
        # Measured (tools/nvml_cost.py): the FIRST call of each NVML query initialises lazily and takes 2-3 ms with a
        # driver-wide lock held -- issued from every rank at the start of the timed region that stalled the kernel
        # launches of all ranks (a 4-GPU sharded step was seen at 33 ms instead of 5 ms; a single-GPU first step at
        # 4-9 ms instead of 2.9).  Later calls cost microseconds and do not move the step time even every 2 ms.
        # So: one sampling rank, the queries warmed up here (before the timed region), 20 Hz inside it.
        self.period = period
        self.samples = []
        self._active = False
        self._stop = threading.Event()
        self._t = None
        self._h = None
        self._nv = None
        try:
            if not enabled:
                raise RuntimeError("clock sampling is done by rank 0 only")
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                h = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = torch_index
                if vis:
                    try:
                        idx = int(vis.split(",")[torch_index])
                    except Exception:
                        idx = torch_index
                h = nv.nvmlDeviceGetHandleByIndex(idx)
            self._h = h
            self._max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self._sample()      # lazy initialisation of both queries happens here, outside the timed region
        except Exception:
            self._nv = None

    def _sample(self):
        nv, h = self._nv, self._h
        try:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            if self._active:
                self.samples.append((sm, int(r)))
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.period)

    def start_thread(self):
        """the sampling thread runs from here on (before the warm-up builds) and only RECORDS inside `with`:
        nothing new -- no thread start, no first NVML call of a thread -- happens at the start of the timed region"""
        if self._nv is not None and self._t is None and not os.environ.get("BENCH_NO_CLOCK_THREAD"):
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __enter__(self):
        self.start_thread()
        self._active = True
        return self

    def __exit__(self, *a):
        if self._nv is not None:
            self._sample()
        self._active = False
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if self._nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(s for s, _ in self.samples)
        seen = set()
        for _, r in self.samples:
            for nm, bit in names.items():
                if r & bit:
                    seen.add(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self._max, "reasons": sorted(seen), "samples": len(sm)}


def collective_spin(build, seconds, world, device):
    """Call build() back to back for about `seconds`; EVERY rank makes the same number of calls (a sharded build is a
    collective: a per-rank clock check would let one rank leave a build earlier than its peers, who then wait at a
    device barrier for a shard that never comes).  Returns (last result, number of calls)."""
    import torch
    # No torch allocation inside the loop: a cudaMalloc by torch's allocator between builds makes the first build after
    # the next device-wide synchronize spend 20-480 ms in the stream-ordered pool (tools/pool_probe.py).
    over = torch.zeros(1, dtype=torch.int32, device=device) if world > 1 else None
    t0 = time.perf_counter()
    m, calls = None, 0
    while True:
        m = build()
        calls += 1
        done = time.perf_counter() - t0 >= seconds
        if world > 1:
            import torch.distributed as dist
            over.fill_(1 if done else 0)
            dist.all_reduce(over, op=dist.ReduceOp.MAX)
            done = bool(int(over.item()))
        if done:
            return m, calls


def algorithmic_bytes(stats, world=1):
    """B_l = 16k + 8c + 41W per level (key read in mark and in filter, redo write, zeroing
    of a and collide, finalize read a/read collide/write a, ranks).  Returns total and per-kernel splits, PER GPU:
    a shard of a sharded build streams 1/world of every level's keys but full-width bit vectors."""
    ks, bits = [k // world for k in stats["keys_in"]], stats["bits"]
    total = 0
    per_pass = []
    fin = []
    for l, (k, b) in enumerate(zip(ks, bits)):
        w = (b + 63) // 64
        c = ks[l + 1] if l + 1 < len(ks) else 0
        total += 16 * k + 8 * c + 41 * w
        # this implementation's pass(l) step: reads the keys that entered level l-1 (level 0: the input),
        # writes level l's list.  pass(0) = 8k; pass(l) = 8 k_{l-1} + 8 k_l.  finalize(l): read a, read collide, write a
        per_pass.append(8 * k if l == 0 else 8 * ks[l - 1] + 8 * k)
        fin.append(24 * w)
    tail = stats["tail_level"] if stats["tail_level"] < len(ks) else len(ks)
    # the persistent cascade kernel = pass + finalize of every level from 1 up to the tail
    cascade = sum(per_pass[1:tail]) + sum(fin[1:tail])
    return total, per_pass, cascade


def workload_config(n_per_gpu, world):
    """the `config` object of BOTH arms (ours and --impl reference): the workload, nothing run specific"""
    n_global = n_per_gpu * world
    if world == 1:
        what = ("Mphf::new_parallel gamma=1.7 on 1e8 distinct random u64 keys (BASELINE configs[1])" if n_per_gpu == 100_000_000
                else "Mphf::new_parallel gamma=1.7 on %d distinct random u64 keys" % n_per_gpu)
    else:
        what = ("ONE Mphf::new_parallel gamma=1.7 over %d distinct random u64 keys, %d per GPU on %d GPUs "
                "(BASELINE configs[1] per GPU, weak scaling towards configs[2])" % (n_global, n_per_gpu, world))
    return {"workload": what, "gamma": GAMMA, "n_keys_per_gpu": n_per_gpu, "n_keys": n_global,
            "l2": "inputs (%.0f MB of keys per GPU and step) are larger than the 126 MB L2" % (8 * n_per_gpu / 1e6)}


def _timed(fn, reps, sync):
    best = 1e30
    for _ in range(reps):
        sync()
        t0 = time.perf_counter()
        r = fn()
        sync()
        best = min(best, time.perf_counter() - t0)
    return best, r


def extra_configs(pkg, L, device, stream, orc):
    """BASELINE configs[0] (benches/build.rs:9-34: 1 M-key `Mphf::new` + a scan of `hash`) and configs[4]
    (src/hashmap.rs:170-173: BoomHashMap over 5e8 2-bit-packed 31-mers, gamma sweep), each next to the oracle on the host:
    single thread for configs[0] as in the reference's bench, a bounded 2e7-key sample for the map."""
    import numpy as np
    import torch
    sync = torch.cuda.synchronize
    out = {}
    # ---- configs[0]
    n0 = 1_000_000
    res0 = {}
    for name, seed, kind, gamma in (("new gamma=1.7, 1M random u64 keys", 1, 0, 1.7),
                                    ("build1_ser: new gamma=2.0, keys 2i, i < 1M (benches/build.rs:11-12)", 0, 2, 2.0)):
        keys = torch.empty(n0, dtype=torch.int64, device="cuda:%d" % device)
        assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n0, seed, kind, device, None) == 0
        res = torch.empty_like(keys)
        pkg.Mphf.new(gamma, keys)
        tb, m = _timed(lambda: pkg.Mphf.new(gamma, keys), 20, sync)
        tq, _ = _timed(lambda: m.hash_batch(keys, out=res, stream=stream), 20, sync)
        bad = C.c_uint64(1)
        assert L.bphf_check_permutation(C.c_void_p(res.data_ptr()), n0, device, C.byref(bad)) == 0
        e = {"gpu_build_ms": tb * 1e3, "gpu_build_Mkeys_s": n0 / tb / 1e6, "gpu_scan_ms": tq * 1e3, "gpu_scan_Mkeys_s": n0 / tq / 1e6,
             "permutation_ok": bad.value == 0}
        if orc is not None:
            hk = orc.keygen(n0, seed=seed, kind=kind)
            tc, ref = _timed(lambda: orc.OracleMphf.new(gamma, hk), 2, lambda: None)
            ts, _ = _timed(lambda: ref.hash_batch(hk, nthreads=1), 2, lambda: None)
            e.update({"cpu_build_ms": tc * 1e3, "cpu_build_Mkeys_s": n0 / tc / 1e6, "cpu_scan_ms": ts * 1e3,
                      "cpu_scan_Mkeys_s": n0 / ts / 1e6, "cpu_cores": 1,
                      "bit_exact_vs_gpu": orc.fingerprint(ref.levels()) == orc.fingerprint(m.levels())})
        res0[name] = e
        del keys, res, m
    out["configs0_new_1M"] = res0
    # ---- configs[3]: Mphf::hash (src/lib.rs:324-335) of 1e9 in-set keys against the 435 MB structure of 1e9 keys
    out["configs3_lookup_1e9"] = config3_lookup(pkg, L, device, stream, orc)
    # ---- configs[4]
    n5 = 500_000_000
    free_b, _ = torch.cuda.mem_get_info(device)
    if free_b < 60e9:
        out["configs4_boomhashmap_5e8"] = {"skipped": "needs ~45 GB of free device memory, %.0f GB free" % (free_b / 1e9)}
        return out
    keys = torch.empty(n5, dtype=torch.int64, device="cuda:%d" % device)
    assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n5, 3, 1, device, None) == 0
    vals = torch.arange(n5, dtype=torch.int64, device=keys.device)
    ko, vo, got = torch.empty_like(keys), torch.empty_like(keys), torch.empty_like(keys)
    found = torch.empty(n5, dtype=torch.uint8, device=keys.device)
    res5 = {"n_keys": n5, "keys": "62-bit bijection of the counter (2-bit packed 31-mers)", "gammas": {}}
    try:
        pkg.Mphf.new_parallel(1.0, keys, None)
        res5["gamma_1.0"] = "NOT rejected"
    except pkg.MphfPanic as e:
        res5["gamma_1.0"] = "rejected like the reference: %s" % e
    for gamma in (1.7, 2.0, 5.0):
        pkg.Mphf.new_parallel(gamma, keys, None)
        tb, m = _timed(lambda: pkg.Mphf.new_parallel(gamma, keys, None), 2, sync)
        tw, tr, _ = m.total_dims()

        def permute():
            assert L.bphf_map_permute_u64(m._h, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), n5,
                                          C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), 1, None) == 0

        def get():
            assert L.bphf_map_get_batch_u64(m._h, C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), n5,
                                            C.c_void_p(keys.data_ptr()), n5, C.c_void_p(got.data_ptr()),
                                            C.c_void_p(found.data_ptr()), 1, None) == 0
        tp, _ = _timed(permute, 2, sync)
        tg, _ = _timed(get, 2, sync)
        ok = bool(found.all().item()) and bool((got == vals).all().item())
        res5["gammas"]["%.1f" % gamma] = {
            "levels": m.num_levels, "bits_per_key_with_ranks": 64.0 * (tw + tr) / n5, "build_ms": tb * 1e3,
            "build_Mkeys_s": n5 / tb / 1e6, "create_map_ms": tp * 1e3, "get_all_ms": tg * 1e3, "get_Mkeys_s": n5 / tg / 1e6,
            "all_found_with_their_values": ok}
        del m
    del keys, vals, ko, vo, got, found
    if orc is not None:
        nc = 20_000_000
        hk = orc.keygen(nc, seed=3, kind=1)
        hv = np.arange(nc, dtype=np.uint64)
        threads = host_threads()
        tc, ref = _timed(lambda: orc.OracleMphf.new_parallel(1.7, hk, None, nthreads=threads), 1, lambda: None)
        tm, (mk, mv) = _timed(lambda: ref.create_map(hk, hv), 1, lambda: None)
        tg, _ = _timed(lambda: ref.map_get_batch(mk, mv, hk), 1, lambda: None)
        res5["cpu_sample"] = {"n_keys": nc, "gamma": 1.7, "cores_build": threads, "build_Mkeys_s": nc / tc / 1e6,
                              "create_map_Mkeys_s_1core": nc / tm / 1e6, "get_Mkeys_s_1core": nc / tg / 1e6,
                              "what": "oracle restatement of BoomHashMap::new_parallel on a bounded sample (create_map and get are "
                                      "single-threaded in the reference, src/hashmap.rs:27-66)"}
    out["configs4_boomhashmap_5e8"] = res5
    return out


def config3_lookup(pkg, L, device, stream, orc, n=1_000_000_000):
    """BASELINE configs[3] on one GPU: n = 1e9 keys built once, then all n looked up, keys and ranks device resident.  Timed
    with CUDA events on `stream` for the plain kernel (every probe of the big levels is a random DRAM sector) and for the
    partitioned kernels (csrc/qpart_kernels.cuh, what the library picks on its own for this size); both outputs must be
    the same permutation of 0..n-1.  CPU: the oracle's hash over a bounded sample against a 1e8-key structure."""
    import torch
    free_b, _ = torch.cuda.mem_get_info(device)
    if free_b < 60e9:
        return {"skipped": "needs ~45 GB of free device memory, %.0f GB free" % (free_b / 1e9)}
    dev = "cuda:%d" % device
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), 0, n, 1, 0, device, None) == 0
    m = pkg.Mphf.new_parallel(GAMMA, keys, None)
    tw, tr, _ = m.total_dims()
    res = torch.empty_like(keys)
    peak, _ = measured_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r = {"n_keys": n, "levels": m.num_levels, "structure_bytes": 8 * (tw + tr)}
    sums = []
    for label, mode in (("plain_kernel", 1), ("partitioned", 0)):
        assert L.bphf_set_query_partition(mode, 0, 0, 0, 0.0) == 0
        m.hash_batch(keys, out=res, stream=stream)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            m.hash_batch(keys, out=res, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bad, sm, xr = C.c_uint64(1), C.c_uint64(), C.c_uint64()
        assert L.bphf_check_permutation(C.c_void_p(res.data_ptr()), n, device, C.byref(bad)) == 0
        assert L.bphf_checksum_u64(C.c_void_p(res.data_ptr()), n, device, C.byref(sm), C.byref(xr)) == 0
        sums.append((sm.value, xr.value))
        r[label] = {"ms": ms, "Mkeys_s": n / ms / 1e3, "permutation_ok": bad.value == 0,
                    "frac_of_hbm": (16.0 * n + 8.0 * (tw + tr)) / (ms * 1e-3) / 1e9 / peak}
    assert L.bphf_set_query_partition(0, 0, 0, 0, 0.0) == 0
    b, f, st = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    assert L.bphf_query_partition_stats(device, C.byref(b), C.byref(f), C.byref(st)) == 0
    r["partitioned"].update({"stages": st.value, "batches_redone_by_the_plain_kernel": f.value,
                             "same_ranks_as_the_plain_kernel": sums[0] == sums[1]})
    r["frac_of_hbm_counts"] = "8 B key in + 8 B rank out per lookup + the structure once, over the measured copy bandwidth"
    del keys, res, m
    if orc is not None:
        nb, nq = 100_000_000, 20_000_000
        threads = host_threads()
        hk = orc.keygen(nb, seed=1, kind=0)
        ref = orc.OracleMphf.new_parallel(GAMMA, hk, None, nthreads=threads)
        t1, _ = _timed(lambda: ref.hash_batch(hk[:nq], nthreads=1), 1, lambda: None)
        tt, _ = _timed(lambda: ref.hash_batch(hk[:nq], nthreads=threads), 1, lambda: None)
        r["cpu_sample"] = {"structure_keys": nb, "lookups": nq, "Mkeys_s_1core": nq / t1 / 1e6, "Mkeys_s_all_cores": nq / tt / 1e6,
                           "cores": threads, "what": "oracle Mphf::hash over a bounded sample (the reference's hash is a "
                                                     "single-threaded call; all cores = one batch split over the host threads)"}
        del ref, hk
    return r


def north_star_1e9(pkg, L, local_rank, rank, world, dev, stream, peak):
    """--gpus 8 only: the north_star target proper -- ONE Mphf over 1e9 keys (1.25e8 per GPU, BASELINE configs[2]) and the
    batched lookup of all 1e9 in-set keys (configs[3]; replicas: every rank looks up its 1.25e8 keys)."""
    import importlib
    import torch
    import torch.distributed as dist
    n = 125_000_000
    total = n * world
    sharded = importlib.import_module(pkg.__name__ + ".sharded")
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), rank * n, n, 1, 0, local_rank, None) == 0
    comm = sharded.ShardComm.from_process_group(local_rank, total, GAMMA)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()
    m = None
    for _ in range(6):
        m = comm.build(GAMMA, keys, total)
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        m = comm.build(GAMMA, keys, total)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    build_ms = float(t.item())
    st = m.build_stats()
    out = torch.empty(n, dtype=torch.int64, device=dev)
    for _ in range(3):
        m.hash_batch(keys, out=out, stream=stream)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        m.hash_batch(keys, out=out, stream=stream)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lookup_ms = float(t.item())
    sm, xr = C.c_uint64(), C.c_uint64()
    assert L.bphf_checksum_u64(C.c_void_p(out.data_ptr()), n, local_rank, C.byref(sm), C.byref(xr)) == 0
    want = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int64, device=dev)
    sw, xw = C.c_uint64(), C.c_uint64()
    assert L.bphf_checksum_u64(C.c_void_p(want.data_ptr()), n, local_rank, C.byref(sw), C.byref(xw)) == 0
    wrap = lambda v: v - (1 << 64) if v >= (1 << 63) else v
    both = torch.tensor([wrap(sm.value), wrap(sw.value)], dtype=torch.int64, device=dev)
    dist.all_reduce(both)
    in_range = torch.tensor([int(bool((out >= 0).all() and (out < total).all()))], dtype=torch.int64, device=dev)
    dist.all_reduce(in_range, op=dist.ReduceOp.MIN)
    tw, tr, _ = m.total_dims()
    total_bytes, _, _ = algorithmic_bytes(st, world)
    res = {"n_keys": total, "n_gpus": world, "build_ms": build_ms, "build_Mkeys_s": total / build_ms / 1e3,
           "build_frac_of_hbm_per_gpu": total_bytes / (build_ms * 1e-3) / 1e9 / peak,
           "levels": st["n_levels"], "exchange_levels": st["bucketed_levels"] // 1000,
           "lookup_ms": lookup_ms, "lookup_Mkeys_s": total / lookup_ms / 1e3,
           "lookup_frac_of_hbm_per_gpu": (16.0 * n + 8.0 * (tw + tr)) / (lookup_ms * 1e-3) / 1e9 / peak,
           "structure_bytes_per_replica": 8 * (tw + tr),
           "outputs_are_a_permutation_of_0_to_n": bool(both[0].item() == both[1].item()) and bool(in_range.item() == 1),
           "what": "ONE Mphf::new_parallel(1.7) over 1e9 keys sharded 8 x 1.25e8 (device time, max over ranks, %d steps) and "
                   "Mphf::hash of all 1e9 keys (each rank looks its shard up in its replica)" % steps}
    del m, out, want, keys
    torch.cuda.synchronize()
    dist.barrier()
    comm.close()
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm (oracle restatement of Mphf::new_parallel; the Rust crate cannot be
    compiled in this image) on the host cores, same config / metric / unit."""
    if rank != 0:
        return
    orc = ge.load_oracle()
    orc.build_lib()
    # bounded sample: the full 1e8-key config per step while K steps of it fit ~2 minutes of host time (K <= 50 at the
    # ~46 Mkeys/s of a 16-core box), fewer keys per step beyond that; --ref-keys overrides
    n = args.ref_keys if args.ref_keys else int(min(args.keys * world, max(2_000_000, 5_000_000_000 // max(args.steps, 1))))
    keys = orc.keygen(n, seed=1, kind=0)
    threads = host_threads()
    for _ in range(args.warmup):
        orc.OracleMphf.new_parallel(GAMMA, keys[: max(n // 10, 1)], None, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m = orc.OracleMphf.new_parallel(GAMMA, keys, None, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    tq = time.perf_counter()
    m.hash_batch(keys[: min(n, 20_000_000)], nthreads=threads)
    lookup = min(n, 20_000_000) / (time.perf_counter() - tq) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.keys, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d keys per step (bounded sample of the workload's key set, same generator), OpenMP restatement of new_parallel (rustc absent); %d threads passed "
                                   "explicitly (OMP_NUM_THREADS=%s in the environment is not used)"
                                   % (n, threads, os.environ.get("OMP_NUM_THREADS", "unset"))},
        "lookup": {"value": lookup, "unit": UNIT, "cores": threads},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # 200 steps = 0.6 s of timed builds: long enough that one host-side hiccup (the first build after the device-wide
    # synchronize in front of the timed region occasionally waits ~10 ms for the memory pool) moves the mean by < 2 %
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--keys", type=int, default=100_000_000, help="keys per GPU (BASELINE configs[1]: 1e8)")
    ap.add_argument("--ref-keys", type=int, default=0,
                    help="keys per step of the CPU reference arm (0: 1e8, scaled down when --steps is large)")
    ap.add_argument("--cpu-baseline-keys", type=int, default=100_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip BASELINE configs[0] / configs[4] (N = 1) and the 1e9-key north-star block (N = 8)")
    ap.add_argument("--no-bit-exact", action="store_true", help="N > 1: skip the oracle build over all N x n keys")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    pkg.build_native()
    L = pkg.lib()
    n = args.keys
    stream = torch.cuda.current_stream(dev)

    # ---- synthetic input, resident in HBM: distinct-by-construction random u64 keys -------------
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    assert L.bphf_keygen_u64(C.c_void_p(keys.data_ptr()), rank * n, n, 1, 0, local_rank, None) == 0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: ONE MPHF over all world * n keys, keys sharded across the GPUs (they never move); per level the
    # shards' (a, collide) pairs are merged by the OR-collide peer-memory kernel.  Every rank ends
    # with a complete replica.  N == 1: the plain single-GPU build.
    comm = None
    n_global = world * n
    if world > 1:
        import importlib
        sharded = importlib.import_module(pkg.__name__ + ".sharded")
        comm = sharded.ShardComm.from_process_group(local_rank, n_global, GAMMA)

    def build():
        if comm is not None:
            return comm.build(GAMMA, keys, n_global)
        return pkg.Mphf.new_parallel(GAMMA, keys, None, stream=stream)

    # ---- warm-up ---------------------------------------------------------------------------------------------
    # clocks ramp from idle for a few hundred ms: spin untimed builds for ~1 s first, then the W warm-up steps
    L.bphf_set_profile(1)
    # NVML is initialised (and its queries warmed up) here, BEFORE the warm-up builds: for a few milliseconds after
    # nvmlInit the driver stalls kernel launches -- measured as a 4-11 ms first timed step when this sat right in
    # front of the timed region (tools/nvml_cost.py)
    clocks = ClockSampler(local_rank, enabled=(rank == 0)).start_thread()
    m = None
    # (a sharded build is a collective: every rank must run the same number of them, so the ranks agree on when the
    # second is over -- a per-rank clock check let one rank leave the loop a build earlier than its peers)
    m, _ = collective_spin(build, 1.0, world, dev)
    for _ in range(max(args.warmup, 3)):
        m = build()
    stats = m.build_stats()

    # ---- timed: K builds, device time via CUDA events on the launching stream -----------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms = [0.0, 0.0]
    cascade_ms = 0.0
    merge0_ms = 0.0
    fin0_ms = 0.0
    launches = 0
    trace_steps = bool(os.environ.get("BENCH_TRACE_STEPS"))  # debugging aid: per-step wall times on stderr
    barrier()
    with clocks:
        e0.record(stream)
        t0 = time.perf_counter()
        step_t = []
        for _ in range(args.steps):
            if trace_steps:
                step_t.append(time.perf_counter())
            m = build()
            st = m.build_stats_raw()  # the struct, not the dict: no per-level Python loops inside the timed region
            pass_ms[0] += st.pass_ms[0]
            pass_ms[1] += st.pass_ms[1]
            cascade_ms += st.cascade_ms
            merge0_ms += st.merge_ms[0]
            fin0_ms += st.fin_ms[0]
            launches += st.kernel_launches
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        if trace_steps and rank == 0:
            step_t.append(time.perf_counter())
            sys.stderr.write("step wall us: " + " ".join("%.0f" % ((b - a) * 1e6) for a, b in zip(step_t[:-1], step_t[1:])) + "\n")
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    # ---- lookup: hash every key (device resident), K passes ------------------------------------------------------
    out = torch.empty(n, dtype=torch.int64, device=dev)
    for _ in range(3):
        m.hash_batch(keys, out=out, stream=stream)
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record(stream)
    for _ in range(args.steps):
        m.hash_batch(keys, out=out, stream=stream)
    q1.record(stream)
    barrier()
    tq = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
    q_ms = float(tq.item()) / args.steps
    lookup_val = world * n / (q_ms * 1e-3) / 1e6
    if world == 1:
        bad = C.c_uint64(1)
        assert L.bphf_check_permutation(C.c_void_p(out.data_ptr()), n, local_rank, C.byref(bad)) == 0
        permutation_ok = bad.value == 0
    else:
        # the shards' outputs together must be exactly 0..n_global-1: compare the order-free checksum (sum of
        # splitmix64 of every value, mod 2^64) of all outputs with that of the range itself
        def csum(t):
            sm, xr = C.c_uint64(), C.c_uint64()
            assert L.bphf_checksum_u64(C.c_void_p(t.data_ptr()), t.numel(), local_rank, C.byref(sm), C.byref(xr)) == 0
            return sm.value
        want = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int64, device=dev)
        both = torch.tensor([csum(out) - (1 << 64) if csum(out) >= (1 << 63) else csum(out),
                             csum(want) - (1 << 64) if csum(want) >= (1 << 63) else csum(want)], dtype=torch.int64, device=dev)
        dist.all_reduce(both)
        in_range = torch.tensor([int(bool((out.view(torch.int64) >= 0).all() and (out.view(torch.int64) < n_global).all()))],
                                dtype=torch.int64, device=dev)
        dist.all_reduce(in_range, op=dist.ReduceOp.MIN)
        permutation_ok = bool(both[0].item() == both[1].item()) and bool(in_range.item() == 1)
        del want

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hk = torch.empty(n, dtype=torch.int64).pin_memory()
        hk.copy_(keys)
        tw, tr, _ = m.total_dims()
        words = torch.empty(tw, dtype=torch.int64).pin_memory()
        ranks = torch.empty(tr, dtype=torch.int64).pin_memory()
        hk_np = hk.numpy().view(np.uint64)

        def e2e_step():
            # every rank hands over its shard of the keys from pinned host memory; the finished structure is read back
            # ONCE (rank 0): the ranks hold identical replicas
            mm = comm.build(GAMMA, hk_np, n_global) if comm is not None else pkg.Mphf.new_parallel(GAMMA, hk_np, None, device=local_rank)
            if rank == 0:
                assert L.bphf_export_all(mm._h, C.c_void_p(words.data_ptr()), C.c_void_p(ranks.data_ptr()), 0) == 0
            return mm

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n / float(dt.item()) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * n * world,
               "d2h_bytes_per_step": 8 * (tw + tr), "ms_per_step": float(dt.item()) * 1e3,
               "what": ("bphf_build_u64(host pinned keys) + bphf_export_all(host) per step, wall clock" if world == 1 else
                        "bphf_build_sharded_u64(each rank's shard from pinned host memory) + bphf_export_all(host) on rank 0 per step, "
                        "wall clock, max over ranks; bytes are totals over all ranks")}
        del hk, words, ranks

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    total_bytes, per_pass, cascade_bytes = algorithmic_bytes(stats, world)
    xchg_levels = stats["bucketed_levels"] // 1000      # sharded: levels built by the exchange (owner computes) kernels
    w0 = (stats["bits"][0] + 63) // 64
    # dominant kernel by time: the level-0 mark kernel, the persistent cascade kernel (N = 1: every level >= 1
    # in one launch) or, in a sharded build, the level-0 steps of the path the level took
    cands = [("mark_list_kernel (level-0 find_collisions)", pass_ms[0] / args.steps, per_pass[0])]
    if world > 1 and xchg_levels > 0:
        # exchange level 0: xscatter reads the shard's keys (8 B), keeps them grouped by owner (8 B) and pushes 4-byte slot
        # offsets; the owner phase reads the n offsets it received twice (mark, survivor bits), its slice of a / collide
        # (zero + finalize: 41 W / G) and stores the final slice into every arena
        cands = [("xscatter_kernel (level 0: route every key's slot offset to the shard that owns the slot)",
                  pass_ms[0] / args.steps, 20 * n),
                 ("xmark + xfinalize + xbits (level 0, owner side: find_collisions on the slice, final a to all shards, survivor bits)",
                  fin0_ms / args.steps, 8 * n + n // 8 + (41 + 8 * (world - 1)) * w0 // world)]
    elif world > 1 and stats["bucketed_levels"] > 0:
        cands = [("scatter0_kernel (level 0 of a sharded build: keys grouped by bit range for the shared-memory mark)",
                  pass_ms[0] / args.steps, 2 * per_pass[0])]
    if world > 1:
        pass  # the cascade kernel of a sharded build only covers the levels after the key collection: not comparable
    elif cascade_ms > 0:
        cands.append(("cascade_kernel (filter + compact + mark + finalize of every level >= 1, one launch)",
                      cascade_ms / args.steps, cascade_bytes))
    elif len(per_pass) > 1:
        cands.append(("pass_kernel level 1 (filter + compact + mark)", pass_ms[1] / args.steps, per_pass[1]))
    k_name, k_ms, k_bytes = max(cands, key=lambda c: c[1])
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    n_struct = m.total_dims()[0] + m.total_dims()[1]
    roofline = {
        "bound": "hbm", "kernel": k_name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": (next((v for k, v in NCU_TRAFFIC_1E8.items() if k_name.startswith(k)), None)
                    if world == 1 else next((v for k, v in NCU_TRAFFIC_XCHG_1E8.items() if k_name.startswith(k)), None))
        if n == 100_000_000 else None,
        "traffic_source": NCU_TRAFFIC_SOURCE if world == 1 else NCU_TRAFFIC_XCHG_SOURCE,
        "peak_source": peak_src, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": k_bytes,
        "kernels": [{"kernel": c[0], "ms": c[1], "algorithmic_bytes": c[2],
                     "achieved": (c[2] / (c[1] * 1e-3) / 1e9) if c[1] > 0 else None} for c in cands],
        "whole_build": {"algorithmic_bytes": total_bytes, "bytes_per_key": total_bytes / n,
                        "achieved": total_bytes / (ms_per_step * 1e-3) / 1e9, "frac": total_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
        "lookup": {"algorithmic_bytes_per_key": 16.0 + 8.0 * n_struct / n,
                   "achieved": (16.0 * n + 8.0 * n_struct) / (q_ms * 1e-3) / 1e9},
    }
    roofline["lookup"]["frac"] = roofline["lookup"]["achieved"] / peak
    # ---- the L2-atomic ceiling north_star asks for, measured on THIS device in this process: every key of every level
    # costs one returning fetch_or (a), ~0.244 of them a non-returning OR (collide, late arrivals at gamma 1.7: the exact
    # count is popcount-derived below) and, from level 1 on, one random probe of the previous level's collide word
    ks = [k // world for k in stats["keys_in"]]
    n_atom = sum(ks)
    n_probe = sum(ks[:-1])
    # late arrivals of a level = keys - occupied slots = keys - (unique + collided slots); unique(l) = k(l) - k(l+1) and
    # collided slots are not exported: use the expectation k - bits * (1 - exp(-k / bits)) per level
    import math
    n_red = sum(int(k - (b / world) * (1.0 - math.exp(-k / (b / world)))) for k, b in zip(ks, stats["bits"]) if b)
    arr_bytes = max(4096, (stats["bits"][0] // world) // 8)
    probe_ops = 50_000_000
    rates = {}
    for kind, nm in ((1, "atom"), (0, "red"), (2, "load")):
        ms = C.c_float()
        assert L.bphf_l2_probe(kind, arr_bytes, probe_ops, local_rank, C.byref(ms)) == 0
        rates[nm] = probe_ops / (ms.value * 1e-3) / 1e9
    floor_ms = (n_atom / rates["atom"] + n_red / rates["red"]) / 1e6      # atomics alone, serial: the L2-side floor
    build_kernel_ms = (pass_ms[0] + cascade_ms) / args.steps if world == 1 else ms_per_step
    roofline["l2_atomic"] = {
        "ceilings_measured_Gops": rates, "array_bytes": arr_bytes,
        "ops_per_build_per_gpu": {"fetch_or": n_atom, "collide_or": n_red, "collide_probe": n_probe},
        "floor_ms": floor_ms, "floor_with_probes_ms": floor_ms + n_probe / rates["load"] / 1e6,
        "measured_ms": build_kernel_ms, "frac_of_ceiling": floor_ms / build_kernel_ms if build_kernel_ms > 0 else None,
        "what": "fetch_or / collide OR counts of one build divided by this device's measured random-op rates on a level-0 sized "
                "array (bphf_l2_probe), against the time of the marking kernels (N = 1: mark_list + cascade; N > 1: whole step)"}
    if world > 1 and xchg_levels > 0:
        out_bytes = (world - 1) / world * 4.0 * n
        roofline["nvlink_level0"] = {"what": "exchange level 0: 4-byte slot offsets pushed by xscatter_kernel (compute and link overlap "
                                             "inside the kernel), final a slices pushed by xfinalize_kernel, 1 bit per key back from xbits_kernel",
                                     "offset_bytes_out_per_gpu": out_bytes, "a_slice_bytes_out_per_gpu": (world - 1) / world * 8.0 * w0,
                                     "bit_bytes_out_per_gpu": (world - 1) / world * n / 8.0,
                                     "xscatter_ms": pass_ms[0] / args.steps, "owner_ms": fin0_ms / args.steps,
                                     "barriers_ms": merge0_ms / args.steps,
                                     "achieved_per_direction": out_bytes / (pass_ms[0] / args.steps * 1e-3) / 1e9 if pass_ms[0] > 0 else None,
                                     "peak_per_direction": 900.0, "unit": "GB/s"}
    elif world > 1 and merge0_ms > 0:
        # the OR-collide fold of level 0 moves (G-1)/G * 16 W bytes into and out of every GPU
        each_way = (world - 1) / world * 16 * w0
        roofline["nvlink_level0"] = {"bytes_in_per_gpu": each_way, "bytes_out_per_gpu": each_way,
                                     "ms": merge0_ms / args.steps, "what": "two device barriers + or_collide_merge_kernel",
                                     "achieved_per_direction": each_way / (merge0_ms / args.steps * 1e-3) / 1e9,
                                     "peak_per_direction": 900.0, "unit": "GB/s"}

    # ---- every rank's replica must be the same structure: order-free checksums of words and ranks ----------------------
    replicas_identical = None
    if world > 1:
        tw_, tr_, _ = m.total_dims()
        wd = torch.empty(tw_, dtype=torch.int64, device=dev)
        rd = torch.empty(tr_, dtype=torch.int64, device=dev)
        assert L.bphf_export_all(m._h, C.c_void_p(wd.data_ptr()), C.c_void_p(rd.data_ptr()), 1) == 0
        sig = torch.stack([wd.sum(), (wd ^ (wd >> 7)).sum(), rd.sum(), (rd ^ (rd >> 11)).sum(),
                           torch.tensor(tw_, device=dev), torch.tensor(tr_, device=dev)])
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        replicas_identical = all(bool((x == sigs[0]).all().item()) for x in sigs)
        del wd, rd

    # ---- CPU baseline (oracle port of new_parallel on the host cores), rank 0 ---------------------------------------
    # N = 1: the 1e8-key workload itself.  N > 1: the whole N x n key set, once -- that build is also the bit-exactness
    # check of the sharded structure (rank 0's replica, word for word).
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        orc = ge.load_oracle()
        threads = host_threads()
        full = world > 1 and not args.no_bit_exact
        nc = n_global if full else (args.cpu_baseline_keys if world == 1 else n)
        ck = orc.keygen(nc, seed=1, kind=0)
        t0 = time.perf_counter()
        ref = orc.OracleMphf.new_parallel(GAMMA, ck, None, nthreads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": nc / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "one new_parallel(1.7) over %d keys of the same generator (%.1f s), %d threads; C/OpenMP restatement, rustc absent"
                         % (nc, dt, threads)}
        if nc == n_global:
            same = orc.fingerprint(ref.levels()) == orc.fingerprint(m.levels())
            cpu["bit_exact_vs_gpu"] = bool(same)
        del ref, ck

    # ---- the other BASELINE configs, next to their CPU numbers ----------------------------------------------------------
    extra = None
    if rank == 0 and world == 1 and not args.no_extra_configs:
        extra = extra_configs(pkg, L, local_rank, stream, None if args.no_cpu_baseline else ge.load_oracle())
    north_star = None
    if world == 8 and not args.no_extra_configs:
        if comm is not None:
            del m
            m = None
            torch.cuda.synchronize()
            dist.barrier()
            comm.close()
            comm = None
        north_star = north_star_1e9(pkg, L, local_rank, rank, world, dev, stream, peak)

    if rank == 0:
        cfg = workload_config(n, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": cfg,
            "build_info": {"levels": stats["n_levels"], "tail_level": stats["tail_level"], "exchange_levels": xchg_levels,
                           "multi_gpu": ("keys sharded across %d GPUs and never moved; levels beyond the L2 are built owner-computes "
                                         "(slot offsets over NVLink, survivor bits back), the others by the OR-collide merge over peer "
                                         "memory; every GPU ends with a replica" % world) if world > 1 else "single GPU"},
            "lookup": {"value": lookup_val, "unit": UNIT, "ms_per_pass": q_ms, "permutation_ok": permutation_ok},
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_launches": launches + args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
            "clocks": clocks.summary(),
        }
        if world > 1:
            line["sharded_bit_exact"] = {"replicas_identical": replicas_identical,
                                         "rank0_equals_oracle_new_parallel_over_all_keys": (cpu or {}).get("bit_exact_vs_gpu")}
        if extra is not None:
            line["other_configs"] = extra
        if north_star is not None:
            line["north_star_1e9"] = north_star
    if comm is not None:
        del m
        torch.cuda.synchronize()
        dist.barrier()
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        # the JSON line must be the LAST thing on stdout: libraries that write through C stdio (NCCL's version
        # banner) sit in a buffer that is otherwise flushed at exit, i.e. after Python's own output
        try:
            C.CDLL(None).fflush(None)
        except Exception:
            pass
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
