"""bench.py contract checks that need no GPU: the reference arm (the oracle's restatement of new_parallel on the host
cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-keys", "300000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mphf_build_throughput" and line["unit"] == "Mkeys/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--ref-keys", "1000"], capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_bench_line_on_the_gpu_has_every_contract_key():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--keys", "4000000", "--steps", "3", "--warmup", "3",
                          "--cpu-baseline-keys", "4000000", "--no-extra-configs"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])   # the JSON line is the last line of stdout
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "roofline", "cpu_baseline", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["gpu_launches"] > 0 and line["dtype"] == "u64"
    assert line["e2e"]["h2d_bytes_per_step"] == 8 * 4000000 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["bit_exact_vs_gpu"] is True
    assert line["lookup"]["permutation_ok"] is True
    la = line["roofline"]["l2_atomic"]   # the L2-atomic ceiling, measured in process
    assert la["ceilings_measured_Gops"]["atom"] > 1 and la["ops_per_build_per_gpu"]["fetch_or"] > 4000000 and 0 < la["frac_of_ceiling"] < 1.5
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


# ---- the warm-up second of a multi-rank run is a collective decision (gloo, world size 2, CPU) -----------------
def _spin_worker(rank, world, port, q):
    try:
        import time
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, ROOT)
        import bench
        if rank == 0:
            time.sleep(0.25)  # rank 0 starts late (it is the one that initialises NVML): its clock runs behind
        _, calls = bench.collective_spin(lambda: time.sleep(0.01 * (1 + rank)), 0.4, world, "cpu")
        q.put((rank, calls))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))


def test_ranks_agree_on_the_number_of_warmup_builds():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_spin_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert isinstance(got[0], int) and got[0] == got[1] and got[0] >= 2, got
