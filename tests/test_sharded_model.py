"""N > 1 host logic on the CPU (gloo, world_size 2 and 3): the sharded schedule of SURVEY.md 8e -- every shard marks
its keys into its own (a, collide), the pairs are folded with the OR-collide monoid slice by slice
(reduce-scatter), the merged slices are gathered, every shard filters its own keys and derives the same k(l+1) --
restated with the oracle's level primitives and torch.distributed, must reproduce the unsharded oracle build.
This is the CPU model of csrc/shard_kernels.cuh (same slicing, same fold order)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, gamma, seed0, how, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge
        orc = ge.load_oracle()
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        keys_all = orc.keygen(n, seed=17, kind=0)
        if how == "even":
            base, rem = divmod(n, world)
            b = [0]
            for r in range(world):
                b.append(b[-1] + base + (1 if r < rem else 0))
        else:  # everything on the last rank
            b = [0] * world + [n]
        keys = keys_all[b[rank]:b[rank + 1]].copy()

        def allsum(v):
            t = torch.tensor([int(v)], dtype=torch.int64)
            dist.all_reduce(t)
            return int(t.item())

        k_global = allsum(len(keys))
        assert k_global == n
        levels, level = [], 0
        while True:
            bits = orc.level_size(gamma, k_global)
            w = (bits + 63) // 64
            wpad = (w + 7) & ~7
            a = np.zeros(wpad, dtype=np.uint64)
            c = np.zeros(wpad, dtype=np.uint64)
            orc.mark_level(keys, seed0 + level, bits, a, c)
            # "peer memory": every shard can read every shard's pair
            both = torch.from_numpy(np.concatenate([a, c]).view(np.int64))
            got = [torch.empty_like(both) for _ in range(world)]
            dist.all_gather(got, both)
            pa = [g.numpy().view(np.uint64)[:wpad] for g in got]
            pc = [g.numpy().view(np.uint64)[wpad:] for g in got]
            # reduce-scatter: this shard folds its slice (pairs of words, as the kernel slices)
            pairs = wpad // 2
            per = (pairs + world - 1) // world
            lo, hi = min(pairs, per * rank) * 2, min(pairs, per * rank + per) * 2
            ma, mc = pa[0][lo:hi].copy(), pc[0][lo:hi].copy()
            for r in range(1, world):
                ma, mc = orc.or_collide(ma, mc, pa[r][lo:hi], pc[r][lo:hi])
            fin = ma & ~mc
            # all-gather of the merged slices (padded to the slice size)
            sl = np.zeros(4 * per, dtype=np.uint64)
            sl[:hi - lo] = fin
            sl[2 * per:2 * per + hi - lo] = mc
            t = torch.from_numpy(sl.view(np.int64))
            got = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(got, t)
            full_a = np.concatenate([g.numpy().view(np.uint64)[:2 * per] for g in got])[:wpad]
            full_c = np.concatenate([g.numpy().view(np.uint64)[2 * per:] for g in got])[:wpad]
            uniq = int(np.unpackbits(full_a.view(np.uint8)).sum())
            levels.append((bits, full_a[:w].copy()))
            keys = orc.filter_level(keys, seed0 + level, bits, full_c)
            k_next = k_global - uniq
            assert allsum(len(keys)) == k_next          # what the merge barrier checks on the device
            level += 1
            assert level <= 101
            k_global = k_next
            if k_global == 0:
                break
        ranks = orc.ranks_from_levels([w for _, w in levels])
        ref = orc.OracleMphf.new_parallel(gamma, keys_all, seed0 if seed0 else None).levels()
        ok = len(ref) == len(levels) and all(
            b0 == b1 and np.array_equal(w0, w1) and np.array_equal(r0, r1)
            for (b0, w0), r0, (b1, w1, r1) in zip(levels, ranks, ref))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), len(levels)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("world,n,gamma,seed0,how", [
    (2, 20000, 1.7, 0, "even"),
    (2, 3001, 2.0, 5, "even"),
    (3, 12345, 1.7, 0, "even"),
    (2, 5000, 1.7, 0, "last"),
])
def test_sharded_schedule_reproduces_the_unsharded_build(world, n, gamma, seed0, how):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, gamma, seed0, how, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert ok, (rank, info)


def test_or_collide_is_associative_commutative_with_identity(oracle):
    rng = np.random.default_rng(5)
    xs = [(rng.integers(0, 2 ** 63, 64, dtype=np.uint64), None) for _ in range(3)]
    xs = [(a, a & rng.integers(0, 2 ** 63, 64, dtype=np.uint64)) for a, _ in xs]  # collide is a subset of a
    (a0, c0), (a1, c1), (a2, c2) = xs
    l = oracle.or_collide(*oracle.or_collide(a0, c0, a1, c1), a2, c2)
    r = oracle.or_collide(a0, c0, *oracle.or_collide(a1, c1, a2, c2))
    assert np.array_equal(l[0], r[0]) and np.array_equal(l[1], r[1])
    s = oracle.or_collide(a1, c1, a0, c0)
    t = oracle.or_collide(a0, c0, a1, c1)
    assert np.array_equal(s[0], t[0]) and np.array_equal(s[1], t[1])
    z = np.zeros(64, dtype=np.uint64)
    i = oracle.or_collide(a0, c0, z, z)
    assert np.array_equal(i[0], a0) and np.array_equal(i[1], c0)


def test_shard_bounds(pkg):
    import importlib
    sh = importlib.import_module(pkg.__name__ + ".sharded")
    assert sh.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert sh.shard_bounds(0, 4) == [0, 0, 0, 0, 0]
    assert sh.shard_bounds(5, 8)[-1] == 5
