"""Sharded (multi-GPU) construction: host side of SURVEY.md 8e over the C ABI.

The reference's `Mphf::new_parallel` (/root/reference/src/lib.rs:363-408) shares one (a, collide) pair per level
between its rayon workers.  Here the key set is split into `world` shards, one per GPU; per level every shard marks
its keys into its own full-width pair and the pairs are folded with the OR-collide monoid by a peer-memory kernel
(csrc/shard_kernels.cuh).  Every shard returns the identical complete `Mphf` (a replica for lookups).

  * one process per GPU (torchrun):  `ShardComm.from_process_group(...)` exchanges the 64-byte IPC handles of the
    shards' arenas with `torch.distributed.all_gather` (any backend; gloo works) and maps the peers;
  * shards as threads of one process (tests; one or several devices): `ShardComm.local_group(...)`.
"""
import ctypes as C
import threading

import numpy as np

from . import _native as N
from .mphf import Mphf, _as_keys, _check

IPC_HANDLE_BYTES = 64


def shard_bounds(n, world):
    """contiguous, near-equal key shards: shard r owns [bounds[r], bounds[r+1])"""
    base, rem = divmod(int(n), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


class ShardComm:
    """One shard's end of the peer-memory communicator (bphf_comm)."""

    def __init__(self, handle, device, rank, world):
        self._h = C.c_void_p(handle)
        self.device, self.rank, self.world = int(device), int(rank), int(world)

    @classmethod
    def create(cls, device, rank, world, max_keys_global, max_gamma=1.7):
        out = C.c_void_p()
        _check(N.lib().bphf_comm_create(int(device), int(rank), int(world), int(max_keys_global), float(max_gamma),
                                        C.byref(out)))
        return cls(out.value, device, rank, world)

    def ipc_handle(self):
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        _check(N.lib().bphf_comm_ipc_handle(self._h, buf))
        return bytes(buf)

    def connect_ipc(self, handles):
        """handles: the ipc_handle() of every shard, in rank order"""
        if len(handles) != self.world or any(len(h) != IPC_HANDLE_BYTES for h in handles):
            raise ValueError("need one %d-byte handle per shard" % IPC_HANDLE_BYTES)
        blob = b"".join(handles)
        _check(N.lib().bphf_comm_connect_ipc(self._h, C.c_char_p(blob)))

    @classmethod
    def from_process_group(cls, device, max_keys_global, max_gamma=1.7, group=None):
        """One process per GPU: exchange the arena handles over torch.distributed and map the peers."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = cls.create(device, rank, world, max_keys_global, max_gamma)
        if world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            dev = torch.device("cuda", device) if on_gpu else torch.device("cpu")
            mine = torch.tensor(list(c.ipc_handle()), dtype=torch.uint8, device=dev)
            got = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(got, mine, group=group)
            c.connect_ipc([bytes(g.cpu().tolist()) for g in got])
            dist.barrier(group)  # every shard has mapped every arena before anyone builds
        return c

    @classmethod
    def local_group(cls, devices, max_keys_global, max_gamma=1.7):
        """All shards in this process (driven by one thread each); devices[r] is shard r's GPU."""
        world = len(devices)
        comms = [cls.create(devices[r], r, world, max_keys_global, max_gamma) for r in range(world)]
        arr = (C.c_void_p * world)(*[c._h.value for c in comms])
        for c in comms:
            _check(N.lib().bphf_comm_connect_local(c._h, arr))
        return comms

    def build(self, gamma, shard_keys, n_global, starting_seed=None):
        """Collective: Mphf::new_parallel(gamma, union of all shards' keys, starting_seed) -> this shard's replica."""
        ptr, n, mem, dev, keep = _as_keys(shard_keys)
        if dev is not None and dev != self.device:
            raise ValueError("shard keys live on another device than the comm")
        out = C.c_void_p()
        rc = N.lib().bphf_build_sharded_u64(self._h, ptr, n, int(n_global), float(gamma),
                                            0 if starting_seed is None else 1,
                                            0 if starting_seed is None else int(starting_seed), mem, C.byref(out))
        del keep
        _check(rc)
        return Mphf(out.value, self.device)

    def close(self):
        if getattr(self, "_h", None):
            N.lib().bphf_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def new_parallel_sharded(gamma, shard_keys, comm, n_global=None, starting_seed=None, group=None):
    """`Mphf::new_parallel` over a key set sharded across the ranks of a torch.distributed group."""
    n_local = int(shard_keys.numel() if hasattr(shard_keys, "numel") else len(shard_keys))
    if n_global is None:
        import torch
        import torch.distributed as dist
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor([n_local], dtype=torch.int64, device=torch.device("cuda", comm.device) if on_gpu else "cpu")
        dist.all_reduce(t, group=group)
        n_global = int(t.item())
    return comm.build(gamma, shard_keys, n_global, starting_seed)


def build_in_threads(comms, gamma, shards, starting_seed=None):
    """Drive the shards of a local_group from one thread each (the build is a blocking collective)."""
    n_global = sum(int(s.numel() if hasattr(s, "numel") else len(s)) for s in shards)
    out, err = [None] * len(comms), [None] * len(comms)

    def run(r):
        try:
            out[r] = comms[r].build(gamma, shards[r], n_global, starting_seed)
        except Exception as e:  # noqa: BLE001
            err[r] = e

    ts = [threading.Thread(target=run, args=(r,)) for r in range(len(comms))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out
