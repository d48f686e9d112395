"""Resource-usage guard on the compiled sm_100a kernels (cuobjdump on the in-tree library, no GPU needed): the u64
hot-path kernels must keep the register counts their occupancy was tuned for, without spills, and the stream-key
variants must be SEPARATE kernels -- folding the stream hash into the u64 kernels as a runtime branch once doubled
their registers (cascade 64 -> 128, pass 40 -> 80)."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def usage(pkg):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "--dump-resource-usage", pkg.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    res, name = {}, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
            name = None
    assert res, "no kernels found in the library"
    return res


def _one(usage, *parts):
    hits = [(k, v) for k, v in usage.items() if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, [k for k, _ in hits])
    return hits[0][1]


@pytest.mark.parametrize("parts,max_reg", [
    (("mark_list_kernelILb0E", "KeyCfg"), 40),     # level 0: 6 CTAs of 256 threads per SM
    (("pass_kernelILb0E", "KeyCfg"), 128),         # level kernels: 2 CTAs of 256 threads per SM (64 KB of warp rings each)
    (("cascade_kernelILb0E",), 128),
    (("query_sync_kernel",), 64),                  # 8 CTAs of 128 threads per SM
    (("tail_kernel", "KeyCfg"), 64),
    (("qprobe_kernel",), 80),                      # partitioned lookup: 3 CTAs of 256 threads per SM, no spills
    (("qfirst_kernel",), 48),                      # 5 CTAs per SM
    (("qrestore_kernelILb0E",), 32),               # a pure gather: 8 CTAs per SM
    (("qwalk_kernel",), 64),
])
def test_hot_kernels_keep_their_registers_and_do_not_spill(usage, parts, max_reg):
    reg, stack, _ = _one(usage, *parts)
    assert reg <= max_reg, (parts, reg)
    # the level kernels may park a few loop-invariant words (<= 32 bytes); anything more means the pipeline state spilled
    assert stack <= (32 if max_reg == 128 else 0), (parts, "spills / stack frame", stack)


def test_stream_key_variants_are_separate_kernels(usage):
    for parts in (("mark_list_kernel", "StreamCfg"), ("pass_kernel", "StreamCfg"), ("tail_kernel", "StreamCfg")):
        _one(usage, *parts)
    assert sum(1 for k in usage if "query_kernel" in k and "StreamCfg" in k) == 2
    assert not any("cascade_kernel" in k and "StreamCfg" in k for k in usage)


def test_level_kernels_keep_their_rings_in_dynamic_shared_memory(usage):
    # per warp: 512 staged survivors + a two-deep TMA key ring of 256-key tiles = 8 KB, 64 KB per CTA, requested at launch
    # (PASS_SMEM_BYTES); the static part is only control words
    for parts in (("cascade_kernelILb0E",), ("pass_kernelILb0E", "KeyCfg")):
        _, _, shared = _one(usage, *parts)
        assert shared <= 2048


def _sass(pkg, usage, *parts):
    name = [k for k in usage if all(p in k for p in parts)]
    assert len(name) == 1, (parts, name)
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", "-fun", name[0], pkg.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    return out.stdout


def test_level_kernels_use_the_bulk_copy_engine_and_fire_and_forget_reductions(pkg, usage):
    """what the design claims, read off the machine code: key tiles arrive through TMA bulk copies completing on an
    mbarrier (UBLKCP + SYNCS), the occupancy bit is a returning atomic OR, the collide bit a non-returning reduction"""
    sass = _sass(pkg, usage, "cascade_kernelILb0E")
    for mnemonic in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "ATOMG.E.OR", "REDG.E.OR"):
        assert mnemonic in sass, mnemonic
    mark = _sass(pkg, usage, "mark_list_kernelILb0E", "KeyCfg")
    assert mark.count("ATOMG.E.OR") >= 8 and mark.count("REDG.E.OR") >= 8     # 8 keys per thread, issued back to back
    look = _sass(pkg, usage, "query_sync_kernel")
    assert "LDG.E.128" in look                                                # one 16-byte granule per probe
    probe = _sass(pkg, usage, "qprobe_kernel")                                # partitioned lookup: units arrive by TMA
    for mnemonic in ("UBLKCP", "SYNCS.PHASECHK", "LDG.E.128", "ATOMS.ADD"):
        assert mnemonic in probe, mnemonic
