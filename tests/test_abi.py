"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "boomphf_cuda.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"BPHF_API[^;(]*?\b(bphf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("bphf_build_u64", "bphf_hash_batch_u64", "bphf_level_export", "bphf_from_levels", "bphf_free",
              "bphf_last_error", "bphf_map_permute_u64", "bphf_map_get_batch_u64"):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(pkg.LIB_PATH)
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    for s in declared:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(pkg.SYMBOLS) == declared


def test_only_c_abi_symbols_are_exported(pkg):
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH], text=True)
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("bphf_") for n in names), names


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "boomphf_cuda.h"\nint main(void){bphf_build_stats s; (void)s; return BPHF_OK;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           str(src), "-o", str(tmp_path / "t.o")])


def test_switch_agrees_with_oracle(pkg, oracle):
    assert pkg.lib().bphf_wy_swap_8byte_tail() == oracle.lib().bo_wy_swap_8byte_tail()


def test_gamma_is_rejected_before_any_device_work(pkg):
    keys = np.arange(10, dtype=np.uint64)
    for g in (1.0, 1.01, float("nan")):
        with pytest.raises(pkg.MphfPanic) as e:
            pkg.Mphf.new(g, keys)
        assert e.value.code == -2 and "gamma > 1.01" in str(e.value)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu(pkg):
    with pytest.raises(pkg.MphfPanic) as e:
        pkg.Mphf.new(1.7, np.arange(10, dtype=np.uint64))
    assert e.value.code == -4  # BPHF_ERR_CUDA


def test_product_package_never_touches_the_oracle():
    pkg_dir = os.path.join(ROOT, "rust-boomphf_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "boomphf_oracle" not in text and "oracle/" not in text, f


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_every_compute_entry_point_fails_loudly_without_a_gpu(pkg):
    import importlib
    sh = importlib.import_module(pkg.__name__ + ".sharded")
    keys = np.arange(100, dtype=np.uint64)
    for build in (lambda: pkg.Mphf.new_parallel(1.7, keys, None),
                  lambda: pkg.Mphf.new(1.7, keys.astype(np.uint32)),
                  lambda: pkg.Mphf.from_chunked_iterator(1.7, [keys[:50], keys[50:]], 100),
                  lambda: pkg.Mphf.from_chunked_iterator_parallel(1.7, [keys], None, 100),
                  lambda: pkg.Mphf.from_levels([(255, np.zeros(4, np.uint64), np.zeros(1, np.uint64))]),
                  lambda: sh.ShardComm.create(0, 0, 2, 1000),
                  lambda: pkg.BoomHashMap.new(keys, keys)):
        with pytest.raises(pkg.MphfPanic) as e:
            build()
        assert e.value.code == -4, str(e.value)   # BPHF_ERR_CUDA: there is no CPU path to fall back to


def test_only_tests_smoke_and_bench_use_the_oracle():
    # oracle/ is test infrastructure: nothing outside tests/, __graft_entry__.smoke()/build() and bench.py may touch it
    offenders = []
    for d in ("tools", "rust-boomphf_b200", "include", "profiles"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if "load_oracle" in text or "boomphf_oracle" in text:
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
