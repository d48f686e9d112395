"""include/boomphf.hpp -- the header-only C++ mirror of the reference's Rust interface -- compiled with g++ against
libboomphf_cuda.so and run as a separate program (tests/cpp/mirror_example.cpp).  Without a GPU the program must stop
at the first constructor with the library's "no CPU fallback" panic; on a B200 it checks the reference's own test
property for six key types, the chunked constructors, the data-model round trip and both panics."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def example(pkg, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "mirror_example")
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_example.cpp"), "-o", exe, "-L", libdir,
                           "-lboomphf_cuda", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_builds_and_has_no_cpu_fallback(example):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([example], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, (out.returncode, out.stderr[-500:])
    assert "panic (-4)" in out.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_the_gpu(example):
    out = subprocess.run([example], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stderr[-1000:])
    assert "all checks passed" in out.stdout
