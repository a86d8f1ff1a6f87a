"""Builds tests/cpp/host_api_test.cpp (the reference's own search tests replayed against the C++ host mirror
include/comet.hpp over the C ABI) and runs it: host-only logic on CPU, every index on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_test")


def build():
    from comet_b200 import capi
    capi.lib()      # makes sure libcomet_b200.so exists
    libdir = os.path.join(ROOT, "comet_b200")
    src = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", libdir, "-lcomet_b200", "-Wl,-rpath," + libdir])


def test_host_logic_cpu():
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([EXE, "--cpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_reference_search_tests_through_cpp_mirror():
    build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
