"""A host that is not Python: tests/c_host/plan_host.c drives the C ABI (include/ctrlv_b200.h) directly — plain
pointers, a cudaStream_t, no torch anywhere — records a launch plan and replays it."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_records_and_replays_a_plan(tmp_path):
    from ctrlv_b200 import _lib
    _lib.load()  # the library must exist (no fallback)
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "plan_host")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(ROOT, "tests", "c_host", "plan_host.c"), "-o", exe,
                           "-L", libdir, "-l:libctrlv_b200.so", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "plan replay == direct result" in out.stdout
