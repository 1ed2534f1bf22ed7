"""The C++ host-side mirrors (csrc/host/alore_host.hpp) compile against include/alore_b200.h and link with
libalore_b200.so (CPU test); on the GPU box the resulting program runs the drop-in path end to end."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "alore_legged_manipulator_b200"


def build(tmp_path):
    exe = tmp_path / "host_smoke"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", str(ROOT / "tests" / "cpp" / "host_smoke.cpp"), "-o", str(exe),
           f"-L{PKG}", "-lalore_b200", f"-Wl,-rpath,{PKG}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_host_compiles_and_links(tmp_path):
    assert build(tmp_path).exists()


@pytest.mark.gpu
def test_cpp_host_runs_drop_in_path(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_SMOKE_OK" in r.stdout, r.stdout + r.stderr
