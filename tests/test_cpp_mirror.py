"""The C++ host mirror (include/p2g/commit.hpp): compiles everywhere (CPU check), runs against the oracle on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    pkg = os.path.join(ROOT, "plonky2_goldibear_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), SRC, "-o", exe,
           "-L", pkg, "-l:libp2g_cuda.so", "-L", os.path.join(ROOT, "oracle"), "-l:libp2o.so",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    import plonky2_goldibear_b200 as p2g
    p2g.load_library()
    from oracle import p2o
    p2o.load()
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([exe], capture_output=True, text=True)
        assert res.returncode == 2 and "no CPU fallback" in res.stderr  # p2g_init refuses, nothing is computed on the CPU


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    from oracle import p2o
    p2o.load()
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "host mirror ok" in res.stdout
