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


def test_ntt_tile_code_on_the_host_matches_the_definition(tmp_path):
    """The headers the NTT kernels are built from (gl_limbs / ntt_tile / ntt_kinds .cuh), compiled with g++ and run
    thread by thread: every round plan (tile sizes 2^0 .. 2^11), one-, two- and three-pass transforms, both directions,
    both fields, coset batching and the bit-reversed leaf order, against a naive O(n^2) evaluation."""
    src = os.path.join(ROOT, "tests", "cpp", "ntt_tile_emul_test.cpp")
    exe = str(tmp_path / "ntt_tile_emul_test")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wno-attributes", "-I", cuda_inc, src, "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "all ok" in res.stdout and "FAIL" not in res.stdout


def test_poseidon2_bb_kernel_code_on_the_host_matches_the_oracle(tmp_path):
    """The headers the BabyBear hash kernels are built from (field_bb.cuh, poseidon2_bb.cuh), compiled with g++: the
    permutation with its signed S-box products and lazy internal rounds, every stated range asserted, against the
    oracle on extreme, random and chained states."""
    from oracle import p2o
    p2o.load()
    src = os.path.join(ROOT, "tests", "cpp", "poseidon2_bb_emul_test.cpp")
    exe = str(tmp_path / "poseidon2_bb_emul_test")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    orc = os.path.join(ROOT, "oracle")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wno-attributes", "-Wno-unknown-pragmas", "-I", cuda_inc, "-I", orc, src, "-o", exe,
                    "-L", orc, "-l:libp2o.so", "-Wl,-rpath," + orc], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, (res.stdout + res.stderr)[-4000:]
    assert "all ok" in res.stdout
