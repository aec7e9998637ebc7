"""The reference's frozen Goldilocks proof (recursion/regression_test_data.rs:93) re-verified with the oracle's
permutation and field constants: Fiat-Shamir transcript + FRI query checks (tools/gl_fri_check.py).  This is what pins
the coset shift, the two-adic root chain and the bit-reversed leaf order of from_values to bytes the reference produced."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gl_fri_check  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "gl_proof_fri.json")


def _oracle_hooks(oracle):
    lib = oracle.load()

    def permute(state):
        return [int(x) for x in oracle.permute(oracle.GL, np.array(state, dtype=np.uint64))]

    return permute, int(lib.p2o_generator(oracle.GL)), lambda b: int(lib.p2o_two_adic_generator(oracle.GL, b))


def test_golden_proof_fri_checks_with_oracle_constants(oracle):
    fx = json.load(open(FIXTURE))
    permute, g, tag = _oracle_hooks(oracle)
    res = gl_fri_check.verify(fx, permute, g, tag)
    assert res["queries_checked"] == len(fx["queries"]) >= 6
    assert res["query_indices"] == fx["query_indices_from_merkle_paths"]


@pytest.mark.parametrize("what", ["generator", "lde_root", "leaf_order"])
def test_golden_proof_rejects_wrong_conventions(oracle, what):
    """The check has teeth: a different coset shift, a different primitive root of the same order, or natural instead of
    bit-reversed leaf order all make the frozen proof fail."""
    fx = json.load(open(FIXTURE))
    permute, g, tag = _oracle_hooks(oracle)
    p = gl_fri_check.P
    if what == "generator":
        args = (permute, 11, tag)
    elif what == "lde_root":
        args = (permute, g, lambda b: pow(tag(b), 3, p))  # another generator of the same 2^b-subgroup
    else:
        orig = gl_fri_check.reverse_bits
        try:
            gl_fri_check.reverse_bits = lambda x, n: x
            with pytest.raises(AssertionError):
                gl_fri_check.verify(fx, permute, g, tag)
        finally:
            gl_fri_check.reverse_bits = orig
        return
    with pytest.raises(AssertionError):
        gl_fri_check.verify(fx, *args)


@pytest.mark.gpu
def test_golden_proof_transcript_with_gpu_permutation(ctx, oracle):
    """Same verification with every sponge permutation executed by the CUDA kernel (p2g_permute)."""
    fx = json.load(open(FIXTURE))
    _, g, tag = _oracle_hooks(oracle)

    def permute(state):
        s = np.array(state, dtype=np.uint64).reshape(1, 12)
        ctx.check(ctx.lib.p2g_permute(ctx.handle, 0, 0, s.ctypes.data, 1))
        return [int(x) for x in s[0]]

    res = gl_fri_check.verify(fx, permute, g, tag)
    assert res["queries_checked"] == len(fx["queries"])
