"""Host mirror of the FRI prover, plonky2/src/fri/prover.rs: `fri_proof` (:22-77) = `fri_committed_trees` (:83-133) +
`fri_proof_of_work` (:136-188) + `fri_prover_query_rounds` (:190-255), same arguments and return values.  Per layer the CUDA library bit-reverses and chunks the
evaluations into a leaf matrix, builds its Merkle tree, folds the coefficients with the challenge and runs the coset
NTT of the folded polynomial (p2g_fri_*); the Fiat-Shamir challenger runs here, between the two calls, exactly where
the reference has it.  The layer trees stay on the GPU (`DeviceMerkleTree`) and serve the query phase
(`fri_prover_query_round`, :219-255) through gather kernels."""
import ctypes
from dataclasses import dataclass, field as dc_field
from typing import List

import numpy as np

from . import _lib
from .fields import Field, REPR_CANONICAL
from .merkle_tree import DeviceMerkleTree, MerkleCap


@dataclass
class FriConfig:
    """fri/mod.rs FriConfig: the members the prover reads."""
    rate_bits: int
    cap_height: int
    proof_of_work_bits: int = 16
    num_query_rounds: int = 28


@dataclass
class FriParams:
    """fri/mod.rs FriParams: config + reduction_arity_bits."""
    config: FriConfig
    reduction_arity_bits: List[int] = dc_field(default_factory=list)


def fri_committed_trees(field: Field, coeffs, values, challenger, fri_params: FriParams, ctx=None, repr: int = REPR_CANONICAL):
    """coeffs, values: [n][D] arrays (lde_polynomial_coeffs / lde_polynomial_values as base-field slices); values may be
    None, in which case the library computes them as coeffs.coset_fft(F::generator()) on the device (what
    prove_openings does on the CPU, fri/oracle.rs:226-231).
    Returns (trees, final_coeffs) like the reference; `challenger` is advanced exactly as the reference advances it."""
    ctx = ctx or _lib.default_context()
    d = field.extension_degree
    coeffs = np.ascontiguousarray(np.asarray(coeffs, dtype=field.dtype))
    if values is not None:
        values = np.ascontiguousarray(np.asarray(values, dtype=field.dtype))
    if coeffs.ndim != 2 or coeffs.shape[1] != d or (values is not None and values.shape != coeffs.shape):
        raise ValueError("coeffs and values must both be [n][%d] arrays" % d)
    n = coeffs.shape[0]
    if n <= 0 or n & (n - 1):
        raise ValueError("Not a power of two: %d" % n)  # log2_strict
    log_n = n.bit_length() - 1
    state = ctypes.c_void_p()
    ctx.check(ctx.lib.p2g_fri_begin(ctx.handle, field.field_id, field.hash_id, repr, d, log_n, coeffs.ctypes.data,
                                    values.ctypes.data if values is not None else None, 0, ctypes.byref(state)))
    trees = []
    try:
        for arity_bits in fri_params.reduction_arity_bits:
            h = fri_params.config.cap_height
            cap = np.empty((1 << h, field.num_hash_out_elts), dtype=field.dtype)
            tree_handle = ctypes.c_void_p()
            rc = ctx.lib.p2g_fri_commit_layer(ctx.handle, state, arity_bits, h, cap.ctypes.data, ctypes.byref(tree_handle))
            if tree_handle:
                trees.append(DeviceMerkleTree(field, ctx, tree_handle, MerkleCap(cap)))
            ctx.check(rc)
            challenger.observe_cap(trees[-1].cap)
            beta = np.ascontiguousarray(challenger.get_extension_challenge(), dtype=field.dtype)
            ctx.check(ctx.lib.p2g_fri_fold(ctx.handle, state, beta.ctypes.data))
        lg, shift = ctypes.c_uint32(), ctypes.c_uint64()
        ctx.check(ctx.lib.p2g_fri_shape(state, ctypes.byref(lg), ctypes.byref(shift)))
        count = (1 << lg.value) >> fri_params.config.rate_bits
        final = np.empty((count, d), dtype=field.dtype)
        got = ctypes.c_uint64()
        ctx.check(ctx.lib.p2g_fri_finish(ctx.handle, state, fri_params.config.rate_bits, final.ctypes.data if count else None,
                                         ctypes.byref(got)))
        assert got.value == count
    finally:
        ctx.lib.p2g_fri_free(ctx.handle, state)
    challenger.observe_extension_elements(final)
    return trees, final


def fri_proof_of_work(field: Field, challenger, config: FriConfig, ctx=None) -> int:
    """fri_proof_of_work (:136-188): grind on the GPU (p2g_fri_pow), then observe the witness and re-check the
    response through the ordinary challenger exactly as the reference does (:181-186)."""
    ctx = ctx or _lib.default_context()
    min_leading_zeros = config.proof_of_work_bits + (64 - field.order.bit_length())
    state = np.ascontiguousarray(challenger.sponge_state, dtype=field.dtype)
    inp = np.ascontiguousarray(np.array(challenger.input_buffer, dtype=field.dtype))
    witness = ctypes.c_uint64()
    ctx.check(ctx.lib.p2g_fri_pow(ctx.handle, field.field_id, field.hash_id, state.ctypes.data,
                                  inp.ctypes.data if inp.size else None, inp.size, min_leading_zeros, ctypes.byref(witness)))
    challenger.observe_element(witness.value)
    pow_response = challenger.get_challenge()
    assert 64 - int(pow_response).bit_length() >= min_leading_zeros
    return int(witness.value)


def fri_prover_query_rounds(initial_merkle_trees, trees, challenger, n: int, fri_params: FriParams):
    """fri_prover_query_rounds / fri_prover_query_round (:190-255).  All rounds' rows and Merkle paths of one tree are
    fetched with one gather launch each (the trees live on the GPU).  Returns, per round,
    {"x_index", "initial_trees_proof": [(leaf row, siblings)], "steps": [(evals [arity][D], siblings)]}."""
    indices = [int(r) % n for r in challenger.get_n_challenges(fri_params.config.num_query_rounds)]
    rounds = [{"x_index": x, "initial_trees_proof": [], "steps": []} for x in indices]
    for t in initial_merkle_trees:
        rows, proofs = _get_and_prove(t, indices)
        for rd, row, pr in zip(rounds, rows, proofs):
            rd["initial_trees_proof"].append((row, pr))
    cur = list(indices)
    for i, tree in enumerate(trees):
        arity_bits = fri_params.reduction_arity_bits[i]
        cur = [x >> arity_bits for x in cur]
        rows, proofs = _get_and_prove(tree, cur)
        for rd, row, pr in zip(rounds, rows, proofs):
            rd["steps"].append((np.asarray(row).reshape(1 << arity_bits, -1), pr))  # unflatten
    return rounds


def _get_and_prove(tree, indices):
    if hasattr(tree, "get_many"):
        return tree.get_many(indices), tree.prove_many(indices)
    return [tree.get(i) for i in indices], [tree.prove(i) for i in indices]


def fri_proof(field: Field, initial_merkle_trees, lde_polynomial_coeffs, lde_polynomial_values, challenger, fri_params: FriParams,
              ctx=None, repr: int = REPR_CANONICAL):
    """fri_proof (:22-77): {"commit_phase_merkle_caps", "query_round_proofs", "final_poly", "pow_witness"}."""
    n = np.asarray(lde_polynomial_coeffs).shape[0]
    assert lde_polynomial_values is None or np.asarray(lde_polynomial_values).shape[0] == n
    trees, final_coeffs = fri_committed_trees(field, lde_polynomial_coeffs, lde_polynomial_values, challenger, fri_params, ctx=ctx, repr=repr)
    pow_witness = fri_proof_of_work(field, challenger, fri_params.config, ctx=ctx)
    rounds = fri_prover_query_rounds(initial_merkle_trees, trees, challenger, n, fri_params)
    proof = {"commit_phase_merkle_caps": [t.cap for t in trees], "query_round_proofs": rounds, "final_poly": final_coeffs,
             "pow_witness": pow_witness}
    for t in trees:
        t.free()
    return proof
