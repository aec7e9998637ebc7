"""Host mirror of the FRI commit phase, `fri_committed_trees` (plonky2/src/fri/prover.rs:83-133): same arguments, same
return value (the layer trees and the final polynomial).  Per layer the CUDA library bit-reverses and chunks the
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
    """fri/mod.rs FriConfig: the two members the commit phase reads."""
    rate_bits: int
    cap_height: int


@dataclass
class FriParams:
    """fri/mod.rs FriParams: config + reduction_arity_bits."""
    config: FriConfig
    reduction_arity_bits: List[int] = dc_field(default_factory=list)


def fri_committed_trees(field: Field, coeffs, values, challenger, fri_params: FriParams, ctx=None, repr: int = REPR_CANONICAL):
    """coeffs, values: [n][D] arrays (lde_polynomial_coeffs / lde_polynomial_values as base-field slices).
    Returns (trees, final_coeffs) like the reference; `challenger` is advanced exactly as the reference advances it."""
    ctx = ctx or _lib.default_context()
    d = field.extension_degree
    coeffs = np.ascontiguousarray(np.asarray(coeffs, dtype=field.dtype))
    values = np.ascontiguousarray(np.asarray(values, dtype=field.dtype))
    if coeffs.ndim != 2 or coeffs.shape[1] != d or values.shape != coeffs.shape:
        raise ValueError("coeffs and values must both be [n][%d] arrays" % d)
    n = coeffs.shape[0]
    if n <= 0 or n & (n - 1):
        raise ValueError("Not a power of two: %d" % n)  # log2_strict
    log_n = n.bit_length() - 1
    state = ctypes.c_void_p()
    ctx.check(ctx.lib.p2g_fri_begin(ctx.handle, field.field_id, field.hash_id, repr, d, log_n, coeffs.ctypes.data,
                                    values.ctypes.data, 0, ctypes.byref(state)))
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
