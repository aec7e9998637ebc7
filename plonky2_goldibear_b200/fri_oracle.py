"""Host mirror of `PolynomialBatch` (plonky2/src/fri/oracle.rs:29-158).  `from_values` / `from_coeffs` keep the
reference's argument meaning and error behaviour; the work happens in p2g_commit on the GPU."""
import ctypes
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .fields import Field, INPUT_COEFFS, INPUT_VALUES, REPR_CANONICAL
from .merkle_tree import DeviceMerkleTree, MerkleCap, MerkleTree, log2_strict, node_slot

SALT_SIZE = 4  # oracle.rs:25

STAGES = ("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree")  # timed! scopes, oracle.rs:76-114


class ProvingProcessInfo:
    """Minimal stand-in for util/proving_process_info.rs: records the four scopes the hot path owns."""

    def __init__(self):
        self.scopes = []

    def record(self, name: str, seconds: float):
        self.scopes.append((name, seconds))

    def as_dict(self):
        out = {}
        for k, v in self.scopes:
            out[k] = out.get(k, 0.0) + v
        return out


def reverse_bits(n: int, num_bits: int) -> int:
    """plonky2/src/util/mod.rs:33-41"""
    r = 0
    for i in range(num_bits):
        r |= ((n >> i) & 1) << (num_bits - 1 - i)
    return r


def _ptr_array(arrays: Sequence[np.ndarray]):
    arr = (ctypes.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr


class PolynomialBatch:
    """Result container, same fields as the reference struct (oracle.rs:29-40)."""

    def __init__(self, field, polynomials, merkle_tree, degree_log, rate_bits, blinding):
        self.field = field
        self.polynomials = polynomials  # list of coefficient vectors (PolynomialCoeffs)
        self.merkle_tree = merkle_tree
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.blinding = blinding
        self.desc = None

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_values(cls, field: Field, values: Sequence[np.ndarray], rate_bits: int, blinding: bool, cap_height: int,
                    timing: Optional[ProvingProcessInfo] = None, fft_root_table=None, **kw) -> "PolynomialBatch":
        """oracle.rs:68-90.  `fft_root_table` is accepted for signature parity and ignored: twiddles are
        resident on the device."""
        return cls._commit(field, values, INPUT_VALUES, rate_bits, blinding, cap_height, timing, **kw)

    @classmethod
    def from_coeffs(cls, field: Field, polynomials: Sequence[np.ndarray], rate_bits: int, blinding: bool,
                    cap_height: int, timing: Optional[ProvingProcessInfo] = None, fft_root_table=None,
                    **kw) -> "PolynomialBatch":
        """oracle.rs:93-123."""
        return cls._commit(field, polynomials, INPUT_COEFFS, rate_bits, blinding, cap_height, timing, **kw)

    @classmethod
    def _commit(cls, field, columns, input_kind, rate_bits, blinding, cap_height, timing, ctx=None,
                repr: int = REPR_CANONICAL, salts: Optional[Sequence[np.ndarray]] = None, rng=None,
                shard_index: int = 0, shard_count: int = 0, want_leaves: bool = True, pinned_out: bool = False,
                keep_on_device: bool = False, leaf_pitch: int = 0, _reuse: Optional[dict] = None):
        ctx = ctx or _lib.default_context()
        if len(columns) == 0:
            raise IndexError("index out of bounds: the len is 0 but the index is 0")  # polynomials[0], oracle.rs:101
        cols = [np.ascontiguousarray(np.asarray(c, dtype=field.dtype)) for c in columns]
        degree = cols[0].shape[0]
        for c in cols:
            if c.ndim != 1 or c.shape[0] != degree:
                raise ValueError("Polynomial degrees inconsistent")  # assert_eq! at oracle.rs:139
        degree_log = log2_strict(degree)
        n_total = degree << rate_bits
        salt_cols: List[np.ndarray] = []
        if blinding:
            # oracle.rs:144-148 draws SALT_SIZE columns of F::rand_vec(N); randomness stays on the host side
            if salts is None:
                rng = rng or np.random.default_rng()
                salts = [rng.integers(0, field.order, size=n_total, dtype=np.uint64).astype(field.dtype)
                         for _ in range(SALT_SIZE)]
            salt_cols = [np.ascontiguousarray(np.asarray(s, dtype=field.dtype)) for s in salts]
            if len(salt_cols) != SALT_SIZE or any(s.shape != (n_total,) for s in salt_cols):
                raise ValueError("blinding needs SALT_SIZE salt columns of N elements")
        width = len(cols) + len(salt_cols)
        g = shard_count or 1
        rows = n_total // g
        lg = log2_strict(g)
        cap_local = max(cap_height - lg, 0)
        d = field.num_hash_out_elts
        alloc = ctx.pinned_empty if pinned_out else (lambda shape, dtype: np.empty(shape, dtype=dtype))
        if _reuse is not None:
            # output buffers handed back by an earlier call of the same shape (latency measurements: pinned allocations
            # are expensive and never freed before the context closes)
            base_alloc = alloc

            def alloc(shape, dtype, _n=[0]):
                key = (_n[0], tuple(shape), np.dtype(dtype).str)
                _n[0] += 1
                if key not in _reuse:
                    _reuse[key] = base_alloc(shape, dtype)
                return _reuse[key]
        if keep_on_device:
            want_leaves = False
        pitch = leaf_pitch or width
        if pitch < width:
            raise ValueError("leaf_pitch smaller than the leaf length")
        leaves = alloc((rows, pitch), field.dtype) if want_leaves else None
        digests = None if keep_on_device else alloc((max(2 * (rows - (1 << cap_local)), 0), d), field.dtype)
        cap = np.empty((1 << cap_local, d), dtype=field.dtype)
        coeffs = None
        if input_kind == INPUT_VALUES and not keep_on_device:
            coeffs = list(alloc((len(cols), degree), field.dtype))  # one block: adjacent columns come back in one copy

        desc = _lib.CommitDesc()
        desc.field, desc.hash, desc.repr, desc.input_kind = field.field_id, field.hash_id, repr, input_kind
        desc.log_n, desc.num_cols, desc.rate_bits, desc.cap_height = degree_log, len(cols), rate_bits, cap_height
        desc.num_salt_cols = len(salt_cols)
        desc.flags = 0 if want_leaves else _lib.FLAG_NO_LEAVES_OUT
        handle = ctypes.c_void_p()
        if keep_on_device:
            desc.flags = _lib.FLAG_KEEP_ON_DEVICE
            desc.batch_out = ctypes.pointer(handle)
        col_ptrs = _ptr_array(cols)
        desc.col_ptrs = ctypes.cast(col_ptrs, ctypes.POINTER(ctypes.c_void_p))
        if salt_cols:
            salt_ptrs = _ptr_array(salt_cols)
            desc.salt_ptrs = ctypes.cast(salt_ptrs, ctypes.POINTER(ctypes.c_void_p))
        if coeffs is not None:
            out_ptrs = _ptr_array(coeffs)
            desc.coeffs_out = ctypes.cast(out_ptrs, ctypes.POINTER(ctypes.c_void_p))
        desc.leaves_out = leaves.ctypes.data if leaves is not None else None
        desc.leaf_pitch = pitch
        desc.digests_out = digests.ctypes.data if digests is not None and digests.size else None
        desc.cap_out = cap.ctypes.data
        desc.shard_index, desc.shard_count = shard_index, shard_count
        t0 = time.perf_counter()
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(desc)))
        wall = time.perf_counter() - t0
        if timing is not None:
            for name, s in zip(STAGES, desc.stage_seconds):
                if name == "IFFT" and input_kind != INPUT_VALUES:
                    continue
                timing.record(name, s)
        polys = coeffs if coeffs is not None else cols
        if keep_on_device:
            tree = DeviceMerkleTree(field, ctx, handle, MerkleCap(cap))
            polys = None if input_kind == INPUT_VALUES else cols  # coefficients stay on the device (download on demand)
        else:
            tree = MerkleTree(field, leaves[:, :width] if leaves is not None and pitch != width else leaves, digests, MerkleCap(cap))
        batch = cls(field, polys, tree, degree_log, rate_bits, blinding)
        batch.desc = desc
        batch.wall_seconds = wall
        return batch

    def download_polynomials(self) -> np.ndarray:
        """coefficients [num_cols][n] of a device-resident batch"""
        tree = self.merkle_tree
        n = 1 << self.degree_log
        cols = tree.leaf_len - (SALT_SIZE if self.blinding else 0)
        out = np.empty((cols, n), dtype=self.field.dtype)
        tree.ctx.check(tree.ctx.lib.p2g_batch_download(tree.ctx.handle, tree.handle, 3, out.ctypes.data))
        return out

    # ------------------------------------------------------------------------------------------
    def get_lde_values(self, index: int, step: int) -> np.ndarray:
        """oracle.rs:153-158"""
        index = reverse_bits(index * step, self.degree_log + self.rate_bits)
        row = self.merkle_tree.get(index)
        return row[: len(row) - (SALT_SIZE if self.blinding else 0)]


def assemble_sharded(field: Field, shards: Sequence[PolynomialBatch], cap_height: int, ctx=None,
                     repr: int = REPR_CANONICAL) -> PolynomialBatch:
    """Join the per-GPU results of a coset-sharded commit (shard_count = len(shards)) into one batch with the
    reference's global `digests` / `cap` layout.  Only shard roots cross GPUs."""
    g = len(shards)
    lg = log2_strict(g)
    d = field.num_hash_out_elts
    first = shards[0]
    rows = first.merkle_tree.leaves.shape[0]
    n_total = rows * g
    leaves = np.concatenate([s.merkle_tree.leaves for s in shards], axis=0)
    if cap_height >= lg:
        digests = np.concatenate([s.merkle_tree.digests for s in shards], axis=0)
        cap = np.concatenate([s.merkle_tree.cap.hashes for s in shards], axis=0)
    else:
        ctx = ctx or _lib.default_context()
        roots = np.ascontiguousarray(np.concatenate([s.merkle_tree.cap.hashes for s in shards], axis=0))
        num_cap = 1 << cap_height
        top = np.empty((2 * (g - num_cap), d), dtype=field.dtype)
        cap = np.empty((num_cap, d), dtype=field.dtype)
        ctx.check(ctx.lib.p2g_merkle_finish_cap(ctx.handle, field.field_id, field.hash_id, repr, roots.ctypes.data,
                                                g, cap_height, top.ctypes.data, cap.ctypes.data))
        digests = np.empty((2 * (n_total - num_cap), d), dtype=field.dtype)
        shards_per_sub = g >> cap_height
        sub_len = digests.shape[0] >> cap_height
        top_sub_len = top.shape[0] >> cap_height
        lw = log2_strict(rows)  # log2(leaves per shard)
        for s in range(num_cap):
            block = digests[s * sub_len:(s + 1) * sub_len]
            for j in range(shards_per_sub):
                sh = shards[s * shards_per_sub + j].merkle_tree.digests
                block[2 * j * rows: 2 * j * rows + sh.shape[0]] = sh
            tblock = top[s * top_sub_len:(s + 1) * top_sub_len]
            for layer in range(log2_strict(shards_per_sub)):
                for j in range(shards_per_sub >> layer):
                    block[node_slot(layer + lw, j)] = tblock[node_slot(layer, j)]
    tree = MerkleTree(field, leaves, digests, MerkleCap(cap))
    return PolynomialBatch(field, first.polynomials, tree, first.degree_log, first.rate_bits, first.blinding)
