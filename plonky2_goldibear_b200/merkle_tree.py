"""Host mirror of `MerkleTree` (plonky2/src/hash/merkle_tree.rs:45-223): same fields, same digest layout, same
`prove` index arithmetic; the tree is built by the CUDA library (p2g_merkle)."""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .fields import Field, REPR_CANONICAL


def log2_strict(n: int) -> int:
    """util/src/lib.rs:30-33: panics unless n is a power of two."""
    if n <= 0 or n & (n - 1):
        raise ValueError("Not a power of two: %d" % n)
    return n.bit_length() - 1


@dataclass
class MerkleCap:
    """merkle_tree.rs:19: the 2^cap_height digests of the cap layer, shape [2^h][NUM_HASH_OUT_ELTS]."""
    hashes: np.ndarray

    def __len__(self):
        return self.hashes.shape[0]

    def height(self) -> int:
        return log2_strict(len(self))

    def flatten(self) -> np.ndarray:
        return self.hashes.reshape(-1)


class MerkleTree:
    def __init__(self, field: Field, leaves: np.ndarray, digests: np.ndarray, cap: MerkleCap):
        self.field = field
        self.leaves = leaves    # [N][leaf_len]
        self.digests = digests  # [2*(N - 2^h)][NUM_HASH_OUT_ELTS], recursive sibling-adjacent layout (:50-58)
        self.cap = cap

    @classmethod
    def new(cls, field: Field, leaves, cap_height: int, ctx=None, repr: int = REPR_CANONICAL) -> "MerkleTree":
        """MerkleTree::new (merkle_tree.rs:152-181). `leaves`: 2-D array [N][leaf_len] (or a list of equal-length rows)."""
        ctx = ctx or _lib.default_context()
        leaves = np.ascontiguousarray(np.asarray(leaves, dtype=field.dtype))
        if leaves.ndim != 2:
            raise ValueError("leaves must be a 2-D array")
        n = leaves.shape[0]
        log_n = log2_strict(n)
        if cap_height > log_n:
            # assert! at merkle_tree.rs:153-157
            raise ValueError("cap_height=%d should be at most log2(leaves.len())=%d" % (cap_height, log_n))
        d = field.num_hash_out_elts
        num_cap = 1 << cap_height
        digests = np.empty((2 * (n - num_cap), d), dtype=field.dtype)
        cap = np.empty((num_cap, d), dtype=field.dtype)
        desc = _lib.MerkleDesc()
        desc.field, desc.hash, desc.repr = field.field_id, field.hash_id, repr
        desc.num_leaves, desc.leaf_len, desc.leaf_pitch = n, leaves.shape[1], leaves.shape[1]
        desc.cap_height = cap_height
        desc.leaves = leaves.ctypes.data if leaves.size else None
        desc.digests_out = digests.ctypes.data if digests.size else None
        desc.cap_out = cap.ctypes.data
        ctx.check(ctx.lib.p2g_merkle(ctx.handle, ctypes.byref(desc)))
        tree = cls(field, leaves, digests, MerkleCap(cap))
        tree.seconds = desc.seconds
        return tree

    def get(self, i: int) -> np.ndarray:
        return self.leaves[i]

    def prove(self, leaf_index: int) -> np.ndarray:
        """MerkleTree::prove (merkle_tree.rs:188-222): sibling digests from the leaf up to the cap."""
        cap_height = log2_strict(len(self.cap))
        num_layers = log2_strict(self.leaves.shape[0]) - cap_height
        assert leaf_index >> (cap_height + num_layers) == 0
        tree_index = leaf_index >> num_layers
        tree_len = self.digests.shape[0] >> cap_height
        digest_tree = self.digests[tree_len * tree_index: tree_len * (tree_index + 1)]
        pair_index = leaf_index & ((1 << num_layers) - 1)
        siblings = []
        for i in range(num_layers):
            parity = pair_index & 1
            pair_index >>= 1
            siblings_index = (pair_index << (i + 1)) + (1 << i) - 1
            siblings.append(digest_tree[2 * siblings_index + (1 - parity)])
        return np.array(siblings, dtype=self.field.dtype).reshape(num_layers, self.field.num_hash_out_elts)


class DeviceMerkleTree:
    """The same tree with leaves and digests left in GPU memory (P2G_FLAG_KEEP_ON_DEVICE): `get` / `prove` are served
    by gather kernels (p2g_batch_get_rows / p2g_batch_prove), only the cap lives on the host.  SURVEY.md 8f, N2."""

    def __init__(self, field: Field, ctx, handle, cap: MerkleCap):
        self.field, self.ctx, self.handle, self.cap = field, ctx, handle, cap
        rows, width, layers, ncap = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint64()
        ctx.check(ctx.lib.p2g_batch_shape(handle, ctypes.byref(rows), ctypes.byref(width), ctypes.byref(layers), ctypes.byref(ncap)))
        self.num_leaves, self.leaf_len, self.num_layers = rows.value, width.value, layers.value

    def get_many(self, indices) -> np.ndarray:
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        out = np.empty((idx.size, self.leaf_len), dtype=self.field.dtype)
        self.ctx.check(self.ctx.lib.p2g_batch_get_rows(self.ctx.handle, self.handle, idx.ctypes.data, idx.size, out.ctypes.data))
        return out

    def get(self, i: int) -> np.ndarray:
        return self.get_many([i])[0]

    def prove_many(self, indices) -> np.ndarray:
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        out = np.empty((idx.size, self.num_layers, self.field.num_hash_out_elts), dtype=self.field.dtype)
        self.ctx.check(self.ctx.lib.p2g_batch_prove(self.ctx.handle, self.handle, idx.ctypes.data, idx.size,
                                                    out.ctypes.data if out.size else None))
        return out

    def prove(self, leaf_index: int) -> np.ndarray:
        return self.prove_many([leaf_index])[0]

    def download(self, what: str) -> np.ndarray:
        """whole array back to the host: 'leaves' | 'digests' | 'cap'"""
        d = self.field.num_hash_out_elts
        shape = {"leaves": (self.num_leaves, self.leaf_len), "digests": (2 * (self.num_leaves - len(self.cap)), d),
                 "cap": (len(self.cap), d)}[what]
        out = np.empty(shape, dtype=self.field.dtype)
        self.ctx.check(self.ctx.lib.p2g_batch_download(self.ctx.handle, self.handle, {"leaves": 0, "digests": 1, "cap": 2}[what],
                                                       out.ctypes.data))
        return out

    def free(self):
        if self.handle:
            self.ctx.lib.p2g_batch_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def node_slot(layer: int, j: int) -> int:
    """Index, inside one cap subtree's digest block, of node j of `layer` (0 = leaf digests)."""
    return ((j >> 1) << (layer + 2)) + (2 << layer) - 2 + (j & 1)
