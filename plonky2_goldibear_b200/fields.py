"""Field descriptors for the two configurations the commit path serves:
Goldilocks + Poseidon (width 12) and BabyBear + Poseidon2 (width 16)."""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Field:
    name: str
    field_id: int  # P2G_FIELD_*
    hash_id: int   # P2G_HASH_*
    order: int
    dtype: type
    num_hash_out_elts: int  # HashOut elements: 4 (GL) / 8 (BB), hash/hash_types.rs:30,66
    two_adicity: int
    generator: int
    extension_degree: int = 2   # D of the binomial extension F[X]/(X^D - W): field/src/lib.rs:14-16
    extension_w: int = 7
    width: int = 12             # permutation width (rate 8 for both hashers, hashing.rs:100-123)

    @property
    def elem_bytes(self) -> int:
        return np.dtype(self.dtype).itemsize


GOLDILOCKS = Field("goldilocks", 0, 0, 0xFFFFFFFF00000001, np.uint64, 4, 32, 7, 2, 7, 12)
BABYBEAR = Field("babybear", 1, 1, 2013265921, np.uint32, 8, 27, 31, 4, 11, 16)

REPR_CANONICAL = 0
REPR_MONTY_R32 = 1
INPUT_VALUES = 0
INPUT_COEFFS = 1
