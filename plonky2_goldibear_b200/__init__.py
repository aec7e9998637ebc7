"""B200-native commit path of telosnetwork/plonky2_goldibear: `PolynomialBatch::from_values` / `from_coeffs`
(plonky2/src/fri/oracle.rs:68-123) and `MerkleTree::new` (plonky2/src/hash/merkle_tree.rs:152-181) as hand-written
CUDA for sm_100a behind a C ABI (include/p2g.h).  This package is the host-side mirror of that interface used by
the tests and the benchmark; there is no CPU fallback -- importing works anywhere, every call needs the built
library and a GPU and fails loudly otherwise."""
from .fields import GOLDILOCKS, BABYBEAR, Field
from ._lib import Context, P2GError, library_path, load_library
from .merkle_tree import MerkleTree, MerkleCap
from .fri_oracle import PolynomialBatch, ProvingProcessInfo, SALT_SIZE
from . import polynomial
from .challenger import Challenger
from .fri_prover import (FriConfig, FriParams, fri_committed_trees, fri_proof, fri_proof_of_work,
                         fri_prover_query_rounds)

__all__ = [
    "GOLDILOCKS", "BABYBEAR", "Field", "Context", "P2GError", "library_path", "load_library",
    "MerkleTree", "MerkleCap", "PolynomialBatch", "ProvingProcessInfo", "SALT_SIZE",
    "polynomial", "Challenger", "FriConfig", "FriParams", "fri_committed_trees", "fri_proof", "fri_proof_of_work",
    "fri_prover_query_rounds",
]
