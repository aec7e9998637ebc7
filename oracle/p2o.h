/*
 * p2o.h -- CPU ORACLE for the plonky2_goldibear polynomial-commitment hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py may load it.  The product library
 * (plonky2_goldibear_b200/libp2g_cuda.so) never includes, links or calls anything in oracle/.
 *
 * It is a restatement, function by function, of the reference CPU path (paths relative to the
 * reference tree):
 *   plonky2/src/fri/oracle.rs:68-150            PolynomialBatch::from_values / from_coeffs / lde_values
 *   field/src/fft.rs:12-205                      fft_root_table, fft_classic, ifft_with_options
 *   field/src/polynomial/mod.rs:201-220,282-295  lde (zero pad), coset_fft_with_options
 *   plonky2/src/util/mod.rs:25-41                transpose, reverse_bits
 *   util/src/lib.rs:57-66                        reverse_index_bits
 *   plonky2/src/hash/merkle_tree.rs:86-222       fill_subtree / fill_digests_buf / MerkleTree::new / prove
 *   plonky2/src/hash/merkle_proofs.rs:55-79      verify_merkle_proof_to_cap
 *   plonky2/src/hash/hashing.rs:76-133           compress, hash_n_to_m_no_pad (overwrite-mode sponge)
 *   plonky2/src/plonk/config.rs:70-84            hash_or_noop
 *   plonky2/src/hash/poseidon_goldilocks.rs      Poseidon width 12 (naive :927-948, fast :889-922)
 *   plonky2/src/gates/poseidon2_babybear.rs:609-680,787-832,903-917  Poseidon2 width 16 (in-tree restatement
 *                                                of the external p3_poseidon2 permutation)
 *   plonky2/src/fri/prover.rs:83-188             fri_committed_trees, fri_proof_of_work
 *   plonky2/src/iop/challenger.rs                 (the transcript itself is host logic: plonky2_goldibear_b200/challenger.py)
 *
 * The reference is Rust and its field / Poseidon2 arithmetic lives in an un-vendored git dependency
 * (telosnetwork/Plonky3, branch goldilocks_improvements, no lockfile); there is no Rust toolchain in the
 * build image, so the reference itself cannot be compiled (no oracle/_ref).  PARITY PINNING:
 *   pinned  : Poseidon-GL permutation (4 literal KATs), bit reversal (literal 256 table),
 *             GL sponge + 2-to-1 compression + Merkle path/cap convention (196 Merkle paths extracted from the
 *             reference's golden proof bytes, recursion/regression_test_data.rs:93, tests/golden/),
 *             GL generator 7, the GL two-adic root chain up to order 2^17, the coset shift and the bit-reversed leaf
 *             order of from_values: the golden proof's Fiat-Shamir transcript and all 28 FRI query rounds re-verify
 *             with this oracle's permutation and constants (tools/gl_fri_check.py, tests/test_golden_fri.py), and
 *             fail with any other shift / root / leaf order.
 *             FRI commit phase (bit reversal + chunking, fold order, shift^arity, extension W = 7): the oracle's layers
 *             pass the same verifier fold checks that re-verify the golden proof (tests/test_fri_commit.py);
 *             grinding: on the golden transcript the smallest witness is valid and <= the recorded pow_witness.
 *   unpinned: Poseidon2-BabyBear numeric output (no literal vector exists in the reference tree; only the
 *             structure is pinned by the in-tree restatement), BabyBear generator 31 and two-adic root
 *             0x1a427a41, Goldilocks two-adic roots of order above 2^17 (the 2^32 root 1753635133440165772 is the
 *             upstream Plonky3 value = 7^((p-1)/2^32); the golden proof fixes its 2^15-th power).
 *             => BabyBear parity is "parity unpinned"; Goldilocks parity is pinned.
 *
 * Element conventions: Goldilocks elements are uint64_t, BabyBear elements are uint32_t; both canonical
 * (< p) on input and output.  A digest is 32 bytes: 4 GL elements or 8 BB elements.
 */
#ifndef P2O_H
#define P2O_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { P2O_FIELD_GL = 0, P2O_FIELD_BB = 1 };
enum { P2O_INPUT_VALUES = 0, P2O_INPUT_COEFFS = 1 };

/* field constants */
uint64_t p2o_order(int field);
uint64_t p2o_generator(int field);                     /* F::generator(): 7 (GL), 31 (BB) */
uint64_t p2o_two_adic_generator(int field, unsigned bits);
uint64_t p2o_mul(int field, uint64_t a, uint64_t b);
uint64_t p2o_pow(int field, uint64_t a, uint64_t e);
uint64_t p2o_inv(int field, uint64_t a);

/* permutations (canonical in/out) */
void p2o_poseidon_gl(uint64_t state[12]);        /* fast form, poseidon_goldilocks.rs:912-922 */
void p2o_poseidon_gl_naive(uint64_t state[12]);  /* naive form, poseidon_goldilocks.rs:927-948 */
void p2o_poseidon2_bb(uint32_t state[16]);
/* witness of one Poseidon2BabyBearGate operation, wire order of gates/poseidon2_babybear.rs:79-147 (num_ops = 1) */
void p2o_poseidon2_bb_gate_wires(const uint32_t in[16], uint32_t swap, uint32_t wires[166]);

/* transforms, in place on one column of 2^log_n elements */
void p2o_fft(int field, void *data, unsigned log_n, unsigned zero_factor); /* fft_with_options(.., Some(zero_factor), None) */
void p2o_ifft(int field, void *data, unsigned log_n);                      /* ifft */
void p2o_coset_fft(int field, void *data, unsigned log_n, uint64_t shift, unsigned zero_factor);
void p2o_coset_ifft(int field, void *data, unsigned log_n, uint64_t shift); /* polynomial/mod.rs:62-72 */
void p2o_reverse_index_bits(void *data, size_t elem_bytes, unsigned log_n);
size_t p2o_reverse_bits(size_t x, unsigned num_bits);

/* hashing */
void p2o_hash_or_noop(int field, const void *elems, size_t len, void *digest32);
void p2o_hash_no_pad(int field, const void *elems, size_t len, void *digest32);
void p2o_two_to_one(int field, const void *left32, const void *right32, void *out32);

/* MerkleTree::new on row-major leaves[num_leaves][leaf_len] (contiguous).
 * digests_out: 2*(num_leaves - 2^cap_height) digests in the reference's recursive layout; cap_out: 2^cap_height.
 * returns 0, or -1 when cap_height > log2(num_leaves) / num_leaves not a power of two (the reference panics). */
int p2o_merkle_tree(int field, const void *leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    void *digests_out, void *cap_out, int threads);
/* MerkleTree::prove: writes log2(num_leaves)-cap_height sibling digests */
void p2o_merkle_prove(const void *digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      void *siblings_out);
/* verify_merkle_proof_to_cap: 1 = ok */
int p2o_verify_merkle_proof_to_cap(int field, const void *leaf, size_t leaf_len, size_t leaf_index,
                                   const void *cap, unsigned cap_height, const void *siblings,
                                   unsigned num_siblings);

typedef struct {
    int field;          /* P2O_FIELD_* */
    int input_kind;     /* P2O_INPUT_* */
    unsigned log_n;     /* degree_log */
    unsigned num_cols;  /* polynomials in the batch */
    unsigned rate_bits;
    unsigned cap_height;
    unsigned num_salt_cols; /* 0, or SALT_SIZE=4 when blinding; salts supplied by the caller */
    int threads;        /* OpenMP threads (<=0: all) */
    const void *input;  /* column-major [num_cols][n] */
    const void *salts;  /* column-major [num_salt_cols][N], may be NULL */
    void *coeffs_out;   /* column-major [num_cols][n]; may be NULL */
    void *leaves_out;   /* row-major [N][num_cols+num_salt_cols]; may be NULL */
    void *digests_out;  /* 2*(N-2^cap_height) digests; may be NULL */
    void *cap_out;      /* 2^cap_height digests; may be NULL */
    double stage_seconds[4]; /* "IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree" */
} p2o_commit_desc;

/* PolynomialBatch::from_values / from_coeffs. returns 0 on success, -1 on a condition the reference panics on. */
int p2o_commit(p2o_commit_desc *desc);

/* FRI commit phase (plonky2/src/fri/prover.rs:83-133).  An extension element is D base elements in a row. */
unsigned p2o_ext_degree(int field);  /* 2 (GL), 4 (BB) */
uint64_t p2o_ext_w(int field);       /* X^D = W: 7 (GL, pinned by the golden FRI proof), 11 (BB, unpinned) */
void p2o_ext_mul(int field, const void *a, const void *b, void *out);
/* leaves of one layer tree: [n >> arity_bits][arity * D] = the bit-reversed values (:103-109) */
void p2o_fri_layer_leaves(int field, const void *values, unsigned log_n, void *leaves_out);
/* after the challenge beta (:114-125): folded coefficients [n >> arity_bits][D], shift^arity, and the folded
 * polynomial's evaluations on the new coset, natural order */
void p2o_fri_fold_layer(int field, const void *coeffs, unsigned log_n, unsigned arity_bits, const void *beta,
                        uint64_t shift_in, void *coeffs_out, void *values_out, uint64_t *shift_out);

/* fri_proof_of_work (fri/prover.rs:136-188): smallest witness whose response has >= min_leading_zeros leading zeros.
 * state: WIDTH canonical elements, input: the challenger's input_len (< 8) buffered elements.  0 = found. */
int p2o_fri_pow(int field, const void *state, const void *input, unsigned input_len, unsigned min_leading_zeros,
                uint64_t *witness);

int p2o_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
