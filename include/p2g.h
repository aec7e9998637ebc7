/*
 * p2g.h -- C ABI of the B200 polynomial-commitment library (libp2g_cuda.so).
 *
 * This is the drop-in boundary for ONE hot path of telosnetwork/plonky2_goldibear: the bodies of
 *   PolynomialBatch::from_values   plonky2/src/fri/oracle.rs:68-90
 *   PolynomialBatch::from_coeffs   plonky2/src/fri/oracle.rs:93-123
 *   MerkleTree::new                plonky2/src/hash/merkle_tree.rs:152-181
 * The reference has no FFI layer; under its (new) `cuda` cargo feature those three functions bind the entry
 * points below (INTEGRATION.md shows the Rust side).  Plain pointers and sizes only; no torch types.
 *
 * Everything here runs on the GPU: there is NO CPU fallback.  If no CUDA device is usable p2g_init fails
 * and every other call returns P2G_ERR_NO_DEVICE.
 *
 * Element conventions
 *   Goldilocks : uint64_t.  Any value accepted on input (treated mod p); outputs canonical (< p).
 *   BabyBear   : uint32_t.  `repr` selects canonical integers or p3-monty-31's in-memory Montgomery form
 *                (R = 2^32; 1 <-> 268435454); outputs use the same representation as the inputs.
 *   Digest     : 32 bytes = 4 Goldilocks or 8 BabyBear elements (HashOut, hash/hash_types.rs:105-108), stored as
 *                elements in the field's representation.
 * Layouts
 *   columns    : column-major, one contiguous array of n = 2^log_n elements per polynomial (mirrors
 *                Vec<PolynomialValues<F>> / Vec<PolynomialCoeffs<F>>).
 *   leaves     : row-major [N][leaf_pitch] with N = n << rate_bits rows; row k holds the evaluations at
 *                g * w_N^{bitrev_N(k)} of every polynomial, then the salt columns (oracle.rs:108-109,153-158).
 *   digests    : 2*(N - 2^cap_height) digests in the reference's recursive sibling-adjacent layout
 *                (merkle_tree.rs:50-58); cap: 2^cap_height digests.
 * Errors: calls return 0 or a negative P2G_ERR_*; p2g_last_error gives the message.  Conditions on which
 * the reference panics (inconsistent shapes, cap_height > log2(leaves), size not a power of two) are
 * P2G_ERR_INVALID -- the Rust shim turns them back into panics.
 * Threading: a context serialises its calls internally (one stream, one workspace); use one context per
 * concurrent prover thread.
 */
#ifndef P2G_H
#define P2G_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2G_ABI_VERSION 1

enum { P2G_FIELD_GOLDILOCKS = 0, P2G_FIELD_BABYBEAR = 1 };
enum { P2G_HASH_POSEIDON_GL12 = 0, P2G_HASH_POSEIDON2_BB16 = 1 };
enum { P2G_REPR_CANONICAL = 0, P2G_REPR_MONTY_R32 = 1 };
enum { P2G_INPUT_VALUES = 0, P2G_INPUT_COEFFS = 1 };

enum {
    P2G_OK = 0,
    P2G_ERR_INVALID = -1,    /* a shape/argument the reference would panic on */
    P2G_ERR_NO_DEVICE = -2,  /* no usable CUDA device / context creation failed */
    P2G_ERR_CUDA = -3,       /* a CUDA runtime call failed */
    P2G_ERR_OOM = -4,        /* device or pinned-host allocation failed */
    P2G_ERR_UNSUPPORTED = -5 /* field/hash combination or size outside what the kernels implement */
};

typedef struct p2g_ctx p2g_ctx;

/* Context.
 * n_devices == 1: one CUDA device, its streams, resident twiddle / round-constant tables, a reusable workspace.
 * n_devices in {2, 4, 8}: a MULTI-DEVICE context for one process (a rayon-style prover with a single address space,
 *   maybe_rayon/src/lib.rs:5-20): one single-device context per entry of device_ids, peer access enabled between every
 *   pair, and a gather buffer per device.  p2g_commit on it runs the whole job: device d computes the inverse NTT of
 *   its column slice and stores the coefficients into every device's gather buffer from the transform's epilogue
 *   (peer stores over NVLink / NVSwitch), an event per device is the cross-device barrier, then device d extends and
 *   hashes LDE cosets [d * 2^rate_bits / n, ...) and writes its leaf rows, digest slice and cap entries straight into
 *   the caller's buffers over its own PCIe link.  Requires 2^rate_bits >= n_devices and
 *   cap_height >= log2(n_devices).  Flags accepted: NO_LEAVES_OUT, KEEP_ON_DEVICE (the batch handle then spans the
 *   devices; get_rows / prove / download route by leaf index) and DEVICE_INPUT (col_ptrs[d] = device pointer ON DEVICE d
 *   to that device's column slice, columns [d * ceil(c / n), ...) x n values).  Every other entry point of a
 *   multi-device context (p2g_merkle, p2g_transform, p2g_permute, p2g_fri_*) is served by its first device.
 *   The same device id may be listed more than once: its shards then share that GPU (how the partition is tested on a
 *   one-GPU box).
 * Multi-PROCESS sharding (one context per GPU and process, exchange by CUDA IPC) stays available through
 * p2g_commit_desc.shard_* and p2g_intt_allgather. */
int32_t p2g_init(const int32_t *device_ids, int32_t n_devices, p2g_ctx **out_ctx);
int32_t p2g_num_devices(p2g_ctx *ctx);
p2g_ctx *p2g_sub_context(p2g_ctx *ctx, int32_t index); /* the single-device context of device_ids[index] (ctx itself when n = 1) */
void p2g_destroy(p2g_ctx *ctx);
const char *p2g_last_error(p2g_ctx *ctx); /* ctx may be NULL: last error of p2g_init on this thread */
int32_t p2g_abi_version(void);

/* Pinned host memory for full-speed PCIe transfers (the Rust shim can back its Vecs with it). */
void *p2g_host_alloc(p2g_ctx *ctx, size_t bytes);
void p2g_host_free(p2g_ctx *ctx, void *ptr);

#define P2G_FLAG_DEVICE_INPUT 1u   /* col_ptrs[0] is a DEVICE pointer to the whole column-major [num_cols][n] block */
#define P2G_FLAG_DEVICE_OUTPUT 2u  /* *_out pointers are DEVICE pointers (coeffs_out[0] = whole block) */
#define P2G_FLAG_NO_LEAVES_OUT 4u  /* do not return leaves (benchmarks / callers that only need the cap) */
#define P2G_FLAG_KEEP_ON_DEVICE 8u /* leaves, digests and coefficients stay in device memory owned by a batch handle
                                      (desc->batch_out); only the cap comes back.  Rows, Merkle paths and whole arrays
                                      are then fetched on demand (p2g_batch_*): the device-resident form of
                                      get_lde_values / MerkleTree::get / MerkleTree::prove */

/* A commit split over column ranges (multi-GPU pipelines: the LDE of the columns that have arrived runs while the
 * exchange of the next ones is still in flight).  Both need P2G_FLAG_DEVICE_OUTPUT and a device leaves_out. */
#define P2G_FLAG_LEAVES_ONLY 16u  /* stop after "transpose LDEs": the num_cols columns given are extended into columns
                                     [0, num_cols) of the rows at leaves_out (leaf_pitch >= num_cols; point leaves_out at
                                     the first column of the range inside a wider leaf matrix); nothing is hashed */
#define P2G_FLAG_LEAVES_READY 32u /* "build Merkle tree" only: leaves_out already holds this shard's rows, num_cols
                                     (+ salt) wide at leaf_pitch; col_ptrs is ignored */

typedef struct p2g_batch p2g_batch;

/* One call = one PolynomialBatch::from_values (input_kind VALUES) or from_coeffs (COEFFS). */
typedef struct {
    int32_t field;       /* P2G_FIELD_* */
    int32_t hash;        /* P2G_HASH_*; must match the field */
    int32_t repr;        /* P2G_REPR_* (BabyBear only; Goldilocks ignores it) */
    int32_t input_kind;  /* P2G_INPUT_* */
    uint32_t log_n;      /* degree_log: every column has n = 2^log_n elements */
    uint32_t num_cols;   /* polynomials in the batch (> 0) */
    uint32_t rate_bits;
    uint32_t cap_height;
    uint32_t num_salt_cols; /* 0, or SALT_SIZE = 4 when `blinding` (oracle.rs:25,134); salts come from the caller
                               so the RNG stays on the Rust side */
    uint32_t flags;         /* P2G_FLAG_* */
    const void *const *col_ptrs;  /* [num_cols] host pointers, n elements each */
    const void *const *salt_ptrs; /* [num_salt_cols] pointers, N elements each, natural (un-bit-reversed) LDE order
                                     exactly like the reference's extra lde_values columns.  HOST or DEVICE memory,
                                     independently of P2G_FLAG_DEVICE_INPUT (which describes col_ptrs only): the
                                     upload resolves its direction from each pointer */
    void *const *coeffs_out; /* [num_cols] host pointers (n each) receiving `polynomials`; may alias col_ptrs;
                                NULL to skip (from_coeffs callers already own them) */
    void *leaves_out;        /* [N][leaf_pitch] row-major; NULL with P2G_FLAG_NO_LEAVES_OUT */
    uint64_t leaf_pitch;     /* elements between rows of leaves_out; 0 = num_cols + num_salt_cols */
    void *digests_out;       /* 2*(N - 2^cap_height) * 32 bytes; may be NULL */
    void *cap_out;           /* 2^cap_height * 32 bytes */
    /* multi-GPU shard (0,0 = whole job).  This context computes LDE cosets [shard_index*2^rate_bits/shard_count,
     * ...) only: leaves rows [shard_index*N/shard_count, ...), the matching slice of `digests` and, when
     * 2^cap_height >= shard_count, the matching cap entries; *_out then point at those slices.  When
     * 2^cap_height < shard_count, cap_out receives the shard's single subtree root and digests_out its
     * local digests; p2g_merkle_finish_cap combines the gathered roots. */
    uint32_t shard_index, shard_count;
    /* device seconds of the reference's four timed! scopes (oracle.rs:76-114):
     * "IFFT", "FFT + blinding", "transpose LDEs" (fused into the previous one here: always 0), "build Merkle tree" */
    double stage_seconds[4];
    double h2d_seconds, d2h_seconds;
    /* finer split of the same device time: [0] inverse-NTT passes, [1] LDE NTT passes (+ salt scatter),
     * [2] leaf sponge kernel, [3] tree-level kernels */
    double kernel_seconds[4];
    p2g_batch **batch_out;   /* with P2G_FLAG_KEEP_ON_DEVICE: receives the handle (free with p2g_batch_free) */
} p2g_commit_desc;

int32_t p2g_commit(p2g_ctx *ctx, p2g_commit_desc *desc);

/* Device-resident batch accessors.  Indices are leaf indices of the (shard-local) tree, i.e. already bit-reversed
 * like MerkleTree::get / prove expect (oracle.rs:153-158 does the reverse_bits before indexing).
 *   get_rows : rows_out[i][0..width) = leaves[row_indices[i]]                    (MerkleTree::get, merkle_tree.rs:183-185)
 *   prove    : siblings_out[i][layer] = sibling digest of leaf_indices[i]        (MerkleTree::prove, merkle_tree.rs:188-222)
 *   download : the whole array: what = 0 leaves [rows][width], 1 digests, 2 cap, 3 coefficients [num_cols][n] */
int32_t p2g_batch_get_rows(p2g_ctx *ctx, p2g_batch *batch, const uint64_t *row_indices, uint64_t count, void *rows_out);
int32_t p2g_batch_prove(p2g_ctx *ctx, p2g_batch *batch, const uint64_t *leaf_indices, uint64_t count, void *siblings_out);
int32_t p2g_batch_download(p2g_ctx *ctx, p2g_batch *batch, int32_t what, void *out);
/* shape of the batch: rows, width (elements per leaf), num_layers (siblings per proof), cap entries */
int32_t p2g_batch_shape(p2g_batch *batch, uint64_t *rows, uint64_t *width, uint32_t *num_layers, uint64_t *num_cap);
void p2g_batch_free(p2g_ctx *ctx, p2g_batch *batch);

/* MerkleTree::new alone (merkle_tree.rs:152-181), e.g. for the FRI commit-phase trees (fri/prover.rs:108). */
typedef struct {
    int32_t field, hash, repr;
    uint32_t flags;        /* P2G_FLAG_DEVICE_INPUT / P2G_FLAG_DEVICE_OUTPUT */
    uint64_t num_leaves;   /* power of two */
    uint64_t leaf_len;     /* elements per leaf; <= 4 (GL) / 8 (BB) takes the no-hash branch of hash_or_noop */
    uint64_t leaf_pitch;   /* elements between leaves; 0 = leaf_len */
    uint32_t cap_height;
    const void *leaves;    /* [num_leaves][leaf_pitch] */
    void *digests_out;     /* 2*(num_leaves - 2^cap_height) digests */
    void *cap_out;         /* 2^cap_height digests */
    double seconds;
} p2g_merkle_desc;

int32_t p2g_merkle(p2g_ctx *ctx, p2g_merkle_desc *desc);

/* Finish the top of a sharded tree (multi-GPU job with 2^cap_height < shard_count): `roots` holds the
 * shard_count subtree roots (32 bytes each, host memory, gathered over NCCL / P2P).  They are the leaves of a
 * tiny tree of digests; the call returns that tree's 2*(shard_count - 2^cap_height) inner digests in the reference
 * layout (top_digests_out) and the 2^cap_height cap entries.  Node j of tiny layer t is node j of layer
 * t + log2(leaves per shard) of the full tree, which is where the caller files it in the full `digests` array.
 * Runs on the context's GPU. */
int32_t p2g_merkle_finish_cap(p2g_ctx *ctx, int32_t field, int32_t hash, int32_t repr, const void *roots,
                              uint32_t shard_count, uint32_t cap_height, void *top_digests_out, void *cap_out);

/* Single permutations / hashes on the GPU (known-answer tests, fixtures). Host pointers.
 * p2g_permute: `count` states of 12 x u64 (GL) or 16 x u32 (BB), permuted in place. */
int32_t p2g_permute(p2g_ctx *ctx, int32_t hash, int32_t repr, void *states, uint64_t count);

/* Batched transforms on column-major data (parity tests for the NTT kernels; with P2G_FLAG_DEVICE_INPUT/OUTPUT the
 * column-sharded inverse transform of a multi-GPU commit):
 * kind 0: ifft (values -> coeffs, field/src/fft.rs:66-94); kind 1: coset LDE of coeffs with shift g and rate_bits,
 * output column-major natural order [num_cols][N] like lde_values (oracle.rs:125-150); kind 2: coset_ifft with shift
 * F::generator() (values on g*H -> coeffs, field/src/polynomial/mod.rs:62-72), the transform that ends
 * compute_quotient_polys (plonk/prover.rs:921-925). */
int32_t p2g_transform(p2g_ctx *ctx, int32_t field, int32_t repr, int32_t kind, uint32_t log_n, uint32_t num_cols,
                      uint32_t rate_bits, uint32_t flags, const void *in, void *out);

/* ---- FRI commit phase: fri_committed_trees, plonky2/src/fri/prover.rs:83-133.
 * The Fiat-Shamir challenger stays with the caller; the library does the work between two challenges:
 *
 *   p2g_fri_begin(coeffs, values)            lde_polynomial_coeffs / lde_polynomial_values, 2^log_n extension
 *                                            elements each, an element being its `degree` base coefficients in a
 *                                            row (F::Extension::as_base_slice); shift starts at F::generator() (:100).
 *                                            values may be NULL: they are then computed on the device as
 *                                            coeffs.coset_fft(F::generator()), which is how prove_openings obtains
 *                                            them (fri/oracle.rs:226-231) -- no host FFT, no upload
 *   for arity_bits in reduction_arity_bits:
 *     p2g_fri_commit_layer(arity_bits, cap_height, cap_out, &tree)
 *                                            reverse_index_bits + chunks(arity) + flatten + MerkleTree::new (:103-110);
 *                                            cap_out (host, 2^cap_height digests) is what challenger.observe_cap needs;
 *                                            `tree` is a device-resident p2g_batch (rows = 2^(log_n - arity_bits)
 *                                            leaves of arity*degree elements) serving p2g_batch_get_rows / _prove /
 *                                            _download for the query phase (fri/prover.rs:219-255); free it with
 *                                            p2g_batch_free
 *     p2g_fri_fold(beta)                     beta = challenger.get_extension_challenge() (`degree` base elements, in
 *                                            `repr`): coeffs <- reduce_with_powers over chunks of arity (:116-123),
 *                                            shift <- shift^arity, values <- coeffs.coset_fft(shift) (:124-125)
 *   p2g_fri_finish(rate_bits, out, &count)   the final polynomial, truncated to len >> rate_bits (:129-131)
 *
 * Errors: the reference's panics map to P2G_ERR_INVALID (cap_height above log2 of the layer's leaf count,
 * merkle_tree.rs:153-157). Extension degrees: 2 for Goldilocks (X^2 = 7), 4 for BabyBear (X^4 = 11). */
typedef struct p2g_fri p2g_fri;
int32_t p2g_fri_begin(p2g_ctx *ctx, int32_t field, int32_t hash, int32_t repr, uint32_t degree, uint32_t log_n,
                      const void *coeffs, const void *values, uint32_t flags, p2g_fri **out);
int32_t p2g_fri_commit_layer(p2g_ctx *ctx, p2g_fri *fri, uint32_t arity_bits, uint32_t cap_height, void *cap_out,
                             p2g_batch **tree_out);
int32_t p2g_fri_fold(p2g_ctx *ctx, p2g_fri *fri, const void *beta);
int32_t p2g_fri_finish(p2g_ctx *ctx, p2g_fri *fri, uint32_t rate_bits, void *final_coeffs_out, uint64_t *count_out);
/* FRI grinding, fri_proof_of_work (fri/prover.rs:136-188): sponge_state = the challenger's
 * 12 (GL) / 16 (BB) state elements, input_buffer = its input_len (< 8) pending elements, all canonical, host memory.
 * Finds a witness w such that, with w observed next, the challenger's next challenge has at least min_leading_zeros
 * leading zero bits as a canonical u64 (min_leading_zeros = proof_of_work_bits + 64 - bits(order), :147).
 * The reference returns whichever candidate its thread pool hits first (find_any); this returns the smallest valid
 * candidate (batches of 2^16..2^22 candidates are searched in increasing order), which is reproducible.  The caller then does
 * challenger.observe_element(w) and re-checks the response as the reference does (:181-186). */
int32_t p2g_fri_pow(p2g_ctx *ctx, int32_t field, int32_t hash, const void *sponge_state, const void *input_buffer,
                    uint32_t input_len, uint32_t min_leading_zeros, uint64_t *witness_out);
/* current codeword length (log2) and coset shift (canonical) */
int32_t p2g_fri_shape(p2g_fri *fri, uint32_t *log_n, uint64_t *shift);
void p2g_fri_free(p2g_ctx *ctx, p2g_fri *fri);

/* ---- multi-GPU plumbing (one process per GPU).
 * p2g_dev_alloc / p2g_ipc_export / p2g_ipc_import: a device buffer that the other ranks map over NVLink (CUDA IPC).
 * p2g_intt_allgather: the column-sharded inverse NTT of `num_cols` columns (device pointer `values_dev`, column-major)
 * whose last pass writes the coefficients at column `col_offset` of EVERY gather buffer: gather_bufs[0] is this rank's
 * own [total_cols][n] buffer, gather_bufs[1..] the peers' (imported) ones.  The all-gather that every rank needs before
 * its coset LDE is thereby fused into the transform's epilogue (stores to peer memory) instead of a separate NCCL call.
 * Asynchronous on the context's stream; the caller puts a cross-GPU barrier between this call and the commit that
 * reads the gathered coefficients (bench.py: a one-element all-reduce on the same stream, double-buffered). */
void *p2g_dev_alloc(p2g_ctx *ctx, size_t bytes);
void p2g_dev_free(p2g_ctx *ctx, void *ptr);
int32_t p2g_ipc_export(p2g_ctx *ctx, void *dev_ptr, uint8_t handle_out[64]);
int32_t p2g_ipc_import(p2g_ctx *ctx, const uint8_t handle[64], void **dev_ptr_out);
int32_t p2g_ipc_close(p2g_ctx *ctx, void *dev_ptr);
/* copy ordered on the context's stream: kind 0 = host->device, 1 = device->host, 2 = device->device */
int32_t p2g_memcpy_async(p2g_ctx *ctx, void *dst, const void *src, size_t bytes, int32_t kind);
int32_t p2g_intt_allgather(p2g_ctx *ctx, int32_t field, int32_t repr, uint32_t log_n, uint32_t num_cols, uint32_t col_offset,
                           const void *values_dev, void *const *gather_bufs, uint32_t num_bufs);

/* The LDE of one shard split by NTT PASS, for multi-GPU pipelines in which the coefficient columns arrive block by
 * block (bench.py --exchange ring).  A commit split by column RANGE (P2G_FLAG_LEAVES_ONLY) makes every block write a narrow
 * strip of every leaf row, which costs the last pass its write locality (16-column blocks: +36..49 % LDE time on one GPU,
 * tools/lde_blocks_probe.py); here only the strided passes (column-major scratch inside the context) run per block and
 * the last pass writes whole leaf rows once:
 *   phase P2G_LDE_STRIDED: coeffs_dev = [num_cols][n] coefficients (device) of columns [col_base, col_base + num_cols)
 *                          of a batch of total_cols; leaves_dev unused.  No-op for sizes whose transform is one pass.
 *   phase P2G_LDE_LAST   : after every column had its STRIDED call: col_base = 0, num_cols = total_cols; writes the
 *                          shard's rows (all of its cosets) to leaves_dev at leaf_pitch.  coeffs_dev = all total_cols
 *                          columns (read by single-pass sizes only).
 * Asynchronous on the context's stream; P2G_FLAG_LEAVES_READY builds the tree afterwards.  The scratch is only valid
 * until the context's next transform, so no other call may come between the STRIDED calls and the LAST one. */
#define P2G_LDE_STRIDED 0
#define P2G_LDE_LAST 1
int32_t p2g_lde_columns(p2g_ctx *ctx, int32_t field, int32_t repr, uint32_t log_n, uint32_t rate_bits, uint32_t shard_index,
                        uint32_t shard_count, uint32_t total_cols, uint32_t col_base, uint32_t num_cols, const void *coeffs_dev,
                        void *leaves_dev, uint64_t leaf_pitch, int32_t phase);

/* Device workspace currently held by the context (bytes) and a way to drop it. */
uint64_t p2g_workspace_bytes(p2g_ctx *ctx);
void p2g_release_workspace(p2g_ctx *ctx);

/* Raw CUDA stream of the context (cudaStream_t) so callers that own device buffers can order work on it. */
void *p2g_stream(p2g_ctx *ctx);
/* Number of kernels the context has launched since creation (bench.py reports it as gpu_launches). */
uint64_t p2g_kernel_launches(p2g_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
