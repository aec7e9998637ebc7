// Internal interfaces between the host orchestration (p2g_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "field_bb.cuh"
#include "field_gl.cuh"
#include "hash_tables.h"

namespace p2g {

enum FieldId { FIELD_GL = 0, FIELD_BB = 1 };
// A CTA transforms at most 2^MAX_TILE_LOG points per column segment in shared memory.  Tiles of up to 2^WIDE_TILE_LOG
// points carry the field's full run width (64-byte runs: 8 Goldilocks / 16 BabyBear columns or offsets); a 2^11-point
// tile halves it (32-byte runs, one DRAM sector) to stay at 64 KB, which makes 2^21 and 2^22 two-pass transforms.
enum { MAX_TILE_LOG = 11, WIDE_TILE_LOG = 10 };
static inline __host__ __device__ unsigned tile_width_log(unsigned field_wlog, unsigned log_np) {
    return log_np > WIDE_TILE_LOG ? field_wlog - (log_np - WIDE_TILE_LOG) : field_wlog;
}
// The two final tile kinds keep their tile COLUMN-major (element (i, w) at x[w * cs + i]): the interleaved transforms of
// such a tile are contiguous runs of global memory, so each arrives as ONE bulk copy of the TMA unit (cp.async.bulk,
// ntt.cu) instead of np element-sized LDGSTS.  cs = np + pad: the pad spreads the 2^lw columns over the 32 banks (lanes
// take consecutive w, then consecutive i: conflict-free for 8-byte words with pad = 16 >> lw, for 4-byte words with
// 32 >> lw) and keeps every column 16-byte aligned, which the bulk copy requires (BabyBear at 16 columns: pad 4, 2-way).
static inline __host__ __device__ unsigned tile_col_stride(unsigned elem_bytes, unsigned log_np, unsigned lw) {
    unsigned pad = (128u / elem_bytes) >> lw;
    const unsigned min_pad = 16u / elem_bytes;
    if (pad < min_pad) pad = min_pad;
    return (1u << log_np) + pad;
}
enum { POW_DIGIT_LOG = 10, POW_DIGIT = 1 << POW_DIGIT_LOG, POW_LEVELS = 3 };

// ---- hash.cu -----------------------------------------------------------------------------------
// Host-side derivation of the Poseidon tables (poseidon_tables.cpp)
void build_poseidon_gl_tables(PoseidonGlTables *out);
void build_poseidon2_bb_tables(Poseidon2BbTables *out);
cudaError_t upload_hash_tables(const PoseidonGlTables *gl, const Poseidon2BbTables *bb);

struct TreeShape {
    // local tree: num_leaves = 2^log_leaves, 2^cap_height cap entries, subtrees of 2^(log_leaves-cap_height) leaves
    uint32_t log_leaves;
    uint32_t cap_height;
    void *digests;  // 2*(2^log_leaves - 2^cap_height) digests of 32 bytes
    void *cap;      // 2^cap_height digests
};

// leaf digests (hash_or_noop of each row) for rows [row_begin, row_begin + row_count) of the local tree
void launch_leaf_hash(int field, int repr, const void *leaves, uint64_t pitch_elems, uint32_t leaf_len,
                      uint64_t row_begin, uint64_t row_count, const TreeShape &tree, cudaStream_t stream,
                      uint64_t *launches);
// all inner levels (two_to_one) up to the cap
void launch_tree_levels(int field, int repr, const TreeShape &tree, cudaStream_t stream, uint64_t *launches);
// device-resident accessors: rows_out[i][:] = leaves[idx[i]][:width];  siblings_out[i][layer] = sibling digest
void launch_gather_rows(int field, const void *leaves, uint64_t pitch, uint32_t width, const uint64_t *idx, uint64_t count,
                        void *rows_out, cudaStream_t stream, uint64_t *launches);
void launch_gather_proofs(const void *digests, uint32_t log_leaves, uint32_t cap_height, const uint64_t *idx,
                          uint64_t count, void *siblings_out, cudaStream_t stream, uint64_t *launches);
// permutation of `count` states in place (device pointer)
void launch_permute(int field, int repr, void *states, uint64_t count, cudaStream_t stream, uint64_t *launches);

// FRI grinding: candidates [start, start + count) capped at last_candidate; *best (device u64, preset to ~0) receives the
// smallest candidate whose response has >= min_leading_zeros leading zeros.  `state` = 12/16 canonical elements (device)
void launch_pow_search(int field, const void *state, uint32_t pos, uint32_t min_leading_zeros, uint64_t start,
                       uint64_t count, uint64_t last_candidate, void *best, cudaStream_t stream, uint64_t *launches);

// ---- ntt.cu ------------------------------------------------------------------------------------
struct NttPass {
    const void *src;
    void *dst;
    uint64_t src_col_stride;  // elements
    uint64_t dst_col_stride;  // elements (column-major outputs) or row pitch (leaf output)
    uint32_t num_cols;
    uint32_t log_n;        // column length
    uint32_t log_min;      // chunk length at this level (log_n for the first pass)
    uint32_t log_np;       // points transformed per tile
    uint32_t lp1, lp2;     // tile sizes of the earlier passes (digit sizes of the chunk index), 0 if unused
    const void *roots;     // direction specific, internal representation: Goldilocks w_{np}^k for k < np (the limb
                           // joins look up w * 2^(24 i) = w_{np}^(k + i * step) in the same table), BabyBear k < np/2
    // One launch can carry several LDE cosets ("batch"): CTAs of the same tile and different cosets are adjacent in
    // the grid, so the coefficient tile they all read comes from DRAM once and from L2 after that.
    uint32_t num_batch;        // >= 1
    uint64_t src_batch_stride; // elements between the batch entries' inputs (0: all read the same input)
    uint64_t dst_batch_stride; // elements between their outputs (column-major outputs)
    uint64_t row_batch_stride; // leaf output: rows between their blocks
    const void *tw_tables[8];     // strided pass, per batch entry: [m_in] inter-pass factors, entry f*m_out + j0 =
                                  // w_{m_in}^(f*j0) (* s^j0 on the first LDE pass); indexed like the store
    const void *scale_tables[8];  // per batch entry: [np] coset factors s^(m_out * i) applied to tile row i on load, or null
    uint64_t final_scale;  // BabyBear: internal representation of 1/n, multiplied into the outputs of a final natural pass
    uint32_t use_final_scale;
    uint64_t join_w[4];    // Goldilocks final natural pass: n^-1 * 2^(24 i) mod p, the scale rides in the last join
    uint64_t row_base;     // leaf output: first row of batch entry 0
    // final natural pass of a column-sharded inverse transform: the coefficients are also written into the gather
    // buffers of the other GPUs (peer pointers mapped over NVLink), i.e. the all-gather is fused into the epilogue
    void *peer_dst[8];
    uint32_t num_peers;
    uint32_t async_ok;       // src base and column stride are 16-byte aligned: vector / asynchronous tile loads
    uint32_t in_user_repr;   // 1: src is in the caller's representation (load conversion), 0: internal
    uint32_t out_user_repr;  // 1: dst is in the caller's representation (canonical / chosen repr)
    uint32_t inverse;        // direction (selects the kernels' compile-time roots of unity)
    uint32_t gl_eps;         // 0xffffffff (Goldilocks 2^32 - 1), a run-time operand on purpose: see gl_limbs.cuh fold96
};
enum NttKind { NTT_STRIDED = 0, NTT_FINAL_NATURAL = 1, NTT_FINAL_LEAVES = 2 };
void launch_ntt_pass(int field, int repr, int kind, const NttPass &p, cudaStream_t stream, uint64_t *launches);
// dst_k[i] = src[i] for every peer buffer k (16-byte pieces; bytes a multiple of 16): the exchange step of a column-sharded
// inverse transform when it is sent as one block (p2g_api.cu, exchange_mode)
void launch_peer_push(const void *src, void *const *peers, unsigned num_peers, size_t bytes, cudaStream_t stream, uint64_t *launches);
// fills tw_tables / scale_tables entries from three-level power tables ([3][1024], base^(k*1024^level)):
//   tw_table[f*m_out + j0] = wpow^(f*j0) * (spow ? spow^j0 : 1);   scale_table[i] = spow^(i << log_mout)
void launch_build_tw_table(int field, void *out, uint32_t log_min, uint32_t log_np, const void *wpow, const void *spow,
                           cudaStream_t stream, uint64_t *launches);
void launch_build_scale_table(int field, void *out, uint32_t log_np, uint32_t log_mout, const void *spow,
                              cudaStream_t stream, uint64_t *launches);
// copies salt columns (column-major natural LDE order) into the leaf matrix: leaves[k][col0 + j] = salt[j][bitrev(k)]
void launch_salt_scatter(int field, const void *salts, uint32_t num_salt, uint32_t log_N, uint64_t row_begin,
                         uint64_t row_count, void *leaves, uint64_t pitch, uint32_t col0, cudaStream_t stream,
                         uint64_t *launches);
// out[col][i] (natural LDE order) = leaves[bitrev(i)][col]  -- only for p2g_transform(kind 1)
void launch_leaves_to_columns(int field, const void *leaves, uint64_t pitch, uint32_t num_cols, uint32_t log_N,
                              void *out, cudaStream_t stream, uint64_t *launches);

// data[c][j] *= base^j for j < 2^log_n (caller's representation in and out); spow = three-level power table of base
void launch_scale_powers(int field, int repr, void *data, uint32_t log_n, uint32_t num_cols, uint64_t col_stride, const void *spow,
                         cudaStream_t stream, uint64_t *launches);

// ---- fri.cu ------------------------------------------------------------------------------------
// out[i] = in[bitrev(i)], 16-byte elements (one extension element: 2 x u64 or 4 x u32)
void launch_fri_bitrev16(const void *in, uint32_t log_n, void *out, cudaStream_t stream, uint64_t *launches);
// cols[k][i] = rows[i][k] for the D coordinates of n extension elements
void launch_fri_split(int field, const void *rows, uint64_t n, void *cols, cudaStream_t stream, uint64_t *launches);
// coeffs_out[k][i] = coordinate k of sum_j coeffs_in[.][i*arity + j] * beta^j; beta and w (X^D = w) in internal form
void launch_fri_fold(int field, int repr, const void *coeffs_in, uint64_t n_in, uint32_t arity_bits,
                     const uint64_t *beta_internal, uint64_t w_internal, void *coeffs_out, cudaStream_t stream,
                     uint64_t *launches);

}  // namespace p2g
