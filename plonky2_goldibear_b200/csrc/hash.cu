// Sponge, 2-to-1 compression and Merkle-tree kernels (one leaf / one node per thread).
//
// Replaces, on the GPU:
//   Hasher::hash_or_noop            plonky2/src/plonk/config.rs:70-84
//   hash_n_to_m_no_pad (overwrite)  plonky2/src/hash/hashing.rs:100-123
//   compress (two_to_one)           plonky2/src/hash/hashing.rs:76-96
//   fill_subtree / fill_digests_buf plonky2/src/hash/merkle_tree.rs:86-149 (same digest layout, level by level
//                                   instead of recursively)
#include <cstdio>

#include "p2g_internal.h"

namespace p2g {
__constant__ PoseidonGlTables c_pgl;
__constant__ Poseidon2BbTables c_p2bb;
}  // namespace p2g

#include "poseidon2_bb.cuh"
#include "poseidon_gl.cuh"

namespace p2g {

cudaError_t upload_hash_tables(const PoseidonGlTables *gl, const Poseidon2BbTables *bb) {
    cudaError_t e = cudaMemcpyToSymbol(c_pgl, gl, sizeof(PoseidonGlTables));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_p2bb, bb, sizeof(Poseidon2BbTables));
}

// ---- hash policies -----------------------------------------------------------------------------
struct HashGl {
    typedef u64 W;
    static const int WIDTH = 12, RATE = 8, DIGEST = 4;
    static __device__ __forceinline__ W load(W x) { return x; }
    static __device__ __forceinline__ W store(W x) { return gl_canon(x); }
    static __device__ __forceinline__ void permute(W *s) { poseidon_gl_permute(s); }
};
template <bool MONTY>
struct HashBb {
    typedef u32 W;
    static const int WIDTH = 16, RATE = 8, DIGEST = 8;
    static __device__ __forceinline__ W load(W x) { return MONTY ? BbDev::from_monty(x) : BbDev::from_canonical(x); }
    static __device__ __forceinline__ W store(W x) { return MONTY ? x : BbDev::to_canonical(x); }
    static __device__ __forceinline__ void permute(W *s) { poseidon2_bb_permute(s); }
};

// position (in digests) of node j of layer `layer` (0 = leaf digests) inside one cap subtree;
// closed form of the recursive layout, cf. MerkleTree::prove merkle_tree.rs:200-217
__device__ __forceinline__ u64 node_slot(u32 layer, u64 j) {
    return ((j >> 1) << (layer + 2)) + ((u64)2 << layer) - 2 + (j & 1);
}

template <class H>
__device__ __forceinline__ void store_digest(typename H::W *dst, const typename H::W *d) {
    // 32 bytes, 32-byte aligned
    if (sizeof(typename H::W) == 8) {
        ulonglong2 a = make_ulonglong2((u64)d[0], (u64)d[1]), b = make_ulonglong2((u64)d[2], (u64)d[3]);
        reinterpret_cast<ulonglong2 *>(dst)[0] = a;
        reinterpret_cast<ulonglong2 *>(dst)[1] = b;
    } else {
        uint4 a = make_uint4((u32)d[0], (u32)d[1], (u32)d[2], (u32)d[3]);
        uint4 b = make_uint4((u32)d[4], (u32)d[5], (u32)d[6], (u32)d[7]);
        reinterpret_cast<uint4 *>(dst)[0] = a;
        reinterpret_cast<uint4 *>(dst)[1] = b;
    }
}

template <class H>
__global__ void __launch_bounds__(128)
leaf_hash_kernel(const typename H::W *__restrict__ leaves, u64 pitch, u32 len, u64 row_begin, u64 row_count,
                 u32 log_sub, typename H::W *__restrict__ digests, typename H::W *__restrict__ cap) {
    typedef typename H::W W;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < row_count;  // no early return: the permutation contains CTA barriers
    u64 k = row_begin + (live ? t : row_count - 1);
    const W *row = leaves + k * pitch;
    W out[H::DIGEST];
    if (len <= (u32)H::DIGEST) {
        // hash_or_noop: short leaves are copied (canonical bytes), zero padded
#pragma unroll
        for (int i = 0; i < H::DIGEST; i++) out[i] = (u32)i < len ? H::store(H::load(row[i])) : (W)0;
    } else {
        W s[H::WIDTH];
#pragma unroll
        for (int i = 0; i < H::WIDTH; i++) s[i] = 0;
#pragma unroll 1
        for (u32 off = 0; off < len; off += H::RATE) {
            // overwrite mode: a partial last chunk replaces only its own words (hashing.rs:107-110)
#pragma unroll
            for (int i = 0; i < H::RATE; i++)
                if (off + i < len) s[i] = H::load(row[off + i]);
            H::permute(s);
        }
#pragma unroll
        for (int i = 0; i < H::DIGEST; i++) out[i] = H::store(s[i]);
    }
    if (!live) return;
    W *dst;
    if (log_sub == 0) {
        dst = cap + k * H::DIGEST;  // tree that is all cap (merkle_tree.rs:124-133)
    } else {
        u64 sub = k >> log_sub, l = k & (((u64)1 << log_sub) - 1);
        u64 sub_digests = 2 * (((u64)1 << log_sub) - 1);
        dst = digests + (sub * sub_digests + node_slot(0, l)) * H::DIGEST;
    }
    store_digest<H>(dst, out);
}

// parents of layer `layer` nodes: one thread per parent
template <class H>
__global__ void __launch_bounds__(128)
tree_level_kernel(typename H::W *__restrict__ digests, typename H::W *__restrict__ cap, u32 layer, u32 log_sub,
                  u64 num_parents) {
    typedef typename H::W W;
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < num_parents;
    if (!live) idx = num_parents - 1;
    u32 lg_pps = log_sub - layer - 1;  // log2(parents per subtree)
    u64 sub = idx >> lg_pps, j = idx & (((u64)1 << lg_pps) - 1);
    u64 sub_digests = 2 * (((u64)1 << log_sub) - 1);
    W *base = digests + sub * sub_digests * H::DIGEST;
    const W *pair = base + node_slot(layer, 2 * j) * H::DIGEST;  // children are adjacent: 64 bytes
    W s[H::WIDTH];
#pragma unroll
    for (int i = 0; i < 2 * H::DIGEST; i++) s[i] = H::load(pair[i]);
#pragma unroll
    for (int i = 2 * H::DIGEST; i < H::WIDTH; i++) s[i] = 0;
    H::permute(s);
    if (!live) return;
    W out[H::DIGEST];
#pragma unroll
    for (int i = 0; i < H::DIGEST; i++) out[i] = H::store(s[i]);
    W *dst = (layer + 1 == log_sub) ? cap + sub * H::DIGEST : base + node_slot(layer + 1, j) * H::DIGEST;
    store_digest<H>(dst, out);
}

template <class H>
__global__ void __launch_bounds__(128) permute_kernel(typename H::W *states, u64 count) {
    typedef typename H::W W;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < count;
    if (!live) t = count - 1;
    W s[H::WIDTH];
#pragma unroll
    for (int i = 0; i < H::WIDTH; i++) s[i] = H::load(states[t * H::WIDTH + i]);
    H::permute(s);
    if (!live) return;
#pragma unroll
    for (int i = 0; i < H::WIDTH; i++) states[t * H::WIDTH + i] = H::store(s[i]);
}

// ---- FRI grinding (fri_proof_of_work, plonky2/src/fri/prover.rs:136-188) ------------------------------------------
// Candidate c = start + thread: overwrite lane `pos` of the half-duplexed sponge state with c, permute, and test the
// LAST squeezed element (lane RATE-1, :175) for enough leading zeros as a canonical u64 (:176-177).  The smallest
// hit of the launch wins (atomicMin): the reference takes whichever rayon finds first (find_any), any hit is a valid
// witness, and "smallest" makes ours reproducible.
template <class H>
__global__ void __launch_bounds__(128) pow_search_kernel(const typename H::W *__restrict__ state, u32 pos, u32 min_leading_zeros,
                                                         u64 start, u64 count, u64 last_candidate, unsigned long long *best) {
    typedef typename H::W W;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 cand = start + t;
    const bool live = t < count && cand <= last_candidate && cand >= start;
    W s[H::WIDTH];
#pragma unroll
    for (int i = 0; i < H::WIDTH; i++) s[i] = H::load(state[i]);
#pragma unroll
    for (int i = 0; i < H::RATE; i++)
        if ((u32)i == pos) s[i] = H::load((W)(live ? cand : 0));
    H::permute(s);
    if (!live) return;
    u64 resp = (u64)H::store(s[H::RATE - 1]);
    u32 lz = resp ? (u32)__clzll((long long)resp) : 64u;
    if (lz >= min_leading_zeros) atomicMin(best, (unsigned long long)cand);
}

// ---- device-resident accessors (MerkleTree::get / MerkleTree::prove on the GPU copy) -----------------------------
template <class E>
__global__ void gather_rows_kernel(const E *__restrict__ leaves, u64 pitch, u32 width, const u64 *__restrict__ idx,
                                   u64 count, E *__restrict__ out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * width) return;
    u64 i = t / width;
    u32 j = (u32)(t - i * width);
    out[t] = leaves[idx[i] * pitch + j];
}
// one thread per (proof, layer); 32-byte digests moved as two 16-byte words.  Index arithmetic of merkle_tree.rs:200-217.
__global__ void gather_proofs_kernel(const uint4 *__restrict__ digests, u32 log_leaves, u32 cap_height,
                                     const u64 *__restrict__ idx, u64 count, uint4 *__restrict__ out) {
    u32 num_layers = log_leaves - cap_height;
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * num_layers) return;
    u64 i = t / num_layers;
    u32 layer = (u32)(t - i * num_layers);
    u64 leaf = idx[i];
    u64 tree_index = leaf >> num_layers;
    u64 tree_len = 2 * (((u64)1 << num_layers) - 1);
    u64 pair_index = (leaf & (((u64)1 << num_layers) - 1)) >> layer;
    u64 parity = pair_index & 1;
    pair_index >>= 1;
    u64 siblings_index = (pair_index << (layer + 1)) + ((u64)1 << layer) - 1;
    u64 slot = tree_len * tree_index + 2 * siblings_index + (1 - parity);
    out[2 * t] = digests[2 * slot];
    out[2 * t + 1] = digests[2 * slot + 1];
}

void launch_gather_rows(int field, const void *leaves, uint64_t pitch, uint32_t width, const uint64_t *idx, uint64_t count,
                        void *rows_out, cudaStream_t stream, uint64_t *launches) {
    u64 total = count * width;
    if (!total) return;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (field == FIELD_GL)
        gather_rows_kernel<u64><<<blocks, 256, 0, stream>>>((const u64 *)leaves, pitch, width, (const u64 *)idx, count, (u64 *)rows_out);
    else
        gather_rows_kernel<u32><<<blocks, 256, 0, stream>>>((const u32 *)leaves, pitch, width, (const u64 *)idx, count, (u32 *)rows_out);
    ++*launches;
}
void launch_gather_proofs(const void *digests, uint32_t log_leaves, uint32_t cap_height, const uint64_t *idx,
                          uint64_t count, void *siblings_out, cudaStream_t stream, uint64_t *launches) {
    u64 total = count * (log_leaves - cap_height);
    if (!total) return;
    gather_proofs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const uint4 *)digests, log_leaves, cap_height,
                                                                            (const u64 *)idx, count, (uint4 *)siblings_out);
    ++*launches;
}

// ---- launchers ---------------------------------------------------------------------------------
template <class H>
static void leaf_hash_t(const void *leaves, u64 pitch, u32 len, u64 row_begin, u64 row_count, const TreeShape &tree,
                        cudaStream_t stream, uint64_t *launches) {
    typedef typename H::W W;
    if (row_count == 0) return;
    u32 log_sub = tree.log_leaves - tree.cap_height;
    unsigned blocks = (unsigned)((row_count + 127) / 128);
    leaf_hash_kernel<H><<<blocks, 128, 0, stream>>>((const W *)leaves, pitch, len, row_begin, row_count, log_sub,
                                                   (W *)tree.digests, (W *)tree.cap);
    ++*launches;
}
template <class H>
static void tree_levels_t(const TreeShape &tree, cudaStream_t stream, uint64_t *launches) {
    typedef typename H::W W;
    u32 log_sub = tree.log_leaves - tree.cap_height;
    for (u32 layer = 0; layer < log_sub; layer++) {
        u64 parents = ((u64)1 << tree.log_leaves) >> (layer + 1);
        unsigned blocks = (unsigned)((parents + 127) / 128);
        tree_level_kernel<H><<<blocks, 128, 0, stream>>>((W *)tree.digests, (W *)tree.cap, layer, log_sub, parents);
        ++*launches;
    }
}
template <class H>
static void permute_t(void *states, u64 count, cudaStream_t stream, uint64_t *launches) {
    typedef typename H::W W;
    if (!count) return;
    permute_kernel<H><<<(unsigned)((count + 127) / 128), 128, 0, stream>>>((W *)states, count);
    ++*launches;
}

void launch_leaf_hash(int field, int repr, const void *leaves, uint64_t pitch, uint32_t len, uint64_t row_begin,
                      uint64_t row_count, const TreeShape &tree, cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        leaf_hash_t<HashGl>(leaves, pitch, len, row_begin, row_count, tree, stream, launches);
    else if (repr)
        leaf_hash_t<HashBb<true>>(leaves, pitch, len, row_begin, row_count, tree, stream, launches);
    else
        leaf_hash_t<HashBb<false>>(leaves, pitch, len, row_begin, row_count, tree, stream, launches);
}
void launch_tree_levels(int field, int repr, const TreeShape &tree, cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        tree_levels_t<HashGl>(tree, stream, launches);
    else if (repr)
        tree_levels_t<HashBb<true>>(tree, stream, launches);
    else
        tree_levels_t<HashBb<false>>(tree, stream, launches);
}
void launch_permute(int field, int repr, void *states, uint64_t count, cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        permute_t<HashGl>(states, count, stream, launches);
    else if (repr)
        permute_t<HashBb<true>>(states, count, stream, launches);
    else
        permute_t<HashBb<false>>(states, count, stream, launches);
}

void launch_pow_search(int field, const void *state, uint32_t pos, uint32_t min_leading_zeros, uint64_t start,
                       uint64_t count, uint64_t last_candidate, void *best, cudaStream_t stream, uint64_t *launches) {
    unsigned blocks = (unsigned)((count + 127) / 128);
    if (field == FIELD_GL)
        pow_search_kernel<HashGl><<<blocks, 128, 0, stream>>>((const u64 *)state, pos, min_leading_zeros, start, count, last_candidate,
                                                              (unsigned long long *)best);
    else
        pow_search_kernel<HashBb<false>><<<blocks, 128, 0, stream>>>((const u32 *)state, pos, min_leading_zeros, start, count,
                                                                     last_candidate, (unsigned long long *)best);
    ++*launches;
}

}  // namespace p2g
