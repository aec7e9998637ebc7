// FRI commit phase on the device: the per-layer work of fri_committed_trees,
// plonky2/src/fri/prover.rs:83-133, between two Fiat-Shamir challenges.
//
//   reverse_index_bits_in_place(values) + chunks(arity) + flatten   (:103-109)  -> fri_bitrev16_kernel for the first
//        layer only; later layers get their values from the coset NTT, whose DIF butterflies already leave them in
//        bit-reversed order, so a layer's leaf matrix [n/arity][arity*D] is the NTT output buffer itself
//   MerkleTree::new(chunked_values, cap_height)                      (:110)      -> hash.cu (leaf sponge + tree levels)
//   coeffs.chunks_exact(arity).map(reduce_with_powers(chunk, beta))  (:116-123)  -> fri_fold_kernel
//   coeffs.coset_fft(shift^arity)                                    (:124-125)  -> ntt.cu, one column per extension
//        coordinate: the twiddles and the coset shift are base-field elements, so the transform of F::Extension
//        elements is D independent base-field transforms (field/src/fft.rs:12-31 builds the root table from
//        F::Extension::two_adic_generator, which is the embedded base-field root for every size that fits)
//
// Extension fields (external p3 BinomialExtensionField, Cargo.toml:19-22): Goldilocks F[X]/(X^2 - 7),
// BabyBear F[X]/(X^4 - 11); an element is its D coefficients, lowest degree first (`as_base_slice`, :107).
// Coefficients are kept on the device one coordinate per column ([D][n], what the NTT kernels take); values are kept
// interleaved ([n][D], what the leaf sponge takes).
#include "p2g_internal.h"

namespace p2g {

namespace {

struct FriGl {
    typedef u64 E;
    static const int D = 2;
    static __device__ __forceinline__ E load_user(E x, int) { return x; }
    static __device__ __forceinline__ E store_user(E x, int) { return gl_canon(x); }
    static __device__ __forceinline__ E mul(E a, E b) { return gl_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return gl_add(a, b); }
};
struct FriBb {
    typedef u32 E;
    static const int D = 4;
    static __device__ __forceinline__ E load_user(E x, int repr) { return repr ? BbDev::from_monty(x) : BbDev::from_canonical(x); }
    static __device__ __forceinline__ E store_user(E x, int repr) { return repr ? x : BbDev::to_canonical(x); }
    static __device__ __forceinline__ E mul(E a, E b) { return bb_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return bb_add(a, b); }
};

template <class F>
struct ExtParams {
    typename F::E beta[F::D];  // internal representation
    typename F::E w;           // X^D = w, internal representation
};

// acc <- acc * b in F[X]/(X^D - w)
template <class F>
__device__ __forceinline__ void ext_mul(typename F::E acc[F::D], const typename F::E b[F::D], typename F::E w) {
    typedef typename F::E E;
    E lo[F::D], hi[F::D];
#pragma unroll
    for (int k = 0; k < F::D; k++) lo[k] = hi[k] = 0;
#pragma unroll
    for (int i = 0; i < F::D; i++)
#pragma unroll
        for (int j = 0; j < F::D; j++) {
            E p = F::mul(acc[i], b[j]);
            if (i + j < F::D)
                lo[i + j] = F::add(lo[i + j], p);
            else
                hi[i + j - F::D] = F::add(hi[i + j - F::D], p);
        }
#pragma unroll
    for (int k = 0; k < F::D; k++) acc[k] = k + 1 < F::D ? F::add(lo[k], F::mul(hi[k], w)) : lo[k];
}

// out[k][i] = sum_j in[k][i*arity + j] * beta^j  (Horner from the last term, plonk_common.rs:124-136)
template <class F>
__global__ void fri_fold_kernel(const typename F::E *__restrict__ in, u64 n_in, u32 arity_bits, ExtParams<F> prm, int repr,
                                typename F::E *__restrict__ out) {
    typedef typename F::E E;
    const u64 n_out = n_in >> arity_bits;
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const u32 arity = 1u << arity_bits;
    E acc[F::D];
#pragma unroll
    for (int k = 0; k < F::D; k++) acc[k] = 0;
    for (u32 j = arity; j-- > 0;) {
        ext_mul<F>(acc, prm.beta, prm.w);
#pragma unroll
        for (int k = 0; k < F::D; k++) acc[k] = F::add(acc[k], F::load_user(in[(u64)k * n_in + (i << arity_bits) + j], repr));
    }
#pragma unroll
    for (int k = 0; k < F::D; k++) out[(u64)k * n_out + i] = F::store_user(acc[k], repr);
}

// out[i] = in[bitrev(i)] for 16-byte elements (one extension element of either field)
__global__ void fri_bitrev16_kernel(const uint4 *__restrict__ in, u32 log_n, uint4 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    u64 r = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    out[i] = in[r];
}

// cols[k][i] = rows[i][k]
template <class E>
__global__ void fri_split_kernel(const E *__restrict__ rows, u64 n, u32 d, E *__restrict__ cols) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * d) return;
    u64 i = t / d;
    u32 k = (u32)(t % d);
    cols[(u64)k * n + i] = rows[t];
}

}  // namespace

void launch_fri_bitrev16(const void *in, uint32_t log_n, void *out, cudaStream_t stream, uint64_t *launches) {
    u64 n = (u64)1 << log_n;
    fri_bitrev16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const uint4 *)in, log_n, (uint4 *)out);
    ++*launches;
}

void launch_fri_split(int field, const void *rows, uint64_t n, void *cols, cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL) {
        u64 t = n * FriGl::D;
        fri_split_kernel<u64><<<(unsigned)((t + 255) / 256), 256, 0, stream>>>((const u64 *)rows, n, FriGl::D, (u64 *)cols);
    } else {
        u64 t = n * FriBb::D;
        fri_split_kernel<u32><<<(unsigned)((t + 255) / 256), 256, 0, stream>>>((const u32 *)rows, n, FriBb::D, (u32 *)cols);
    }
    ++*launches;
}

void launch_fri_fold(int field, int repr, const void *coeffs_in, uint64_t n_in, uint32_t arity_bits,
                     const uint64_t *beta_internal, uint64_t w_internal, void *coeffs_out, cudaStream_t stream,
                     uint64_t *launches) {
    u64 n_out = n_in >> arity_bits;
    unsigned blocks = (unsigned)((n_out + 127) / 128);
    if (field == FIELD_GL) {
        ExtParams<FriGl> prm;
        for (int k = 0; k < FriGl::D; k++) prm.beta[k] = beta_internal[k];
        prm.w = w_internal;
        fri_fold_kernel<FriGl><<<blocks, 128, 0, stream>>>((const u64 *)coeffs_in, n_in, arity_bits, prm, repr, (u64 *)coeffs_out);
    } else {
        ExtParams<FriBb> prm;
        for (int k = 0; k < FriBb::D; k++) prm.beta[k] = (u32)beta_internal[k];
        prm.w = (u32)w_internal;
        fri_fold_kernel<FriBb><<<blocks, 128, 0, stream>>>((const u32 *)coeffs_in, n_in, arity_bits, prm, repr, (u32 *)coeffs_out);
    }
    ++*launches;
}

}  // namespace p2g
