// Batched number-theoretic transforms for the commit path, shared-memory staged.
//
// Replaces, on the GPU:
//   ifft_with_options / fft_classic      field/src/fft.rs:66-94,168-205     (per-column radix-2 FFT on the CPU)
//   PolynomialCoeffs::lde + coset_fft    field/src/polynomial/mod.rs:201-220,282-295
//   transpose + reverse_index_bits       plonky2/src/util/mod.rs:25-31, util/src/lib.rs:189-238
//
// Structure (per column of n = 2^L elements, all columns in one launch):
//   * n is split as n = n_1 * ... * n_P with every n_i <= 2^10 ("four-step" decomposition, P <= 3).
//     Pass i < P loads, for W consecutive offsets j0, the n_i points x[j0 + m*k] (m = chunk/n_i), transforms
//     them in shared memory, multiplies by the inter-pass twiddle w_chunk^(f*j0) and stores frequency f at
//     chunk_base + f*m + j0, so that the next pass again works on contiguous chunks.  Every global access is
//     a run of W consecutive elements (64 bytes).
//   * the last pass transforms contiguous chunks.  For the inverse transform it writes natural order
//     (coefficients, scaled by 1/n).  For the LDE it writes the row-major leaf matrix directly: the decimation-
//     in-frequency butterflies leave frequency f at position bitrev(f), which is exactly the row order
//     the reference obtains with transpose + reverse_index_bits, so no separate permutation pass exists.
//   * the LDE never builds the zero-padded size-N input: rows [b*n, (b+1)*n) are the size-n transform of the
//     coefficients scaled by s_b^j, s_b = g * w_N^{bitrev_r(b)}; the scaling is fused into the
//     first pass' load.
// Twiddles: in-tile roots live in shared memory; inter-pass twiddles and coset powers come from three-level
// 1024-entry power tables (w^e = T0[e & 1023] * T1[(e >> 10) & 1023] * T2[e >> 20]) that stay in L1/L2.
#include "p2g_internal.h"

namespace p2g {

struct NttGl {
    typedef u64 E;
    static const int W_LOG = 3;  // 8 x 8 B = 64-byte runs
    static const int PAD = 4;
    static __device__ __forceinline__ E load_user(E x) { return x; }
    static __device__ __forceinline__ E store_user(E x) { return gl_canon(x); }
    static __device__ __forceinline__ E mul(E a, E b) { return gl_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return gl_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return gl_sub(a, b); }
};
template <bool MONTY>
struct NttBb {
    typedef u32 E;
    static const int W_LOG = 4;  // 16 x 4 B = 64-byte runs
    static const int PAD = 2;
    static __device__ __forceinline__ E load_user(E x) { return MONTY ? BbDev::from_monty(x) : BbDev::from_canonical(x); }
    static __device__ __forceinline__ E store_user(E x) { return MONTY ? x : BbDev::to_canonical(x); }
    static __device__ __forceinline__ E mul(E a, E b) { return bb_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return bb_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return bb_sub(a, b); }
};

template <class F>
__device__ __forceinline__ typename F::E pow3(const typename F::E *__restrict__ t, u32 levels, u32 e) {
    typename F::E r = t[e & (POW_DIGIT - 1)];
    if (levels > 1) r = F::mul(r, t[POW_DIGIT + ((e >> POW_DIGIT_LOG) & (POW_DIGIT - 1))]);
    if (levels > 2) r = F::mul(r, t[2 * POW_DIGIT + (e >> (2 * POW_DIGIT_LOG))]);
    return r;
}

// Decimation-in-frequency butterflies over 2^lw interleaved transforms of np = 2^lp points held in shared memory,
// register blocked: a thread takes the 2^K points that differ in the K position bits of a round, runs K stages on
// them in registers and puts them back, so shared memory is touched once per round (<= 4 stages) and the CTA
// synchronises once per round instead of once per stage.
// element (i, w) lives at x[i * si + w * sw]; after the call frequency f sits at i = bitrev_lp(f).
// W_FAST selects which index varies fastest across threads (w for the strided tiles, i for the contiguous ones).
template <class F, int K, bool W_FAST>
__device__ __forceinline__ void dif_round(typename F::E *x, const typename F::E *roots, u32 lp, u32 lw, u32 si, u32 sw,
                                          u32 s0) {
    typedef typename F::E E;
    constexpr u32 R = 1u << K;
    const u32 lowbits = lp - s0 - K;           // position bits below the round's K bits
    const u32 units = 1u << (lp - K + lw);     // (hi, low, w) combinations
    for (u32 u = threadIdx.x; u < units; u += blockDim.x) {
        u32 w, t;
        if (W_FAST) {
            w = u & ((1u << lw) - 1);
            t = u >> lw;
        } else {
            t = u & ((1u << (lp - K)) - 1);
            w = u >> (lp - K);
        }
        const u32 low = t & ((1u << lowbits) - 1);
        const u32 hi = t >> lowbits;
        const u32 p0 = (hi << (lp - s0)) | low;
        E *base = x + p0 * si + w * sw;
        const u32 stride = si << lowbits;
        E v[R];
#pragma unroll
        for (u32 j = 0; j < R; j++) v[j] = base[j * stride];
#pragma unroll
        for (int q = 0; q < K; q++) {
            const u32 hj = R >> (q + 1);  // distance between butterfly partners, in j
            const u32 s = s0 + q;
#pragma unroll
            for (u32 j = 0; j < R; j++) {
                if (j & hj) continue;
                // twiddle w_{np}^{(position mod half) << s}, position = (j mod hj) << lowbits | low
                const u32 e = (((j & (hj - 1)) << lowbits) | low) << s;
                E a = v[j], b = v[j + hj];
                v[j] = F::add(a, b);
                v[j + hj] = F::mul(F::sub(a, b), roots[e]);
            }
        }
#pragma unroll
        for (u32 j = 0; j < R; j++) base[j * stride] = v[j];
    }
    __syncthreads();
}

template <class F, bool W_FAST>
__device__ __forceinline__ void dif_stages(typename F::E *x, const typename F::E *roots, u32 lp, u32 lw, u32 si, u32 sw) {
    u32 s0 = 0;
    // rounds of 4 stages, the remainder split so that no round is a lone stage when it can be avoided
    while (lp - s0 >= 4 && lp - s0 != 5) {
        dif_round<F, 4, W_FAST>(x, roots, lp, lw, si, sw, s0);
        s0 += 4;
    }
    u32 rem = lp - s0;
    if (rem == 5) {
        dif_round<F, 3, W_FAST>(x, roots, lp, lw, si, sw, s0);
        dif_round<F, 2, W_FAST>(x, roots, lp, lw, si, sw, s0 + 3);
    } else if (rem == 3) {
        dif_round<F, 3, W_FAST>(x, roots, lp, lw, si, sw, s0);
    } else if (rem == 2) {
        dif_round<F, 2, W_FAST>(x, roots, lp, lw, si, sw, s0);
    } else if (rem == 1) {
        dif_round<F, 1, W_FAST>(x, roots, lp, lw, si, sw, s0);
    }
}

// digit reversal of the chunk index produced by the earlier passes: q = f1 * 2^lp2 + f2  ->  f1 + 2^lp1 * f2
__device__ __forceinline__ u32 chunk_to_natural(u32 q, u32 lp1, u32 lp2) {
    u32 f2 = q & ((1u << lp2) - 1), f1 = q >> lp2;
    return f1 + (f2 << lp1);
}
__device__ __forceinline__ u32 natural_to_chunk(u32 d, u32 lp1, u32 lp2) {
    u32 f1 = d & ((1u << lp1) - 1), f2 = d >> lp1;
    return (f1 << lp2) + f2;
}

template <class F>
__global__ void __launch_bounds__(256) ntt_strided_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lmo = p.log_min - lp;                           // log2(m_out)
    const u32 lw = lmo < (u32)F::W_LOG ? lmo : (u32)F::W_LOG;  // tile width (runs of 2^lw elements)
    const u32 W = 1u << lw;
    E *roots_s = x + ((size_t)np << lw);
    const u32 tiles_per_chunk = 1u << (lmo - lw);
    const u32 tiles_per_col = tiles_per_chunk << (p.log_n - p.log_min);
    const u32 col = blockIdx.x / tiles_per_col;
    const u32 t = blockIdx.x - col * tiles_per_col;
    const u32 q = t >> (lmo - lw), jt = t & (tiles_per_chunk - 1);
    const E *src = (const E *)p.src + (size_t)col * p.src_col_stride + ((size_t)q << p.log_min);
    E *dst = (E *)p.dst + (size_t)col * p.dst_col_stride + ((size_t)q << p.log_min);
    const E *scale = (const E *)p.scale3;
    const E *tw = (const E *)p.tw3;
    const u32 total = np << lw;
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 w = idx & (W - 1), i = idx >> lw;
        u32 j = (i << lmo) + (jt << lw) + w;
        E v = src[j];
        if (p.in_user_repr) v = F::load_user(v);
        if (scale) v = F::mul(v, pow3<F>(scale, p.scale_levels, j));
        x[idx] = v;
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    __syncthreads();
    dif_stages<F, true>(x, roots_s, lp, lw, W, 1);
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 w = idx & (W - 1), pp = idx >> lw;
        u32 f = __brev(pp) >> (32 - lp);
        u32 j0 = (jt << lw) + w;
        E v = F::mul(x[idx], pow3<F>(tw, p.tw_levels, f * j0));
        dst[((size_t)f << lmo) + j0] = v;  // internal representation
    }
}

template <class F>
__global__ void __launch_bounds__(256) ntt_final_natural_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lc = p.log_n - lp;                              // log2(chunks per column)
    const u32 lw = lc < (u32)F::W_LOG ? lc : (u32)F::W_LOG;  // chunks per tile
    const u32 W = 1u << lw;
    const u32 S = np + F::PAD;
    E *roots_s = x + (size_t)S * W;
    const u32 groups_per_col = 1u << (lc - lw);
    const u32 col = blockIdx.x / groups_per_col;
    const u32 g = blockIdx.x - col * groups_per_col;
    const u32 d0 = g << lw;
    const E *src = (const E *)p.src + (size_t)col * p.src_col_stride;
    E *dst = (E *)p.dst + (size_t)col * p.dst_col_stride;
    const u32 total = np << lw;
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 i = idx & (np - 1), w = idx >> lp;
        u32 q = natural_to_chunk(d0 + w, p.lp1, p.lp2);
        E v = src[((size_t)q << lp) + i];
        if (p.in_user_repr) v = F::load_user(v);
        x[w * S + i] = v;
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    __syncthreads();
    dif_stages<F, false>(x, roots_s, lp, lw, 1, S);
    const E fs = (E)p.final_scale;
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 w = idx & (W - 1), pp = idx >> lw;
        u32 f = lp ? (__brev(pp) >> (32 - lp)) : 0;
        E v = x[w * S + pp];
        if (p.use_final_scale) v = F::mul(v, fs);
        if (p.out_user_repr) v = F::store_user(v);
        dst[(size_t)(d0 + w) + ((size_t)f << lc)] = v;
    }
}

template <class F>
__global__ void __launch_bounds__(256) ntt_final_leaves_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lc = p.log_n - lp;
    const u32 W = 1u << F::W_LOG;  // columns per tile
    const u32 S = np + F::PAD;
    E *roots_s = x + (size_t)S * W;
    const u32 col_groups = (p.num_cols + W - 1) >> F::W_LOG;
    const u32 q = blockIdx.x / col_groups;
    const u32 cg = blockIdx.x - q * col_groups;
    const u32 col0 = cg << F::W_LOG;
    const u32 ncols = min(W, p.num_cols - col0);
    const E *src = (const E *)p.src + (size_t)col0 * p.src_col_stride + ((size_t)q << lp);
    const E *scale = (const E *)p.scale3;
    const u32 total = np * ncols;
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 i = idx & (np - 1), w = idx >> lp;
        E v = src[(size_t)w * p.src_col_stride + i];
        if (p.in_user_repr) v = F::load_user(v);
        if (scale) v = F::mul(v, pow3<F>(scale, p.scale_levels, i));  // single-pass LDE: q == 0
        x[w * S + i] = v;
    }
    // unused column slots of a ragged last tile must still hold defined values for the butterflies
    for (u32 idx = total + threadIdx.x; idx < (np << F::W_LOG); idx += blockDim.x) {
        u32 i = idx & (np - 1), w = idx >> lp;
        x[w * S + i] = 0;
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    __syncthreads();
    dif_stages<F, false>(x, roots_s, lp, F::W_LOG, 1, S);
    // frequency f sits at position bitrev(f), and the leaf order wants row offset bitrev(f): rows are positions.
    const u32 dq = chunk_to_natural(q, p.lp1, p.lp2);
    const u32 rq = lc ? (__brev(dq) >> (32 - lc)) : 0;
    E *rows = (E *)p.dst + (p.row_base + ((size_t)rq << lp)) * p.dst_col_stride + col0;
    for (u32 idx = threadIdx.x; idx < (np << F::W_LOG); idx += blockDim.x) {
        u32 w = idx & (W - 1), pp = idx >> F::W_LOG;
        if (w < ncols) {
            E v = x[w * S + pp];
            if (p.out_user_repr) v = F::store_user(v);
            rows[(size_t)pp * p.dst_col_stride + w] = v;
        }
    }
}

template <class E>
__global__ void salt_scatter_kernel(const E *__restrict__ salts, u32 num_salt, u32 log_N, u64 row_begin, u64 row_count,
                                    E *__restrict__ leaves, u64 pitch, u32 col0) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_count * num_salt) return;
    u32 j = (u32)(idx % num_salt);
    u64 k = row_begin + idx / num_salt;
    u64 src = log_N ? (__brevll(k) >> (64 - log_N)) : 0;
    leaves[(k - row_begin) * pitch + col0 + j] = salts[((u64)j << log_N) + src];
}

template <class E>
__global__ void leaves_to_columns_kernel(const E *__restrict__ leaves, u64 pitch, u32 num_cols, u32 log_N,
                                         E *__restrict__ out) {
    u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 N = (u64)1 << log_N;
    if (idx >= N * num_cols) return;
    u64 i = idx & (N - 1);
    u32 col = (u32)(idx >> log_N);
    u64 k = log_N ? (__brevll(i) >> (64 - log_N)) : 0;
    out[idx] = leaves[k * pitch + col];
}

// ---- launchers ---------------------------------------------------------------------------------
template <class F>
static void launch_pass_t(int kind, const NttPass &p, cudaStream_t stream, uint64_t *launches) {
    typedef typename F::E E;
    const u32 lp = p.log_np, np = 1u << lp;
    if (kind == NTT_STRIDED) {
        u32 lmo = p.log_min - lp;
        u32 lw = lmo < (u32)F::W_LOG ? lmo : (u32)F::W_LOG;
        size_t smem = (((size_t)np << lw) + (np >> 1)) * sizeof(E);
        u64 tiles_per_col = ((u64)1 << (lmo - lw)) << (p.log_n - p.log_min);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(ntt_strided_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_set = true;
        }
        ntt_strided_kernel<F><<<(unsigned)(tiles_per_col * p.num_cols), 256, smem, stream>>>(p);
    } else if (kind == NTT_FINAL_NATURAL) {
        u32 lc = p.log_n - lp;
        u32 lw = lc < (u32)F::W_LOG ? lc : (u32)F::W_LOG;
        size_t smem = ((size_t)(np + F::PAD) * (1u << lw) + (np >> 1)) * sizeof(E);
        u64 groups = (u64)1 << (lc - lw);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(ntt_final_natural_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_set = true;
        }
        ntt_final_natural_kernel<F><<<(unsigned)(groups * p.num_cols), 256, smem, stream>>>(p);
    } else {
        u32 lc = p.log_n - lp;
        size_t smem = ((size_t)(np + F::PAD) * (1u << F::W_LOG) + (np >> 1)) * sizeof(E);
        u64 col_groups = (p.num_cols + (1u << F::W_LOG) - 1) >> F::W_LOG;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(ntt_final_leaves_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            attr_set = true;
        }
        ntt_final_leaves_kernel<F><<<(unsigned)(col_groups << lc), 256, smem, stream>>>(p);
    }
    ++*launches;
}

void launch_ntt_pass(int field, int repr, int kind, const NttPass &p, cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        launch_pass_t<NttGl>(kind, p, stream, launches);
    else if (repr)
        launch_pass_t<NttBb<true>>(kind, p, stream, launches);
    else
        launch_pass_t<NttBb<false>>(kind, p, stream, launches);
}

void launch_salt_scatter(int field, const void *salts, uint32_t num_salt, uint32_t log_N, uint64_t row_begin,
                         uint64_t row_count, void *leaves, uint64_t pitch, uint32_t col0, cudaStream_t stream,
                         uint64_t *launches) {
    u64 total = row_count * num_salt;
    if (!total) return;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (field == FIELD_GL)
        salt_scatter_kernel<u64><<<blocks, 256, 0, stream>>>((const u64 *)salts, num_salt, log_N, row_begin, row_count,
                                                          (u64 *)leaves, pitch, col0);
    else
        salt_scatter_kernel<u32><<<blocks, 256, 0, stream>>>((const u32 *)salts, num_salt, log_N, row_begin, row_count,
                                                          (u32 *)leaves, pitch, col0);
    ++*launches;
}

void launch_leaves_to_columns(int field, const void *leaves, uint64_t pitch, uint32_t num_cols, uint32_t log_N,
                              void *out, cudaStream_t stream, uint64_t *launches) {
    u64 total = ((u64)num_cols) << log_N;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (field == FIELD_GL)
        leaves_to_columns_kernel<u64><<<blocks, 256, 0, stream>>>((const u64 *)leaves, pitch, num_cols, log_N, (u64 *)out);
    else
        leaves_to_columns_kernel<u32><<<blocks, 256, 0, stream>>>((const u32 *)leaves, pitch, num_cols, log_N, (u32 *)out);
    ++*launches;
}

}  // namespace p2g
