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
// Twiddles: in-tile roots live in shared memory.  Inter-pass twiddles (with the coset factor s^j0 merged in on the
// first LDE pass) come from a per-pass table indexed exactly like the store, so they are read in the same 64-byte
// runs; the remaining coset factor s^(m*i) is one small per-row table.  Both are built once per (field, size, coset)
// by a kernel from three-level 1024-entry power tables and cached in the context.
#include <mutex>
#include <set>
#include <utility>

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
    static const int PAD = 4;    // keeps every padded row 16-byte aligned for cp.async
    // The transforms are linear and every twiddle / coset factor / scale is a Montgomery-form constant, so
    // mul(word, constant) = word * constant in the WORD's own representation: canonical words go through the passes as
    // canonical words and Montgomery words as Montgomery words -- no conversion on the way in or out, only the range
    // fix that lets any u32 in (two conditional subtractions).  (The hash kernels, which are not linear, do convert.)
    static __device__ __forceinline__ E load_user(E x) { return BbDev::from_monty(x); }
    static __device__ __forceinline__ E store_user(E x) { return x; }
    static __device__ __forceinline__ E mul(E a, E b) { return bb_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return bb_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return bb_sub(a, b); }
};

// base^e from the three-level power table (table building only)
template <class F>
__device__ __forceinline__ typename F::E pow3(const typename F::E *__restrict__ t, u32 e) {
    typename F::E r = t[e & (POW_DIGIT - 1)];
    if (e >> POW_DIGIT_LOG) r = F::mul(r, t[POW_DIGIT + ((e >> POW_DIGIT_LOG) & (POW_DIGIT - 1))]);
    if (e >> (2 * POW_DIGIT_LOG)) r = F::mul(r, t[2 * POW_DIGIT + (e >> (2 * POW_DIGIT_LOG))]);
    return r;
}

template <class F>
__global__ void build_tw_table_kernel(typename F::E *out, u32 log_min, u32 log_np, const typename F::E *wpow,
                                      const typename F::E *spow) {
    u32 pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >> log_min) return;
    u32 lmo = log_min - log_np;
    u32 f = pos >> lmo, j0 = pos & ((1u << lmo) - 1);
    typename F::E v = pow3<F>(wpow, f * j0);
    if (spow) v = F::mul(v, pow3<F>(spow, j0));
    out[pos] = v;
}
template <class F>
__global__ void build_scale_table_kernel(typename F::E *out, u32 log_np, u32 log_mout, const typename F::E *spow) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_np) return;
    out[i] = pow3<F>(spow, i << log_mout);
}

// 16-byte asynchronous global->shared copies (LDGSTS): the tile load needs no registers and no per-element round trip
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// what the first round does to an element as it picks it up from the staged tile
template <class F>
struct LoadFix {
    const typename F::E *scale;  // per tile row factor (coset scaling), or null
    bool user_repr;              // staged data is in the caller's representation
    __device__ __forceinline__ typename F::E operator()(typename F::E v, u32 row) const {
        if (user_repr) v = F::load_user(v);
        if (scale) v = F::mul(v, scale[row]);
        return v;
    }
    // the same for the R rows base + (j << shift) of one register-blocked unit: the factors are fetched through one
    // per-thread pointer, a batch ahead of the multiplications
    template <u32 R>
    __device__ __forceinline__ void unit(typename F::E (&v)[R], u32 base, u32 shift) const {
        if (user_repr) {
#pragma unroll
            for (u32 j = 0; j < R; j++) v[j] = F::load_user(v[j]);
        }
        if (scale) {
            const typename F::E *sp = scale + base;
            typename F::E f[R];
#pragma unroll
            for (u32 j = 0; j < R; j++) f[j] = __ldg(sp + (j << shift));
#pragma unroll
            for (u32 j = 0; j < R; j++) v[j] = F::mul(v[j], f[j]);
        }
    }
};

// Decimation-in-frequency butterflies over 2^lw interleaved transforms of np = 2^lp points held in shared memory,
// register blocked: a thread takes the 2^K points that differ in the K position bits of a round, runs K stages on
// them in registers and puts them back, so shared memory is touched once per round (<= 4 stages) and the CTA
// synchronises once per round instead of once per stage.
// element (i, w) lives at x[i * si + w * sw]; after the call frequency f sits at i = bitrev_lp(f).
// W_FAST selects which index varies fastest across threads (w for the strided tiles, i for the contiguous ones).
template <class F, int K, bool W_FAST, bool FIRST>
__device__ __forceinline__ void dif_round(typename F::E *x, const typename F::E *roots, u32 lp, u32 lw, u32 si, u32 sw,
                                          u32 s0, const LoadFix<F> &fix) {
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
        if (FIRST) fix.template unit<R>(v, p0, lowbits);
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
__device__ __forceinline__ void dif_stages(typename F::E *x, const typename F::E *roots, u32 lp, u32 lw, u32 si, u32 sw,
                                           const LoadFix<F> &fix) {
    // rounds of 4 stages, the remainder split so that no round is a lone stage when it can be avoided;
    // the first round also applies the load conversion (representation change, coset factor)
    u32 first = lp >= 4 && lp != 5 ? 4 : (lp == 5 ? 3 : lp);
    if (first == 4)
        dif_round<F, 4, W_FAST, true>(x, roots, lp, lw, si, sw, 0, fix);
    else if (first == 3)
        dif_round<F, 3, W_FAST, true>(x, roots, lp, lw, si, sw, 0, fix);
    else if (first == 2)
        dif_round<F, 2, W_FAST, true>(x, roots, lp, lw, si, sw, 0, fix);
    else if (first == 1)
        dif_round<F, 1, W_FAST, true>(x, roots, lp, lw, si, sw, 0, fix);
    else {
        // lp == 0: a one-point transform still owes the conversion
        for (u32 u = threadIdx.x; u < (1u << lw); u += blockDim.x) x[u * sw] = fix(x[u * sw], 0);
        __syncthreads();
        return;
    }
    u32 s0 = first;
    while (lp - s0 >= 4 && lp - s0 != 5) {
        dif_round<F, 4, W_FAST, false>(x, roots, lp, lw, si, sw, s0, fix);
        s0 += 4;
    }
    u32 rem = lp - s0;
    if (rem == 5) {
        dif_round<F, 3, W_FAST, false>(x, roots, lp, lw, si, sw, s0, fix);
        dif_round<F, 2, W_FAST, false>(x, roots, lp, lw, si, sw, s0 + 3, fix);
    } else if (rem == 3) {
        dif_round<F, 3, W_FAST, false>(x, roots, lp, lw, si, sw, s0, fix);
    } else if (rem == 2) {
        dif_round<F, 2, W_FAST, false>(x, roots, lp, lw, si, sw, s0, fix);
    } else if (rem == 1) {
        dif_round<F, 1, W_FAST, false>(x, roots, lp, lw, si, sw, s0, fix);
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

// epilogue of a strided pass: tile element idx = (pp << lw) + w holds frequency f = bitrev(pp) of run element w; it is
// multiplied by the inter-pass factor and stored at chunk + (f << lmo) + jt*W + w.  Factors are read a batch ahead of
// the multiplications.  GUARD = false when the tile is a whole number of batches (every large tile).
template <class F, bool GUARD>
__device__ __forceinline__ void strided_store(const typename F::E *x, typename F::E *dst, const typename F::E *tw, u32 total,
                                              u32 lp, u32 lw, u32 lmo, u32 jt) {
    typedef typename F::E E;
    constexpr u32 B = 8;
    const u32 W = 1u << lw;
    for (u32 idx0 = threadIdx.x; idx0 < total; idx0 += blockDim.x * B) {
        E t[B];
        u32 pos[B];
#pragma unroll
        for (u32 k = 0; k < B; k++) {
            u32 idx = idx0 + k * blockDim.x;
            u32 w = idx & (W - 1), pp = idx >> lw;
            u32 f = __brev(pp) >> (32 - lp);
            pos[k] = (f << lmo) + (jt << lw) + w;
            if (!GUARD || idx < total) t[k] = tw[pos[k]];
        }
#pragma unroll
        for (u32 k = 0; k < B; k++) {
            u32 idx = idx0 + k * blockDim.x;
            if (!GUARD || idx < total) dst[pos[k]] = F::mul(x[idx], t[k]);  // internal representation
        }
    }
}

template <class F>
__global__ void __launch_bounds__(256, 3) ntt_strided_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lmo = p.log_min - lp;                           // log2(m_out)
    const u32 wl = tile_width_log(F::W_LOG, lp);
    const u32 lw = lmo < wl ? lmo : wl;  // tile width (runs of 2^lw elements)
    const u32 W = 1u << lw;
    E *roots_s = x + ((size_t)np << lw);
    const u32 tiles_per_chunk = 1u << (lmo - lw);
    const u32 tiles_per_col = tiles_per_chunk << (p.log_n - p.log_min);
    const u32 col = blockIdx.x / tiles_per_col;
    const u32 t = blockIdx.x - col * tiles_per_col;
    const u32 q = t >> (lmo - lw), jt = t & (tiles_per_chunk - 1);
    const E *src = (const E *)p.src + (size_t)col * p.src_col_stride + ((size_t)q << p.log_min);
    E *dst = (E *)p.dst + (size_t)col * p.dst_col_stride + ((size_t)q << p.log_min);
    const E *tw = (const E *)p.tw_table;
    const u32 total = np << lw;
    // stage the raw tile: row i is a run of W consecutive elements at src[(i << lmo) + jt*W]
    const u32 run_bytes = (u32)sizeof(E) << lw;
    if (p.async_ok && run_bytes >= 16) {
        // 16-byte pieces; a thread keeps its piece of the run and walks down the rows (block size is a multiple of the
        // pieces per row), so the loop is a copy and two pointer increments
        const u32 lpr = 31 - __clz(run_bytes >> 4);  // log2(pieces per row)
        const u32 part = threadIdx.x & ((1u << lpr) - 1), rows_step = blockDim.x >> lpr;
        const char *g = (const char *)(src + (jt << lw)) + part * 16 + (((size_t)(threadIdx.x >> lpr) << lmo) * sizeof(E));
        char *sm = (char *)x + (size_t)(threadIdx.x >> lpr) * run_bytes + part * 16;
        const size_t g_step = ((size_t)rows_step << lmo) * sizeof(E);
        const u32 s_step = rows_step * run_bytes;
#pragma unroll 4
        for (u32 i = threadIdx.x >> lpr; i < np; i += rows_step) {
            cp_async16(sm, g);
            g += g_step;
            sm += s_step;
        }
    } else {
        for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
            u32 w = idx & (W - 1), i = idx >> lw;
            x[idx] = src[((size_t)i << lmo) + (jt << lw) + w];
        }
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    cp_async_wait_all();
    __syncthreads();
    LoadFix<F> fix = {(const E *)p.scale_table, p.in_user_repr != 0};
    dif_stages<F, true>(x, roots_s, lp, lw, W, 1, fix);
    // store with the inter-pass factor; the factor table is indexed like the store, read a batch ahead
    if (total % (blockDim.x * 8) == 0)
        strided_store<F, false>(x, dst, tw, total, lp, lw, lmo, jt);
    else
        strided_store<F, true>(x, dst, tw, total, lp, lw, lmo, jt);
}

template <class F>
__global__ void __launch_bounds__(256, 3) ntt_final_natural_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lc = p.log_n - lp;                              // log2(chunks per column)
    const u32 wl = tile_width_log(F::W_LOG, lp);
    const u32 lw = lc < wl ? lc : wl;  // chunks per tile
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
    const u32 chunk_bytes = (u32)sizeof(E) << lp;
    if (p.async_ok && chunk_bytes >= 16) {
        const u32 per_chunk = chunk_bytes >> 4;
        for (u32 w = 0; w < W; w++) {
            const u32 q = natural_to_chunk(d0 + w, p.lp1, p.lp2);
            const char *g = (const char *)(src + ((size_t)q << lp)) + threadIdx.x * 16;
            char *sm = (char *)(x + (size_t)w * S) + threadIdx.x * 16;
#pragma unroll 4
            for (u32 part = threadIdx.x; part < per_chunk; part += blockDim.x) {
                cp_async16(sm, g);
                g += blockDim.x * 16;
                sm += blockDim.x * 16;
            }
        }
    } else {
        for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
            u32 i = idx & (np - 1), w = idx >> lp;
            u32 q = natural_to_chunk(d0 + w, p.lp1, p.lp2);
            x[w * S + i] = src[((size_t)q << lp) + i];
        }
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    cp_async_wait_all();
    __syncthreads();
    LoadFix<F> fix = {nullptr, p.in_user_repr != 0};
    dif_stages<F, false>(x, roots_s, lp, lw, 1, S, fix);
    const E fs = (E)p.final_scale;
    for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
        u32 w = idx & (W - 1), pp = idx >> lw;
        u32 f = lp ? (__brev(pp) >> (32 - lp)) : 0;
        E v = x[w * S + pp];
        if (p.use_final_scale) v = F::mul(v, fs);
        if (p.out_user_repr) v = F::store_user(v);
        const size_t o = (size_t)(d0 + w) + ((size_t)f << lc);
        dst[o] = v;
        // fused all-gather: the same 64-byte run goes to every peer's buffer (plain stores to NVLink-mapped memory)
        for (u32 k = 0; k < p.num_peers; k++) ((E *)p.peer_dst[k] + (size_t)col * p.dst_col_stride)[o] = v;
    }
}

template <class F>
__global__ void __launch_bounds__(256, 3) ntt_final_leaves_kernel(NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const u32 lp = p.log_np, np = 1u << lp;
    const u32 lc = p.log_n - lp;
    const u32 wl = tile_width_log(F::W_LOG, lp);
    const u32 W = 1u << wl;  // columns per tile
    const u32 S = np + F::PAD;
    E *roots_s = x + (size_t)S * W;
    const u32 col_groups = (p.num_cols + W - 1) >> wl;
    const u32 q = blockIdx.x / col_groups;
    const u32 cg = blockIdx.x - q * col_groups;
    const u32 col0 = cg << wl;
    const u32 ncols = min(W, p.num_cols - col0);
    const E *src = (const E *)p.src + (size_t)col0 * p.src_col_stride + ((size_t)q << lp);
    const u32 total = np * ncols;
    const u32 chunk_bytes = (u32)sizeof(E) << lp;
    if (p.async_ok && chunk_bytes >= 16) {
        const u32 per_chunk = chunk_bytes >> 4;
        for (u32 w = 0; w < ncols; w++) {
            const char *g = (const char *)(src + (size_t)w * p.src_col_stride) + threadIdx.x * 16;
            char *sm = (char *)(x + (size_t)w * S) + threadIdx.x * 16;
#pragma unroll 4
            for (u32 part = threadIdx.x; part < per_chunk; part += blockDim.x) {
                cp_async16(sm, g);
                g += blockDim.x * 16;
                sm += blockDim.x * 16;
            }
        }
    } else {
        for (u32 idx = threadIdx.x; idx < total; idx += blockDim.x) {
            u32 i = idx & (np - 1), w = idx >> lp;
            x[w * S + i] = src[(size_t)w * p.src_col_stride + i];
        }
    }
    // unused column slots of a ragged last tile must still hold defined values for the butterflies
    for (u32 idx = total + threadIdx.x; idx < (np << wl); idx += blockDim.x) {
        u32 i = idx & (np - 1), w = idx >> lp;
        x[w * S + i] = 0;
    }
    const E *roots = (const E *)p.roots;
    for (u32 k = threadIdx.x; k < (np >> 1); k += blockDim.x) roots_s[k] = roots[k];
    cp_async_wait_all();
    __syncthreads();
    LoadFix<F> fix = {(const E *)p.scale_table, p.in_user_repr != 0};  // scale: single-pass LDE (q == 0), row = position
    dif_stages<F, false>(x, roots_s, lp, wl, 1, S, fix);
    // frequency f sits at position bitrev(f), and the leaf order wants row offset bitrev(f): rows are positions.
    const u32 dq = chunk_to_natural(q, p.lp1, p.lp2);
    const u32 rq = lc ? (__brev(dq) >> (32 - lc)) : 0;
    E *rows = (E *)p.dst + (p.row_base + ((size_t)rq << lp)) * p.dst_col_stride + col0;
    {
        // a thread keeps its column and walks down the rows (block size is a multiple of W): one shared load, the
        // representation change and one store per element, pointers advanced by constants
        const u32 w = threadIdx.x & (W - 1), pp0 = threadIdx.x >> wl, step = blockDim.x >> wl;
        if (w < ncols) {
            const E *sm = x + w * S + pp0;
            E *out = rows + (size_t)pp0 * p.dst_col_stride + w;
            const size_t out_step = (size_t)step * p.dst_col_stride;
            if (p.out_user_repr) {
#pragma unroll 4
                for (u32 pp = pp0; pp < np; pp += step) {
                    *out = F::store_user(*sm);
                    sm += step;
                    out += out_step;
                }
            } else {
#pragma unroll 4
                for (u32 pp = pp0; pp < np; pp += step) {
                    *out = *sm;
                    sm += step;
                    out += out_step;
                }
            }
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
// one thread per 16 tile elements (a 4-stage round gives every thread exactly one unit), between 64 and 256 threads
static unsigned block_threads(u64 tile_elems) {
    u64 t = tile_elems >> 4;
    return (unsigned)(t < 64 ? 64 : (t > 256 ? 256 : t));
}

// opt a kernel into > 48 KB of dynamic shared memory, once per (kernel, device): the attribute is per device and one
// process may drive several
static void allow_large_smem(const void *kernel) {
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done.insert(std::make_pair(kernel, dev)).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

template <class F>
static void launch_pass_t(int kind, const NttPass &p, cudaStream_t stream, uint64_t *launches) {
    typedef typename F::E E;
    const u32 lp = p.log_np, np = 1u << lp;
    if (kind == NTT_STRIDED) {
        u32 lmo = p.log_min - lp;
        u32 wl = tile_width_log(F::W_LOG, lp);
        u32 lw = lmo < wl ? lmo : wl;
        size_t smem = (((size_t)np << lw) + (np >> 1)) * sizeof(E);
        u64 tiles_per_col = ((u64)1 << (lmo - lw)) << (p.log_n - p.log_min);
        allow_large_smem((const void *)ntt_strided_kernel<F>);
        ntt_strided_kernel<F><<<(unsigned)(tiles_per_col * p.num_cols), block_threads((u64)np << lw), smem, stream>>>(p);
    } else if (kind == NTT_FINAL_NATURAL) {
        u32 lc = p.log_n - lp;
        u32 wl = tile_width_log(F::W_LOG, lp);
        u32 lw = lc < wl ? lc : wl;
        size_t smem = ((size_t)(np + F::PAD) * (1u << lw) + (np >> 1)) * sizeof(E);
        u64 groups = (u64)1 << (lc - lw);
        allow_large_smem((const void *)ntt_final_natural_kernel<F>);
        ntt_final_natural_kernel<F><<<(unsigned)(groups * p.num_cols), block_threads((u64)np << lw), smem, stream>>>(p);
    } else {
        u32 lc = p.log_n - lp;
        u32 wl = tile_width_log(F::W_LOG, lp);
        size_t smem = ((size_t)(np + F::PAD) * (1u << wl) + (np >> 1)) * sizeof(E);
        u64 col_groups = (p.num_cols + (1u << wl) - 1) >> wl;
        allow_large_smem((const void *)ntt_final_leaves_kernel<F>);
        ntt_final_leaves_kernel<F><<<(unsigned)(col_groups << lc), block_threads((u64)np << wl), smem, stream>>>(p);
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

template <class F>
static void build_tables_t(void *out, u32 log_min, u32 log_np, const void *wpow, const void *spow, bool scale_only,
                           u32 log_mout, cudaStream_t stream) {
    typedef typename F::E E;
    if (scale_only) {
        u32 n = 1u << log_np;
        build_scale_table_kernel<F><<<(n + 255) / 256, 256, 0, stream>>>((E *)out, log_np, log_mout, (const E *)spow);
    } else {
        u32 n = 1u << log_min;
        build_tw_table_kernel<F><<<(n + 255) / 256, 256, 0, stream>>>((E *)out, log_min, log_np, (const E *)wpow, (const E *)spow);
    }
}
void launch_build_tw_table(int field, void *out, uint32_t log_min, uint32_t log_np, const void *wpow, const void *spow,
                           cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        build_tables_t<NttGl>(out, log_min, log_np, wpow, spow, false, 0, stream);
    else
        build_tables_t<NttBb<true>>(out, log_min, log_np, wpow, spow, false, 0, stream);
    ++*launches;
}
void launch_build_scale_table(int field, void *out, uint32_t log_np, uint32_t log_mout, const void *spow,
                              cudaStream_t stream, uint64_t *launches) {
    if (field == FIELD_GL)
        build_tables_t<NttGl>(out, 0, log_np, nullptr, spow, true, log_mout, stream);
    else
        build_tables_t<NttBb<true>>(out, 0, log_np, nullptr, spow, true, log_mout, stream);
    ++*launches;
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

// data[c][j] *= base^j (base given as a three-level power table): the trailing step of coset_ifft
// (field/src/polynomial/mod.rs:62-72), where base = shift^-1
template <class F>
__global__ void scale_powers_kernel(typename F::E *data, u32 log_n, u32 num_cols, u64 col_stride, const typename F::E *__restrict__ spow) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ((u64)num_cols << log_n)) return;
    u32 j = (u32)(t & (((u64)1 << log_n) - 1));
    u64 c = t >> log_n;
    typename F::E *p = data + c * col_stride + j;
    *p = F::store_user(F::mul(F::load_user(*p), pow3<F>(spow, j)));
}

void launch_scale_powers(int field, int repr, void *data, uint32_t log_n, uint32_t num_cols, uint64_t col_stride, const void *spow,
                         cudaStream_t stream, uint64_t *launches) {
    u64 total = ((u64)num_cols) << log_n;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (field == FIELD_GL)
        scale_powers_kernel<NttGl><<<blocks, 256, 0, stream>>>((u64 *)data, log_n, num_cols, col_stride, (const u64 *)spow);
    else if (repr)
        scale_powers_kernel<NttBb<true>><<<blocks, 256, 0, stream>>>((u32 *)data, log_n, num_cols, col_stride, (const u32 *)spow);
    else
        scale_powers_kernel<NttBb<false>><<<blocks, 256, 0, stream>>>((u32 *)data, log_n, num_cols, col_stride, (const u32 *)spow);
    ++*launches;
}

}  // namespace p2g
