// Batched number-theoretic transforms for the commit path, shared-memory staged.
//
// Replaces, on the GPU:
//   ifft_with_options / fft_classic      field/src/fft.rs:66-94,168-205     (per-column radix-2 FFT on the CPU)
//   PolynomialCoeffs::lde + coset_fft    field/src/polynomial/mod.rs:201-220,282-295
//   transpose + reverse_index_bits       plonky2/src/util/mod.rs:25-31, util/src/lib.rs:189-238
//
// Structure (per column of n = 2^L elements, all columns -- and, for the LDE, several cosets -- in one launch):
//   * n = n_1 * ... * n_P with every n_i <= 2^11 (P <= 3).  Pass i < P ("strided") loads, for W consecutive offsets j0,
//     the n_i points x[j0 + m*k], transforms them in shared memory, multiplies by the inter-pass twiddle w_chunk^(f*j0)
//     and stores frequency f at chunk_base + f*m + j0.  Every global access is a run of W consecutive elements (64 B).
//   * the last pass transforms contiguous chunks.  The inverse transform writes natural order (coefficients, times
//     1/n); the LDE writes the row-major leaf matrix directly: decimation-in-frequency leaves frequency f at position
//     bitrev(f), which is exactly the row order the reference obtains with transpose + reverse_index_bits.
//   * the LDE never builds the zero-padded size-N input: rows [b*n, (b+1)*n) are the size-n transform of the
//     coefficients scaled by s_b^j, s_b = g * w_N^{bitrev_r(b)}; the scaling is fused into the first
//     pass' load.  CTAs of the same tile and different cosets sit next to each other in the grid, so the coefficients
//     cross the DRAM interface once per batch of cosets, not once per coset.
// The tile itself (layout, rounds, epilogues) is in ntt_tile.cuh / ntt_kinds.cuh / gl_limbs.cuh; this file holds the
// field policies, the device-only tile loads, the kernels and their launchers.
// Twiddles: in-tile roots live in shared memory.  Inter-pass twiddles (with the coset factor s^j0 merged in on the
// first LDE pass) come from a per-pass table indexed exactly like the store, so they are read in the same 64-byte
// runs; the remaining coset factor s^(m*i) is one small per-row table.  Both are built once per (field, size, coset)
// by a kernel from three-level 1024-entry power tables and cached in the context.
#include <mutex>
#include <set>
#include <utility>

#include "ntt_kinds.cuh"
#include "p2g_internal.h"

// Resident CTAs per SM the Goldilocks kernels are compiled for.  A 16-point unit holds 64 limb registers; at 2 CTAs
// (128 registers) nothing spills, at 3 CTAs (80 registers) ptxas spills 90-470 bytes per thread, and the extra eight
// warps still win: LDE of 2^20 x 135 20.0 against 20.8 ms (profiles/r2_ntt_variants.txt).
#ifndef P2G_NTT_GL_MIN_CTAS
#define P2G_NTT_GL_MIN_CTAS 3
#endif

namespace p2g {

struct NttGl {
    typedef u64 E;
    static const int W_LOG = 3;  // 8 x 8 B = 64-byte runs
    static const bool LIMBS = true;
    static const int MIN_CTAS = P2G_NTT_GL_MIN_CTAS;
    static const unsigned TW_PREFETCH = 0;  // inter-pass factors fetched ahead of a unit's arithmetic (ntt_tile.cuh, StridedEpi)
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
    static const bool LIMBS = false;
    static const int MIN_CTAS = 3;
    static const unsigned TW_PREFETCH = 16;
    // The transforms are linear and every twiddle / coset factor / scale is a Montgomery-form constant, so
    // mul(word, constant) = word * constant in the WORD's own representation: canonical words go through the passes as
    // canonical words and Montgomery words as Montgomery words -- no conversion on the way in or out, only the range
    // fix that lets any u32 in (two conditional subtractions).  (The hash kernels, which are not linear, do convert.)
    static __device__ __forceinline__ E load_user(E x) { return BbDev::from_monty(x); }
    static __device__ __forceinline__ E store_user(E x) { return x; }
    static __device__ __forceinline__ E mul(E a, E b) { return bb_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return bb_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return bb_sub(a, b); }
    // (a - b) * w^e for a butterfly: table entry e = (w, w' = floor(w 2^32 / p)) with w CANONICAL -- Shoup's form of a
    // product by a constant: q = floor(d w' / 2^32) is the quotient of d w by p up to one, so d w - q p lies in [0, 2p)
    // for ANY d < 2^32, which lets the difference stay unreduced (a - b + p, one three-input add).  IMAD.HI + 2 IMAD
    // (8 multiplier-pipe cycles) and one conditional subtraction, against IMAD.WIDE + IMAD + IMAD.HI (10) and two for
    // the Montgomery product of a reduced difference.  The product is exact in integers, so words keep the
    // representation they came in, as with the Montgomery-form constants of the other factors.
    static __device__ __forceinline__ E mul_root(E a, E b, const E *roots, u32 e) {
        const uint2 t = reinterpret_cast<const uint2 *>(roots)[e];
        const u32 d = a - b + P2G_BB_P;
        const u32 q = __umulhi(d, t.y);
        const u32 r = d * t.x - q * P2G_BB_P;
        return bb_min_u32(r, r - P2G_BB_P);
    }
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

// BabyBear keeps its in-tile roots ((w^e, floor(w^e 2^32 / p)), e < np/2) in shared memory behind the tile; Goldilocks reads its 32-byte join
// entries from the global table through the read-only path (tile_load_w4)
template <class F>
__host__ __device__ inline size_t roots_smem_bytes(u32 lp) {
    return F::LIMBS ? 0 : (size_t)((1u << lp) >> 1) * 2 * sizeof(typename F::E);  // (w, w') pairs, NttBb::mul_root
}
template <class F>
__device__ __forceinline__ const typename F::E *stage_roots(const NttPass &p, typename F::E *roots_s, u32 lp) {
    typedef typename F::E E;
    const E *roots = (const E *)p.roots;
    if (F::LIMBS) return roots;
    const u32 cnt = (1u << lp) >> 1;  // entries of two words
    for (u32 k = threadIdx.x; k < cnt; k += blockDim.x)
        reinterpret_cast<uint2 *>(roots_s)[k] = __ldg(reinterpret_cast<const uint2 *>(roots) + k);
    return roots_s;
}

// the rounds of a staged tile, one CTA barrier between rounds
template <class Tile, class E>
__device__ __forceinline__ void run_rounds(const Tile &t, E *x, const E *roots_s) {
    for (u32 r = 0; r < t.pl.nrounds; r++) {
        t.round(r, x, roots_s, threadIdx.x, blockDim.x);
        if (r + 1 < t.pl.nrounds) __syncthreads();
    }
}

// ---- bulk tile load of the two final kinds (TMA unit, cp.async.bulk + mbarrier) -------------------------------------------
// A final tile is 2^lw contiguous runs of np elements (chunks of a column / the same chunk of adjacent columns) and is
// kept column-major in shared memory (ntt_tile.cuh), so each run is ONE bulk copy issued by one thread: no per-element
// LDGSTS (np x 2^lw of them per tile before: 6 % of the Goldilocks and 18 % of the BabyBear final kernels' issue slots,
// shared-memory store wavefronts on top), no registers, and the copy engine of the SM completes the mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

// Starts the load of a column-major tile: column w (< live_w) = np elements at base_of(w), the other columns zero.
// bulk: every run is 16-byte aligned and a multiple of 16 bytes long (NttPass::async_ok and np * sizeof(E) % 16 == 0);
// otherwise plain loads.  tile_load_wait() (all threads) before the first round.
template <class E, class BaseOf>
__device__ __forceinline__ void tile_load_columns(E *x, unsigned long long *bar, u32 lp, u32 lw, u32 cs, u32 live_w, bool bulk,
                                                  const BaseOf &base_of) {
    const u32 np = 1u << lp, W = 1u << lw;
    if (bulk) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const u32 bytes = np * (u32)sizeof(E);
            mbar_arrive_expect_tx(bar, live_w * bytes);
            for (u32 w = 0; w < live_w; w++) bulk_load(x + (size_t)w * cs, base_of(w), bytes, bar);
        }
        for (u32 idx = threadIdx.x; idx < ((W - live_w) << lp); idx += blockDim.x)
            x[(size_t)(live_w + (idx >> lp)) * cs + (idx & (np - 1))] = (E)0;
    } else {
        for (u32 idx = threadIdx.x; idx < (W << lp); idx += blockDim.x) {
            const u32 w = idx >> lp, i = idx & (np - 1);
            x[(size_t)w * cs + i] = w < live_w ? base_of(w)[i] : (E)0;
        }
    }
}
__device__ __forceinline__ void tile_load_wait(unsigned long long *bar, bool bulk) {
    if (bulk) mbar_wait(bar, 0);
    __syncthreads();
}
template <class F>
__host__ __device__ inline bool bulk_ok(const NttPass &p) {
    return p.async_ok && (((size_t)sizeof(typename F::E) << p.log_np) % 16 == 0);
}
// the mbarrier sits behind the tile and the staged roots (both multiples of 8 bytes)
template <class F>
__host__ __device__ inline size_t final_smem_bytes(u32 lp, u32 lw) {
    return ((size_t)tile_col_stride((u32)sizeof(typename F::E), lp, lw) << lw) * sizeof(typename F::E) + roots_smem_bytes<F>(lp) + 8;
}

template <class F, int DIR>
__global__ void __launch_bounds__(256, F::MIN_CTAS) ntt_strided_kernel(const __grid_constant__ NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const StridedTile<F, DIR> t(p, blockIdx.x);
    const u32 np = 1u << t.lp;
    E *roots_s = x + t.tile_elems();
    // stage the raw tile: row i is a run of W consecutive elements at src[(i << lmo) + jt*W]
    const u32 run_bytes = (u32)sizeof(E) << t.lw;
    if (p.async_ok && run_bytes >= 16) {
        // 16-byte pieces; a thread keeps its piece of the run and walks down the rows (block size is a multiple of the
        // pieces per row), so the loop is a copy and two pointer increments
        const u32 lpr = 31 - __clz(run_bytes >> 4);  // log2(pieces per row)
        const u32 part = threadIdx.x & ((1u << lpr) - 1), rows_step = blockDim.x >> lpr;
        const char *g = (const char *)(t.src + ((size_t)t.jt << t.lw)) + part * 16 + (((size_t)(threadIdx.x >> lpr) << t.lmo) * sizeof(E));
        char *sm = (char *)x + (size_t)(threadIdx.x >> lpr) * run_bytes + part * 16;
        const size_t g_step = ((size_t)rows_step << t.lmo) * sizeof(E);
        const u32 s_step = rows_step * run_bytes;
#pragma unroll 4
        for (u32 i = threadIdx.x >> lpr; i < np; i += rows_step) {
            cp_async16(sm, g);
            g += g_step;
            sm += s_step;
        }
    } else {
        t.load_scalar(x, threadIdx.x, blockDim.x);
    }
    const E *roots = stage_roots<F>(p, roots_s, t.lp);
    cp_async_wait_all();
    __syncthreads();
    run_rounds(t, x, roots);
}

template <class F, int DIR>
__global__ void __launch_bounds__(256, F::MIN_CTAS) ntt_final_natural_kernel(const __grid_constant__ NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const NaturalTile<F, DIR> t(p, blockIdx.x);
    E *roots_s = x + t.tile_elems();
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + final_smem_bytes<F>(t.lp, t.lw) - 8);
    const bool bulk = bulk_ok<F>(p);
    tile_load_columns<E>(x, bar, t.lp, t.lw, t.cs, 1u << t.lw, bulk, [&](u32 w) { return t.chunk(w); });
    const E *roots = stage_roots<F>(p, roots_s, t.lp);
    tile_load_wait(bar, bulk);
    run_rounds(t, x, roots);
}

template <class F, int DIR>
__global__ void __launch_bounds__(256, F::MIN_CTAS) ntt_final_leaves_kernel(const __grid_constant__ NttPass p) {
    typedef typename F::E E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *x = reinterpret_cast<E *>(smem_raw);
    const LeavesTile<F, DIR> t(p, blockIdx.x);
    E *roots_s = x + t.tile_elems();
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + final_smem_bytes<F>(t.lp, t.lw) - 8);
    const bool bulk = bulk_ok<F>(p);
    tile_load_columns<E>(x, bar, t.lp, t.lw, t.cs, t.ncols, bulk, [&](u32 w) { return t.src + (size_t)w * t.src_col_stride; });
    const E *roots = stage_roots<F>(p, roots_s, t.lp);
    tile_load_wait(bar, bulk);
    run_rounds(t, x, roots);
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
    const u32 nb = p.num_batch ? p.num_batch : 1;
    const size_t roots_bytes = roots_smem_bytes<F>(lp);
    if (kind == NTT_STRIDED) {
        u32 lmo = p.log_min - lp;
        u32 wl = tile_width_log(F::W_LOG, lp);
        u32 lw = lmo < wl ? lmo : wl;
        size_t smem = ((size_t)np << lw) * sizeof(E) + roots_bytes;
        u64 tiles = (u64)StridedTile<F, 0>::num_tiles(p) * nb;
        if (p.inverse) {
            allow_large_smem((const void *)ntt_strided_kernel<F, 1>);
            ntt_strided_kernel<F, 1><<<(unsigned)tiles, block_threads((u64)np << lw), smem, stream>>>(p);
        } else {
            allow_large_smem((const void *)ntt_strided_kernel<F, 0>);
            ntt_strided_kernel<F, 0><<<(unsigned)tiles, block_threads((u64)np << lw), smem, stream>>>(p);
        }
    } else if (kind == NTT_FINAL_NATURAL) {
        u32 lc = p.log_n - lp;
        u32 wl = tile_width_log(F::W_LOG, lp);
        u32 lw = lc < wl ? lc : wl;
        size_t smem = final_smem_bytes<F>(lp, lw);
        u64 tiles = NaturalTile<F, 1>::num_tiles(p);
        // natural order is the inverse transform's output, the leaf matrix the forward transform's: one direction each
        allow_large_smem((const void *)ntt_final_natural_kernel<F, 1>);
        ntt_final_natural_kernel<F, 1><<<(unsigned)tiles, block_threads((u64)np << lw), smem, stream>>>(p);
    } else {
        u32 wl = tile_width_log(F::W_LOG, lp);
        size_t smem = final_smem_bytes<F>(lp, wl);
        u64 tiles = (u64)LeavesTile<F, 0>::num_tiles(p) * nb;
        allow_large_smem((const void *)ntt_final_leaves_kernel<F, 0>);
        ntt_final_leaves_kernel<F, 0><<<(unsigned)tiles, block_threads((u64)np << wl), smem, stream>>>(p);
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

// the finished coefficient block to every peer: each thread moves 4 x 16 bytes per trip, a warp instruction stores 512
// contiguous bytes into one peer's buffer (plain stores to peer-mapped memory, NVLink / NVSwitch)
struct PeerList {
    uint4 *p[8];
    unsigned n;
};
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4 *__restrict__ src, PeerList peers, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
        for (unsigned k = 0; k < peers.n; k++) {
            uint4 *q = peers.p[k];
            q[i] = a;
            q[i + stride] = b;
            q[i + 2 * stride] = c;
            q[i + 3 * stride] = d;
        }
    }
    for (; i < n16; i += stride) {
        const uint4 a = __ldg(src + i);
        for (unsigned k = 0; k < peers.n; k++) peers.p[k][i] = a;
    }
}
void launch_peer_push(const void *src, void *const *peers, unsigned num_peers, size_t bytes, cudaStream_t stream, uint64_t *launches) {
    PeerList pl;
    pl.n = num_peers < 8 ? num_peers : 8;
    for (unsigned k = 0; k < pl.n; k++) pl.p[k] = (uint4 *)peers[k];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 255) / 256;
    if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
    if (!blocks) return;
    peer_push_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint4 *)src, pl, n16);
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
