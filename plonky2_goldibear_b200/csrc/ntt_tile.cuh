// One shared-memory tile of a batched NTT pass: the decimation-in-frequency rounds and the three tile kinds
// (strided pass, final pass in natural order, final pass into the leaf matrix).
//
// Replaces, on the GPU, the butterfly loops of fft_classic (field/src/fft.rs:168-205) together with the
// transpose + reverse_index_bits step of from_coeffs (fri/oracle.rs:108-118); see ntt.cu for the pass structure.
//
// Tile layout: i = point index inside the tile's transform (0 .. np-1), w = which of the 2^lw interleaved transforms
// (columns / offsets / chunks); consecutive threads take consecutive w in every kind.
//   * strided kind, row-major: element (i, w) at x[(i << lw) + w].  A row i is one 64-byte run of global memory, every
//     shared-memory access of a warp is a set of whole rows: conflict-free in every round but the last one (two rows of
//     equal parity per half-warp there: 2-way).
//   * final kinds, column-major with a padded column stride: x[w * cs + i], cs = np + pad (tile_col_stride).  A column is
//     one contiguous run of global memory and arrives as one bulk copy; the pad puts the columns a bank group apart, so
//     a warp (8 w x 2 or 4 consecutive i) is conflict-free exactly where the row-major form is.  (The round-1 kernels
//     were column-major WITHOUT the pad and with lanes along i: 8-way conflicts in two of three rounds.)
//
// A round = a thread takes the 2^K points that differ in the K position bits of the round ("unit"), transforms them in
// registers and puts them back; the LAST round hands its outputs to the tile kind's epilogue, which stores them
// straight to global memory (no separate store loop: a unit's outputs are whole 64-byte runs already).
//   * Goldilocks: a unit is a multiplication-free 2^K-point DFT on 24-bit limbs (gl_limbs.cuh) and the round's twiddle
//     w_L^(low k) rides in the join;
//   * BabyBear: K radix-2 stages; the product (a - b) * w with a precomputed quotient (Shoup form, F::mul_root).
//
// The code in this header is plain C++ so that tests/cpp/ntt_tile_emul_test.cpp can run every tile kind on the host,
// thread by thread, against a naive transform; the asynchronous / vectorised loads are the only device-only parts.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "gl_limbs.cuh"

#if defined(__CUDACC__)
#define P2G_TILE_FN __device__ __forceinline__
#else
#define P2G_TILE_FN inline
#endif

namespace p2g {

typedef unsigned long long u64;
typedef unsigned int u32;

P2G_TILE_FN u32 tile_brev32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
P2G_TILE_FN u32 tile_bitrev(u32 x, u32 bits) { return bits ? tile_brev32(x) >> (32 - bits) : 0; }

// rounds of a 2^lp-point tile
struct TilePlan {
    u32 nrounds;
    u32 k0, k1, k2;  // stages of round 0, 1, 2 (separate fields: an indexed array would live in local memory)
    P2G_TILE_FN u32 K(u32 r) const { return r == 0 ? k0 : (r == 1 ? k1 : k2); }
    P2G_TILE_FN u32 s0(u32 r) const { return r == 0 ? 0 : (r == 1 ? k0 : k0 + k1); }  // stages done before round r
};
// ceil(lp / 4) rounds, as even as possible, larger rounds first: 10 -> 4,3,3; 11 -> 4,4,3; 9 -> 3,3,3; 5 -> 3,2.
// (A round's fixed costs -- pick the tile up, split / join or twiddle, put it back -- are per element, so a 2-stage
// round costs almost as much as a 4-stage one: the round-1 plan 4,4,2 spent 31% of a pass in its last two stages.)
P2G_TILE_FN TilePlan tile_plan(u32 lp) {
    TilePlan pl;
    const u32 nr = lp <= 4 ? 1 : (lp <= 8 ? 2 : 3);
    pl.k0 = (lp + nr - 1) / nr;
    pl.k1 = nr >= 2 ? (lp - pl.k0 + nr - 2) / (nr - 1) : 0;
    pl.k2 = lp - pl.k0 - pl.k1;
    // a one-point transform (lp = 0) still owes the load conversion and the epilogue: one round of zero stages
    pl.nrounds = nr;
    return pl;
}

// what the first round does to an element as it picks it up from the staged tile
template <class F>
struct LoadFix {
    typedef typename F::E E;
    const E *scale;  // per tile row factor (coset scaling), or null
    bool user_repr;  // staged data is in the caller's representation
    // the R rows base + (j << shift) of one unit: the factors are fetched a batch ahead of the multiplications
    template <u32 R>
    P2G_TILE_FN void unit(E (&v)[R], u32 base, u32 shift) const {
        if (user_repr) {
#pragma unroll
            for (u32 j = 0; j < R; j++) v[j] = F::load_user(v[j]);
        }
        if (scale) {
            const E *sp = scale + base;
            E f[R];
#pragma unroll
            for (u32 j = 0; j < R; j++) f[j] = sp[(size_t)j << shift];
#pragma unroll
            for (u32 j = 0; j < R; j++) v[j] = F::mul(v[j], f[j]);
        }
    }
};

template <class E>
P2G_TILE_FN E tile_ldg(const E *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// 32-byte twiddle entry (Goldilocks): global memory, read-only path, two 16-byte loads.  (Staging the table in shared
// memory behind the tile was measured slower: 23.5 against 20.8 ms for the LDE of 2^20 x 135 -- 32 KB more per CTA to
// copy, and the copy sits on the critical path of every tile.  Cache hints -- L1::no_allocate on the read-once
// inter-pass factors, L1::evict_last on these entries -- change nothing: 19.74-19.82 ms with and without.)
P2G_TILE_FN void tile_load_w4(const u64 *p, u64 (&W)[4]) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(p) + 1);
    W[0] = a.x;
    W[1] = a.y;
    W[2] = b.x;
    W[3] = b.y;
#else
    W[0] = p[0];
    W[1] = p[1];
    W[2] = p[2];
    W[3] = p[3];
#endif
}

// twiddle context of a tile.  roots: Goldilocks [np][4] in GLOBAL memory, entry e = w^e * 2^(24 i), i < 4 (the W_i of
// the limb join, gl_limbs.cuh); BabyBear (w^e, floor(w^e 2^32 / p)) for e < np/2 in shared memory (F::mul_root).
template <class F>
struct TwCtx {
    typedef typename F::E E;
    const E *roots;
    u32 lp;
    E join_w[4];  // Goldilocks: W_i of the last round's join when the epilogue is SCALED (n^-1 * 2^(24 i))
    u32 eps;      // Goldilocks: 2^32 - 1 as a run-time value (gl_limbs.cuh, fold96)
};



// ---- unit transforms ---------------------------------------------------------------------------------------------------
// in: v[j] = point p0 + (j << lowbits) of the current block (block length 2^(lp - s0)); out: v[bitrev_K(k)] =
// frequency digit k, already multiplied by this round's twiddle w_L^(low k)
template <class F, bool LIMBS>
struct UnitTransform;

// what a Goldilocks unit does with a finished output (slot s = bitrev_K(k), limbs): join it with the round's twiddle and
// put it where the round wants it
template <class F, int K, bool LAST, bool SCALED>
struct GlEmit {
    const TwCtx<F> &tc;
    u64 *out;      // LAST: the register array handed to the epilogue; else the tile, at the unit's base
    u32 stride;    // tile stride between the unit's points (not LAST)
    u32 d;         // low << s0: output digit k takes the twiddle w_np^(k * d), k * d < np
    P2G_TILE_FN void operator()(int s, const int (&l)[4]) const {
        if (LAST) {
            out[s] = SCALED ? gll::join(l, tc.join_w[0], tc.join_w[1], tc.join_w[2], tc.join_w[3], tc.eps) : gll::join_one(l, tc.eps);
        } else {
            const u32 k = gll::bitrev_small((u32)s, K);
            u64 r;
            if (k == 0) {
                r = gll::join_one(l, tc.eps);
            } else {
                // table entry e = the four W_i = w_np^e * 2^(24 i) of the join, 32 bytes, read-only path
                u64 W[4];
                tile_load_w4(tc.roots + (size_t)(k * d) * 4, W);
                r = gll::join(l, W[0], W[1], W[2], W[3], tc.eps);
            }
            out[(size_t)s * stride] = r;
        }
    }
};

template <class F>
struct UnitTransform<F, true> {  // Goldilocks
    // v: the unit's points; LAST: outputs come back in v.  Not LAST: outputs are stored to the tile (store_base, stride)
    template <int K, int DIR, bool LAST, bool SCALED>
    static P2G_TILE_FN void run(u64 (&v)[1 << K], const TwCtx<F> &tc, u32 low, u32 lowbits, u32 s0, u64 *store_base, u32 stride) {
        (void)lowbits;
        constexpr int R = 1 << K;
        int l[R][4];
#pragma unroll
        for (int j = 0; j < R; j++) gll::split(v[j], l[j]);
        gll::add_bias(l[0]);  // reaches every output (gl_limbs.cuh)
        GlEmit<F, K, LAST, SCALED> emit = {tc, LAST ? &v[0] : store_base, stride, low << s0};
        gll::dft_emit<K, DIR, R>(l, 0, emit);
    }
};

template <class F>
struct UnitTransform<F, false> {  // BabyBear: K radix-2 stages
    template <int K, int DIR, bool LAST, bool SCALED>
    static P2G_TILE_FN void run(typename F::E (&v)[1 << K], const TwCtx<F> &tc, u32 low, u32 lowbits, u32 s0,
                                typename F::E *store_base, u32 stride) {
        typedef typename F::E E;
        constexpr u32 R = 1u << K;
        const E *roots = tc.roots;
#pragma unroll
        for (int q = 0; q < K; q++) {
            const u32 hj = R >> (q + 1);
            const u32 s = s0 + q;
#pragma unroll
            for (u32 j = 0; j < R; j++) {
                if (j & hj) continue;
                // twiddle w_np^((position mod half) << s), position = (j mod hj) << lowbits | low
                const u32 e = (((j & (hj - 1)) << lowbits) | low) << s;
                const E a = v[j], b = v[j + hj];
                v[j] = F::add(a, b);
                v[j + hj] = F::mul_root(a, b, roots, e);  // (a - b) * w_np^e
            }
        }
        if (!LAST) {
#pragma unroll
            for (u32 j = 0; j < R; j++) store_base[(size_t)j * stride] = v[j];
        }
    }
};

// cs = 0: row-major tile, element (i, w) at x[(i << lw) + w] (strided kind); cs > 0: column-major, x[w * cs + i] (the
// final kinds, p2g_internal.h tile_col_stride).  The thread -> (w, unit) map is the same for both.
template <class F, int K, int DIR, bool FIRST, bool LAST, bool COLMAJ, class Epi>
P2G_TILE_FN void tile_round(typename F::E *x, const TwCtx<F> &tc, u32 lw, u32 cs, u32 s0, const LoadFix<F> &fix, const Epi &epi,
                            u32 tid, u32 nthreads) {
    typedef typename F::E E;
    constexpr u32 R = 1u << K;
    const u32 lp = tc.lp;
    const u32 lowbits = lp - s0 - K;        // position bits below the round's K bits
    const u32 units = 1u << (lp - K + lw);  // (hi, low, w) combinations
    for (u32 u = tid; u < units; u += nthreads) {
        const u32 w = u & ((1u << lw) - 1), t = u >> lw;
        const u32 low = t & ((1u << lowbits) - 1), hi = t >> lowbits;
        const u32 p0 = (hi << (lp - s0)) | low;
        E *base = COLMAJ ? x + (size_t)w * cs + p0 : x + ((size_t)p0 << lw) + w;
        const u32 stride = COLMAJ ? 1u << lowbits : 1u << (lowbits + lw);
        E v[R];
        typename Epi::template Pre<LAST ? K : 0> pre;
        if (LAST) epi.template prefetch<K>(p0, w, pre);
#pragma unroll
        for (u32 j = 0; j < R; j++) v[j] = base[(size_t)j * stride];
        if (FIRST) fix.template unit<R>(v, p0, lowbits);
        UnitTransform<F, F::LIMBS>::template run<K, DIR, LAST, Epi::SCALED>(v, tc, low, lowbits, s0, base, stride);
        if (LAST) epi.template unit<K>(p0, w, v, pre);
    }
}

// round r of the tile's plan (the caller separates rounds by a CTA barrier)
template <class F, int DIR, bool COLMAJ, class Epi>
P2G_TILE_FN void tile_run_round(u32 r, const TilePlan &pl, typename F::E *x, const TwCtx<F> &tc, u32 lw, u32 cs,
                                const LoadFix<F> &fix, const Epi &epi, u32 tid, u32 nthreads) {
    const bool first = r == 0, last = r + 1 == pl.nrounds;
    const u32 s0 = pl.s0(r), K = pl.K(r);
#define P2G_ROUND(KK, FI, LA) tile_round<F, KK, DIR, FI, LA, COLMAJ, Epi>(x, tc, lw, cs, s0, fix, epi, tid, nthreads)
    if (first && last) {
        if (K == 4) P2G_ROUND(4, true, true);
        else if (K == 3) P2G_ROUND(3, true, true);
        else if (K == 2) P2G_ROUND(2, true, true);
        else if (K == 1) P2G_ROUND(1, true, true);
        else P2G_ROUND(0, true, true);
    } else if (first) {
        if (K == 4) P2G_ROUND(4, true, false);
        else P2G_ROUND(3, true, false);
    } else if (last) {
        if (K == 4) P2G_ROUND(4, false, true);
        else if (K == 3) P2G_ROUND(3, false, true);
        else P2G_ROUND(2, false, true);
    } else {
        if (K == 4) P2G_ROUND(4, false, false);
        else P2G_ROUND(3, false, false);
    }
#undef P2G_ROUND
}

// ---- epilogues: where the last round's outputs go ----------------------------------------------------------------------
// unit<K>(row0, w, v): v[s] is tile row row0 + s (row0 a multiple of 2^K) of interleaved transform w; row pp holds
// frequency bitrev_lp(pp).

// strided pass: frequency f of offset j0 = jt*W + w goes to chunk + (f << lmo) + j0, times the inter-pass factor (the
// factor table is indexed like the store, so it is read in the same 64-byte runs, a batch ahead of the multiplications)
template <class F>
struct StridedEpi {
    typedef typename F::E E;
    static const bool SCALED = false;
    E *dst;
    const E *tw;
    u32 lp, lw, lmo, jt;
    // F::TW_PREFETCH of a unit's factors are fetched before its arithmetic, the rest at the store: registers against
    // latency.  Goldilocks 0 (a 16-point unit already holds 64 limb registers; fetching all 16 factors ahead costs 288
    // bytes of spills per thread at 3 CTAs per SM: LDE of 2^20 x 135 17.70 ms, 8 ahead 17.64, none 17.40); BabyBear 16
    // (2^22 x 256: 63.3 ms against 63.8 for 8 or none).
    static const u32 NPRE = F::TW_PREFETCH;
    template <int K>
    struct Pre {
        E t[(1u << K) < NPRE ? (1u << K) : (NPRE ? NPRE : 1u)];
    };
    template <int K>
    P2G_TILE_FN size_t pos(u32 row0, u32 w, u32 s) const {
        const u32 f0 = tile_bitrev(row0, lp);  // < 2^(lp - K)
        return ((size_t)f0 << lmo) + ((size_t)jt << lw) + w + ((size_t)gll::bitrev_small(s, K) << (lp - K + lmo));
    }
    template <int K, int KP>
    P2G_TILE_FN void prefetch(u32 row0, u32 w, Pre<KP> &pre) const {
#pragma unroll
        for (u32 s = 0; s < (1u << K) && s < (1u << KP) && (int)s < (int)NPRE; s++) pre.t[s] = tile_ldg(tw + pos<K>(row0, w, s));
    }
    template <int K, int KP>
    P2G_TILE_FN void unit(u32 row0, u32 w, E (&v)[1 << K], const Pre<KP> &pre) const {
#pragma unroll
        for (u32 s = 0; s < (1u << K); s++) {
            const size_t o = pos<K>(row0, w, s);
            dst[o] = F::mul(v[s], (int)s < (int)NPRE ? pre.t[s] : tile_ldg(tw + o));  // internal repr
        }
    }
};

// final pass of the inverse transform: natural order, scaled by 1/n (Goldilocks: inside the join), the caller's
// representation; optionally the same run also goes to the gather buffers of the other GPUs (fused all-gather)
template <class F>
struct NaturalEpi {
    typedef typename F::E E;
    static const bool SCALED = F::LIMBS;
    E *dst;
    E *const *peers;  // num_peers gather buffers of other GPUs (same layout as dst's buffer) ...
    size_t peer_off;  // ... and this column's offset inside them
    u32 num_peers;
    u32 lp, lc, d0;
    E fs;
    bool use_fs, out_user;
    template <int K>
    struct Pre {};
    template <int K, int KP>
    P2G_TILE_FN void prefetch(u32, u32, Pre<KP> &) const {}
    template <int K, int KP>
    P2G_TILE_FN void unit(u32 row0, u32 w, E (&v)[1 << K], const Pre<KP> &) const {
        constexpr u32 R = 1u << K;
        const u32 f0 = tile_bitrev(row0, lp);
        const size_t o0 = (size_t)(d0 + w) + ((size_t)f0 << lc);
        const u32 sh = lp - K + lc;
#pragma unroll
        for (u32 s = 0; s < R; s++) {
            E val = v[s];
            if (use_fs) val = F::mul(val, fs);
            if (out_user) val = F::store_user(val);
            v[s] = val;
            dst[o0 + ((size_t)gll::bitrev_small(s, K) << sh)] = val;
        }
        for (u32 k = 0; k < num_peers; k++) {
            E *pd = peers[k] + peer_off;
#pragma unroll
            for (u32 s = 0; s < R; s++) pd[o0 + ((size_t)gll::bitrev_small(s, K) << sh)] = v[s];
        }
    }
};

// final pass of the LDE: tile row pp IS leaf row (row offset of the chunk) + pp, w is the column
template <class F>
struct LeavesEpi {
    typedef typename F::E E;
    static const bool SCALED = false;
    E *rows;  // leaf (row offset of the chunk, col0)
    size_t pitch;
    u32 ncols;
    bool out_user;
    template <int K>
    struct Pre {};
    template <int K, int KP>
    P2G_TILE_FN void prefetch(u32, u32, Pre<KP> &) const {}
    template <int K, int KP>
    P2G_TILE_FN void unit(u32 row0, u32 w, E (&v)[1 << K], const Pre<KP> &) const {
        constexpr u32 R = 1u << K;
        if (w >= ncols) return;
        E *out = rows + (size_t)row0 * pitch + w;
#pragma unroll
        for (u32 s = 0; s < R; s++) out[(size_t)s * pitch] = out_user ? F::store_user(v[s]) : v[s];
    }
};

// digit reversal of the chunk index produced by the earlier passes: q = f1 * 2^lp2 + f2  ->  f1 + 2^lp1 * f2
P2G_TILE_FN u32 chunk_to_natural(u32 q, u32 lp1, u32 lp2) {
    u32 f2 = q & ((1u << lp2) - 1), f1 = q >> lp2;
    return f1 + (f2 << lp1);
}
P2G_TILE_FN u32 natural_to_chunk(u32 d, u32 lp1, u32 lp2) {
    u32 f1 = d & ((1u << lp1) - 1), f2 = d >> lp1;
    return (f1 << lp2) + f2;
}

}  // namespace p2g
