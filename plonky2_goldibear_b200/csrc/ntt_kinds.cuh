// The three tile kinds of a batched NTT pass: which global elements make up the tile of CTA `block`, and where the last
// round's outputs go.  Geometry only -- the rounds are in ntt_tile.cuh, the device-only fast loads in ntt.cu.
// Plain C++ (see ntt_tile.cuh): tests/cpp/ntt_tile_emul_test.cpp runs these on the host.
//
// Pass structure (per column of n = 2^L elements): n = n_1 * ... * n_P with every n_i <= 2^11.  Pass i < P ("strided")
// takes, for W consecutive offsets j0, the n_i points x[j0 + m*k] (m = chunk / n_i), transforms them, multiplies by
// the inter-pass twiddle w_chunk^(f*j0) and stores frequency f at chunk_base + f*m + j0, so that the next pass again
// works on contiguous chunks.  The last pass transforms contiguous chunks; the inverse transform writes natural order
// (fft.rs:66-94 semantics), the LDE writes the row-major, bit-reversed leaf matrix directly (fri/oracle.rs:108-118).
#pragma once
#include "ntt_tile.cuh"
#include "p2g_internal.h"

#if defined(__CUDACC__)
#define P2G_TILE_HOST_FN __host__ __device__ inline  // called by the launchers
#else
#define P2G_TILE_HOST_FN inline
#endif

namespace p2g {

// CTA index -> (batch entry, tile): batch entries of the same tile are adjacent in the grid
struct BatchSplit {
    u32 b, tile;
    P2G_TILE_FN BatchSplit(const NttPass &p, u32 block) {
        const u32 nb = p.num_batch ? p.num_batch : 1;
        tile = block / nb;
        b = block - tile * nb;
    }
};

template <class F, int DIR>
struct StridedTile {
    typedef typename F::E E;
    const E *src;  // chunk base of this tile's column (+ batch offset)
    E *dst;
    const E *tw, *scale;
    u32 lp, lw, lmo, jt;
    bool user_repr;
    u32 eps;
    TilePlan pl;

    P2G_TILE_FN StridedTile(const NttPass &p, u32 block) {
        const BatchSplit bs(p, block);
        lp = p.log_np;
        lmo = p.log_min - lp;  // log2(m_out)
        const u32 wl = tile_width_log(F::W_LOG, lp);
        lw = lmo < wl ? lmo : wl;  // tile width (runs of 2^lw elements)
        const u32 tiles_per_chunk = 1u << (lmo - lw);
        const u32 tiles_per_col = tiles_per_chunk << (p.log_n - p.log_min);
        const u32 col = bs.tile / tiles_per_col;
        const u32 t = bs.tile - col * tiles_per_col;
        const u32 q = t >> (lmo - lw);
        jt = t & (tiles_per_chunk - 1);
        src = (const E *)p.src + (size_t)bs.b * p.src_batch_stride + (size_t)col * p.src_col_stride + ((size_t)q << p.log_min);
        dst = (E *)p.dst + (size_t)bs.b * p.dst_batch_stride + (size_t)col * p.dst_col_stride + ((size_t)q << p.log_min);
        tw = (const E *)p.tw_tables[bs.b];
        scale = (const E *)p.scale_tables[bs.b];
        user_repr = p.in_user_repr != 0;
        eps = p.gl_eps;
        pl = tile_plan(lp);
    }
    static P2G_TILE_HOST_FN u32 num_tiles(const NttPass &p) {  // per batch entry
        const u32 lp = p.log_np, lmo = p.log_min - lp;
        const u32 wl = tile_width_log(F::W_LOG, lp);
        const u32 lw = lmo < wl ? lmo : wl;
        return ((1u << (lmo - lw)) << (p.log_n - p.log_min)) * p.num_cols;
    }
    P2G_TILE_FN u32 tile_elems() const { return 1u << (lp + lw); }
    // tile row i is the run of 2^lw consecutive elements at src[(i << lmo) + jt*W]
    P2G_TILE_FN void load_scalar(E *x, u32 tid, u32 nthreads) const {
        const u32 total = tile_elems(), W = 1u << lw;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            const u32 w = idx & (W - 1), i = idx >> lw;
            x[idx] = src[((size_t)i << lmo) + ((size_t)jt << lw) + w];
        }
    }
    P2G_TILE_FN void round(u32 r, E *x, const E *roots_s, u32 tid, u32 nthreads) const {
        TwCtx<F> tc;
        tc.roots = roots_s;
        tc.lp = lp;
        tc.eps = eps;
        LoadFix<F> fix = {scale, user_repr};
        StridedEpi<F> epi = {dst, tw, lp, lw, lmo, jt};
        tile_run_round<F, DIR, false>(r, pl, x, tc, lw, 0, fix, epi, tid, nthreads);
    }
};

template <class F, int DIR>
struct NaturalTile {
    typedef typename F::E E;
    const E *src;  // column base
    E *dst;
    const NttPass *pp;
    size_t peer_off;
    u32 lp, lc, lw, cs, d0, lp1, lp2;
    bool user_repr;
    u32 eps;
    TilePlan pl;

    P2G_TILE_FN NaturalTile(const NttPass &p, u32 block) {
        pp = &p;
        lp = p.log_np;
        lc = p.log_n - lp;  // log2(chunks per column)
        const u32 wl = tile_width_log(F::W_LOG, lp);
        lw = lc < wl ? lc : wl;  // chunks per tile
        cs = tile_col_stride((u32)sizeof(E), lp, lw);
        const u32 groups_per_col = 1u << (lc - lw);
        const u32 col = block / groups_per_col;
        const u32 g = block - col * groups_per_col;
        d0 = g << lw;
        src = (const E *)p.src + (size_t)col * p.src_col_stride;
        dst = (E *)p.dst + (size_t)col * p.dst_col_stride;
        peer_off = (size_t)col * p.dst_col_stride;
        lp1 = p.lp1;
        lp2 = p.lp2;
        user_repr = p.in_user_repr != 0;
        eps = p.gl_eps;
        pl = tile_plan(lp);
    }
    static P2G_TILE_HOST_FN u32 num_tiles(const NttPass &p) {
        const u32 lp = p.log_np, lc = p.log_n - lp;
        const u32 wl = tile_width_log(F::W_LOG, lp);
        const u32 lw = lc < wl ? lc : wl;
        return (1u << (lc - lw)) * p.num_cols;
    }
    P2G_TILE_FN u32 tile_elems() const { return cs << lw; }  // storage: 2^lw padded columns
    // interleaved transform w is the contiguous chunk natural_to_chunk(d0 + w) of the column
    P2G_TILE_FN const E *chunk(u32 w) const { return src + ((size_t)natural_to_chunk(d0 + w, lp1, lp2) << lp); }
    P2G_TILE_FN void load_scalar(E *x, u32 tid, u32 nthreads) const {
        const u32 total = 1u << (lp + lw), W = 1u << lw;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            const u32 w = idx & (W - 1), i = idx >> lw;
            x[(size_t)w * cs + i] = chunk(w)[i];
        }
    }
    P2G_TILE_FN void round(u32 r, E *x, const E *roots_s, u32 tid, u32 nthreads) const {
        const NttPass &p = *pp;
        TwCtx<F> tc;
        tc.roots = roots_s;
        tc.lp = lp;
        tc.eps = eps;
#pragma unroll
        for (int i = 0; i < 4; i++) tc.join_w[i] = (E)p.join_w[i];
        LoadFix<F> fix = {nullptr, user_repr};
        NaturalEpi<F> epi = {dst, (E *const *)p.peer_dst, peer_off, p.num_peers, lp, lc, d0, (E)p.final_scale,
                             p.use_final_scale != 0, p.out_user_repr != 0};
        tile_run_round<F, DIR, true>(r, pl, x, tc, lw, cs, fix, epi, tid, nthreads);
    }
};

template <class F, int DIR>
struct LeavesTile {
    typedef typename F::E E;
    const E *src;  // (col0, chunk q)
    E *rows;       // leaf (first row of the chunk, col0)
    const E *scale;
    size_t src_col_stride, pitch;
    u32 lp, lw, cs, ncols;
    bool user_repr, out_user;
    u32 eps;
    TilePlan pl;

    P2G_TILE_FN LeavesTile(const NttPass &p, u32 block) {
        const BatchSplit bs(p, block);
        lp = p.log_np;
        const u32 lc = p.log_n - lp;
        lw = tile_width_log(F::W_LOG, lp);  // columns per tile
        cs = tile_col_stride((u32)sizeof(E), lp, lw);
        const u32 W = 1u << lw;
        const u32 col_groups = (p.num_cols + W - 1) >> lw;
        const u32 q = bs.tile / col_groups;
        const u32 cg = bs.tile - q * col_groups;
        const u32 col0 = cg << lw;
        ncols = p.num_cols - col0 < W ? p.num_cols - col0 : W;
        src_col_stride = p.src_col_stride;
        src = (const E *)p.src + (size_t)bs.b * p.src_batch_stride + (size_t)col0 * p.src_col_stride + ((size_t)q << lp);
        // frequency f sits at tile row bitrev(f), and the leaf order wants row offset bitrev(f): rows are tile rows
        const u32 dq = chunk_to_natural(q, p.lp1, p.lp2);
        const u32 rq = tile_bitrev(dq, lc);
        pitch = p.dst_col_stride;
        rows = (E *)p.dst + (p.row_base + (size_t)bs.b * p.row_batch_stride + ((size_t)rq << lp)) * pitch + col0;
        scale = (const E *)p.scale_tables[bs.b];  // single-pass LDE only (q == 0): tile row = position
        user_repr = p.in_user_repr != 0;
        out_user = p.out_user_repr != 0;
        eps = p.gl_eps;
        pl = tile_plan(lp);
    }
    static P2G_TILE_HOST_FN u32 num_tiles(const NttPass &p) {
        const u32 lp = p.log_np, lc = p.log_n - lp;
        const u32 lw = tile_width_log(F::W_LOG, lp);
        return ((p.num_cols + (1u << lw) - 1) >> lw) << lc;
    }
    P2G_TILE_FN u32 tile_elems() const { return cs << lw; }  // storage: 2^lw padded columns
    // unused column slots of a ragged last tile hold zeros
    P2G_TILE_FN void load_scalar(E *x, u32 tid, u32 nthreads) const {
        const u32 total = 1u << (lp + lw), W = 1u << lw;
        for (u32 idx = tid; idx < total; idx += nthreads) {
            const u32 w = idx & (W - 1), i = idx >> lw;
            x[(size_t)w * cs + i] = w < ncols ? src[(size_t)w * src_col_stride + i] : (E)0;
        }
    }
    P2G_TILE_FN void round(u32 r, E *x, const E *roots_s, u32 tid, u32 nthreads) const {
        TwCtx<F> tc;
        tc.roots = roots_s;
        tc.lp = lp;
        tc.eps = eps;
        LoadFix<F> fix = {scale, user_repr};
        LeavesEpi<F> epi = {rows, pitch, ncols, out_user};
        tile_run_round<F, DIR, true>(r, pl, x, tc, lw, cs, fix, epi, tid, nthreads);
    }
};

}  // namespace p2g
