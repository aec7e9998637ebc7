// Host emulation of the NTT tile code (plonky2_goldibear_b200/csrc/{gl_limbs,ntt_tile,ntt_kinds}.cuh): the very
// headers the CUDA kernels are built from are compiled here with g++ and run "thread by thread" (a CTA barrier =
// the end of a loop over the thread index), pass by pass, against a naive O(n^2) transform.  It pins, without a GPU:
// the limb arithmetic of the Goldilocks butterflies, the twiddle indexing of every round plan (tile sizes 2^0..2^11),
// the tile geometry of the strided / natural / leaves kinds, coset batching and the bit-reversed leaf order
// (row k of coset b = P(s_b * w_n^bitrev(k)), fri/oracle.rs:108-118 with fri/verifier.rs:198-199).
// Test infrastructure only; build: g++ -O2 -std=c++17 -I/usr/local/cuda/include (tests/test_cpp_mirror.py runs it).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../plonky2_goldibear_b200/csrc/ntt_kinds.cuh"

using namespace p2g;
typedef unsigned __int128 u128;

static const u64 GL_P = 0xFFFFFFFF00000001ull;
static const u32 BB_P = 2013265921u;

struct GlH {
    typedef u64 E;
    static const int W_LOG = 3;
    static const bool LIMBS = true;
    static const unsigned TW_PREFETCH = 0;
    static E load_user(E x) { return x; }
    static E store_user(E x) { return x >= GL_P ? x - GL_P : x; }
    static E mul(E a, E b) { return (u64)((u128)a * b % GL_P); }
    static E add(E a, E b) { return (u64)(((u128)a + b) % GL_P); }
    static E sub(E a, E b) { return (u64)(((u128)a + GL_P - b % GL_P) % GL_P); }
    // host-side table arithmetic on canonical values
    static u64 P() { return GL_P; }
    static u64 internal(u64 c) { return c; }
    static u64 canonical(u64 i) { return i % GL_P; }
    static u64 gen() { return 7; }
    static u64 root(unsigned bits) {
        u64 g = 1753635133440165772ull;
        for (unsigned i = bits; i < 32; i++) g = mul(g, g);
        return g;
    }
};
struct BbH {
    typedef u32 E;
    static const int W_LOG = 4;
    static const bool LIMBS = false;
    static const unsigned TW_PREFETCH = 16;
    static u32 redc(u64 x) {
        u32 m = (u32)x * 0x77ffffffu;
        u64 t = x + (u64)m * BB_P;
        u32 r = (u32)(t >> 32);
        return r >= BB_P ? r - BB_P : r;
    }
    static E load_user(E x) {
        while (x >= BB_P) x -= BB_P;
        return x;
    }
    static E store_user(E x) { return x; }
    static E mul(E a, E b) { return redc((u64)a * b); }
    static E add(E a, E b) { return (u32)(((u64)a + b) % BB_P); }
    static E sub(E a, E b) { return (u32)(((u64)a + BB_P - b) % BB_P); }
    // the device policy's Shoup-form butterfly product (csrc/ntt.cu, NttBb::mul_root), same 32-bit arithmetic
    static E mul_root(E a, E b, const E *roots, u32 e) {
        const u32 w = roots[2 * e], wq = roots[2 * e + 1];
        const u32 d = a - b + BB_P;
        const u32 q = (u32)(((u64)d * wq) >> 32);
        const u32 r = d * w - q * BB_P;
        if (r >= 2 * BB_P) {
            std::printf("FAIL: Shoup product out of range\n");
            std::exit(1);
        }
        return r >= BB_P ? r - BB_P : r;
    }
    static u64 P() { return BB_P; }
    static u64 internal(u64 c) { return c * 268435454ull % BB_P; }  // Montgomery form
    static u64 canonical(u64 i) { return redc(i); }
    static u64 gen() { return 31; }
    static u64 root(unsigned bits) {
        u64 g = 0x1a427a41ull;
        for (unsigned i = bits; i < 27; i++) g = g * g % BB_P;
        return g;
    }
};

template <class F>
static u64 fmul(u64 a, u64 b) { return (u64)((u128)a * b % F::P()); }
template <class F>
static u64 fpow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = fmul<F>(r, a);
        a = fmul<F>(a, a);
        e >>= 1;
    }
    return r;
}
static unsigned brev(unsigned x, unsigned bits) {
    unsigned r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <class F, class Tile>
static void emulate(const NttPass &p, u32 tiles, u32 nthreads) {
    typedef typename F::E E;
    for (u32 block = 0; block < tiles; block++) {
        Tile t(p, block);
        std::vector<E> x(t.tile_elems());
        for (u32 tid = 0; tid < nthreads; tid++) t.load_scalar(x.data(), tid, nthreads);
        for (u32 r = 0; r < t.pl.nrounds; r++)
            for (u32 tid = 0; tid < nthreads; tid++) t.round(r, x.data(), (const E *)p.roots, tid, nthreads);
    }
}

// one transform over all columns with the pass plan lps[0..P): mirrors run_transform_batch (csrc/p2g_api.cu)
template <class F>
static int run_case(unsigned log_n, const std::vector<unsigned> &lps, unsigned num_cols, bool inverse, unsigned nb, u32 nthreads,
                    unsigned pitch_extra) {
    typedef typename F::E E;
    const u64 n = (u64)1 << log_n;
    const unsigned P = (unsigned)lps.size();
    const unsigned rate_bits = 3;
    std::vector<E> in((size_t)num_cols * n), work;
    u64 seed = 0x9E3779B97F4A7C15ull * (log_n + 3 * num_cols + 7 * nb + (inverse ? 1 : 0));
    for (auto &v : in) {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        v = (E)(seed >> 11);
        if (sizeof(E) == 4) v = (E)(v % BB_P);  // Montgomery words of arbitrary values
    }
    if (sizeof(E) == 8 && in.size() > 1) {
        in[0] = (E)~(E)0;  // non-canonical Goldilocks words are accepted
        in[1] = (E)GL_P;
    }
    work = in;
    const u64 pitch = num_cols + pitch_extra;
    std::vector<E> tmp((size_t)nb * num_cols * n), out_cols((size_t)num_cols * n), leaves((size_t)nb * n * pitch, (E)0);
    std::vector<u64> shifts(nb);
    const u64 wN = F::root(log_n + rate_bits);
    for (unsigned b = 0; b < nb; b++) shifts[b] = fmul<F>(F::gen(), fpow<F>(wN, brev(b, rate_bits)));

    std::vector<std::vector<E>> keep;  // tables
    unsigned log_min = log_n;
    const void *cur_src = work.data();
    u64 cur_src_bs = 0;
    void *stage = inverse ? (void *)work.data() : (void *)tmp.data();
    for (unsigned i = 0; i < P; i++) {
        NttPass p;
        memset(&p, 0, sizeof p);
        p.num_cols = num_cols;
        p.log_n = log_n;
        p.log_min = log_min;
        p.log_np = lps[i];
        p.inverse = inverse;
        p.gl_eps = 0xffffffffu;
        p.num_batch = nb;
        const unsigned lp = lps[i];
        const u64 np = (u64)1 << lp;
        {
            u64 w = F::root(lp);
            if (inverse) w = fpow<F>(w, F::P() - 2);
            size_t cnt = F::LIMBS ? 4 * np : 2 * (np > 1 ? np / 2 : 1);
            keep.emplace_back(cnt);
            u64 cur = 1;
            for (size_t k = 0; k < (F::LIMBS ? np : cnt / 2); k++) {
                if (F::LIMBS) {
                    for (int i = 0; i < 4; i++) keep.back()[4 * k + i] = (E)fmul<F>(cur, fpow<F>(2, 24 * i));
                } else {  // (w, floor(w 2^32 / p)), w canonical
                    keep.back()[2 * k] = (E)cur;
                    keep.back()[2 * k + 1] = (E)((cur << 32) / F::P());
                }
                cur = fmul<F>(cur, w);
            }
            p.roots = keep.back().data();
        }
        p.src = cur_src;
        p.src_col_stride = n;
        p.src_batch_stride = cur_src_bs;
        p.in_user_repr = i == 0;
        const unsigned lmo = log_min - lp;
        if (i == 0 && !inverse)
            for (unsigned b = 0; b < nb; b++) {
                keep.emplace_back(np);
                for (u64 r = 0; r < np; r++) keep.back()[r] = (E)F::internal(fpow<F>(shifts[b], r << lmo));
                p.scale_tables[b] = keep.back().data();
            }
        if (i + 1 < P) {
            u64 w = F::root(log_min);
            if (inverse) w = fpow<F>(w, F::P() - 2);
            for (unsigned b = 0; b < nb; b++) {
                keep.emplace_back((size_t)1 << log_min);
                for (u64 f = 0; f < np; f++)
                    for (u64 j0 = 0; j0 < ((u64)1 << lmo); j0++) {
                        u64 v = fpow<F>(w, f * j0);
                        if (i == 0 && !inverse) v = fmul<F>(v, fpow<F>(shifts[b], j0));
                        keep.back()[(f << lmo) + j0] = (E)F::internal(v);
                    }
                p.tw_tables[b] = keep.back().data();
            }
            p.dst = stage;
            p.dst_col_stride = n;
            p.dst_batch_stride = (u64)num_cols * n;
            if (inverse)
                emulate<F, StridedTile<F, 1>>(p, StridedTile<F, 1>::num_tiles(p) * nb, nthreads);
            else
                emulate<F, StridedTile<F, 0>>(p, StridedTile<F, 0>::num_tiles(p) * nb, nthreads);
            cur_src = stage;
            cur_src_bs = (u64)num_cols * n;
            log_min -= lp;
        } else {
            p.lp1 = P >= 2 ? lps[0] : 0;
            p.lp2 = P >= 3 ? lps[1] : 0;
            p.out_user_repr = 1;
            if (inverse) {
                p.dst = out_cols.data();
                p.dst_col_stride = n;
                const u64 n_inv = fpow<F>(fpow<F>(2, F::P() - 2), log_n);
                if (F::LIMBS)
                    for (int k = 0; k < 4; k++) p.join_w[k] = fmul<F>(n_inv, fpow<F>(2, 24 * k));
                else {
                    p.use_final_scale = 1;
                    p.final_scale = F::internal(n_inv);
                }
                emulate<F, NaturalTile<F, 1>>(p, NaturalTile<F, 1>::num_tiles(p), nthreads);
            } else {
                p.dst = leaves.data();
                p.dst_col_stride = pitch;
                p.row_base = 0;
                p.row_batch_stride = n;
                emulate<F, LeavesTile<F, 0>>(p, LeavesTile<F, 0>::num_tiles(p) * nb, nthreads);
            }
        }
    }
    // ---- check against the definition on a sample of outputs (all of them for small n)
    int bad = 0;
    const u64 w_n = F::root(log_n), w_inv = fpow<F>(w_n, F::P() - 2);
    const u64 n_inv = fpow<F>(n % F::P(), F::P() - 2);
    const u64 checks = n <= 256 ? n : 48;
    for (unsigned c = 0; c < num_cols; c++) {
        std::vector<u64> val(n);
        for (u64 i = 0; i < n; i++) val[i] = F::canonical(F::load_user(in[(size_t)c * n + i]));
        for (u64 q = 0; q < checks; q++) {
            const u64 k = n <= 256 ? q : (q * 2654435761ull + c) % n;
            if (inverse) {
                u64 acc = 0, wk = fpow<F>(w_inv, k), cur = 1;
                for (u64 i = 0; i < n; i++) {
                    acc = (u64)(((u128)acc + fmul<F>(val[i], cur)) % F::P());
                    cur = fmul<F>(cur, wk);
                }
                acc = fmul<F>(acc, n_inv);
                E got = out_cols[(size_t)c * n + k];
                if ((sizeof(E) == 8 && got >= F::P()) || F::canonical(got) != acc) {
                    if (bad++ < 3) printf("  inverse mismatch col %u k %llu\n", c, (unsigned long long)k);
                }
            } else {
                for (unsigned b = 0; b < nb; b++) {
                    const u64 xpt = fmul<F>(shifts[b], fpow<F>(w_n, brev((unsigned)k, log_n)));
                    u64 acc = 0;
                    for (u64 j = n; j-- > 0;) acc = (u64)(((u128)fmul<F>(acc, xpt) + val[j]) % F::P());
                    E got = leaves[((size_t)b * n + k) * pitch + c];
                    if ((sizeof(E) == 8 && got >= F::P()) || F::canonical(got) != acc) {
                        if (bad++ < 3) printf("  leaf mismatch col %u coset %u row %llu\n", c, b, (unsigned long long)k);
                    }
                }
            }
        }
    }
    // columns beyond the batch must stay untouched (pitch padding)
    if (!inverse)
        for (size_t r = 0; r < (size_t)nb * n; r++)
            for (u64 c = num_cols; c < pitch; c++)
                if (leaves[r * pitch + c] != 0) bad++;
    printf("%s log_n=%u plan=", sizeof(E) == 8 ? "GL" : "BB", log_n);
    for (unsigned v : lps) printf("%u ", v);
    printf("cols=%u %s nb=%u threads=%u : %s\n", num_cols, inverse ? "inverse" : "lde", nb, nthreads, bad ? "FAIL" : "ok");
    return bad;
}

template <class F>
static int suite() {
    int bad = 0;
    // single-tile plans of every size (all round plans), both directions
    for (unsigned lp = 0; lp <= 11; lp++) {
        unsigned cols = lp >= 9 ? 3 : 9;
        bad += run_case<F>(lp, {lp}, cols, true, 1, 64, 0);
        bad += run_case<F>(lp, {lp}, cols, false, lp >= 9 ? 1 : 2, 128, 1);
    }
    // two- and three-pass plans (small tiles, as P2G_TEST_MAX_TILE_LOG forces them on the GPU)
    bad += run_case<F>(7, {4, 3}, 5, true, 1, 64, 0);
    bad += run_case<F>(7, {4, 3}, 5, false, 2, 64, 0);
    bad += run_case<F>(9, {5, 4}, 17, false, 4, 256, 2);
    bad += run_case<F>(9, {5, 4}, 17, true, 1, 256, 0);
    bad += run_case<F>(8, {3, 3, 2}, 3, false, 2, 64, 0);
    bad += run_case<F>(8, {3, 3, 2}, 3, true, 1, 64, 0);
    bad += run_case<F>(12, {6, 6}, 2, false, 1, 256, 0);
    bad += run_case<F>(12, {6, 6}, 2, true, 1, 256, 0);
    bad += run_case<F>(13, {5, 4, 4}, 1, false, 1, 128, 0);
    bad += run_case<F>(13, {5, 4, 4}, 1, true, 1, 128, 0);
    bad += run_case<F>(14, {7, 7}, 1, false, 1, 256, 0);
    return bad;
}

int main() {
    int bad = suite<GlH>() + suite<BbH>();
    printf(bad ? "FAILED (%d mismatches)\n" : "all ok\n", bad);
    return bad != 0;
}
