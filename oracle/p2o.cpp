// p2o.cpp -- CPU oracle for the commit hot path.  TEST INFRASTRUCTURE ONLY (see p2o.h header).
// Every function cites the reference file:line it restates (paths relative to the reference tree).
#include "p2o.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "p2o_constants.h"

typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------
// Fields.  The reference imports these from Plonky3 (p3-goldilocks, p3-baby-bear); restated here as
// plain canonical modular arithmetic.
// ------------------------------------------------------------------------------------------------
struct GL {
    typedef uint64_t E;
    static constexpr uint64_t P = 0xFFFFFFFF00000001ULL;  // 2^64 - 2^32 + 1
    static constexpr uint64_t EPS = 0xFFFFFFFFULL;        // 2^64 mod P
    static constexpr unsigned TWO_ADICITY = 32;
    static constexpr uint64_t GENERATOR = 7;  // pinned by k_is in the golden common data
    // two_adic_generator(32): upstream Plonky3 value = 7^((P-1)/2^32); parity unpinned by a literal.
    static constexpr uint64_t TWO_ADIC_ROOT = 1753635133440165772ULL;
    static constexpr unsigned DIGEST_ELEMS = 4;
    static constexpr unsigned WIDTH = 12;

    // same identity the reference uses (poseidon_goldilocks.rs:254-268): 2^64 = EPS, 2^96 = -1 (mod P).
    // Branch-free (the carry of t0 + t1 is a coin flip, a branch there mispredicts half the time); result < 2^64,
    // not necessarily canonical.
    static inline uint64_t reduce128_weak(u128 x) {
        uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
        uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
        uint64_t t0 = lo - hi_hi;
        t0 -= EPS & (0 - (uint64_t)(lo < hi_hi));
        uint64_t t1 = hi_lo * EPS;
        uint64_t t2 = t0 + t1;
        t2 += EPS & (0 - (uint64_t)(t2 < t1));
        return t2;
    }
    static inline E reduce128(u128 x) {
        uint64_t t2 = reduce128_weak(x);
        return t2 - (P & (0 - (uint64_t)(t2 >= P)));
    }
    // product of arbitrary 64-bit representatives, result an arbitrary representative (used inside the permutation)
    static inline uint64_t mul_weak(uint64_t a, uint64_t b) { return reduce128_weak((u128)a * b); }
    static inline E add(E a, E b) {  // canonical in, canonical out, branch-free
        uint64_t s = a + b;
        return s - (P & (0 - (uint64_t)((s < a) | (s >= P))));
    }
    static inline E sub(E a, E b) { return a - b + (P & (0 - (uint64_t)(a < b))); }
    static inline E mul(E a, E b) { return reduce128((u128)a * b); }
    static inline E from_u64(uint64_t x) { return x >= P ? x - P : x; }
};

struct BB {
    typedef uint32_t E;
    static constexpr uint64_t P = 2013265921ULL;  // 15 * 2^27 + 1
    static constexpr unsigned TWO_ADICITY = 27;
    static constexpr uint64_t GENERATOR = 31;            // upstream Plonky3 value; parity unpinned
    static constexpr uint64_t TWO_ADIC_ROOT = 0x1a427a41;  // two_adic_generator(27); parity unpinned
    static constexpr unsigned DIGEST_ELEMS = 8;
    static constexpr unsigned WIDTH = 16;
    static inline E add(E a, E b) {
        uint32_t s = a + b;
        return s >= P ? s - (uint32_t)P : s;
    }
    static inline E sub(E a, E b) { return a >= b ? a - b : a + ((uint32_t)P - b); }
    static inline E mul(E a, E b) { return (E)(((uint64_t)a * b) % P); }
    static inline E from_u64(uint64_t x) { return (E)(x % P); }
};

template <class F>
static typename F::E f_pow(typename F::E a, uint64_t e) {
    typename F::E r = 1;
    while (e) {
        if (e & 1) r = F::mul(r, a);
        a = F::mul(a, a);
        e >>= 1;
    }
    return r;
}
template <class F>
static typename F::E f_inv(typename F::E a) {
    return f_pow<F>(a, F::P - 2);
}
// TwoAdicField::two_adic_generator(bits) (call site field/src/fft.rs:16)
template <class F>
static typename F::E two_adic_generator(unsigned bits) {
    typename F::E g = (typename F::E)F::TWO_ADIC_ROOT;
    for (unsigned i = bits; i < F::TWO_ADICITY; i++) g = F::mul(g, g);
    return g;
}

// ------------------------------------------------------------------------------------------------
// Bit reversal: util/src/lib.rs:57-66 (reverse_index_bits), plonky2/src/util/mod.rs:33-41 (reverse_bits)
// ------------------------------------------------------------------------------------------------
static inline size_t reverse_bits(size_t x, unsigned num_bits) {
    size_t r = 0;
    for (unsigned i = 0; i < num_bits; i++) r |= ((x >> i) & 1) << (num_bits - 1 - i);
    return r;
}
template <class T>
static void reverse_index_bits_in_place(T *a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = reverse_bits(i, log_n);
        if (i < j) std::swap(a[i], a[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// FFT: field/src/fft.rs
// ------------------------------------------------------------------------------------------------
template <class F>
struct RootTable {
    std::vector<std::vector<typename F::E>> rows;  // rows[k][j] = w_{2^{k+1}}^j, j < max(2^k, 2)
};

// fft_root_table, field/src/fft.rs:12-31
template <class F>
static RootTable<F> build_root_table(unsigned lg_n) {
    typedef typename F::E E;
    RootTable<F> t;
    std::vector<E> bases(lg_n);
    E base = two_adic_generator<F>(lg_n);
    for (unsigned i = 0; i < lg_n; i++) {
        bases[i] = base;
        base = F::mul(base, base);
    }
    t.rows.resize(lg_n);
    for (unsigned lg_m = 1; lg_m <= lg_n; lg_m++) {
        size_t half_m = (size_t)1 << (lg_m - 1);
        E b = bases[lg_n - lg_m];
        size_t len = std::max<size_t>(half_m, 2);
        std::vector<E> &row = t.rows[lg_m - 1];
        row.resize(len);
        E cur = 1;
        for (size_t j = 0; j < len; j++) {
            row[j] = cur;
            cur = F::mul(cur, b);
        }
    }
    return t;
}

// The reference rebuilds the table per column when none is passed (fft.rs:39-43); the values are a pure
// function of (field, lg_n), so the oracle caches them.
template <class F>
static const RootTable<F> &root_table(unsigned lg_n) {
    static std::mutex mu;
    static std::map<unsigned, RootTable<F> *> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(lg_n);
    if (it != cache.end()) return *it->second;
    RootTable<F> *t = new RootTable<F>(build_root_table<F>(lg_n));
    cache[lg_n] = t;
    return *t;
}

// fft_classic, field/src/fft.rs:168-205 with the scalar form of fft_classic_simd (:98-160):
// bit-reverse, replicate for the first r layers, then radix-2 DIT layers u +- w*v.
template <class F>
static void fft_classic(typename F::E *values, unsigned lg_n, unsigned r, const RootTable<F> &table) {
    typedef typename F::E E;
    size_t n = (size_t)1 << lg_n;
    reverse_index_bits_in_place(values, lg_n);
    if (r > 0) {
        size_t mask = ~(((size_t)1 << r) - 1);
        for (size_t i = 0; i < n; i++) values[i] = values[i & mask];
    }
    for (unsigned lg_half_m = r; lg_half_m < lg_n; lg_half_m++) {
        size_t half_m = (size_t)1 << lg_half_m;
        size_t m = half_m * 2;
        const E *omega = table.rows[lg_half_m].data();
        for (size_t k = 0; k < n; k += m) {
            for (size_t j = 0; j < half_m; j++) {
                E t = F::mul(omega[j], values[k + half_m + j]);
                E u = values[k + j];
                values[k + j] = F::add(u, t);
                values[k + half_m + j] = F::sub(u, t);
            }
        }
    }
}

// fft_with_options, field/src/fft.rs:55-63
template <class F>
static void fft_with_options(typename F::E *buf, unsigned lg_n, unsigned zero_factor) {
    if (lg_n == 0) return;
    fft_classic<F>(buf, lg_n, std::min(zero_factor, lg_n), root_table<F>(lg_n));
}

// ifft_with_options, field/src/fft.rs:71-94
template <class F>
static void ifft(typename F::E *buf, unsigned lg_n) {
    typedef typename F::E E;
    size_t n = (size_t)1 << lg_n;
    E n_inv = f_pow<F>(f_inv<F>(2), lg_n);
    fft_with_options<F>(buf, lg_n, 0);
    buf[0] = F::mul(buf[0], n_inv);
    if (n >= 2) buf[n / 2] = F::mul(buf[n / 2], n_inv);
    for (size_t i = 1; i < n / 2; i++) {
        size_t j = n - i;
        E ci = F::mul(buf[j], n_inv);
        E cj = F::mul(buf[i], n_inv);
        buf[i] = ci;
        buf[j] = cj;
    }
}

// coset_fft_with_options, field/src/polynomial/mod.rs:282-295
template <class F>
static void coset_fft(typename F::E *buf, unsigned lg_n, typename F::E shift, unsigned zero_factor) {
    typedef typename F::E E;
    size_t n = (size_t)1 << lg_n;
    E pw = 1;
    for (size_t j = 0; j < n; j++) {
        buf[j] = F::mul(pw, buf[j]);
        pw = F::mul(pw, shift);
    }
    fft_with_options<F>(buf, lg_n, zero_factor);
}

// coset_ifft, field/src/polynomial/mod.rs:62-72
template <class F>
static void coset_ifft(typename F::E *buf, unsigned lg_n, typename F::E shift) {
    typedef typename F::E E;
    ifft<F>(buf, lg_n);
    E r = f_inv<F>(shift), pw = 1;
    for (size_t j = 0; j < ((size_t)1 << lg_n); j++) {
        buf[j] = F::mul(buf[j], pw);
        pw = F::mul(pw, r);
    }
}

// ------------------------------------------------------------------------------------------------
// Poseidon over Goldilocks, width 12: plonky2/src/hash/poseidon_goldilocks.rs
// ------------------------------------------------------------------------------------------------
namespace pgl {
static const unsigned W = 12, HALF_FULL = 4, N_PARTIAL = 22;

static inline uint64_t sbox(uint64_t x) {  // sbox_monomial :840-846
    uint64_t x2 = GL::mul_weak(x, x), x4 = GL::mul_weak(x2, x2), x3 = GL::mul_weak(x, x2);
    return GL::mul(x3, x4);
}
// mds_layer semantics (:497-527 computes the same product through a frequency-domain shortcut;
// mds_row_shf :547-557 is the definition): out[r] = sum_i circ[i] * s[(i+r)%12] + diag[r]*s[r]
// Like the reference's mds_layer (:497-527) the product is taken on the 32-bit halves of the state so that every
// accumulator fits 64 bits; the doubled arrays turn the circulant indexing into plain sliding windows that the
// compiler vectorises.
static inline void mds_layer(uint64_t s[W]) {
    alignas(64) uint64_t lo[32], hi[32], acc_lo[16], acc_hi[16];
    for (unsigned i = 0; i < W; i++) {
        lo[i] = lo[i + W] = (uint32_t)s[i];
        hi[i] = hi[i + W] = s[i] >> 32;
    }
    for (unsigned i = 2 * W; i < 32; i++) lo[i] = hi[i] = 0;
    for (unsigned r = 0; r < 16; r++) acc_lo[r] = acc_hi[r] = 0;
    for (unsigned i = 0; i < W; i++) {
        const uint64_t c = P2O_GL_MDS_CIRC[i];
#pragma GCC unroll 16
        for (unsigned r = 0; r < 16; r++) {  // 16 lanes (12 used): two AVX-512 or four AVX2 vectors per half
            acc_lo[r] += lo[r + i] * c;
            acc_hi[r] += hi[r + i] * c;
        }
    }
    acc_lo[0] += lo[0] * P2O_GL_MDS_DIAG[0];
    acc_hi[0] += hi[0] * P2O_GL_MDS_DIAG[0];
    for (unsigned r = 0; r < W; r++) s[r] = GL::reduce128((u128)acc_lo[r] + ((u128)acc_hi[r] << 32));
}
static inline void constant_layer(uint64_t s[W], unsigned round) {
    for (unsigned i = 0; i < W; i++) s[i] = GL::add(s[i], GL::from_u64(P2O_GL_POSEIDON_RC[i + W * round]));
}
static void full_rounds(uint64_t s[W], unsigned &round) {  // :889-897
    for (unsigned k = 0; k < HALF_FULL; k++) {
        constant_layer(s, round);
        for (unsigned i = 0; i < W; i++) s[i] = sbox(s[i]);
        mds_layer(s);
        round++;
    }
}
// poseidon_naive :927-948
static void permute_naive(uint64_t s[W]) {
    unsigned round = 0;
    full_rounds(s, round);
    for (unsigned k = 0; k < N_PARTIAL; k++) {
        constant_layer(s, round);
        s[0] = sbox(s[0]);
        mds_layer(s);
        round++;
    }
    full_rounds(s, round);
}

// --- fast partial rounds (:899-909, :657-683, :718-744).  The reference ships the sparse-factor tables
// as literals produced by an external sage script; the oracle re-derives them from the round constants and
// the MDS matrix (same construction: push constants up through M^-1, factor M = M'' * M' backwards through
// the partial rounds) and checks fast == naive at load time (p2o self-test) and in tests/.
struct FastTables {
    uint64_t first_const[W];
    uint64_t k[N_PARTIAL];          // added to s0 after its S-box
    uint64_t m00[N_PARTIAL];        // (row 0, col 0) of the sparse factor
    uint64_t row[N_PARTIAL][W - 1];  // d = m00*s0 + sum row[i]*s[i+1]
    uint64_t col[N_PARTIAL][W - 1];  // s[i+1] += s0*col[i]
    uint64_t init[W - 1][W - 1];     // dense (W-1)x(W-1) block applied before the first partial round
};

typedef std::vector<std::vector<uint64_t>> Mat;
static Mat mat_mul(const Mat &a, const Mat &b) {
    size_t n = a.size(), m = b[0].size(), k = b.size();
    Mat c(n, std::vector<uint64_t>(m, 0));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < m; j++) {
            uint64_t acc = 0;
            for (size_t t = 0; t < k; t++) acc = GL::add(acc, GL::mul(a[i][t], b[t][j]));
            c[i][j] = acc;
        }
    return c;
}
static Mat mat_inv(Mat a) {
    size_t n = a.size();
    Mat inv(n, std::vector<uint64_t>(n, 0));
    for (size_t i = 0; i < n; i++) inv[i][i] = 1;
    for (size_t c = 0; c < n; c++) {
        size_t p = c;
        while (p < n && a[p][c] == 0) p++;
        if (p == n) {
            fprintf(stderr, "p2o: singular matrix in Poseidon table derivation\n");
            abort();
        }
        std::swap(a[p], a[c]);
        std::swap(inv[p], inv[c]);
        uint64_t d = f_inv<GL>(a[c][c]);
        for (size_t j = 0; j < n; j++) {
            a[c][j] = GL::mul(a[c][j], d);
            inv[c][j] = GL::mul(inv[c][j], d);
        }
        for (size_t r = 0; r < n; r++) {
            if (r == c || a[r][c] == 0) continue;
            uint64_t f = a[r][c];
            for (size_t j = 0; j < n; j++) {
                a[r][j] = GL::sub(a[r][j], GL::mul(f, a[c][j]));
                inv[r][j] = GL::sub(inv[r][j], GL::mul(f, inv[c][j]));
            }
        }
    }
    return inv;
}

static FastTables derive_fast_tables() {
    FastTables T;
    Mat M(W, std::vector<uint64_t>(W));
    for (unsigned r = 0; r < W; r++)
        for (unsigned c = 0; c < W; c++) M[r][c] = P2O_GL_MDS_CIRC[(c + W - r) % W] + (r == c ? P2O_GL_MDS_DIAG[r] : 0);
    Mat Minv = mat_inv(M);
    // constants of the partial rounds
    std::vector<std::vector<uint64_t>> c(N_PARTIAL, std::vector<uint64_t>(W));
    for (unsigned r = 0; r < N_PARTIAL; r++)
        for (unsigned i = 0; i < W; i++) c[r][i] = GL::from_u64(P2O_GL_POSEIDON_RC[(HALF_FULL + r) * W + i]);
    T.k[N_PARTIAL - 1] = 0;
    for (unsigned r = N_PARTIAL - 1; r >= 1; r--) {
        uint64_t e[W];
        for (unsigned i = 0; i < W; i++) {
            uint64_t acc = 0;
            for (unsigned j = 0; j < W; j++) acc = GL::add(acc, GL::mul(Minv[i][j], c[r][j]));
            e[i] = acc;
        }
        T.k[r - 1] = e[0];
        for (unsigned i = 1; i < W; i++) c[r - 1][i] = GL::add(c[r - 1][i], e[i]);
    }
    for (unsigned i = 0; i < W; i++) T.first_const[i] = c[0][i];
    // sparse factors, last partial round first
    Mat A = M;
    for (int r = N_PARTIAL - 1; r >= 0; r--) {
        Mat Ahat(W - 1, std::vector<uint64_t>(W - 1));
        for (unsigned i = 1; i < W; i++)
            for (unsigned j = 1; j < W; j++) Ahat[i - 1][j - 1] = A[i][j];
        Mat AhatInv = mat_inv(Ahat);
        T.m00[r] = A[0][0];
        for (unsigned j = 1; j < W; j++) {  // u^T = a_row^T * Ahat^-1
            uint64_t acc = 0;
            for (unsigned i = 1; i < W; i++) acc = GL::add(acc, GL::mul(A[0][i], AhatInv[i - 1][j - 1]));
            T.row[r][j - 1] = acc;
        }
        for (unsigned i = 1; i < W; i++) T.col[r][i - 1] = A[i][0];
        Mat Mp(W, std::vector<uint64_t>(W, 0));
        Mp[0][0] = 1;
        for (unsigned i = 1; i < W; i++)
            for (unsigned j = 1; j < W; j++) Mp[i][j] = Ahat[i - 1][j - 1];
        if (r > 0)
            A = mat_mul(Mp, M);
        else
            for (unsigned i = 1; i < W; i++)
                for (unsigned j = 1; j < W; j++) T.init[i - 1][j - 1] = Ahat[i - 1][j - 1];
    }
    return T;
}
static const FastTables &fast_tables() {
    static const FastTables T = derive_fast_tables();
    return T;
}

// poseidon :912-922 with partial_rounds :899-909
static void permute_fast(uint64_t s[W]) {
    const FastTables &T = fast_tables();
    unsigned round = 0;
    full_rounds(s, round);
    for (unsigned i = 0; i < W; i++) s[i] = GL::add(s[i], T.first_const[i]);  // partial_first_constant_layer
    {
        uint64_t out[W];  // mds_partial_layer_init
        out[0] = s[0];
        for (unsigned i = 1; i < W; i++) {
            u128 acc = 0;
            uint64_t hi = 0;  // u160 accumulator like the reference's add_u160_u128
            for (unsigned j = 1; j < W; j++) {
                u128 prod = (u128)T.init[i - 1][j - 1] * s[j];
                u128 nacc = acc + prod;
                if (nacc < acc) hi++;
                acc = nacc;
            }
            uint64_t mid = GL::reduce128(((u128)hi << 64) | (uint64_t)(acc >> 64));
            out[i] = GL::reduce128(((u128)mid << 64) | (uint64_t)acc);
        }
        memcpy(s, out, sizeof(out));
    }
    for (unsigned r = 0; r < N_PARTIAL; r++) {
        uint64_t s0 = GL::add(sbox(s[0]), T.k[r]);
        u128 acc = (u128)s0 * T.m00[r];
        uint64_t hi = 0;
        for (unsigned i = 1; i < W; i++) {
            u128 prod = (u128)s[i] * T.row[r][i - 1];
            u128 nacc = acc + prod;
            if (nacc < acc) hi++;
            acc = nacc;
        }
        uint64_t mid = GL::reduce128(((u128)hi << 64) | (uint64_t)(acc >> 64));
        uint64_t d = GL::reduce128(((u128)mid << 64) | (uint64_t)acc);
        for (unsigned i = 1; i < W; i++) s[i] = GL::add(s[i], GL::mul(s0, T.col[r][i - 1]));
        s[0] = d;
    }
    round += N_PARTIAL;
    full_rounds(s, round);
}
}  // namespace pgl

// ------------------------------------------------------------------------------------------------
// Poseidon2 over BabyBear, width 16.  The reference calls the external p3_poseidon2 crate
// (hash/poseidon2_babybear.rs:69-84,150-159); the in-tree restatement of that permutation is the gate's
// witness generator (gates/poseidon2_babybear.rs:609-680) with explicit linear layers (:787-832, :903-917).
// ------------------------------------------------------------------------------------------------
namespace p2bb {
static const unsigned W = 16;
static inline uint32_t sbox(uint32_t x) {  // exp_const_u64::<7>
    uint32_t x2 = BB::mul(x, x), x3 = BB::mul(x2, x), x4 = BB::mul(x2, x2);
    return BB::mul(x3, x4);
}
static inline void apply_mat4(uint32_t x[4]) {  // :903-917
    uint32_t t01 = BB::add(x[0], x[1]);
    uint32_t t23 = BB::add(x[2], x[3]);
    uint32_t t0123 = BB::add(t01, t23);
    uint32_t t01123 = BB::add(t0123, x[1]);
    uint32_t t01233 = BB::add(t0123, x[3]);
    uint32_t x0 = x[0], x2 = x[2];
    x[3] = BB::add(t01233, BB::add(x0, x0));
    x[1] = BB::add(t01123, BB::add(x2, x2));
    x[0] = BB::add(t01123, t01);
    x[2] = BB::add(t01233, t23);
}
static inline void external(uint32_t s[W]) {  // permute_external_mut :804-832
    for (unsigned i = 0; i < W; i += 4) apply_mat4(s + i);
    uint32_t sums[4];
    for (unsigned k = 0; k < 4; k++) {
        uint32_t a = 0;
        for (unsigned j = 0; j < W; j += 4) a = BB::add(a, s[j + k]);
        sums[k] = a;
    }
    for (unsigned i = 0; i < W; i++) s[i] = BB::add(s[i], sums[i % 4]);
}
static inline void internal(uint32_t s[W]) {  // permute_internal_mut :787-802
    const uint32_t MONTY_INV = 943718400u;      // 2^-32 mod p, gates/poseidon2_babybear.rs:776,790
    for (unsigned i = 0; i < W; i++) s[i] = BB::mul(s[i], MONTY_INV);
    uint32_t part = 0;
    for (unsigned i = 1; i < W; i++) part = BB::add(part, s[i]);
    uint32_t full = BB::add(part, s[0]);
    s[0] = BB::sub(part, s[0]);
    for (unsigned i = 0; i < W - 1; i++)
        s[i + 1] = BB::add(full, BB::mul(s[i + 1], (uint32_t)(1u << P2O_BB_P2_DIAG_SHIFTS[i])));
}
static void permute(uint32_t s[W]) {  // run_once :609-671 without the wire writes
    external(s);
    for (unsigned r = 0; r < 4; r++) {
        for (unsigned i = 0; i < W; i++) s[i] = sbox(BB::add(s[i], P2O_BB_P2_EXTERNAL_RC[r][i]));
        external(s);
    }
    for (unsigned r = 0; r < 13; r++) {
        s[0] = sbox(BB::add(s[0], P2O_BB_P2_INTERNAL_RC[r]));
        internal(s);
    }
    for (unsigned r = 4; r < 8; r++) {
        for (unsigned i = 0; i < W; i++) s[i] = sbox(BB::add(s[i], P2O_BB_P2_EXTERNAL_RC[r][i]));
        external(s);
    }
}
// The witness of one Poseidon2BabyBearGate operation (num_ops = 1), wire order of gates/poseidon2_babybear.rs:79-147:
// inputs 0..15, outputs 16..31, swap 32, delta 33..40, S-box inputs of full rounds 1..3 (41..88), of the 13 partial
// rounds (89..101), of full rounds 4..7 (102..165).  Follows the generator run_once :573-680 with the permutation
// code above, so that tests can substitute the oracle's round-by-round states into the gate's constraint equations
// (eval_unfiltered :204-309), the reference's second in-tree definition of the permutation.
enum { GATE_WIRES = 166 };
static void gate_wires(const uint32_t in[W], uint32_t swap, uint32_t wires[GATE_WIRES]) {
    uint32_t s[W];
    for (unsigned i = 0; i < W; i++) wires[i] = s[i] = in[i];
    wires[32] = swap;
    for (unsigned i = 0; i < 8; i++) wires[33 + i] = BB::mul(swap, BB::sub(s[i + 8], s[i]));  // :590-601
    if (swap == 1)
        for (unsigned i = 0; i < 8; i++) std::swap(s[i], s[i + 8]);  // :603-607
    external(s);
    for (unsigned r = 0; r < 4; r++) {
        for (unsigned i = 0; i < W; i++) s[i] = BB::add(s[i], P2O_BB_P2_EXTERNAL_RC[r][i]);
        if (r != 0)
            for (unsigned i = 0; i < W; i++) wires[41 + 16 * (r - 1) + i] = s[i];  // wire_full_sbox_0
        for (unsigned i = 0; i < W; i++) s[i] = sbox(s[i]);
        external(s);
    }
    for (unsigned r = 0; r < 13; r++) {
        s[0] = BB::add(s[0], P2O_BB_P2_INTERNAL_RC[r]);
        wires[89 + r] = s[0];  // wire_partial_sbox
        s[0] = sbox(s[0]);
        internal(s);
    }
    for (unsigned r = 4; r < 8; r++) {
        for (unsigned i = 0; i < W; i++) s[i] = BB::add(s[i], P2O_BB_P2_EXTERNAL_RC[r][i]);
        for (unsigned i = 0; i < W; i++) wires[102 + 16 * (r - 4) + i] = s[i];  // wire_full_sbox_1
        for (unsigned i = 0; i < W; i++) s[i] = sbox(s[i]);
        external(s);
    }
    for (unsigned i = 0; i < W; i++) wires[16 + i] = s[i];
}
}  // namespace p2bb

// ------------------------------------------------------------------------------------------------
// Sponge / compression / hash_or_noop: hash/hashing.rs:76-133, plonk/config.rs:70-84
// ------------------------------------------------------------------------------------------------
template <class F>
struct Perm;
template <>
struct Perm<GL> {
    static inline void permute(uint64_t *s) { pgl::permute_fast(s); }
};
template <>
struct Perm<BB> {
    static inline void permute(uint32_t *s) { p2bb::permute(s); }
};
static const unsigned RATE = 8;

// hash_n_to_m_no_pad with num_outputs = NUM_HASH_OUT_ELTS (hashing.rs:100-123): zero state, overwrite the first
// chunk.len() words with each RATE-chunk, permute after every chunk (also a partial last one), squeeze.
template <class F>
static void hash_no_pad(const typename F::E *in, size_t len, typename F::E *out) {
    typename F::E st[F::WIDTH];
    for (unsigned i = 0; i < F::WIDTH; i++) st[i] = 0;
    for (size_t off = 0; off < len; off += RATE) {
        size_t m = std::min<size_t>(RATE, len - off);
        for (size_t i = 0; i < m; i++) st[i] = in[off + i];
        Perm<F>::permute(st);
    }
    for (unsigned i = 0; i < F::DIGEST_ELEMS; i++) out[i] = st[i];
}
// hash_or_noop (config.rs:70-84)
template <class F>
static void hash_or_noop(const typename F::E *in, size_t len, typename F::E *out) {
    if (len <= F::DIGEST_ELEMS) {
        for (unsigned i = 0; i < F::DIGEST_ELEMS; i++) out[i] = i < len ? in[i] : 0;
    } else {
        hash_no_pad<F>(in, len, out);
    }
}
// compress (hashing.rs:76-96)
template <class F>
static void two_to_one(const typename F::E *l, const typename F::E *r, typename F::E *out) {
    typename F::E st[F::WIDTH];
    for (unsigned i = 0; i < F::WIDTH; i++) st[i] = 0;
    for (unsigned i = 0; i < F::DIGEST_ELEMS; i++) {
        st[i] = l[i];
        st[F::DIGEST_ELEMS + i] = r[i];
    }
    Perm<F>::permute(st);
    for (unsigned i = 0; i < F::DIGEST_ELEMS; i++) out[i] = st[i];
}

// ------------------------------------------------------------------------------------------------
// Merkle tree: hash/merkle_tree.rs
// ------------------------------------------------------------------------------------------------
// fill_subtree :86-113.  digests_buf holds 2*(num_leaves-1) digests laid out as
// left subtree || left digest || right digest || right subtree.
template <class F>
static void fill_subtree(typename F::E *digests_buf, size_t digests_len, const typename F::E *leaves,
                         size_t num_leaves, size_t leaf_len, typename F::E *out) {
    const unsigned D = F::DIGEST_ELEMS;
    if (digests_len == 0) {
        hash_or_noop<F>(leaves, leaf_len, out);
        return;
    }
    size_t half = digests_len / 2;
    typename F::E *left_buf = digests_buf;                 // half-1 digests
    typename F::E *left_digest = digests_buf + (half - 1) * D;
    typename F::E *right_digest = digests_buf + half * D;
    typename F::E *right_buf = digests_buf + (half + 1) * D;
    typename F::E l[8], r[8];
    fill_subtree<F>(left_buf, half - 1, leaves, num_leaves / 2, leaf_len, l);
    fill_subtree<F>(right_buf, half - 1, leaves + (num_leaves / 2) * leaf_len, num_leaves / 2, leaf_len, r);
    memcpy(left_digest, l, D * sizeof(typename F::E));
    memcpy(right_digest, r, D * sizeof(typename F::E));
    two_to_one<F>(l, r, out);
}

// MerkleTree::new :152-181 + fill_digests_buf :115-149
template <class F>
static int merkle_tree(const typename F::E *leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                       typename F::E *digests, typename F::E *cap, int threads) {
    const unsigned D = F::DIGEST_ELEMS;
    if (num_leaves == 0 || (num_leaves & (num_leaves - 1))) return -1;  // log2_strict panics
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < num_leaves) log_n++;
    if (cap_height > log_n) return -1;  // assert! at :153-157
    size_t num_cap = (size_t)1 << cap_height;
    size_t num_digests = 2 * (num_leaves - num_cap);
    (void)threads;
    if (num_digests == 0) {
#pragma omp parallel for schedule(static) num_threads(threads)
        for (size_t i = 0; i < num_leaves; i++) hash_or_noop<F>(leaves + i * leaf_len, leaf_len, cap + i * D);
        return 0;
    }
    size_t sub_digests = num_digests >> cap_height;
    size_t sub_leaves = num_leaves >> cap_height;
    // The reference runs the 2^cap_height subtrees as rayon tasks and splits each recursively with
    // rayon::join.  To use all cores with few subtrees, split every subtree `extra` levels further down.
    unsigned extra = 0;
    while (((size_t)1 << extra) * num_cap < (size_t)4 * (size_t)std::max(1, threads) && ((size_t)2 << extra) <= sub_leaves)
        extra++;
    size_t pieces = (size_t)1 << extra;  // per subtree
    size_t piece_leaves = sub_leaves / pieces;
    size_t total_pieces = pieces * num_cap;
    std::vector<typename F::E> piece_root(total_pieces * D);
    // piece p of subtree s covers leaves [s*sub_leaves + p*piece_leaves, ...); its digest block inside the
    // subtree's buffer starts at p * 2*piece_leaves digests (each aligned block of 2^k leaves owns the
    // 2*(2^k) digest slots starting at 2*first_leaf, the last two of which belong to its ancestors).
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (size_t q = 0; q < total_pieces; q++) {
        size_t s = q / pieces, p = q % pieces;
        typename F::E *sub = digests + s * sub_digests * D;
        fill_subtree<F>(sub + p * 2 * piece_leaves * D, 2 * (piece_leaves - 1),
                        leaves + (s * sub_leaves + p * piece_leaves) * leaf_len, piece_leaves, leaf_len,
                        piece_root.data() + q * D);
    }
    // finish the top `extra` levels of every subtree sequentially (tiny)
    for (size_t s = 0; s < num_cap; s++) {
        typename F::E *sub = digests + s * sub_digests * D;
        std::vector<typename F::E> cur(piece_root.begin() + s * pieces * D, piece_root.begin() + (s + 1) * pieces * D);
        size_t width = piece_leaves;  // leaves under each node of `cur`
        size_t cnt = pieces;
        while (cnt > 1) {
            std::vector<typename F::E> nxt((cnt / 2) * D);
            for (size_t j = 0; j < cnt / 2; j++) {
                // children j*2, j*2+1 each cover `width` leaves; the pair sits in the middle of the parent's block
                size_t first_leaf = j * 2 * width;
                typename F::E *slot = sub + (2 * first_leaf + 2 * width - 2) * D;
                memcpy(slot, cur.data() + (2 * j) * D, D * sizeof(typename F::E));
                memcpy(slot + D, cur.data() + (2 * j + 1) * D, D * sizeof(typename F::E));
                two_to_one<F>(cur.data() + (2 * j) * D, cur.data() + (2 * j + 1) * D, nxt.data() + j * D);
            }
            cur.swap(nxt);
            cnt /= 2;
            width *= 2;
        }
        memcpy(cap + s * D, cur.data(), D * sizeof(typename F::E));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// PolynomialBatch::from_values / from_coeffs: fri/oracle.rs:68-150
// ------------------------------------------------------------------------------------------------
static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class F>
static int commit(p2o_commit_desc *d) {
    typedef typename F::E E;
    const unsigned lg_n = d->log_n, r = d->rate_bits;
    const size_t n = (size_t)1 << lg_n, N = n << r;
    const size_t c = d->num_cols, salt = d->num_salt_cols, width = c + salt;
    int threads = d->threads > 0 ? d->threads : p2o_max_threads();
    if (c == 0) return -1;  // polynomials[0] panics on an empty batch (oracle.rs:101)
    if (lg_n + r > F::TWO_ADICITY) return -1;
    if (d->cap_height > lg_n + r) return -1;
    for (int i = 0; i < 4; i++) d->stage_seconds[i] = 0;

    // "IFFT" (oracle.rs:76-80)
    std::vector<E> coeffs(c * n);
    memcpy(coeffs.data(), d->input, c * n * sizeof(E));
    double t0 = now_s();
    if (d->input_kind == P2O_INPUT_VALUES) {
        root_table<F>(lg_n);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (size_t j = 0; j < c; j++) ifft<F>(coeffs.data() + j * n, lg_n);
    }
    double t1 = now_s();
    d->stage_seconds[0] = t1 - t0;
    if (d->coeffs_out) memcpy(d->coeffs_out, coeffs.data(), c * n * sizeof(E));

    // "FFT + blinding" (lde_values, oracle.rs:125-150): zero-pad to N, coset FFT with shift g, zero_factor = r
    std::vector<E> lde(width * N);
    if (lg_n + r > 0) root_table<F>(lg_n + r);
    t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (size_t j = 0; j < width; j++) {
        E *col = lde.data() + j * N;
        if (j < c) {
            memcpy(col, coeffs.data() + j * n, n * sizeof(E));
            memset(col + n, 0, (N - n) * sizeof(E));
            coset_fft<F>(col, lg_n + r, (E)F::GENERATOR, r);
        } else {
            memcpy(col, (const E *)d->salts + (j - c) * N, N * sizeof(E));
        }
    }
    t1 = now_s();
    d->stage_seconds[1] = t1 - t0;

    // "transpose LDEs" + reverse_index_bits_in_place on the rows (oracle.rs:108-109)
    E *leaves = (E *)d->leaves_out;
    std::vector<E> leaves_own;
    if (!leaves) {
        leaves_own.resize(N * width);
        leaves = leaves_own.data();
    }
    t0 = now_s();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (size_t k = 0; k < N; k++) {
        size_t src = reverse_bits(k, lg_n + r);
        for (size_t j = 0; j < width; j++) leaves[k * width + j] = lde[j * N + src];
    }
    t1 = now_s();
    d->stage_seconds[2] = t1 - t0;
    lde.clear();
    lde.shrink_to_fit();

    // "build Merkle tree" (oracle.rs:110-114)
    const unsigned D = F::DIGEST_ELEMS;
    size_t num_cap = (size_t)1 << d->cap_height;
    std::vector<E> dig_own, cap_own;
    E *digests = (E *)d->digests_out, *cap = (E *)d->cap_out;
    if (!digests) {
        dig_own.resize(2 * (N - num_cap) * D + 1);
        digests = dig_own.data();
    }
    if (!cap) {
        cap_own.resize(num_cap * D);
        cap = cap_own.data();
    }
    t0 = now_s();
    int rc = merkle_tree<F>(leaves, N, width, d->cap_height, digests, cap, threads);
    t1 = now_s();
    d->stage_seconds[3] = t1 - t0;
    return rc;
}

// verify_merkle_proof_to_cap, hash/merkle_proofs.rs:55-79
template <class F>
static int verify_to_cap(const typename F::E *leaf, size_t leaf_len, size_t index, const typename F::E *cap,
                         unsigned cap_height, const typename F::E *sib, unsigned ns) {
    const unsigned D = F::DIGEST_ELEMS;
    typename F::E cur[8], nxt[8];
    hash_or_noop<F>(leaf, leaf_len, cur);
    for (unsigned i = 0; i < ns; i++) {
        unsigned bit = index & 1;
        index >>= 1;
        if (bit)
            two_to_one<F>(sib + i * D, cur, nxt);
        else
            two_to_one<F>(cur, sib + i * D, nxt);
        memcpy(cur, nxt, sizeof(cur));
    }
    if (index >= ((size_t)1 << cap_height)) return 0;
    return memcmp(cur, cap + index * D, D * sizeof(typename F::E)) == 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// FRI commit phase, plonky2/src/fri/prover.rs:83-133.
// Extension fields come from p3's BinomialExtensionField (external): F[X]/(X^D - W) with D = 2, W = 7 for Goldilocks
// (pinned: the golden proof's FRI query rounds only verify with W = 7, tools/gl_fri_check.py) and D = 4, W = 11 for
// BabyBear (upstream Plonky3 value; parity unpinned like the other BabyBear constants).
template <class F>
struct Ext;
template <>
struct Ext<GL> {
    static const unsigned D = 2;
    static const uint64_t W = 7;
};
template <>
struct Ext<BB> {
    static const unsigned D = 4;
    static const uint64_t W = 11;
};
template <class F>
static void ext_mul(const typename F::E *a, const typename F::E *b, typename F::E *out) {
    typedef typename F::E E;
    const unsigned D = Ext<F>::D;
    E r[2 * Ext<F>::D] = {0};
    for (unsigned i = 0; i < D; i++)
        for (unsigned j = 0; j < D; j++) r[i + j] = F::add(r[i + j], F::mul(a[i], b[j]));
    for (unsigned k = 0; k < D; k++) out[k] = k + 1 < D ? F::add(r[k], F::mul((E)Ext<F>::W, r[k + D])) : r[k];
}

// One loop iteration after the challenge (:114-125): coeffs <- chunks_exact(arity).map(reduce_with_powers(chunk, beta))
// (plonk_common.rs:124-136: Horner from the last term), shift <- shift^arity, values <- coeffs.coset_fft(shift).
// The FFT of extension elements uses base-field roots (fft.rs:12-31), so it acts on each coordinate separately.
template <class F>
static void fri_fold_layer(const typename F::E *coeffs, unsigned log_n, unsigned arity_bits, const typename F::E *beta,
                           uint64_t shift_in, typename F::E *coeffs_out, typename F::E *values_out, uint64_t *shift_out) {
    typedef typename F::E E;
    const unsigned D = Ext<F>::D;
    const size_t n_out = (size_t)1 << (log_n - arity_bits), arity = (size_t)1 << arity_bits;
    for (size_t i = 0; i < n_out; i++) {
        E sum[Ext<F>::D] = {0}, t[Ext<F>::D];
        for (size_t j = arity; j-- > 0;) {
            ext_mul<F>(sum, beta, t);
            for (unsigned k = 0; k < D; k++) sum[k] = F::add(t[k], coeffs[(i * arity + j) * D + k]);
        }
        for (unsigned k = 0; k < D; k++) coeffs_out[i * D + k] = sum[k];
    }
    E shift = f_pow<F>((E)shift_in, arity);
    *shift_out = shift;
    std::vector<E> col(n_out);
    for (unsigned k = 0; k < D; k++) {
        for (size_t i = 0; i < n_out; i++) col[i] = coeffs_out[i * D + k];
        coset_fft<F>(col.data(), log_n - arity_bits, shift, 0);
        for (size_t i = 0; i < n_out; i++) values_out[i * D + k] = col[i];
    }
}


// fri_proof_of_work, fri/prover.rs:136-188: the reference searches 0..=p-1 in parallel and takes any hit; the oracle
// returns the SMALLEST hit (sequential semantics), which is what the GPU path is specified to return too.
template <class F>
static int fri_pow(const typename F::E *state, const typename F::E *input, unsigned input_len, unsigned min_leading_zeros,
                   uint64_t *witness) {
    typedef typename F::E E;
    E base[F::WIDTH];
    for (unsigned i = 0; i < F::WIDTH; i++) base[i] = state[i];
    for (unsigned i = 0; i < input_len; i++) base[i] = input[i];  // duplex_intermediate_state :166-168
    const uint64_t last = F::P - 1, chunk = 1 << 14;
    for (uint64_t start = 0; start <= last; start += chunk) {
        uint64_t end = std::min<uint64_t>(last, start + chunk - 1);
        uint64_t best = ~0ull;
#pragma omp parallel for reduction(min : best)
        for (uint64_t c = start; c <= end; c++) {
            E s[F::WIDTH];
            for (unsigned i = 0; i < F::WIDTH; i++) s[i] = base[i];
            s[input_len] = (E)c;
            Perm<F>::permute(s);
            uint64_t resp = s[RATE - 1];  // squeeze().last() :175
            unsigned lz = resp ? (unsigned)__builtin_clzll(resp) : 64u;
            if (lz >= min_leading_zeros && c < best) best = c;
        }
        if (best != ~0ull) {
            *witness = best;
            return 0;
        }
        if (start + chunk < start) break;
    }
    return -1;
}

extern "C" {

int p2o_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

uint64_t p2o_order(int field) { return field == P2O_FIELD_GL ? GL::P : BB::P; }
uint64_t p2o_generator(int field) { return field == P2O_FIELD_GL ? GL::GENERATOR : BB::GENERATOR; }
uint64_t p2o_two_adic_generator(int field, unsigned bits) {
    return field == P2O_FIELD_GL ? two_adic_generator<GL>(bits) : two_adic_generator<BB>(bits);
}
uint64_t p2o_mul(int field, uint64_t a, uint64_t b) {
    return field == P2O_FIELD_GL ? GL::mul(a, b) : BB::mul((uint32_t)a, (uint32_t)b);
}
uint64_t p2o_pow(int field, uint64_t a, uint64_t e) {
    return field == P2O_FIELD_GL ? f_pow<GL>(a, e) : f_pow<BB>((uint32_t)a, e);
}
uint64_t p2o_inv(int field, uint64_t a) { return field == P2O_FIELD_GL ? f_inv<GL>(a) : f_inv<BB>((uint32_t)a); }

void p2o_poseidon_gl(uint64_t s[12]) { pgl::permute_fast(s); }
void p2o_poseidon_gl_naive(uint64_t s[12]) { pgl::permute_naive(s); }
void p2o_poseidon2_bb(uint32_t s[16]) { p2bb::permute(s); }
void p2o_poseidon2_bb_gate_wires(const uint32_t in[16], uint32_t swap, uint32_t wires[166]) { p2bb::gate_wires(in, swap, wires); }

void p2o_fft(int field, void *data, unsigned log_n, unsigned zero_factor) {
    if (field == P2O_FIELD_GL)
        fft_with_options<GL>((uint64_t *)data, log_n, zero_factor);
    else
        fft_with_options<BB>((uint32_t *)data, log_n, zero_factor);
}
void p2o_ifft(int field, void *data, unsigned log_n) {
    if (field == P2O_FIELD_GL)
        ifft<GL>((uint64_t *)data, log_n);
    else
        ifft<BB>((uint32_t *)data, log_n);
}
void p2o_coset_fft(int field, void *data, unsigned log_n, uint64_t shift, unsigned zero_factor) {
    if (field == P2O_FIELD_GL)
        coset_fft<GL>((uint64_t *)data, log_n, shift, zero_factor);
    else
        coset_fft<BB>((uint32_t *)data, log_n, (uint32_t)shift, zero_factor);
}
void p2o_coset_ifft(int field, void *data, unsigned log_n, uint64_t shift) {
    if (field == P2O_FIELD_GL)
        coset_ifft<GL>((uint64_t *)data, log_n, shift);
    else
        coset_ifft<BB>((uint32_t *)data, log_n, (uint32_t)shift);
}
void p2o_reverse_index_bits(void *data, size_t elem_bytes, unsigned log_n) {
    if (elem_bytes == 8)
        reverse_index_bits_in_place((uint64_t *)data, log_n);
    else if (elem_bytes == 4)
        reverse_index_bits_in_place((uint32_t *)data, log_n);
    else
        abort();
}
size_t p2o_reverse_bits(size_t x, unsigned num_bits) { return reverse_bits(x, num_bits); }

void p2o_hash_or_noop(int field, const void *e, size_t len, void *out) {
    if (field == P2O_FIELD_GL)
        hash_or_noop<GL>((const uint64_t *)e, len, (uint64_t *)out);
    else
        hash_or_noop<BB>((const uint32_t *)e, len, (uint32_t *)out);
}
void p2o_hash_no_pad(int field, const void *e, size_t len, void *out) {
    if (field == P2O_FIELD_GL)
        hash_no_pad<GL>((const uint64_t *)e, len, (uint64_t *)out);
    else
        hash_no_pad<BB>((const uint32_t *)e, len, (uint32_t *)out);
}
void p2o_two_to_one(int field, const void *l, const void *r, void *out) {
    if (field == P2O_FIELD_GL)
        two_to_one<GL>((const uint64_t *)l, (const uint64_t *)r, (uint64_t *)out);
    else
        two_to_one<BB>((const uint32_t *)l, (const uint32_t *)r, (uint32_t *)out);
}

int p2o_merkle_tree(int field, const void *leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    void *digests_out, void *cap_out, int threads) {
    if (threads <= 0) threads = p2o_max_threads();
    if (field == P2O_FIELD_GL)
        return merkle_tree<GL>((const uint64_t *)leaves, num_leaves, leaf_len, cap_height, (uint64_t *)digests_out,
                               (uint64_t *)cap_out, threads);
    return merkle_tree<BB>((const uint32_t *)leaves, num_leaves, leaf_len, cap_height, (uint32_t *)digests_out,
                           (uint32_t *)cap_out, threads);
}

// MerkleTree::prove, merkle_tree.rs:188-222 (digests are 32 bytes for both fields)
void p2o_merkle_prove(const void *digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      void *siblings_out) {
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < num_leaves) log_n++;
    unsigned num_layers = log_n - cap_height;
    size_t num_digests = 2 * (num_leaves - ((size_t)1 << cap_height));
    size_t tree_index = leaf_index >> num_layers;
    size_t tree_len = num_digests >> cap_height;
    const uint8_t *tree = (const uint8_t *)digests + tree_len * tree_index * 32;
    size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
    for (unsigned i = 0; i < num_layers; i++) {
        size_t parity = pair_index & 1;
        pair_index >>= 1;
        size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
        size_t sibling_index = 2 * siblings_index + (1 - parity);
        memcpy((uint8_t *)siblings_out + 32 * i, tree + 32 * sibling_index, 32);
    }
}

int p2o_verify_merkle_proof_to_cap(int field, const void *leaf, size_t leaf_len, size_t leaf_index, const void *cap,
                                   unsigned cap_height, const void *siblings, unsigned num_siblings) {
    if (field == P2O_FIELD_GL)
        return verify_to_cap<GL>((const uint64_t *)leaf, leaf_len, leaf_index, (const uint64_t *)cap, cap_height,
                                 (const uint64_t *)siblings, num_siblings);
    return verify_to_cap<BB>((const uint32_t *)leaf, leaf_len, leaf_index, (const uint32_t *)cap, cap_height,
                             (const uint32_t *)siblings, num_siblings);
}

int p2o_commit(p2o_commit_desc *d) { return d->field == P2O_FIELD_GL ? commit<GL>(d) : commit<BB>(d); }

unsigned p2o_ext_degree(int field) { return field == P2O_FIELD_GL ? Ext<GL>::D : Ext<BB>::D; }
uint64_t p2o_ext_w(int field) { return field == P2O_FIELD_GL ? Ext<GL>::W : Ext<BB>::W; }
void p2o_ext_mul(int field, const void *a, const void *b, void *out) {
    if (field == P2O_FIELD_GL)
        ext_mul<GL>((const uint64_t *)a, (const uint64_t *)b, (uint64_t *)out);
    else
        ext_mul<BB>((const uint32_t *)a, (const uint32_t *)b, (uint32_t *)out);
}
void p2o_fri_fold_layer(int field, const void *coeffs, unsigned log_n, unsigned arity_bits, const void *beta,
                        uint64_t shift_in, void *coeffs_out, void *values_out, uint64_t *shift_out) {
    if (field == P2O_FIELD_GL)
        fri_fold_layer<GL>((const uint64_t *)coeffs, log_n, arity_bits, (const uint64_t *)beta, shift_in,
                           (uint64_t *)coeffs_out, (uint64_t *)values_out, shift_out);
    else
        fri_fold_layer<BB>((const uint32_t *)coeffs, log_n, arity_bits, (const uint32_t *)beta, shift_in,
                           (uint32_t *)coeffs_out, (uint32_t *)values_out, shift_out);
}
// reverse_index_bits_in_place(values) + chunks(arity) + flatten (:103-109): an extension element is 16 bytes in both
// configurations, and flattening consecutive elements is the identity on memory, so the leaf matrix
// [n >> arity_bits][arity * D] is the bit-reversed copy of the values.
void p2o_fri_layer_leaves(int field, const void *values, unsigned log_n, void *leaves_out) {
    (void)field;
    struct E16 {
        uint64_t a, b;
    };
    size_t n = (size_t)1 << log_n;
    memcpy(leaves_out, values, n * sizeof(E16));
    reverse_index_bits_in_place((E16 *)leaves_out, log_n);
}
int p2o_fri_pow(int field, const void *state, const void *input, unsigned input_len, unsigned min_leading_zeros,
                uint64_t *witness) {
    if (field == P2O_FIELD_GL)
        return fri_pow<GL>((const uint64_t *)state, (const uint64_t *)input, input_len, min_leading_zeros, witness);
    return fri_pow<BB>((const uint32_t *)state, (const uint32_t *)input, input_len, min_leading_zeros, witness);
}
}  // extern "C"
