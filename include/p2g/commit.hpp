// C++17 host-side mirror of the reference interface for the commit path, over the C ABI in p2g.h.
//
// The reference is compiled code (Rust); with no Rust toolchain in the build image this header is the compiled-language
// host side: same type names, argument meaning and error behaviour as
//   PolynomialBatch::from_values / from_coeffs / get_lde_values   plonky2/src/fri/oracle.rs:29-158
//   MerkleTree::new / get / prove, MerkleCap                      plonky2/src/hash/merkle_tree.rs:19-222
// Conditions on which the reference panics (inconsistent degrees oracle.rs:139, non-power-of-two sizes util/src/lib.rs:30-33,
// cap_height > log2(leaves) merkle_tree.rs:153-157, empty batch) throw std::invalid_argument / std::out_of_range with
// the reference's message; library failures (no device, CUDA error, OOM) throw std::runtime_error -- there is no CPU
// fallback.  Header only; link against libp2g_cuda.so.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../p2g.h"

namespace p2g {

struct Goldilocks {
    using Elem = uint64_t;
    static constexpr int field = P2G_FIELD_GOLDILOCKS, hash = P2G_HASH_POSEIDON_GL12;
    static constexpr unsigned NUM_HASH_OUT_ELTS = 4;
};
struct BabyBear {
    using Elem = uint32_t;
    static constexpr int field = P2G_FIELD_BABYBEAR, hash = P2G_HASH_POSEIDON2_BB16;
    static constexpr unsigned NUM_HASH_OUT_ELTS = 8;
};

constexpr unsigned SALT_SIZE = 4;  // oracle.rs:25

inline unsigned log2_strict(size_t n) {  // util/src/lib.rs:30-33
    if (n == 0 || (n & (n - 1))) throw std::invalid_argument("Not a power of two: " + std::to_string(n));
    unsigned l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}
inline size_t reverse_bits(size_t x, unsigned num_bits) {  // plonky2/src/util/mod.rs:33-41
    size_t r = 0;
    for (unsigned i = 0; i < num_bits; i++) r |= ((x >> i) & 1) << (num_bits - 1 - i);
    return r;
}

class Context {
public:
    explicit Context(int device = 0) {
        int32_t dev = device;
        int32_t rc = p2g_init(&dev, 1, &ctx_);
        if (rc != P2G_OK) throw std::runtime_error(std::string("p2g_init: ") + p2g_last_error(nullptr));
    }
    ~Context() { p2g_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    p2g_ctx *raw() const { return ctx_; }
    void check(int32_t rc) const {
        if (rc == P2G_OK) return;
        std::string msg = p2g_last_error(ctx_);
        if (rc == P2G_ERR_INVALID) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

private:
    p2g_ctx *ctx_ = nullptr;
};

template <class F>
using HashOut = std::array<typename F::Elem, F::NUM_HASH_OUT_ELTS>;

template <class F>
struct MerkleCap {
    std::vector<HashOut<F>> hashes;
    size_t len() const { return hashes.size(); }
    unsigned height() const { return log2_strict(hashes.size()); }
};

template <class F>
struct MerkleTree {
    using E = typename F::Elem;
    std::vector<std::vector<E>> leaves;
    std::vector<HashOut<F>> digests;  // recursive sibling-adjacent layout, merkle_tree.rs:50-58
    MerkleCap<F> cap;

    // MerkleTree::new, merkle_tree.rs:152-181 (all leaves of one tree have the same length on the commit path)
    static MerkleTree new_(Context &ctx, std::vector<std::vector<E>> leaves, size_t cap_height, int repr = P2G_REPR_CANONICAL) {
        const unsigned log2_leaves_len = log2_strict(leaves.size());
        if (cap_height > log2_leaves_len)
            throw std::invalid_argument("cap_height=" + std::to_string(cap_height) + " should be at most log2(leaves.len())=" +
                                        std::to_string(log2_leaves_len));
        const size_t len = leaves[0].size();
        std::vector<E> flat(leaves.size() * len);
        for (size_t i = 0; i < leaves.size(); i++) {
            if (leaves[i].size() != len) throw std::invalid_argument("leaves of one tree must have equal length");
            std::copy(leaves[i].begin(), leaves[i].end(), flat.begin() + i * len);
        }
        MerkleTree t;
        t.digests.resize(2 * (leaves.size() - (size_t(1) << cap_height)));
        t.cap.hashes.resize(size_t(1) << cap_height);
        p2g_merkle_desc d{};
        d.field = F::field;
        d.hash = F::hash;
        d.repr = repr;
        d.num_leaves = leaves.size();
        d.leaf_len = len;
        d.cap_height = (uint32_t)cap_height;
        d.leaves = flat.data();
        d.digests_out = t.digests.data();
        d.cap_out = t.cap.hashes.data();
        ctx.check(p2g_merkle(ctx.raw(), &d));
        t.leaves = std::move(leaves);
        return t;
    }
    const std::vector<E> &get(size_t i) const { return leaves.at(i); }
    // MerkleTree::prove, merkle_tree.rs:188-222
    std::vector<HashOut<F>> prove(size_t leaf_index) const {
        const unsigned cap_height = log2_strict(cap.len());
        const unsigned num_layers = log2_strict(leaves.size()) - cap_height;
        if (leaf_index >> (cap_height + num_layers)) throw std::out_of_range("leaf index out of range");
        const size_t tree_index = leaf_index >> num_layers, tree_len = digests.size() >> cap_height;
        const HashOut<F> *digest_tree = digests.data() + tree_len * tree_index;
        size_t pair_index = leaf_index & ((size_t(1) << num_layers) - 1);
        std::vector<HashOut<F>> siblings;
        for (unsigned i = 0; i < num_layers; i++) {
            const size_t parity = pair_index & 1;
            pair_index >>= 1;
            const size_t siblings_index = (pair_index << (i + 1)) + (size_t(1) << i) - 1;
            siblings.push_back(digest_tree[2 * siblings_index + (1 - parity)]);
        }
        return siblings;
    }
};

template <class F>
struct PolynomialBatch {
    using E = typename F::Elem;
    std::vector<std::vector<E>> polynomials;  // coefficients
    MerkleTree<F> merkle_tree;
    size_t degree_log = 0, rate_bits = 0;
    bool blinding = false;
    std::array<double, 4> stage_seconds{};  // "IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree" (oracle.rs:76-114)

    // oracle.rs:68-90.  `salts`: SALT_SIZE columns of N elements when blinding (the RNG stays with the caller).
    static PolynomialBatch from_values(Context &ctx, std::vector<std::vector<E>> values, size_t rate_bits, bool blinding,
                                       size_t cap_height, const std::vector<std::vector<E>> *salts = nullptr,
                                       int repr = P2G_REPR_CANONICAL) {
        return commit(ctx, std::move(values), P2G_INPUT_VALUES, rate_bits, blinding, cap_height, salts, repr);
    }
    // oracle.rs:93-123
    static PolynomialBatch from_coeffs(Context &ctx, std::vector<std::vector<E>> polynomials, size_t rate_bits, bool blinding,
                                       size_t cap_height, const std::vector<std::vector<E>> *salts = nullptr,
                                       int repr = P2G_REPR_CANONICAL) {
        return commit(ctx, std::move(polynomials), P2G_INPUT_COEFFS, rate_bits, blinding, cap_height, salts, repr);
    }
    // oracle.rs:153-158
    std::vector<E> get_lde_values(size_t index, size_t step) const {
        const size_t k = reverse_bits(index * step, (unsigned)(degree_log + rate_bits));
        const std::vector<E> &row = merkle_tree.leaves.at(k);
        return std::vector<E>(row.begin(), row.end() - (blinding ? SALT_SIZE : 0));
    }

private:
    static PolynomialBatch commit(Context &ctx, std::vector<std::vector<E>> cols, int kind, size_t rate_bits, bool blinding,
                                  size_t cap_height, const std::vector<std::vector<E>> *salts, int repr) {
        if (cols.empty()) throw std::out_of_range("index out of bounds: the len is 0 but the index is 0");  // polynomials[0]
        const size_t degree = cols[0].size();
        for (auto &c : cols)
            if (c.size() != degree) throw std::invalid_argument("Polynomial degrees inconsistent");  // oracle.rs:139
        const unsigned degree_log = log2_strict(degree);
        const size_t N = degree << rate_bits, width = cols.size() + (blinding ? SALT_SIZE : 0);
        if (blinding && (!salts || salts->size() != SALT_SIZE)) throw std::invalid_argument("blinding needs SALT_SIZE salt columns");
        std::vector<const void *> col_ptrs, salt_ptrs;
        for (auto &c : cols) col_ptrs.push_back(c.data());
        if (blinding)
            for (auto &s : *salts) {
                if (s.size() != N) throw std::invalid_argument("salt columns need N elements");
                salt_ptrs.push_back(s.data());
            }
        PolynomialBatch b;
        std::vector<std::vector<E>> coeffs;
        std::vector<void *> coeff_ptrs;
        if (kind == P2G_INPUT_VALUES) {
            coeffs.assign(cols.size(), std::vector<E>(degree));
            for (auto &c : coeffs) coeff_ptrs.push_back(c.data());
        }
        std::vector<E> flat(N * width);
        const size_t ncap = size_t(1) << cap_height;
        b.merkle_tree.digests.resize(N >= ncap ? 2 * (N - ncap) : 0);
        b.merkle_tree.cap.hashes.resize(ncap);
        p2g_commit_desc d{};
        d.field = F::field;
        d.hash = F::hash;
        d.repr = repr;
        d.input_kind = kind;
        d.log_n = degree_log;
        d.num_cols = (uint32_t)cols.size();
        d.rate_bits = (uint32_t)rate_bits;
        d.cap_height = (uint32_t)cap_height;
        d.num_salt_cols = blinding ? SALT_SIZE : 0;
        d.col_ptrs = col_ptrs.data();
        d.salt_ptrs = blinding ? salt_ptrs.data() : nullptr;
        d.coeffs_out = kind == P2G_INPUT_VALUES ? coeff_ptrs.data() : nullptr;
        d.leaves_out = flat.data();
        d.leaf_pitch = width;
        d.digests_out = b.merkle_tree.digests.data();
        d.cap_out = b.merkle_tree.cap.hashes.data();
        ctx.check(p2g_commit(ctx.raw(), &d));
        b.merkle_tree.leaves.resize(N);
        for (size_t k = 0; k < N; k++) b.merkle_tree.leaves[k].assign(flat.begin() + k * width, flat.begin() + (k + 1) * width);
        b.polynomials = kind == P2G_INPUT_VALUES ? std::move(coeffs) : std::move(cols);
        b.degree_log = degree_log;
        b.rate_bits = rate_bits;
        b.blinding = blinding;
        for (int i = 0; i < 4; i++) b.stage_seconds[i] = d.stage_seconds[i];
        return b;
    }
};

}  // namespace p2g
