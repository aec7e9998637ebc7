// C++17 host-side mirror of the FRI prover pieces that run through the C ABI (p2g.h), with the reference's names:
//   Challenger                         plonky2/src/iop/challenger.rs:15-150
//   fri_committed_trees                plonky2/src/fri/prover.rs:83-133
//   fri_proof_of_work                  plonky2/src/fri/prover.rs:136-188
//   DeviceMerkleTree::get / prove      plonky2/src/hash/merkle_tree.rs:183-222, served from GPU memory for
//                                      fri_prover_query_round (prover.rs:219-255)
// The Fiat-Shamir transcript is host logic (a handful of permutations, done on the GPU through p2g_permute); the layer
// trees, the folding and the coset NTTs are the library's.  Panics of the reference become std::invalid_argument.
#pragma once
#include <algorithm>

#include "commit.hpp"

namespace p2g {

template <class F>
struct Ext;
template <>
struct Ext<Goldilocks> {
    static constexpr unsigned D = 2, WIDTH = 12;
    static constexpr uint64_t ORDER = 0xFFFFFFFF00000001ULL;
};
template <>
struct Ext<BabyBear> {
    static constexpr unsigned D = 4, WIDTH = 16;
    static constexpr uint64_t ORDER = 2013265921ULL;
};

template <class F>
class Challenger {
public:
    using E = typename F::Elem;
    static constexpr unsigned RATE = 8;
    explicit Challenger(Context &ctx) : ctx_(ctx), sponge_state(Ext<F>::WIDTH, 0) {}
    void observe_element(E element) {
        output_buffer.clear();
        input_buffer.push_back((E)(element % Ext<F>::ORDER));
        if (input_buffer.size() == RATE) duplexing();
    }
    void observe_elements(const std::vector<E> &elements) {
        for (E e : elements) observe_element(e);
    }
    void observe_cap(const MerkleCap<F> &cap) {
        for (const auto &hash : cap.hashes)
            for (E e : hash) observe_element(e);
    }
    E get_challenge() {
        if (!input_buffer.empty() || output_buffer.empty()) duplexing();
        E c = output_buffer.back();
        output_buffer.pop_back();
        return c;
    }
    std::vector<E> get_n_challenges(size_t n) {
        std::vector<E> v(n);
        for (auto &c : v) c = get_challenge();
        return v;
    }
    std::vector<E> get_extension_challenge() { return get_n_challenges(Ext<F>::D); }

    std::vector<E> sponge_state, input_buffer, output_buffer;

private:
    void duplexing() {
        std::copy(input_buffer.begin(), input_buffer.end(), sponge_state.begin());
        input_buffer.clear();
        ctx_.check(p2g_permute(ctx_.raw(), F::hash, P2G_REPR_CANONICAL, sponge_state.data(), 1));
        output_buffer.assign(sponge_state.begin(), sponge_state.begin() + RATE);
    }
    Context &ctx_;
};

// A Merkle tree whose leaves and digests stay in GPU memory (a FRI layer tree, or a commit made with
// P2G_FLAG_KEEP_ON_DEVICE); only the cap is on the host.
template <class F>
class DeviceMerkleTree {
public:
    using E = typename F::Elem;
    DeviceMerkleTree(Context &ctx, p2g_batch *b, MerkleCap<F> cap_) : cap(std::move(cap_)), ctx_(&ctx), b_(b) {
        uint64_t rows, width, ncap;
        uint32_t layers;
        ctx.check(p2g_batch_shape(b, &rows, &width, &layers, &ncap));
        num_leaves = rows;
        leaf_len = width;
        num_layers = layers;
    }
    DeviceMerkleTree(DeviceMerkleTree &&o) noexcept
        : cap(std::move(o.cap)), num_leaves(o.num_leaves), leaf_len(o.leaf_len), num_layers(o.num_layers), ctx_(o.ctx_), b_(o.b_) {
        o.b_ = nullptr;
    }
    DeviceMerkleTree(const DeviceMerkleTree &) = delete;
    ~DeviceMerkleTree() {
        if (b_) p2g_batch_free(ctx_->raw(), b_);
    }
    std::vector<E> get(size_t i) const {
        uint64_t idx = i;
        std::vector<E> row(leaf_len);
        ctx_->check(p2g_batch_get_rows(ctx_->raw(), b_, &idx, 1, row.data()));
        return row;
    }
    std::vector<E> prove(size_t leaf_index) const {  // siblings, F::NUM_HASH_OUT_ELTS elements each
        uint64_t idx = leaf_index;
        std::vector<E> sib((size_t)num_layers * F::NUM_HASH_OUT_ELTS);
        ctx_->check(p2g_batch_prove(ctx_->raw(), b_, &idx, 1, sib.empty() ? nullptr : sib.data()));
        return sib;
    }
    std::vector<E> download(int what, size_t elems) const {  // 0 leaves, 1 digests, 2 cap
        std::vector<E> out(elems);
        ctx_->check(p2g_batch_download(ctx_->raw(), b_, what, out.data()));
        return out;
    }
    MerkleCap<F> cap;
    size_t num_leaves = 0, leaf_len = 0, num_layers = 0;

private:
    Context *ctx_;
    p2g_batch *b_;
};

struct FriConfig {
    size_t rate_bits, cap_height, proof_of_work_bits = 16, num_query_rounds = 28;
};
struct FriParams {
    FriConfig config;
    std::vector<size_t> reduction_arity_bits;
};

template <class F>
struct FriCommitedTrees {
    std::vector<DeviceMerkleTree<F>> trees;
    std::vector<typename F::Elem> final_coeffs;  // [len][D]
};

// coeffs / values: 2^k extension elements each, D base elements in a row
template <class F>
FriCommitedTrees<F> fri_committed_trees(Context &ctx, const std::vector<typename F::Elem> &coeffs,
                                        const std::vector<typename F::Elem> &values, Challenger<F> &challenger,
                                        const FriParams &fri_params, int repr = P2G_REPR_CANONICAL) {
    using E = typename F::Elem;
    constexpr unsigned D = Ext<F>::D;
    if (coeffs.size() != values.size() || coeffs.size() % D) throw std::invalid_argument("coeffs and values must hold the same number of extension elements");
    const unsigned log_n = log2_strict(coeffs.size() / D);
    p2g_fri *state = nullptr;
    ctx.check(p2g_fri_begin(ctx.raw(), F::field, F::hash, repr, D, log_n, coeffs.data(), values.data(), 0, &state));
    struct Guard {
        Context &c;
        p2g_fri *s;
        ~Guard() { p2g_fri_free(c.raw(), s); }
    } guard{ctx, state};
    FriCommitedTrees<F> out;
    for (size_t arity_bits : fri_params.reduction_arity_bits) {
        MerkleCap<F> cap;
        cap.hashes.resize(size_t(1) << fri_params.config.cap_height);
        p2g_batch *tree = nullptr;
        int32_t rc = p2g_fri_commit_layer(ctx.raw(), state, (uint32_t)arity_bits, (uint32_t)fri_params.config.cap_height, cap.hashes.data(), &tree);
        if (tree) out.trees.emplace_back(ctx, tree, cap);
        ctx.check(rc);
        challenger.observe_cap(out.trees.back().cap);
        std::vector<E> beta = challenger.get_extension_challenge();
        ctx.check(p2g_fri_fold(ctx.raw(), state, beta.data()));
    }
    uint32_t lg = 0;
    ctx.check(p2g_fri_shape(state, &lg, nullptr));
    uint64_t count = (uint64_t(1) << lg) >> fri_params.config.rate_bits;
    out.final_coeffs.resize(count * D);
    ctx.check(p2g_fri_finish(ctx.raw(), state, (uint32_t)fri_params.config.rate_bits, count ? out.final_coeffs.data() : nullptr, &count));
    challenger.observe_elements(out.final_coeffs);  // observe_extension_elements
    return out;
}

template <class F>
typename F::Elem fri_proof_of_work(Context &ctx, Challenger<F> &challenger, const FriConfig &config) {
    unsigned order_bits = 0;
    while (order_bits < 64 && (Ext<F>::ORDER >> order_bits)) order_bits++;
    const uint32_t min_leading_zeros = (uint32_t)config.proof_of_work_bits + (64 - order_bits);
    uint64_t witness = 0;
    ctx.check(p2g_fri_pow(ctx.raw(), F::field, F::hash, challenger.sponge_state.data(),
                          challenger.input_buffer.empty() ? nullptr : challenger.input_buffer.data(),
                          (uint32_t)challenger.input_buffer.size(), min_leading_zeros, &witness));
    challenger.observe_element((typename F::Elem)witness);
    uint64_t response = challenger.get_challenge();
    unsigned lz = response ? (unsigned)__builtin_clzll(response) : 64u;
    if (lz < min_leading_zeros) throw std::runtime_error("proof-of-work response lost its leading zeros");  // assert! :185
    return (typename F::Elem)witness;
}

}  // namespace p2g
