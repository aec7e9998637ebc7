// Parity test of the C++ host mirror (include/p2g/commit.hpp) against the CPU oracle (oracle/p2o.h): reads like the
// reference's own unit tests (hash/merkle_tree.rs:257-304): random data, build, every leaf's proof verifies; plus
// from_values against the oracle's from_values, and the panics.  Needs a GPU; built and run by tests/test_cpp_mirror.py.
#include <cstdio>
#include <cstring>
#include <random>

#include "p2g/commit.hpp"
#include "p2g/fri.hpp"
#include "p2o.h"

template <class F>
static std::vector<std::vector<typename F::Elem>> random_data(size_t n, size_t k, uint64_t order, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::vector<std::vector<typename F::Elem>> v(n, std::vector<typename F::Elem>(k));
    for (auto &row : v)
        for (auto &x : row) x = (typename F::Elem)(rng() % order);
    return v;
}

#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                   \
        }                                                               \
    } while (0)

template <class F>
static int run(p2g::Context &ctx, int oracle_field, uint64_t order) {
    using E = typename F::Elem;
    // verify_all_leaves (merkle_tree.rs:257-304) with cap heights 1 and log n
    for (size_t cap_height : {size_t(1), size_t(8)}) {
        auto leaves = random_data<F>(256, 7, order, 11 + cap_height);
        auto tree = p2g::MerkleTree<F>::new_(ctx, leaves, cap_height);
        for (size_t i = 0; i < leaves.size(); i++) {
            auto proof = tree.prove(i);
            CHECK(p2o_verify_merkle_proof_to_cap(oracle_field, leaves[i].data(), leaves[i].size(), i, tree.cap.hashes.data(),
                                                 (unsigned)cap_height, proof.data(), (unsigned)proof.size()) == 1);
        }
    }
    // test_cap_height_too_big: should panic
    bool threw = false;
    try {
        p2g::MerkleTree<F>::new_(ctx, random_data<F>(256, 7, order, 3), 9);
    } catch (const std::invalid_argument &e) {
        threw = strstr(e.what(), "cap_height=9 should be at most log2(leaves.len())=8") != nullptr;
    }
    CHECK(threw);
    // from_values against the oracle
    const unsigned log_n = 9, c = 21, r = 3, h = 4;
    const size_t n = size_t(1) << log_n, N = n << r;
    auto cols = random_data<F>(c, n, order, 5);
    std::vector<E> flat_in(c * n), coeffs(c * n), leaves(N * c), digests(2 * (N - 16) * F::NUM_HASH_OUT_ELTS), cap(16 * F::NUM_HASH_OUT_ELTS);
    for (unsigned j = 0; j < c; j++) std::copy(cols[j].begin(), cols[j].end(), flat_in.begin() + j * n);
    p2o_commit_desc od{};
    od.field = oracle_field;
    od.input_kind = P2O_INPUT_VALUES;
    od.log_n = log_n;
    od.num_cols = c;
    od.rate_bits = r;
    od.cap_height = h;
    od.input = flat_in.data();
    od.coeffs_out = coeffs.data();
    od.leaves_out = leaves.data();
    od.digests_out = digests.data();
    od.cap_out = cap.data();
    CHECK(p2o_commit(&od) == 0);
    auto batch = p2g::PolynomialBatch<F>::from_values(ctx, cols, r, false, h);
    CHECK(batch.degree_log == log_n && batch.rate_bits == r && !batch.blinding);
    for (unsigned j = 0; j < c; j++) CHECK(memcmp(batch.polynomials[j].data(), coeffs.data() + j * n, n * sizeof(E)) == 0);
    for (size_t k = 0; k < N; k++) CHECK(memcmp(batch.merkle_tree.leaves[k].data(), leaves.data() + k * c, c * sizeof(E)) == 0);
    CHECK(memcmp(batch.merkle_tree.digests.data(), digests.data(), digests.size() * sizeof(E)) == 0);
    CHECK(memcmp(batch.merkle_tree.cap.hashes.data(), cap.data(), cap.size() * sizeof(E)) == 0);
    auto row = batch.get_lde_values(3, 1);
    CHECK(memcmp(row.data(), leaves.data() + p2g::reverse_bits(3, log_n + r) * c, c * sizeof(E)) == 0);
    // from_coeffs of the result gives the same cap
    auto again = p2g::PolynomialBatch<F>::from_coeffs(ctx, batch.polynomials, r, false, h);
    CHECK(memcmp(again.merkle_tree.cap.hashes.data(), cap.data(), cap.size() * sizeof(E)) == 0);
    // "Polynomial degrees inconsistent"
    threw = false;
    try {
        auto bad = cols;
        bad[1].resize(n / 2);
        p2g::PolynomialBatch<F>::from_coeffs(ctx, bad, r, false, h);
    } catch (const std::invalid_argument &e) {
        threw = std::string(e.what()) == "Polynomial degrees inconsistent";
    }
    CHECK(threw);
    return 0;
}

// fri_committed_trees + fri_proof_of_work against the oracle's restatement, the transcript driven by the same
// Challenger on both sides (its permutation runs on the GPU; parity of that permutation is checked elsewhere).
template <class F>
static int run_fri(p2g::Context &ctx, int oracle_field, uint64_t order) {
    using E = typename F::Elem;
    constexpr unsigned D = p2g::Ext<F>::D;
    const unsigned degree_bits = 9, r = 3, h = 2, lde_bits = degree_bits + r;
    const size_t n = size_t(1) << lde_bits;
    std::mt19937_64 rng(77);
    std::vector<E> coeffs(n * D, 0), values(n * D);
    for (size_t i = 0; i < (size_t(1) << degree_bits) * D; i++) coeffs[i] = (E)(rng() % order);
    for (unsigned k = 0; k < D; k++) {  // values = coset_fft(F::generator()) of each coordinate
        std::vector<E> col(n);
        for (size_t i = 0; i < n; i++) col[i] = coeffs[i * D + k];
        p2o_coset_fft(oracle_field, col.data(), lde_bits, p2o_generator(oracle_field), 0);
        for (size_t i = 0; i < n; i++) values[i * D + k] = col[i];
    }
    p2g::FriParams params{{r, h, 9, 4}, {4, 3, 1}};
    p2g::Challenger<F> ch(ctx), ch_ref(ctx);
    ch.observe_elements({1, 2, 3});
    ch_ref.observe_elements({1, 2, 3});
    auto got = p2g::fri_committed_trees<F>(ctx, coeffs, values, ch, params);
    // oracle side, layer by layer
    std::vector<E> oc = coeffs, ov = values;
    uint64_t shift = p2o_generator(oracle_field);
    unsigned lg = lde_bits;
    CHECK(got.trees.size() == params.reduction_arity_bits.size());
    for (size_t i = 0; i < params.reduction_arity_bits.size(); i++) {
        unsigned a = (unsigned)params.reduction_arity_bits[i];
        size_t rows = size_t(1) << (lg - a), width = (size_t)D << a;
        std::vector<E> leaves(rows * width), digests(2 * (rows - (size_t(1) << h)) * F::NUM_HASH_OUT_ELTS), cap((size_t(1) << h) * F::NUM_HASH_OUT_ELTS);
        p2o_fri_layer_leaves(oracle_field, ov.data(), lg, leaves.data());
        CHECK(p2o_merkle_tree(oracle_field, leaves.data(), rows, width, h, digests.data(), cap.data(), 0) == 0);
        CHECK(memcmp(got.trees[i].cap.hashes.data(), cap.data(), cap.size() * sizeof(E)) == 0);
        CHECK(got.trees[i].download(0, leaves.size()) == leaves);
        CHECK(got.trees[i].download(1, digests.size()) == digests);
        auto row = got.trees[i].get(rows - 1);
        CHECK(memcmp(row.data(), leaves.data() + (rows - 1) * width, width * sizeof(E)) == 0);
        auto proof = got.trees[i].prove(rows - 1);
        CHECK(p2o_verify_merkle_proof_to_cap(oracle_field, row.data(), width, rows - 1, cap.data(), h, proof.data(),
                                             (unsigned)(proof.size() / F::NUM_HASH_OUT_ELTS)) == 1);
        ch_ref.observe_elements(cap);
        std::vector<E> beta = ch_ref.get_extension_challenge();
        std::vector<E> nc((size_t(1) << (lg - a)) * D), nv(nc.size());
        p2o_fri_fold_layer(oracle_field, oc.data(), lg, a, beta.data(), shift, nc.data(), nv.data(), &shift);
        oc.swap(nc);
        ov.swap(nv);
        lg -= a;
    }
    oc.resize(((size_t(1) << lg) >> r) * D);
    CHECK(got.final_coeffs == oc);
    ch_ref.observe_elements(oc);
    // grinding
    uint64_t exp_w = 0;
    CHECK(p2o_fri_pow(oracle_field, ch_ref.sponge_state.data(), ch_ref.input_buffer.data(), (unsigned)ch_ref.input_buffer.size(),
                      (unsigned)(9 + (order >> 32 ? 0 : 33)), &exp_w) == 0);
    CHECK((uint64_t)p2g::fri_proof_of_work<F>(ctx, ch, params.config) == exp_w);
    // cap_height above the last layer's leaf count panics in the reference
    bool threw = false;
    try {
        p2g::Challenger<F> c2(ctx);
        p2g::fri_committed_trees<F>(ctx, coeffs, values, c2, p2g::FriParams{{r, 9, 9, 4}, {4}});
    } catch (const std::invalid_argument &e) {
        threw = strstr(e.what(), "cap_height=9 should be at most log2(leaves.len())=8") != nullptr;
    }
    CHECK(threw);
    return 0;
}

int main() {
    try {
        p2g::Context ctx(0);
        if (run<p2g::Goldilocks>(ctx, P2O_FIELD_GL, 0xFFFFFFFF00000001ULL)) return 1;
        if (run<p2g::BabyBear>(ctx, P2O_FIELD_BB, 2013265921ULL)) return 1;
        if (run_fri<p2g::Goldilocks>(ctx, P2O_FIELD_GL, 0xFFFFFFFF00000001ULL)) return 1;
        if (run_fri<p2g::BabyBear>(ctx, P2O_FIELD_BB, 2013265921ULL)) return 1;
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    printf("host mirror ok\n");
    return 0;
}
