"""FRI commit phase (plonky2/src/fri/prover.rs:83-133).

CPU tests pin the oracle's restatement: extension arithmetic and the fold against their definitions, and -- the strong
one -- a whole oracle commit phase against the verifier's folding checks (tools/gl_fri_check.fri_fold_chain), the very
code that re-verifies the 28 FRI query rounds of the reference's frozen proof (tests/test_golden_fri.py).  A prover
that chunked, bit-reversed, folded or shifted differently from the reference's would not pass them.
GPU tests (-m gpu) compare the CUDA path with the oracle bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def ext_params(field):
    return (2, 7) if field == 0 else (4, 11)


def py_ext_mul(a, b, p, d, w):
    r = [0] * (2 * d)
    for i in range(d):
        for j in range(d):
            r[i + j] += int(a[i]) * int(b[j])
    return [(r[k] + w * r[k + d]) % p for k in range(d)]


def random_ext(rng, oracle, field, n):
    d, _ = ext_params(field)
    return rng.integers(0, oracle.ORDER[field], size=(n, d), dtype=np.uint64).astype(oracle.DTYPE[field])


def lde_pair(rng, oracle, field, degree_bits, rate_bits):
    """(lde_polynomial_coeffs, lde_polynomial_values) of a random polynomial of degree < 2^degree_bits over the extension
    field, as the prover hands them to fri_proof (fri/oracle.rs:223-231): coefficients zero-padded to the LDE size,
    values = coset_fft with shift F::generator()."""
    d, _ = ext_params(field)
    n = 1 << (degree_bits + rate_bits)
    coeffs = np.zeros((n, d), dtype=oracle.DTYPE[field])
    coeffs[: 1 << degree_bits] = random_ext(rng, oracle, field, 1 << degree_bits)
    g = int(oracle.load().p2o_generator(field))
    values = np.stack([oracle.coset_fft(field, coeffs[:, k].copy(), g) for k in range(d)], axis=1)
    return coeffs, np.ascontiguousarray(values)


def oracle_challenger(oracle, field_desc, field):
    import plonky2_goldibear_b200 as p2g
    return p2g.Challenger(field_desc, permute=lambda s: oracle.permute(field, s))


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field", [0, 1])
def test_extension_multiplication_is_polynomial_multiplication_mod_binomial(oracle, field):
    d, w = ext_params(field)
    assert oracle.ext_degree(field) == d and int(oracle.load().p2o_ext_w(field)) == w
    rng = np.random.default_rng(5)
    p = oracle.ORDER[field]
    for _ in range(50):
        a, b = random_ext(rng, oracle, field, 2)
        assert [int(x) for x in oracle.ext_mul(field, a, b)] == py_ext_mul(a, b, p, d, w)
    one = np.zeros(d, dtype=oracle.DTYPE[field])
    one[0] = 1
    a = random_ext(rng, oracle, field, 1)[0]
    assert (oracle.ext_mul(field, a, one) == a).all()


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("arity_bits", [1, 2, 4])
def test_fold_layer_matches_its_definition(oracle, field, arity_bits):
    d, w = ext_params(field)
    p = oracle.ORDER[field]
    rng = np.random.default_rng(17 + arity_bits)
    log_n = 6
    coeffs = random_ext(rng, oracle, field, 1 << log_n)
    beta = random_ext(rng, oracle, field, 1)[0]
    shift = 7 if field == 0 else 31
    co, vo, so = oracle.fri_fold_layer(field, coeffs, arity_bits, beta, shift)
    arity, n_out = 1 << arity_bits, (1 << log_n) >> arity_bits
    # reduce_with_powers: sum_j chunk[j] * beta^j
    for i in range(n_out):
        acc = [0] * d
        pw = [1] + [0] * (d - 1)
        for j in range(arity):
            t = py_ext_mul(coeffs[i * arity + j], pw, p, d, w)
            acc = [(x + y) % p for x, y in zip(acc, t)]
            pw = py_ext_mul(pw, beta, p, d, w)
        assert [int(x) for x in co[i]] == acc
    assert so == pow(shift, arity, p)
    # values_out[i] = P'(shift' * w_n'^i), natural order, coordinate by coordinate
    wn = int(oracle.load().p2o_two_adic_generator(field, log_n - arity_bits))
    for i in range(n_out):
        x = so * pow(wn, i, p) % p
        for k in range(d):
            acc = 0
            for c in reversed(co[:, k]):
                acc = (acc * x + int(c)) % p
            assert int(vo[i, k]) == acc


def test_package_challenger_replays_the_pinned_transcript_class(oracle):
    """plonky2_goldibear_b200.Challenger == tools/gl_fri_check.Challenger (the one that reproduces the golden proof's
    query indices) on an arbitrary observe / challenge sequence."""
    import gl_fri_check as chk
    import plonky2_goldibear_b200 as p2g
    perm = lambda s: [int(x) for x in oracle.permute(0, np.array(s, dtype=np.uint64))]
    a = chk.Challenger(perm)
    b = oracle_challenger(oracle, p2g.GOLDILOCKS, 0)
    rng = np.random.default_rng(3)
    for step in range(40):
        k = int(rng.integers(0, 3))
        if k == 0:
            xs = [int(x) for x in rng.integers(0, 2**63, size=int(rng.integers(1, 20)))]
            a.observe_many(xs)
            b.observe_elements(xs)
        elif k == 1:
            assert a.challenge() == b.get_challenge()
        else:
            assert list(a.ext_challenge()) == [int(x) for x in b.get_extension_challenge()]


@pytest.mark.parametrize("arities", [[4, 4], [1, 2, 3], [3], []])
def test_oracle_commit_phase_passes_the_pinned_verifier_fold_checks(oracle, arities):
    import gl_fri_check as chk
    import plonky2_goldibear_b200 as p2g
    rng = np.random.default_rng(23)
    degree_bits, rate_bits, cap_height = 9, 3, 1  # sum(arities) <= degree_bits, as the reference's strategies guarantee
    lde_bits = degree_bits + rate_bits
    coeffs, values = lde_pair(rng, oracle, 0, degree_bits, rate_bits)
    ch = oracle_challenger(oracle, p2g.GOLDILOCKS, 0)
    ch.observe_elements([1, 2, 3])
    trees, final = oracle.fri_committed_trees(0, coeffs, values, ch, arities, rate_bits, cap_height)
    assert final.shape[0] == (1 << (lde_bits - sum(arities))) >> rate_bits
    # the verifier's side of the transcript
    vc = oracle_challenger(oracle, p2g.GOLDILOCKS, 0)
    vc.observe_elements([1, 2, 3])
    betas = []
    for _, _, cap in trees:
        vc.observe_cap(cap)
        betas.append(tuple(int(x) for x in vc.get_extension_challenge()))
    vc.observe_extension_elements(final)
    assert vc.get_challenge() == ch.get_challenge()
    tag = lambda e: (int(e[0]), int(e[1]))
    two_adic = lambda bits: int(oracle.load().p2o_two_adic_generator(0, bits))
    for x_index in [0, 1, 5, (1 << lde_bits) - 1] + [int(x) for x in rng.integers(0, 1 << lde_bits, size=6)]:
        steps, idx = [], x_index
        for (leaves, _, _), ab in zip(trees, arities):
            row = leaves[idx >> ab].reshape(1 << ab, 2)
            steps.append([tag(e) for e in row])
            idx >>= ab
        # values are natural order; leaf position x_index holds the evaluation at g * w^bitrev(x_index)
        first = tag(values[chk.reverse_bits(x_index, lde_bits)])
        chk.fri_fold_chain(x_index, first, steps, arities, betas, [tag(e) for e in final], lde_bits, 7, two_adic)
    if arities:
        bad = list(betas)
        bad[0] = ((bad[0][0] + 1) % chk.P, bad[0][1])
        steps, idx = [], 3
        for (leaves, _, _), ab in zip(trees, arities):
            steps.append([tag(e) for e in leaves[idx >> ab].reshape(1 << ab, 2)])
            idx >>= ab
        with pytest.raises(AssertionError):
            chk.fri_fold_chain(3, tag(values[chk.reverse_bits(3, lde_bits)]), steps, arities, bad, [tag(e) for e in final], lde_bits, 7, two_adic)


def leading_zeros64(x):
    return 64 - int(x).bit_length()


@pytest.mark.parametrize("field,bits", [(0, 6), (0, 10), (1, 6), (1, 9)])
def test_oracle_grinding_returns_the_smallest_valid_witness(oracle, field, bits):
    import plonky2_goldibear_b200 as p2g
    f = {0: p2g.GOLDILOCKS, 1: p2g.BABYBEAR}[field]
    rng = np.random.default_rng(31 + bits)
    min_lz = bits + (64 - f.order.bit_length())
    for input_len in (0, 3, 7):
        ch = oracle_challenger(oracle, f, field)
        ch.observe_elements(rng.integers(0, f.order, size=16 + input_len, dtype=np.uint64))  # two duplexings, then input_len pending
        assert len(ch.input_buffer) == input_len
        w = oracle.fri_pow(field, ch.sponge_state, ch.input_buffer, min_lz)

        def response(c):
            s = ch.sponge_state.copy()
            s[:input_len] = ch.input_buffer
            s[input_len] = c
            return int(oracle.permute(field, s)[7])
        assert leading_zeros64(response(w)) >= min_lz
        assert all(leading_zeros64(response(c)) < min_lz for c in range(w))
        # and the ordinary challenger agrees with the half-duplex shortcut (prover.rs:181-186)
        ch.observe_element(w)
        assert leading_zeros64(ch.get_challenge()) >= min_lz


def test_oracle_grinding_on_the_golden_transcript(oracle):
    """At the point where the reference ground its proof of work (the frozen proof's transcript, replayed by the pinned
    verifier code), the oracle's smallest witness is valid and not above the reference's (the reference keeps whichever
    hit its thread pool finds first)."""
    import json
    import gl_fri_check as chk
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "gl_proof_fri.json")))
    perm = lambda s: [int(x) for x in oracle.permute(0, np.array(s, dtype=np.uint64))]
    two_adic = lambda bits: int(oracle.load().p2o_two_adic_generator(0, bits))
    res = chk.verify(fx, perm, 7, two_adic)
    min_lz = res["proof_of_work_bits"] + 0  # 64 - bits(order) = 0 for Goldilocks
    w = oracle.fri_pow(0, res["pre_pow_state"], res["pre_pow_input"], min_lz)
    assert w <= res["pow_witness"]
    s = list(res["pre_pow_state"])
    k = len(res["pre_pow_input"])
    s[:k] = res["pre_pow_input"]
    for cand in (w, res["pow_witness"]):
        t = list(s)
        t[k] = cand
        assert leading_zeros64(perm(t)[7]) >= min_lz


# ---------------------------------------------------------------------------------------------------------------------
def gpu_fields():
    import plonky2_goldibear_b200 as p2g
    return {0: p2g.GOLDILOCKS, 1: p2g.BABYBEAR}


@pytest.mark.gpu
@pytest.mark.parametrize("field", [0, 1])
def test_gpu_challenger_matches_oracle_permutation(ctx, oracle, field):
    import plonky2_goldibear_b200 as p2g
    f = gpu_fields()[field]
    a, b = p2g.Challenger(f, ctx=ctx), oracle_challenger(oracle, f, field)
    rng = np.random.default_rng(9)
    for _ in range(12):
        xs = rng.integers(0, f.order, size=int(rng.integers(1, 30)), dtype=np.uint64)
        a.observe_elements(xs)
        b.observe_elements(xs)
        assert a.get_n_challenges(3) == b.get_n_challenges(3)
        assert (a.get_extension_challenge() == b.get_extension_challenge()).all()


@pytest.mark.gpu
@pytest.mark.parametrize("field,degree_bits,rate_bits,cap_height,arities", [
    (0, 9, 3, 4, [4, 4]),
    (0, 11, 3, 3, [4, 4, 3]),
    (0, 12, 3, 4, [4, 4]),            # the reference's ConstantArityBits(4, 5) shape at degree 2^12; folded layer 2^11: two-pass NTT
    (0, 12, 3, 3, [4, 4, 4]),
    (0, 14, 1, 0, [1, 2, 3, 4, 1]),   # mixed arities, cap of one root
    (0, 5, 3, 0, [3, 3, 2]),          # folds down to a single coefficient (over-reduced: the final polynomial is empty)
    (0, 6, 2, 3, []),                 # no reduction at all: only the final polynomial
    (0, 4, 1, 3, [2]),                # cap_height == log2(leaves): the tree is all cap
    (1, 9, 3, 4, [4, 4]),
    (1, 12, 3, 3, [4, 4, 4]),
    (1, 13, 1, 1, [1, 3, 2, 4]),
])
def test_gpu_commit_phase_matches_oracle(ctx, oracle, field, degree_bits, rate_bits, cap_height, arities):
    import plonky2_goldibear_b200 as p2g
    f = gpu_fields()[field]
    rng = np.random.default_rng(1000 * field + degree_bits)
    coeffs, values = lde_pair(rng, oracle, field, degree_bits, rate_bits)
    ch_o = oracle_challenger(oracle, f, field)
    ch_o.observe_elements([5, 6, 7])
    exp_trees, exp_final = oracle.fri_committed_trees(field, coeffs, values, ch_o, arities, rate_bits, cap_height)
    ch_g = p2g.Challenger(f, ctx=ctx)
    ch_g.observe_elements([5, 6, 7])
    params = p2g.FriParams(p2g.FriConfig(rate_bits, cap_height), list(arities))
    trees, final = p2g.fri_committed_trees(f, coeffs, values, ch_g, params, ctx=ctx)
    assert len(trees) == len(exp_trees)
    assert final.shape == exp_final.shape and (final == exp_final).all()
    for t, (leaves, digests, cap) in zip(trees, exp_trees):
        assert (t.cap.hashes == cap).all()
        assert (t.download("leaves") == leaves).all()
        assert (t.download("digests") == digests).all()
        # query-phase accessors (fri/prover.rs:236-252)
        idx = sorted(set(int(x) for x in rng.integers(0, leaves.shape[0], size=5)) | {0, leaves.shape[0] - 1})
        assert (t.get_many(idx) == leaves[idx]).all()
        proofs = t.prove_many(idx)
        for i, pr in zip(idx, proofs):
            assert (pr == oracle.merkle_prove(field, digests, leaves.shape[0], cap_height, i)).all()
    assert ch_g.get_challenge() == ch_o.get_challenge()  # both transcripts ended in the same state
    for t in trees:
        t.free()


@pytest.mark.gpu
def test_gpu_commit_phase_babybear_montgomery_representation(ctx, oracle):
    import plonky2_goldibear_b200 as p2g
    from test_gpu_parity import from_monty, to_monty
    f = p2g.BABYBEAR
    rng = np.random.default_rng(77)
    coeffs, values = lde_pair(rng, oracle, 1, 8, 3)
    arities, h = [4, 3], 2
    ch_o = oracle_challenger(oracle, f, 1)
    exp_trees, exp_final = oracle.fri_committed_trees(1, coeffs, values, ch_o, arities, 3, h)

    class MontyChallenger(p2g.Challenger):
        """transcript over canonical values, hands beta to the library in Montgomery form like p3's BabyBear memory"""
        def observe_cap(self, cap):
            super().observe_cap(from_monty(np.asarray(cap.hashes)))

        def get_extension_challenge(self):
            return to_monty(super().get_extension_challenge())

        def observe_extension_elements(self, e):
            super().observe_extension_elements(from_monty(np.asarray(e)))

    ch_g = MontyChallenger(f, ctx=ctx)
    params = p2g.FriParams(p2g.FriConfig(3, h), arities)
    trees, final = p2g.fri_committed_trees(f, to_monty(coeffs), to_monty(values), ch_g, params, ctx=ctx, repr=1)
    assert (from_monty(final) == exp_final).all()
    for t, (leaves, digests, cap) in zip(trees, exp_trees):
        assert (from_monty(t.cap.hashes) == cap).all()
        assert (from_monty(t.download("leaves")) == leaves).all()
        assert (from_monty(t.download("digests")) == digests).all()
        t.free()


@pytest.mark.gpu
def test_gpu_commit_phase_errors_are_the_reference_panics(ctx, oracle):
    import ctypes
    import plonky2_goldibear_b200 as p2g
    f = p2g.GOLDILOCKS
    rng = np.random.default_rng(2)
    coeffs, values = lde_pair(rng, oracle, 0, 3, 1)  # 16 extension elements
    # cap_height above log2(leaves): MerkleTree::new asserts (merkle_tree.rs:153-157)
    with pytest.raises(p2g.P2GError) as e:
        p2g.fri_committed_trees(f, coeffs, values, p2g.Challenger(f, ctx=ctx), p2g.FriParams(p2g.FriConfig(1, 3), [2]), ctx=ctx)
    assert "cap_height=3 should be at most log2(leaves.len())=2" in str(e.value)
    # shapes
    with pytest.raises(ValueError):
        p2g.fri_committed_trees(f, coeffs[:12], values[:12], p2g.Challenger(f, ctx=ctx), p2g.FriParams(p2g.FriConfig(1, 0), []), ctx=ctx)
    # protocol order on the raw ABI: fold before a layer was committed
    state = ctypes.c_void_p()
    ctx.check(ctx.lib.p2g_fri_begin(ctx.handle, 0, 0, 0, 2, 4, coeffs.ctypes.data, values.ctypes.data, 0, ctypes.byref(state)))
    beta = np.array([1, 2], dtype=np.uint64)
    assert ctx.lib.p2g_fri_fold(ctx.handle, state, beta.ctypes.data) == -1
    ctx.lib.p2g_fri_free(ctx.handle, state)
    # wrong extension degree for the field
    assert ctx.lib.p2g_fri_begin(ctx.handle, 0, 0, 0, 4, 4, coeffs.ctypes.data, values.ctypes.data, 0, ctypes.byref(state)) == -5


@pytest.mark.gpu
@pytest.mark.parametrize("field,bits", [(0, 4), (0, 12), (0, 18), (1, 5), (1, 14)])
def test_gpu_grinding_matches_oracle(ctx, oracle, field, bits):
    import plonky2_goldibear_b200 as p2g
    f = gpu_fields()[field]
    rng = np.random.default_rng(bits)
    for input_len in (0, 1, 7):
        ch_o, ch_g = oracle_challenger(oracle, f, field), p2g.Challenger(f, ctx=ctx)
        xs = rng.integers(0, f.order, size=8 + input_len, dtype=np.uint64)
        ch_o.observe_elements(xs)
        ch_g.observe_elements(xs)
        exp = oracle.fri_pow(field, ch_o.sponge_state, ch_o.input_buffer, bits + 64 - f.order.bit_length())
        got = p2g.fri_proof_of_work(f, ch_g, p2g.FriConfig(3, 4, proof_of_work_bits=bits), ctx=ctx)
        assert got == exp
        ch_o.observe_element(exp)
        ch_o.get_challenge()
        assert ch_o.get_challenge() == ch_g.get_challenge()


@pytest.mark.gpu
def test_gpu_fri_proof_is_accepted_by_the_pinned_verifier_checks(ctx, oracle):
    """fri_proof end to end on the GPU (commit phase, grinding, query rounds served from device-resident trees), then the
    verifier's side: transcript replay, proof-of-work check, Merkle paths against the caps, folding chain."""
    import gl_fri_check as chk
    import plonky2_goldibear_b200 as p2g
    f = p2g.GOLDILOCKS
    rng = np.random.default_rng(99)
    degree_bits, rate_bits, cap_height, arities = 10, 3, 2, [4, 3, 1]
    lde_bits = degree_bits + rate_bits
    coeffs, values = lde_pair(rng, oracle, 0, degree_bits, rate_bits)
    params = p2g.FriParams(p2g.FriConfig(rate_bits, cap_height, proof_of_work_bits=11, num_query_rounds=9), arities)
    # one "initial" oracle: the codeword itself committed as a 2-column batch, kept on the device
    initial = p2g.MerkleTree.new(f, oracle.fri_layer_leaves(0, values, 0), cap_height, ctx=ctx)
    ch = p2g.Challenger(f, ctx=ctx)
    ch.observe_cap(initial.cap)
    proof = p2g.fri_proof(f, [initial], coeffs, values, ch, params, ctx=ctx)
    # ---- verifier
    vc = oracle_challenger(oracle, f, 0)
    vc.observe_cap(initial.cap)
    betas = []
    for cap in proof["commit_phase_merkle_caps"]:
        vc.observe_cap(cap)
        betas.append(tuple(int(x) for x in vc.get_extension_challenge()))
    vc.observe_extension_elements(proof["final_poly"])
    vc.observe_element(proof["pow_witness"])
    assert 64 - int(vc.get_challenge()).bit_length() >= 11
    indices = [c % (1 << lde_bits) for c in vc.get_n_challenges(9)]
    assert indices == [r["x_index"] for r in proof["query_round_proofs"]]
    tag = lambda e: (int(e[0]), int(e[1]))
    two_adic = lambda bits: int(oracle.load().p2o_two_adic_generator(0, bits))
    for rd in proof["query_round_proofs"]:
        x = rd["x_index"]
        (row, path), = rd["initial_trees_proof"]
        assert oracle.verify_merkle_proof_to_cap(0, row, x, initial.cap.hashes, path)
        idx, steps = x, []
        for (evals, path), cap, ab in zip(rd["steps"], proof["commit_phase_merkle_caps"], arities):
            idx >>= ab
            assert oracle.verify_merkle_proof_to_cap(0, evals.reshape(-1), idx, cap.hashes, path)
            steps.append([tag(e) for e in evals])
        chk.fri_fold_chain(x, tag(row), steps, arities, betas, [tag(e) for e in proof["final_poly"]], lde_bits, 7, two_adic)


# ---- the transforms around the quotient stage and prove_openings ------------------------------------
@pytest.mark.parametrize("field", [0, 1])
def test_oracle_coset_ifft_inverts_coset_fft(oracle, field):
    rng = np.random.default_rng(8)
    g = int(oracle.load().p2o_generator(field))
    for log_n in (0, 1, 5, 9):
        coeffs = rng.integers(0, oracle.ORDER[field], size=1 << log_n, dtype=np.uint64).astype(oracle.DTYPE[field])
        vals = oracle.coset_fft(field, coeffs, g)
        assert (oracle.coset_ifft(field, vals, g) == coeffs).all()


@pytest.mark.gpu
@pytest.mark.parametrize("field,log_n,cols", [(0, 0, 1), (0, 3, 2), (0, 10, 16), (0, 13, 5), (0, 21, 2), (1, 4, 3), (1, 12, 48), (1, 22, 1)])
def test_gpu_coset_ifft_matches_oracle(ctx, oracle, field, log_n, cols):
    import plonky2_goldibear_b200 as p2g
    f = gpu_fields()[field]
    rng = np.random.default_rng(log_n)
    vals = rng.integers(0, f.order, size=(cols, 1 << log_n), dtype=np.uint64).astype(f.dtype)
    got = p2g.polynomial.coset_ifft(f, vals, ctx=ctx)
    for j in range(cols):
        assert (got[j] == oracle.coset_ifft(field, vals[j], f.generator)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("field,degree_bits,rate_bits,arities", [(0, 9, 3, [4, 4]), (0, 13, 1, [3, 2]), (0, 3, 0, [1]), (1, 10, 3, [4, 3])])
def test_gpu_commit_phase_computes_the_lde_values_itself(ctx, oracle, field, degree_bits, rate_bits, arities):
    """prove_openings hands fri_proof `lde_final_poly.coset_fft(g)` (fri/oracle.rs:226-231); with values=None the library
    computes that codeword on the device and the result equals the run fed with the oracle's values."""
    import plonky2_goldibear_b200 as p2g
    f = gpu_fields()[field]
    rng = np.random.default_rng(4)
    coeffs, values = lde_pair(rng, oracle, field, degree_bits, rate_bits)
    params = p2g.FriParams(p2g.FriConfig(rate_bits, 1), arities)
    ch_a, ch_b = p2g.Challenger(f, ctx=ctx), p2g.Challenger(f, ctx=ctx)
    trees_a, final_a = p2g.fri_committed_trees(f, coeffs, values, ch_a, params, ctx=ctx)
    trees_b, final_b = p2g.fri_committed_trees(f, coeffs, None, ch_b, params, ctx=ctx)
    assert (final_a == final_b).all() and ch_a.get_challenge() == ch_b.get_challenge()
    for a, b in zip(trees_a, trees_b):
        assert (a.cap.hashes == b.cap.hashes).all()
        assert (a.download("leaves") == b.download("leaves")).all()
        a.free()
        b.free()
