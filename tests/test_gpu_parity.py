"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs -- bit-exact for every coefficient, leaf, digest and cap entry."""
import json
import os

import numpy as np
import pytest

from util import BB_P, GL_P, synthetic_columns

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def fields():
    import plonky2_goldibear_b200 as p2g
    return {0: p2g.GOLDILOCKS, 1: p2g.BABYBEAR}


def to_monty(a):
    return ((a.astype(np.uint64) << np.uint64(32)) % np.uint64(BB_P)).astype(np.uint32)


def from_monty(a):
    return ((a.astype(np.uint64) * np.uint64(943718400)) % np.uint64(BB_P)).astype(np.uint32)


# ---- permutations ------------------------------------------------------------------------------------------
def test_poseidon_gl_known_answers_on_gpu(ctx):
    kat = json.load(open(os.path.join(GOLDEN, "poseidon_gl_kat.json")))
    states = np.array([[int(x, 16) for x in v["input"]] for v in kat["vectors"]], dtype=np.uint64)
    exp = np.array([[int(x, 16) for x in v["output"]] for v in kat["vectors"]], dtype=np.uint64)
    ctx.check(ctx.lib.p2g_permute(ctx.handle, 0, 0, states.ctypes.data, states.shape[0]))
    assert (states == exp).all()


@pytest.mark.parametrize("field", [0, 1])
def test_permutation_matches_oracle(ctx, oracle, field):
    rng = np.random.default_rng(11)
    p, w = oracle.ORDER[field], (12 if field == 0 else 16)
    states = rng.integers(0, p, size=(3000, w), dtype=np.uint64).astype(oracle.DTYPE[field])
    states[0] = 0
    states[1] = p - 1
    states[2] = np.arange(w)
    exp = np.stack([oracle.permute(field, s) for s in states])
    got = states.copy()
    ctx.check(ctx.lib.p2g_permute(ctx.handle, field, 0, got.ctypes.data, got.shape[0]))
    assert (got == exp).all()
    if field == 1:  # Montgomery representation in and out
        gm = to_monty(states)
        ctx.check(ctx.lib.p2g_permute(ctx.handle, 1, 1, gm.ctypes.data, gm.shape[0]))
        assert (from_monty(gm) == exp).all()
    else:  # non-canonical Goldilocks inputs (>= p) are accepted and treated mod p
        nc = states.copy()
        small = nc < np.uint64(0xFFFFFFFF)
        nc[small] += np.uint64(GL_P)
        assert small.any()
        ctx.check(ctx.lib.p2g_permute(ctx.handle, 0, 0, nc.ctypes.data, nc.shape[0]))
        assert (nc == exp).all()


# ---- Merkle tree -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("n,leaf_len,cap_height", [
    (1, 7, 0), (2, 3, 0), (2, 3, 1), (256, 7, 1), (256, 7, 8), (256, 4, 0), (256, 8, 3), (256, 9, 3), (64, 1, 2),
    (1024, 135, 4), (8192, 135, 0), (512, 32, 4), (128, 16, 7), (4096, 20, 4), (16, 0, 2),
])
def test_merkle_tree_matches_oracle(ctx, oracle, field, n, leaf_len, cap_height):
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    rng = np.random.default_rng(n * 1000 + leaf_len)
    leaves = rng.integers(0, F.order, size=(n, leaf_len), dtype=np.uint64).astype(F.dtype)
    tree = p2g.MerkleTree.new(F, leaves, cap_height, ctx=ctx)
    digests, cap = oracle.merkle_tree(field, leaves, cap_height)
    assert (tree.cap.hashes == cap).all()
    assert (tree.digests == digests).all()
    for i in {0, n - 1, n // 2, (n * 5) // 7}:  # prove() indexes the layout exactly like the reference
        sib = tree.prove(i)
        assert (sib == oracle.merkle_prove(field, digests, n, cap_height, i)).all()
        assert oracle.verify_merkle_proof_to_cap(field, leaves[i], i, cap, sib)


def test_merkle_errors_match_reference_panics(ctx):
    import plonky2_goldibear_b200 as p2g
    F = p2g.GOLDILOCKS
    with pytest.raises(ValueError):  # cap_height > log2(leaves)  (merkle_tree.rs:153-157)
        p2g.MerkleTree.new(F, np.zeros((256, 7), dtype=np.uint64), 9, ctx=ctx)
    with pytest.raises(ValueError):  # not a power of two (log2_strict)
        p2g.MerkleTree.new(F, np.zeros((100, 7), dtype=np.uint64), 0, ctx=ctx)
    # the same conditions reach the C ABI as P2G_ERR_INVALID
    from plonky2_goldibear_b200 import _lib
    import ctypes
    d = _lib.MerkleDesc()
    d.num_leaves, d.leaf_len, d.cap_height = 256, 7, 9
    buf = np.zeros((256, 7), dtype=np.uint64)
    d.leaves = buf.ctypes.data
    assert ctx.lib.p2g_merkle(ctx.handle, ctypes.byref(d)) == -1
    assert b"cap_height" in ctx.lib.p2g_last_error(ctx.handle)


def test_golden_proof_paths_through_gpu_hashes(ctx):
    """Rebuild the digests along the reference's golden Merkle paths with the GPU permutation."""
    g = json.load(open(os.path.join(GOLDEN, "gl_proof_merkle_paths.json")))
    caps = [np.array([[int(x, 16) for x in h] for h in c], dtype=np.uint64) for c in g["caps"]]
    import plonky2_goldibear_b200 as p2g
    F = p2g.GOLDILOCKS
    for p in g["paths"][:12]:
        leaf = np.array([int(x, 16) for x in p["leaf"]], dtype=np.uint64)
        # a 1-leaf tree with cap_height 0 is hash_or_noop(leaf)
        cur = p2g.MerkleTree.new(F, leaf.reshape(1, -1), 0, ctx=ctx).cap.hashes[0]
        idx = p["leaf_index"]
        for sib in p["siblings"]:
            s = np.array([int(x, 16) for x in sib], dtype=np.uint64)
            pair = np.stack([s, cur]) if idx & 1 else np.stack([cur, s])
            cur = p2g.MerkleTree.new(F, pair, 0, ctx=ctx).cap.hashes[0]  # two 4-element leaves: noop + two_to_one
            idx >>= 1
        assert (cur == caps[p["cap_id"]][idx]).all(), p["tree"]


# ---- transforms --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13, 16])
def test_ifft_matches_oracle(ctx, oracle, field, log_n):
    F = fields()[field]
    c = 5
    vals = synthetic_columns(F.order, F.dtype, c, 1 << log_n, seed=log_n)
    out = np.empty_like(vals)
    ctx.check(ctx.lib.p2g_transform(ctx.handle, field, 0, 0, log_n, c, 0, 0, vals.ctypes.data, out.ctypes.data))
    exp = np.stack([oracle.ifft(field, v) for v in vals])
    assert (out == exp).all()


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n,rate_bits", [(0, 0), (0, 3), (1, 1), (3, 3), (6, 2), (10, 3), (11, 1), (13, 3), (15, 0)])
def test_coset_lde_matches_oracle(ctx, oracle, field, log_n, rate_bits):
    F = fields()[field]
    c = 3
    n, N = 1 << log_n, 1 << (log_n + rate_bits)
    coeffs = synthetic_columns(F.order, F.dtype, c, n, seed=100 + log_n)
    out = np.empty((c, N), dtype=F.dtype)
    ctx.check(ctx.lib.p2g_transform(ctx.handle, field, 0, 1, log_n, c, rate_bits, 0, coeffs.ctypes.data, out.ctypes.data))
    for j in range(c):
        padded = np.concatenate([coeffs[j], np.zeros(N - n, dtype=F.dtype)])
        exp = oracle.coset_fft(field, padded, F.generator, rate_bits)  # lde + coset_fft_with_options(g, Some(r))
        assert (out[j] == exp).all()


# ---- the whole commit --------------------------------------------------------------------------------------
COMMIT_CASES = [
    # (field, log_n, cols, rate_bits, cap_height, blinding)
    (0, 3, 85, 3, 4, False),   # fibonacci-sized oracle: latency-bound, tiny tree
    (0, 4, 135, 3, 4, False),
    (0, 1, 2, 3, 4, False),    # N == 2^cap_height: the tree is all cap
    (0, 0, 3, 2, 2, False),
    (0, 0, 1, 0, 0, False),    # a single leaf
    (0, 12, 135, 3, 4, False),  # recursion-bench sized wires oracle
    (0, 12, 20, 3, 4, True),    # Zs + partial products with blinding salts
    (0, 10, 16, 3, 4, False),
    (0, 13, 7, 1, 0, False),
    (0, 11, 9, 0, 3, False),
    (0, 14, 135, 3, 4, False),
    (1, 12, 167, 3, 4, False),  # BabyBear narrow config wires
    (1, 12, 36, 3, 4, True),
    (1, 5, 256, 3, 4, False),
    (1, 14, 33, 2, 5, False),
    (1, 0, 9, 3, 1, False),
]


@pytest.mark.parametrize("field,log_n,c,r,h,blinding", COMMIT_CASES)
def test_from_values_matches_oracle(ctx, oracle, field, log_n, c, r, h, blinding):
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n)
    salts = synthetic_columns(F.order, F.dtype, 4, N, seed=7) if blinding else None
    timing = p2g.ProvingProcessInfo()
    batch = p2g.PolynomialBatch.from_values(F, list(vals), r, blinding, h, timing, None, ctx=ctx, salts=salts)
    ref = oracle.commit(field, vals, oracle.VALUES, r, h, salts=salts)
    assert (np.stack(batch.polynomials) == ref["coeffs"]).all()
    assert (batch.merkle_tree.leaves == ref["leaves"]).all()
    assert (batch.merkle_tree.digests == ref["digests"]).all()
    assert (batch.merkle_tree.cap.hashes == ref["cap"]).all()
    assert [s for s, _ in timing.scopes] == ["IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"]
    assert batch.degree_log == log_n and batch.rate_bits == r and batch.blinding == blinding
    # get_lde_values strips the salt and undoes the row bit reversal (oracle.rs:153-158)
    lib = oracle.load()
    for idx in (0, N // 3):
        row = batch.get_lde_values(idx, 1)
        assert (row == ref["leaves"][lib.p2o_reverse_bits(idx, log_n + r)][:c]).all()
    # from_coeffs on the returned coefficients reproduces the same tree and skips the IFFT scope
    t2 = p2g.ProvingProcessInfo()
    again = p2g.PolynomialBatch.from_coeffs(F, batch.polynomials, r, blinding, h, t2, None, ctx=ctx, salts=salts)
    assert (again.merkle_tree.cap.hashes == ref["cap"]).all() and (again.merkle_tree.digests == ref["digests"]).all()
    assert "IFFT" not in t2.as_dict()


def test_edge_inputs(ctx, oracle):
    """all zero, all p-1, single nonzero entry, the reference's i*1337 % 100 pattern."""
    import plonky2_goldibear_b200 as p2g
    for field in (0, 1):
        F = fields()[field]
        n, c = 1 << 9, 11
        cases = [np.zeros((c, n), dtype=F.dtype), np.full((c, n), F.order - 1, dtype=F.dtype)]
        one = np.zeros((c, n), dtype=F.dtype)
        one[3, 77] = 5
        cases.append(one)
        cases.append(np.tile((np.arange(n) * 1337 % 100).astype(F.dtype), (c, 1)))
        for vals in cases:
            batch = p2g.PolynomialBatch.from_values(F, list(vals), 3, False, 4, None, None, ctx=ctx)
            ref = oracle.commit(field, vals, oracle.VALUES, 3, 4)
            assert (np.stack(batch.polynomials) == ref["coeffs"]).all()
            assert (batch.merkle_tree.leaves == ref["leaves"]).all()
            assert (batch.merkle_tree.digests == ref["digests"]).all()
            assert (batch.merkle_tree.cap.hashes == ref["cap"]).all()


def test_babybear_montgomery_representation(ctx, oracle):
    import plonky2_goldibear_b200 as p2g
    from plonky2_goldibear_b200.fields import REPR_MONTY_R32
    F = p2g.BABYBEAR
    vals = synthetic_columns(F.order, F.dtype, 20, 1 << 10)
    ref = oracle.commit(1, vals, oracle.VALUES, 2, 3)
    batch = p2g.PolynomialBatch.from_values(F, list(to_monty(vals)), 2, False, 3, None, None, ctx=ctx, repr=REPR_MONTY_R32)
    assert (from_monty(np.stack(batch.polynomials)) == ref["coeffs"]).all()
    assert (from_monty(batch.merkle_tree.leaves) == ref["leaves"]).all()
    assert (from_monty(batch.merkle_tree.digests) == ref["digests"]).all()
    assert (from_monty(batch.merkle_tree.cap.hashes) == ref["cap"]).all()


def test_commit_errors_match_reference_panics(ctx):
    import plonky2_goldibear_b200 as p2g
    F = p2g.GOLDILOCKS
    a, b = np.zeros(16, dtype=np.uint64), np.zeros(32, dtype=np.uint64)
    with pytest.raises(ValueError, match="Polynomial degrees inconsistent"):  # oracle.rs:139
        p2g.PolynomialBatch.from_coeffs(F, [a, b], 3, False, 4, ctx=ctx)
    with pytest.raises(ValueError, match="Not a power of two"):
        p2g.PolynomialBatch.from_values(F, [np.zeros(24, dtype=np.uint64)], 3, False, 4, ctx=ctx)
    with pytest.raises(IndexError):
        p2g.PolynomialBatch.from_values(F, [], 3, False, 4, ctx=ctx)
    with pytest.raises(p2g.P2GError) as e:  # cap_height > log2(N)
        p2g.PolynomialBatch.from_values(F, [a], 1, False, 6, ctx=ctx)
    assert e.value.code == -1


@pytest.mark.parametrize("field,log_n,c,r,h,g", [(0, 10, 17, 3, 4, 2), (0, 10, 17, 3, 4, 8), (0, 9, 5, 3, 1, 4),
                                                 (0, 9, 5, 3, 0, 8), (1, 11, 40, 3, 4, 4), (1, 8, 9, 2, 0, 4)])
def test_coset_sharded_commit_assembles_to_the_same_tree(ctx, oracle, field, log_n, c, r, h, g):
    """The multi-GPU partition emulated on one GPU: every shard computes only its cosets; leaves,
    digests and cap assemble to exactly the single-GPU / CPU result, with only shard roots crossing shards."""
    import plonky2_goldibear_b200 as p2g
    from plonky2_goldibear_b200.fri_oracle import assemble_sharded
    F = fields()[field]
    vals = synthetic_columns(F.order, F.dtype, c, 1 << log_n, seed=5)
    ref = oracle.commit(field, vals, oracle.VALUES, r, h)
    shards = [p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx, shard_index=i, shard_count=g)
              for i in range(g)]
    whole = assemble_sharded(F, shards, h, ctx=ctx)
    assert (whole.merkle_tree.leaves == ref["leaves"]).all()
    assert (whole.merkle_tree.cap.hashes == ref["cap"]).all()
    assert (whole.merkle_tree.digests == ref["digests"]).all()


# ---- size-independent properties at a larger size ----------------------------------------------------------
def test_large_commit_properties(ctx, oracle):
    """2^18 x 24 Goldilocks (leaves 403 MB): checked by properties instead of a full CPU recomputation:
    sampled Merkle paths verify against the cap with the oracle hashes, sampled rows equal direct evaluation
    of the returned coefficients, and iNTT followed by the un-shifted evaluation reproduces the input."""
    import plonky2_goldibear_b200 as p2g
    F = p2g.GOLDILOCKS
    log_n, c, r, h = 18, 24, 3, 4
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=9)
    batch = p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx)
    lib = oracle.load()
    rng = np.random.default_rng(2)
    coeffs = np.stack(batch.polynomials)
    for j in (0, c - 1):  # round trip: fft(ifft(v)) == v on the CPU oracle
        assert (oracle.fft(0, coeffs[j]) == vals[j]).all()
    wN = lib.p2o_two_adic_generator(0, log_n + r)
    for k in rng.integers(0, N, size=6):
        k = int(k)
        sib = batch.merkle_tree.prove(k)
        assert oracle.verify_merkle_proof_to_cap(0, batch.merkle_tree.leaves[k], k, batch.merkle_tree.cap.hashes, sib)
        x = 7 * pow(wN, lib.p2o_reverse_bits(k, log_n + r), GL_P) % GL_P
        col = int(rng.integers(0, c))
        acc = 0
        for cf in coeffs[col][::-1].tolist():
            acc = (acc * x + cf) % GL_P
        assert acc == int(batch.merkle_tree.leaves[k][col])


# ---- FRI commit-phase tree shapes (fri/prover.rs:83-133 calls MerkleTree::new on these) ------------
@pytest.mark.parametrize("field,d_ext,arity_bits,log_n", [(0, 2, 4, 14), (0, 2, 1, 12), (1, 4, 4, 13), (1, 4, 1, 10), (1, 4, 3, 9)])
def test_fri_layer_tree_shapes(ctx, oracle, field, d_ext, arity_bits, log_n):
    """Leaves of a FRI layer tree: the extension-field codeword in bit-reversed order, chunked by the arity and
    flattened to base elements (leaf length arity*D); arity 2 leaves have 4 (GL) / 8 (BB) elements and take the
    no-hash branch of hash_or_noop.  cap_height as in the FRI config (4), clamped like the prover does."""
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    arity = 1 << arity_bits
    rng = np.random.default_rng(log_n)
    codeword = rng.integers(0, F.order, size=(1 << log_n, d_ext), dtype=np.uint64).astype(F.dtype)
    idx = np.array([oracle.load().p2o_reverse_bits(i, log_n) for i in range(1 << log_n)])
    leaves = codeword[idx].reshape((1 << log_n) // arity, arity * d_ext)
    cap_height = min(4, log_n - arity_bits)
    tree = p2g.MerkleTree.new(F, leaves, cap_height, ctx=ctx)
    digests, cap = oracle.merkle_tree(field, leaves, cap_height)
    assert (tree.digests == digests).all() and (tree.cap.hashes == cap).all()


# ---- BASELINE.json full sizes, checked through size-independent properties -------------------------------------
def _check_commit_properties(ctx, oracle, field, log_n, c, r, h, samples=5, seed=3):
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    p = F.order
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(p, F.dtype, c, n, seed=seed)
    batch = p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx)
    lib = oracle.load()
    rng = np.random.default_rng(seed)
    tree = batch.merkle_tree
    assert tree.leaves.shape == (N, c) and tree.digests.shape == (2 * (N - (1 << h)), F.num_hash_out_elts)
    # (1) interpolation round trip on sampled columns: the CPU FFT of the returned coefficients gives back the input
    for j in {0, c // 2, c - 1}:
        assert (oracle.fft(field, batch.polynomials[j]) == vals[j]).all()
    # (2) sampled rows are the evaluations at g * w_N^bitrev(k) (Horner on the CPU), (3) their Merkle paths verify
    wN = lib.p2o_two_adic_generator(field, log_n + r)
    for k in [0, N - 1] + [int(x) for x in rng.integers(0, N, size=samples)]:
        sib = tree.prove(k)
        assert oracle.verify_merkle_proof_to_cap(field, tree.leaves[k], k, tree.cap.hashes, sib)
        x = F.generator * pow(wN, lib.p2o_reverse_bits(k, log_n + r), p) % p
        col = int(rng.integers(0, c))
        acc = 0
        for cf in batch.polynomials[col][::-1].tolist():
            acc = (acc * x + cf) % p
        assert acc == int(tree.leaves[k][col])
    # (4) the cap is the root layer of the digests: recompute the top of every cap subtree with the oracle's compression
    sub_len = tree.digests.shape[0] >> h
    for s in {0, (1 << h) - 1}:
        blk = tree.digests[s * sub_len:(s + 1) * sub_len]
        mid = sub_len // 2
        assert (oracle.two_to_one(field, blk[mid - 1], blk[mid]) == tree.cap.hashes[s]).all()
    # (5) linearity of the LDE: leaves(a) + leaves(b) == leaves(a + b) on a sampled row, via a 2-column side commit
    a, b = vals[0].astype(np.uint64), vals[1].astype(np.uint64)
    s_col = ((a + b) % np.uint64(p)).astype(F.dtype) if p < (1 << 62) else np.array(
        [(int(x) + int(y)) % p for x, y in zip(a[:0], b[:0])], dtype=F.dtype)
    if p < (1 << 62):
        side = p2g.PolynomialBatch.from_values(F, [s_col], r, False, 0, ctx=ctx)
        for k in (1, N // 2 + 3):
            assert (int(tree.leaves[k][0]) + int(tree.leaves[k][1])) % p == int(side.merkle_tree.leaves[k][0])


def test_config2_full_size_goldilocks(ctx, oracle):
    """BASELINE config 2: Goldilocks 2^20 x 135, rate_bits 3, cap_height 4 (leaves 9.06 GB)."""
    _check_commit_properties(ctx, oracle, 0, 20, 135, 3, 4)


def test_config3_babybear_reduced_rows(ctx, oracle):
    """BASELINE config 3 shape (BabyBear, 256 columns, rate_bits 3, cap_height 4) at 2^20 rows, every output on the host
    (8.6 GB of leaves); the full 2^22 rows are checked device-resident in the next test."""
    _check_commit_properties(ctx, oracle, 1, 20, 256, 3, 4)


def test_config3_full_size_babybear_device_resident(ctx, oracle):
    """BASELINE config 3 at its real size: BabyBear 2^22 x 256, rate_bits 3, cap_height 4 (34 GB of leaves, two-pass NTT
    with 2^11-point tiles).  The outputs stay on the device (P2G_FLAG_KEEP_ON_DEVICE; the host would need 41 GB for
    them); sampled rows, their Merkle paths and sampled coefficient columns come back through the batch accessors and
    are checked against the definition: row k = P(g * w_N^bitrev(k)), the path hashes up to the cap with the oracle's
    sponge / compression, and the CPU FFT of a returned coefficient column gives back the input column."""
    import plonky2_goldibear_b200 as p2g
    F = fields()[1]
    p = F.order
    log_n, c, r, h = 22, 256, 3, 4
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = np.empty((c, n), dtype=F.dtype)
    for cb in range(0, c, 8):
        vals[cb:cb + 8] = synthetic_columns(p, F.dtype, 8, n, seed=3, col_begin=cb)
    batch = p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx, keep_on_device=True)
    tree = batch.merkle_tree
    try:
        lib = oracle.load()
        rng = np.random.default_rng(22)
        ks = [0, N - 1, n, n - 1] + [int(x) for x in rng.integers(0, N, size=4)]
        rows = tree.get_many(ks)
        sibs = tree.prove_many(ks)
        assert rows.shape == (len(ks), c) and sibs.shape == (len(ks), log_n + r - h, F.num_hash_out_elts)
        for row, sib, k in zip(rows, sibs, ks):
            assert oracle.verify_merkle_proof_to_cap(1, row, k, tree.cap.hashes, sib)
        coeffs = batch.download_polynomials()
        for j in (0, 131, c - 1):
            assert (oracle.fft(1, coeffs[j]) == vals[j]).all()
        wN = lib.p2o_two_adic_generator(1, log_n + r)
        for row, k in list(zip(rows, ks))[:5]:
            x = F.generator * pow(wN, lib.p2o_reverse_bits(k, log_n + r), p) % p
            col = int(rng.integers(0, c))
            # P(x) with vectorised powers: x^j for j < n by doubling, then a dot product below 2^53
            pw = np.ones(n, dtype=np.uint64)
            step, length = x, 1
            while length < n:
                pw[length:2 * length] = pw[:length] * np.uint64(step) % np.uint64(p)
                step = step * step % p
                length *= 2
            acc = int(np.sum(coeffs[col].astype(np.uint64) * pw % np.uint64(p), dtype=np.uint64)) % p
            assert acc == int(row[col])
    finally:
        tree.free()


def test_three_pass_ntt_plan_at_production_size(ctx, oracle):
    """2^23 rows is a genuine three-pass transform (8/8/7 tiles): bit-exact against the oracle on a few columns --
    coefficients, every leaf, every digest, the cap.  (The forced-small-tile test below walks the same code at sizes
    where every shape can be tried; this one pins the production plan itself.)"""
    import plonky2_goldibear_b200 as p2g
    F = fields()[0]
    log_n, c, r, h = 23, 5, 1, 3
    vals = synthetic_columns(F.order, F.dtype, c, 1 << log_n, seed=23)
    batch = p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx)
    ref = oracle.commit(0, vals, oracle.VALUES, r, h)
    assert (np.stack(batch.polynomials) == ref["coeffs"]).all()
    assert (batch.merkle_tree.leaves == ref["leaves"]).all()
    assert (batch.merkle_tree.cap.hashes == ref["cap"]).all() and (batch.merkle_tree.digests == ref["digests"]).all()


def test_wide_tile_two_pass_ntt_plan(ctx, oracle):
    """log n = 21, 22 are two-pass transforms with 2^11-point tiles (11/10, 11/11; half-width runs): full parity on a few
    columns."""
    import plonky2_goldibear_b200 as p2g
    for field, log_n in ((0, 21), (1, 22)):
        F = fields()[field]
        vals = synthetic_columns(F.order, F.dtype, 3, 1 << log_n, seed=log_n)
        batch = p2g.PolynomialBatch.from_values(F, list(vals), 1, False, 2, ctx=ctx)
        ref = oracle.commit(field, vals, oracle.VALUES, 1, 2)
        assert (np.stack(batch.polynomials) == ref["coeffs"]).all()
        assert (batch.merkle_tree.leaves == ref["leaves"]).all()
        assert (batch.merkle_tree.cap.hashes == ref["cap"]).all() and (batch.merkle_tree.digests == ref["digests"]).all()


@pytest.mark.parametrize("max_tile,field,log_n,c", [(3, 0, 9, 9), (4, 1, 11, 17), (5, 0, 13, 3), (2, 1, 6, 20), (6, 0, 12, 8)])
def test_three_pass_ntt_plan_forced_small_tiles(ctx, oracle, max_tile, field, log_n, c):
    """The three-pass plan (2^23 rows and up in production) forced at small sizes through the library's test hook
    P2G_TEST_MAX_TILE_LOG: inverse NTT, coset LDE, leaf order and tree bit-exact against the oracle."""
    import os
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    vals = synthetic_columns(F.order, F.dtype, c, 1 << log_n, seed=100 + log_n)
    ref = oracle.commit(field, vals, oracle.VALUES, 2, 3)
    os.environ["P2G_TEST_MAX_TILE_LOG"] = str(max_tile)
    try:
        batch = p2g.PolynomialBatch.from_values(F, list(vals), 2, False, 3, ctx=ctx)
    finally:
        del os.environ["P2G_TEST_MAX_TILE_LOG"]
    assert (np.stack(batch.polynomials) == ref["coeffs"]).all()
    assert (batch.merkle_tree.leaves == ref["leaves"]).all()
    assert (batch.merkle_tree.cap.hashes == ref["cap"]).all() and (batch.merkle_tree.digests == ref["digests"]).all()


# ---- device-resident batch: get / prove / get_lde_values served from GPU memory ---------------------
@pytest.mark.parametrize("field,log_n,c,r,h,blinding,kind", [(0, 12, 135, 3, 4, False, "values"), (0, 10, 20, 3, 4, True, "coeffs"),
                                                            (1, 11, 167, 3, 4, False, "values"), (0, 1, 3, 3, 4, False, "values")])
def test_device_resident_batch_accessors(ctx, oracle, field, log_n, c, r, h, blinding, kind):
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n)
    salts = synthetic_columns(F.order, F.dtype, 4, N, seed=7) if blinding else None
    ref = oracle.commit(field, vals, oracle.VALUES if kind == "values" else oracle.COEFFS, r, h, salts=salts)
    fn = p2g.PolynomialBatch.from_values if kind == "values" else p2g.PolynomialBatch.from_coeffs
    batch = fn(F, list(vals), r, blinding, h, ctx=ctx, salts=salts, keep_on_device=True)
    tree = batch.merkle_tree
    assert (tree.cap.hashes == ref["cap"]).all()
    rng = np.random.default_rng(1)
    idx = [0, N - 1] + [int(x) for x in rng.integers(0, N, size=20)]
    assert (tree.get_many(idx) == ref["leaves"][idx]).all()
    proofs = tree.prove_many(idx)
    for i, k in enumerate(idx):
        assert (proofs[i] == oracle.merkle_prove(field, ref["digests"], N, h, k)).all()
        assert oracle.verify_merkle_proof_to_cap(field, ref["leaves"][k], k, ref["cap"], proofs[i])
    lib = oracle.load()
    for i in (0, 5 % N, N - 1):
        assert (batch.get_lde_values(i, 1) == ref["leaves"][lib.p2o_reverse_bits(i, log_n + r)][:c]).all()
    assert (tree.download("leaves") == ref["leaves"]).all()
    assert (tree.download("digests") == ref["digests"]).all()
    assert (batch.download_polynomials() == ref["coeffs"]).all()
    with pytest.raises(p2g.P2GError):
        tree.get(N)  # out of range, like the reference's slice index panic
    tree.free()


def test_fused_intt_allgather_emulated_on_one_gpu(ctx, oracle):
    """The fused inverse-NTT + all-gather (p2g_intt_allgather) with both 'ranks' on one GPU: each rank transforms its
    column share and its epilogue writes the coefficients into BOTH gather buffers; afterwards every buffer holds every
    column, bit-identical to the oracle's ifft.  (Across GPUs the peer buffers are CUDA-IPC mappings; bench.py --gpus N.)"""
    import ctypes
    import torch
    for field, log_n, c in ((0, 12, 7), (1, 13, 10), (0, 5, 3)):
        F = fields()[field]
        n = 1 << log_n
        world, cpr = 2, (c + 1) // 2
        vals = synthetic_columns(F.order, F.dtype, c, n, seed=21)
        tdt = torch.int64 if field == 0 else torch.int32
        bufs = [torch.zeros((world * cpr, n), dtype=tdt, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for rank in range(world):
            c0, c1 = rank * cpr, min((rank + 1) * cpr, c)
            share = torch.from_numpy(np.ascontiguousarray(vals[c0:c1]).view(np.int64 if field == 0 else np.int32)).cuda()
            torch.cuda.synchronize()
            order = [bufs[rank]] + [bufs[r] for r in range(world) if r != rank]
            ptrs = (ctypes.c_void_p * world)(*[b.data_ptr() for b in order])
            ctx.check(ctx.lib.p2g_intt_allgather(ctx.handle, field, 0, log_n, c1 - c0, c0, share.data_ptr(), ptrs, world))
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(ctx.lib.p2g_stream(ctx.handle))
        stream.synchronize()
        exp = np.stack([oracle.ifft(field, v) for v in vals])
        for b in bufs:
            got = b.cpu().numpy().view(F.dtype)[:c]
            assert (got == exp).all()


@pytest.mark.parametrize("field,log_n,c,r,h,g,groups", [(0, 11, 19, 3, 4, 1, (8, 8, 3)), (0, 10, 21, 3, 4, 4, (16, 5)),
                                                        (1, 11, 40, 3, 4, 2, (32, 8)), (0, 12, 9, 2, 1, 1, (9,))])
def test_commit_split_over_column_ranges(ctx, oracle, field, log_n, c, r, h, g, groups):
    """P2G_FLAG_LEAVES_ONLY / LEAVES_READY: the LDE of column ranges written straight into one device-resident leaf
    matrix (as the pipelined multi-GPU exchange does, group by group), then the tree over the finished rows -- leaves,
    digests and cap equal the one-shot commit's (and the oracle's) for the shard."""
    import ctypes
    import torch
    from plonky2_goldibear_b200 import _lib
    F = fields()[field]
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=33)
    ref = oracle.commit(field, vals, oracle.VALUES, r, h)
    coeffs = np.ascontiguousarray(ref["coeffs"])
    tdt, ndt = (torch.int64, np.int64) if field == 0 else (torch.int32, np.int32)
    d_el = F.num_hash_out_elts
    lg = g.bit_length() - 1
    rows_local, cap_local = N // g, max(h - lg, 0)
    d_coeffs = torch.from_numpy(coeffs.view(ndt)).cuda()
    keep = []
    for shard in range(g):
        d_leaves = torch.zeros((rows_local, c), dtype=tdt, device="cuda")
        d_dig = torch.zeros((max(2 * (rows_local - (1 << cap_local)), 1), d_el), dtype=tdt, device="cuda")
        d_cap = torch.zeros((1 << cap_local, d_el), dtype=tdt, device="cuda")
        torch.cuda.synchronize()

        def desc_for(num_cols, flags):
            d = _lib.CommitDesc()
            d.field, d.hash, d.repr, d.input_kind = F.field_id, F.hash_id, 0, 1
            d.log_n, d.num_cols, d.rate_bits, d.cap_height = log_n, num_cols, r, h
            d.flags = _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT | flags
            d.shard_index, d.shard_count = (shard, g) if g > 1 else (0, 0)
            d.leaf_pitch = c
            return d
        col = 0
        for gc in groups:
            d = desc_for(gc, _lib.FLAG_LEAVES_ONLY)
            arr = (ctypes.c_void_p * 1)(d_coeffs.data_ptr() + col * n * F.elem_bytes)
            keep.append(arr)
            d.col_ptrs = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
            d.leaves_out = d_leaves.data_ptr() + col * F.elem_bytes
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(d)))
            col += gc
        assert col == c
        d = desc_for(c, _lib.FLAG_LEAVES_READY)
        d.leaves_out, d.digests_out, d.cap_out = d_leaves.data_ptr(), d_dig.data_ptr(), d_cap.data_ptr()
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(d)))
        torch.cuda.synchronize()
        r0 = shard * rows_local
        assert (d_leaves.cpu().numpy().view(F.dtype) == ref["leaves"][r0:r0 + rows_local]).all()
        if g == 1:
            assert (d_cap.cpu().numpy().view(F.dtype) == ref["cap"]).all()
            assert (d_dig.cpu().numpy().view(F.dtype)[:len(ref["digests"])] == ref["digests"]).all()
        else:
            one = ctx_commit_shard(ctx, F, vals, r, h, shard, g)
            assert (d_cap.cpu().numpy().view(F.dtype) == one.merkle_tree.cap.hashes).all()
            assert (d_dig.cpu().numpy().view(F.dtype)[:len(one.merkle_tree.digests)] == one.merkle_tree.digests).all()


@pytest.mark.parametrize("field,log_n,c,r,h,g,tile_log", [(0, 12, 19, 3, 4, 1, 0), (0, 12, 135, 3, 4, 8, 0), (0, 9, 21, 3, 4, 4, 0),
                                                          (1, 13, 40, 3, 4, 2, 0), (0, 12, 11, 2, 2, 2, 4), (1, 12, 33, 3, 3, 4, 5)])
def test_lde_split_by_pass(ctx, oracle, monkeypatch, field, log_n, c, r, h, g, tile_log):
    """p2g_lde_columns: the strided passes block by block in the order a ring exchange delivers the blocks (own block
    first, then the others), the last pass once over all columns, then the tree (P2G_FLAG_LEAVES_READY) -- leaves,
    digests and cap of every shard equal the oracle's.  Covers one-, two- and (through the tile-size hook) three-pass
    plans, ragged and empty blocks, several cosets per shard."""
    import ctypes
    import torch
    from plonky2_goldibear_b200 import _lib
    from plonky2_goldibear_b200.sharding import ring_blocks
    if tile_log:
        monkeypatch.setenv("P2G_TEST_MAX_TILE_LOG", str(tile_log))
    F = fields()[field]
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=35)
    ref = oracle.commit(field, vals, oracle.VALUES, r, h)
    coeffs = np.ascontiguousarray(ref["coeffs"])
    tdt, ndt = (torch.int64, np.int64) if field == 0 else (torch.int32, np.int32)
    d_el = F.num_hash_out_elts
    lg = g.bit_length() - 1
    rows_local, cap_local = N // g, max(h - lg, 0)
    d_coeffs = torch.from_numpy(coeffs.view(ndt)).cuda()
    # odd column counts: even shares to the column (what bench.py uses); the others: whole NTT column tiles
    blocks = ring_blocks(c, max(g, 4), 1 if c % 2 else (8 if field == 0 else 16))
    es = F.elem_bytes
    nd, per_cap = len(ref["digests"]) // g, (1 << h) // g
    for shard in range(g):
        d_leaves = torch.zeros((rows_local, c + 3), dtype=tdt, device="cuda")  # a pitch wider than the rows
        d_dig = torch.zeros((max(2 * (rows_local - (1 << cap_local)), 1), d_el), dtype=tdt, device="cuda")
        d_cap = torch.zeros((1 << cap_local, d_el), dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        order = [(shard - s) % len(blocks) for s in range(len(blocks))]
        for b in order:
            lo, hi = blocks[b]
            if hi > lo:
                ctx.check(ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, log_n, r, shard, g if g > 1 else 0, c, lo, hi - lo,
                                                  d_coeffs.data_ptr() + lo * n * es, None, 0, 0))
        ctx.check(ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, log_n, r, shard, g if g > 1 else 0, c, 0, c,
                                          d_coeffs.data_ptr(), d_leaves.data_ptr(), c + 3, 1))
        d = _lib.CommitDesc()
        d.field, d.hash, d.repr, d.input_kind = F.field_id, F.hash_id, 0, 1
        d.log_n, d.num_cols, d.rate_bits, d.cap_height = log_n, c, r, h
        d.flags = _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT | _lib.FLAG_LEAVES_READY
        d.shard_index, d.shard_count = (shard, g) if g > 1 else (0, 0)
        d.leaf_pitch = c + 3
        d.leaves_out, d.digests_out, d.cap_out = d_leaves.data_ptr(), d_dig.data_ptr(), d_cap.data_ptr()
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(d)))
        torch.cuda.synchronize()
        r0 = shard * rows_local
        assert (d_leaves.cpu().numpy().view(F.dtype)[:, :c] == ref["leaves"][r0:r0 + rows_local]).all()
        assert (d_dig.cpu().numpy().view(F.dtype)[:nd] == ref["digests"][shard * nd:(shard + 1) * nd]).all()
        assert (d_cap.cpu().numpy().view(F.dtype) == ref["cap"][shard * per_cap:(shard + 1) * per_cap]).all()
    # misuse is refused
    assert ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, log_n, r, 0, 0, c, 1, c, d_coeffs.data_ptr(), None, 0, 0) != 0
    assert ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, log_n, r, 0, 0, c, 0, c, d_coeffs.data_ptr(), None, 0, 1) != 0


def ctx_commit_shard(ctx, F, vals, r, h, shard, g):
    import plonky2_goldibear_b200 as p2g
    return p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=ctx, shard_index=shard, shard_count=g)


@pytest.mark.parametrize("field,log_n,c,r,h,kind,g,pitch_extra", [
    (0, 11, 41, 3, 4, "values", 1, 0), (0, 11, 41, 3, 4, "coeffs", 1, 3), (0, 12, 19, 3, 2, "values", 8, 0),
    (1, 11, 70, 3, 4, "values", 1, 2), (1, 12, 40, 2, 3, "coeffs", 4, 0), (0, 10, 135, 3, 4, "values", 2, 1)])
def test_pinned_pipelined_commit_matches_oracle(ctx, oracle, field, log_n, c, r, h, kind, g, pitch_extra):
    """The end-to-end headline path: PINNED host columns (more than two tiles of them) take the pipelined route of
    p2g_commit -- column groups uploaded on the h2d stream under the transforms of the previous group, coefficients and
    finished cosets sent home early on the copy stream.  Everything that comes back (coefficients, leaves at a pitch,
    digests, cap) must equal the oracle's, also when the same context is reused: a wrong event order would return stale
    host leaves or coefficients under a correct (device-computed) cap."""
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    n = 1 << log_n
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=77 + c)
    pinned = ctx.pinned_empty((c, n), F.dtype)
    for rep in range(3):
        if rep:
            vals = synthetic_columns(F.order, F.dtype, c, n, seed=1000 + rep)
        pinned[...] = vals
        okind = oracle.VALUES if kind == "values" else oracle.COEFFS
        ref = oracle.commit(field, vals, okind, r, h)
        lg = g.bit_length() - 1
        rows_local = (n << r) // g
        for shard in range(g if rep == 0 else 1):
            fn = p2g.PolynomialBatch.from_values if kind == "values" else p2g.PolynomialBatch.from_coeffs
            b = fn(F, list(pinned), r, False, h, ctx=ctx, pinned_out=True, leaf_pitch=c + pitch_extra,
                   shard_index=shard if g > 1 else 0, shard_count=g if g > 1 else 0)
            r0 = shard * rows_local
            assert (b.merkle_tree.leaves == ref["leaves"][r0:r0 + rows_local]).all()
            if kind == "values":
                assert (np.stack(b.polynomials) == ref["coeffs"]).all()
            if g == 1:
                assert (b.merkle_tree.digests == ref["digests"]).all()
                assert (b.merkle_tree.cap.hashes == ref["cap"]).all()
            elif h >= lg:
                per = (1 << h) // g
                assert (b.merkle_tree.cap.hashes == ref["cap"][shard * per:(shard + 1) * per]).all()
                nd = len(ref["digests"]) // g
                assert (b.merkle_tree.digests == ref["digests"][shard * nd:(shard + 1) * nd]).all()


@pytest.mark.parametrize("field,log_n,c", [(0, 10, 11), (1, 9, 20)])
def test_device_input_with_host_salts(ctx, oracle, field, log_n, c):
    """P2G_FLAG_DEVICE_INPUT + blinding: the columns are device resident, the salt columns come from the HOST (the RNG
    stays on the caller's side, fri/oracle.rs:144-148) -- the salt upload resolves its direction from the pointer."""
    import ctypes
    import torch
    from plonky2_goldibear_b200 import _lib
    F = fields()[field]
    r, h = 3, 2
    n, N = 1 << log_n, 1 << (log_n + r)
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=5)
    salts = synthetic_columns(F.order, F.dtype, 4, N, seed=6)
    ref = oracle.commit(field, vals, oracle.VALUES, r, h, salts=salts)
    ndt = np.int64 if field == 0 else np.int32
    d_vals = torch.from_numpy(np.ascontiguousarray(vals).view(ndt)).cuda()
    width = c + 4
    d_el = F.num_hash_out_elts
    leaves = np.empty((N, width), dtype=F.dtype)
    digests = np.empty((2 * (N - (1 << h)), d_el), dtype=F.dtype)
    cap = np.empty((1 << h, d_el), dtype=F.dtype)
    coeffs = np.empty((c, n), dtype=F.dtype)
    for salts_on_device in (False, True):
        d = _lib.CommitDesc()
        d.field, d.hash, d.repr, d.input_kind = F.field_id, F.hash_id, 0, 0
        d.log_n, d.num_cols, d.rate_bits, d.cap_height = log_n, c, r, h
        d.num_salt_cols = 4
        d.flags = _lib.FLAG_DEVICE_INPUT
        col = (ctypes.c_void_p * 1)(d_vals.data_ptr())
        d.col_ptrs = ctypes.cast(col, ctypes.POINTER(ctypes.c_void_p))
        d_salts = torch.from_numpy(np.ascontiguousarray(salts).view(ndt)).cuda() if salts_on_device else None
        sp = (ctypes.c_void_p * 4)(*[(d_salts.data_ptr() + j * N * F.elem_bytes) if salts_on_device else salts[j].ctypes.data
                                     for j in range(4)])
        d.salt_ptrs = ctypes.cast(sp, ctypes.POINTER(ctypes.c_void_p))
        outp = (ctypes.c_void_p * c)(*[coeffs[j].ctypes.data for j in range(c)])
        d.coeffs_out = ctypes.cast(outp, ctypes.POINTER(ctypes.c_void_p))
        d.leaves_out, d.leaf_pitch = leaves.ctypes.data, width
        d.digests_out, d.cap_out = digests.ctypes.data, cap.ctypes.data
        leaves[...] = 0
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(d)))
        assert (leaves == ref["leaves"]).all() and (digests == ref["digests"]).all() and (cap == ref["cap"]).all()
        assert (coeffs == ref["coeffs"]).all()


# ---- multi-device context: one process, p2g_init(device_ids, n), the library shards the commit -------
def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0, 0]]  # the partition on one GPU: the shards share it
    for g in (2, 4, 8):
        if n >= g:
            lists.append(list(range(g)))
    return lists


@pytest.mark.parametrize("field,log_n,c,r,h,kind", [(0, 12, 135, 3, 4, "values"), (1, 11, 70, 3, 3, "values"),
                                                    (0, 11, 9, 3, 4, "coeffs"), (0, 13, 3, 2, 2, "values")])
def test_multi_device_context_commit_matches_oracle(oracle, field, log_n, c, r, h, kind):
    """p2g_commit on a multi-device context: host columns in, the whole job's coefficients / leaves / digests / cap out,
    bit-exact against the oracle -- the column-sharded inverse NTT with the all-gather fused into its epilogue (peer
    stores), the event barrier, the coset-sharded LDE + trees writing their slices of the caller's buffers.  On real
    multi-GPU boxes the device lists [0..g) run over NVLink; [0, 0] and [0, 0, 0, 0] run the same code on one GPU."""
    import plonky2_goldibear_b200 as p2g
    F = fields()[field]
    n = 1 << log_n
    vals = synthetic_columns(F.order, F.dtype, c, n, seed=51)
    okind = oracle.VALUES if kind == "values" else oracle.COEFFS
    ref = oracle.commit(field, vals, okind, r, h)
    for devs in _device_lists():
        if len(devs) > (1 << r) or (1 << h) < len(devs):
            continue
        mctx = p2g.Context(devs)
        try:
            assert mctx.lib.p2g_num_devices(mctx.handle) == len(devs)
            fn = p2g.PolynomialBatch.from_values if kind == "values" else p2g.PolynomialBatch.from_coeffs
            for rep in range(2):  # the context is reused: gather buffers and events must be clean for the next call
                b = fn(F, list(vals), r, False, h, ctx=mctx, pinned_out=bool(rep))
                if kind == "values":
                    assert (np.stack(b.polynomials) == ref["coeffs"]).all(), devs
                assert (b.merkle_tree.leaves == ref["leaves"]).all(), devs
                assert (b.merkle_tree.digests == ref["digests"]).all(), devs
                assert (b.merkle_tree.cap.hashes == ref["cap"]).all(), devs
            # device-resident variant: the batch spans the devices, accessors route by leaf index
            kb = fn(F, list(vals), r, False, h, ctx=mctx, keep_on_device=True)
            tree = kb.merkle_tree
            N = n << r
            ks = [0, N - 1, N // len(devs), N // len(devs) - 1, 12345 % N]
            rows = tree.get_many(ks)
            sibs = tree.prove_many(ks)
            assert (tree.cap.hashes == ref["cap"]).all()
            for row, sib, k in zip(rows, sibs, ks):
                assert (row == ref["leaves"][k]).all()
                assert oracle.verify_merkle_proof_to_cap(field, row, k, ref["cap"], sib)
            assert (tree.download("digests") == ref["digests"]).all() and (tree.download("leaves") == ref["leaves"]).all()
            if kind == "values":
                assert (kb.download_polynomials() == ref["coeffs"]).all()
            tree.free()
        finally:
            mctx.close()


@pytest.mark.parametrize("mode", ["fused", "push", "ce"])
def test_multi_device_exchange_modes_match_oracle(oracle, monkeypatch, mode):
    """The three ways the column-sharded inverse NTT hands its coefficients to the other devices (peer stores fused into
    the last pass, a 16-byte-per-lane copy kernel behind it, the copy engines -- P2G_EXCHANGE), each forced, on 2 and 4
    shards: identical outputs.  Odd sizes (a 2-row Goldilocks column is one 16-byte piece, a 1-row one is not sendable
    as a block and must fall back to the fused stores)."""
    import plonky2_goldibear_b200 as p2g
    monkeypatch.setenv("P2G_EXCHANGE", mode)
    for field, log_n, c, r, h in [(0, 12, 135, 3, 4), (1, 11, 70, 3, 3), (0, 1, 5, 2, 2), (0, 0, 7, 2, 2), (1, 1, 6, 2, 2)]:
        F = fields()[field]
        vals = synthetic_columns(F.order, F.dtype, c, 1 << log_n, seed=77)
        ref = oracle.commit(field, vals, oracle.VALUES, r, h)
        for devs in _device_lists():
            if len(devs) > (1 << r) or (1 << h) < len(devs):
                continue
            mctx = p2g.Context(devs)
            try:
                b = p2g.PolynomialBatch.from_values(F, list(vals), r, False, h, ctx=mctx)
                assert (np.stack(b.polynomials) == ref["coeffs"]).all(), (mode, devs)
                assert (b.merkle_tree.leaves == ref["leaves"]).all(), (mode, devs)
                assert (b.merkle_tree.cap.hashes == ref["cap"]).all(), (mode, devs)
            finally:
                mctx.close()


def test_multi_device_context_rejects_what_it_cannot_shard(oracle):
    import plonky2_goldibear_b200 as p2g
    F = fields()[0]
    vals = synthetic_columns(F.order, F.dtype, 3, 1 << 6, seed=1)
    mctx = p2g.Context([0, 0, 0, 0])
    try:
        with pytest.raises(p2g.P2GError) as e:  # cap entries would span devices
            p2g.PolynomialBatch.from_values(F, list(vals), 3, False, 1, ctx=mctx)
        assert e.value.code == -5
        with pytest.raises(p2g.P2GError) as e:  # more devices than cosets
            p2g.PolynomialBatch.from_values(F, list(vals), 1, False, 4, ctx=mctx)
        assert e.value.code == -5
    finally:
        mctx.close()
    with pytest.raises(p2g.P2GError):
        p2g.Context([0, 0, 0])  # 1, 2, 4 or 8 devices
