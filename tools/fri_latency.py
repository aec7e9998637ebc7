import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import plonky2_goldibear_b200 as p2g
ctx = p2g.Context(0)
f = p2g.GOLDILOCKS
rng = np.random.default_rng(1)
for logn in (14, 15):
    n = 1 << logn
    coeffs = np.zeros((n, 2), dtype=np.uint64); coeffs[: n >> 3] = rng.integers(0, f.order, size=(n >> 3, 2), dtype=np.uint64)
    params = p2g.FriParams(p2g.FriConfig(3, 4, proof_of_work_bits=16, num_query_rounds=28), [4, 4] if logn == 15 else [4, 3])
    cols = [rng.integers(0, f.order, size=n >> 3, dtype=np.uint64) for _ in range(135)]
    T = {"commit": 0, "trees": 0, "pow": 0, "query": 0, "perm": 0}
    R = 30
    for it in range(R + 3):
        if it == 3: T = {k: 0 for k in T}
        t0 = time.perf_counter()
        b = p2g.PolynomialBatch.from_values(f, cols, 3, False, 4, ctx=ctx, keep_on_device=True)
        t1 = time.perf_counter()
        ch = p2g.Challenger(f, ctx=ctx)
        ch.observe_cap(b.merkle_tree.cap)
        trees, final = p2g.fri_committed_trees(f, coeffs, None, ch, params, ctx=ctx)
        t2 = time.perf_counter()
        w = p2g.fri_proof_of_work(f, ch, params.config, ctx=ctx)
        t3 = time.perf_counter()
        rounds = p2g.fri_prover_query_rounds([b.merkle_tree], trees, ch, n, params)
        t4 = time.perf_counter()
        s = np.zeros(12, dtype=np.uint64)
        for _ in range(10): ctx.check(ctx.lib.p2g_permute(ctx.handle, 0, 0, s.ctypes.data, 1))
        t5 = time.perf_counter()
        for t in trees: t.free()
        b.merkle_tree.free()
        T["commit"] += t1 - t0; T["trees"] += t2 - t1; T["pow"] += t3 - t2; T["query"] += t4 - t3; T["perm"] += (t5 - t4) / 10
    print("2^%d:" % logn, {k: "%.3f ms" % (v / R * 1e3) for k, v in T.items()})
