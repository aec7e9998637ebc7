"""Smallest end-to-end case for compute-sanitizer runs (one tool per gpurun call): both fields, two NTT plans, blinding,
device-resident accessors; every output checked against the oracle."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plonky2_goldibear_b200 as p2g
from oracle import p2o
from util import synthetic_columns

ctx = p2g.Context(0)
for F, fid, log_n, c, blinding in ((p2g.GOLDILOCKS, p2o.GL, 6, 20, False), (p2g.GOLDILOCKS, p2o.GL, 11, 9, True),
                                   (p2g.BABYBEAR, p2o.BB, 7, 33, False), (p2g.BABYBEAR, p2o.BB, 11, 17, False)):
    n, N = 1 << log_n, 1 << (log_n + 2)
    vals = synthetic_columns(F.order, F.dtype, c, n)
    salts = synthetic_columns(F.order, F.dtype, 4, N, seed=7) if blinding else None
    ref = p2o.commit(fid, vals, p2o.VALUES, 2, 3, salts=salts)
    b = p2g.PolynomialBatch.from_values(F, list(vals), 2, blinding, 3, ctx=ctx, salts=salts)
    assert (b.merkle_tree.leaves == ref["leaves"]).all() and (b.merkle_tree.digests == ref["digests"]).all()
    assert (b.merkle_tree.cap.hashes == ref["cap"]).all() and (np.stack(b.polynomials) == ref["coeffs"]).all()
    k = p2g.PolynomialBatch.from_values(F, list(vals), 2, blinding, 3, ctx=ctx, salts=salts, keep_on_device=True)
    assert (k.merkle_tree.get_many([0, N - 1, 5]) == ref["leaves"][[0, N - 1, 5]]).all()
    assert (k.merkle_tree.prove(5) == p2o.merkle_prove(fid, ref["digests"], N, 3, 5)).all()
    k.merkle_tree.free()
    print("ok", F.name, log_n, c, blinding, flush=True)
ctx.close()
print("sanitize case done")
