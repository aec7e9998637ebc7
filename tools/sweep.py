"""Size sweep of the Goldilocks commit (BASELINE.json config 5: 2^16 .. 2^23 rows x 135 columns, rate_bits 3, cap_height 4)
on one GPU: device-side stage times as reported by p2g_commit (CUDA events), leaves left out of the D2H.
usage: sweep.py [lo hi] -> one JSON line per size on stdout"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonky2_goldibear_b200 as p2g

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 23
c = 135
f = p2g.GOLDILOCKS
ctx = p2g.Context(0)
rng = np.random.default_rng(42)
for log_n in range(lo, hi + 1):
    n = 1 << log_n
    block = rng.integers(0, f.order, size=(c, n), dtype=np.uint64)
    cols = list(block)
    best = None
    for rep in range(2 if log_n >= 22 else 3):
        b = p2g.PolynomialBatch.from_values(f, cols, 3, False, 4, ctx=ctx, keep_on_device=True)
        st = list(b.desc.stage_seconds)
        ks = list(b.desc.kernel_seconds)
        b.merkle_tree.free()
        dev = sum(st)
        if best is None or dev < best[0]:
            best = (dev, ks)
    dev, ks = best
    print(json.dumps({"log_rows": log_n, "cols": c, "device_ms": dev * 1e3, "rows_cols_per_s": c * n / dev,
                      "stage_ms": {"inverse_ntt": ks[0] * 1e3, "lde_ntt": ks[1] * 1e3, "leaf_sponge": ks[2] * 1e3, "tree_levels": ks[3] * 1e3},
                      "device_gb": ctx.workspace_bytes() / 1e9}), flush=True)
    del block, cols
    ctx.release_workspace()
