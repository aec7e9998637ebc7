"""Quick stage timing of one commit (device seconds per stage as reported by p2g_commit)."""
import sys
import os
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import plonky2_goldibear_b200 as p2g
from util import synthetic_columns

field = p2g.GOLDILOCKS if (len(sys.argv) < 2 or sys.argv[1] == "gl") else p2g.BABYBEAR
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = int(sys.argv[3]) if len(sys.argv) > 3 else 135
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
repr_ = int(os.environ.get("P2G_QB_REPR", "0"))  # 1: BabyBear values are p3 Montgomery words
ctx = p2g.Context(0)
vals = synthetic_columns(field.order, field.dtype, c, 1 << log_n)
cols = list(vals)
for i in range(reps):
    t0 = time.perf_counter()
    b = p2g.PolynomialBatch.from_values(field, cols, 3, False, 4, ctx=ctx, want_leaves=False, repr=repr_)
    wall = time.perf_counter() - t0
    st = list(b.desc.stage_seconds)
    dev = sum(st)
    print("%s 2^%d x %d: h2d %.1f ms | ifft %.2f | lde %.2f | merkle %.2f | d2h %.1f ms | device %.2f ms -> %.3e rows*cols/s | wall %.1f ms | launches %d" % (
        field.name, log_n, c, b.desc.h2d_seconds * 1e3, st[0] * 1e3, st[1] * 1e3, st[3] * 1e3, b.desc.d2h_seconds * 1e3,
        dev * 1e3, (c << log_n) / dev, wall * 1e3, ctx.kernel_launches()), flush=True)
print("cap[0]", b.merkle_tree.cap.hashes[0])
