"""Stability loop: 300 x (device-resident commit + fri_proof on it); prints free device memory and workspace size."""
import sys, os, numpy as np, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import plonky2_goldibear_b200 as p2g
from oracle import p2o
ctx = p2g.Context(0)
f = p2g.GOLDILOCKS
rng = np.random.default_rng(1)
n = 1 << 14
coeffs = np.zeros((n, 2), dtype=np.uint64); coeffs[: n >> 3] = rng.integers(0, f.order, size=(n >> 3, 2), dtype=np.uint64)
params = p2g.FriParams(p2g.FriConfig(3, 2, proof_of_work_bits=8, num_query_rounds=5), [4, 3, 2])
def free_mb(): torch.cuda.synchronize(); return torch.cuda.mem_get_info()[0] / 1e6
cols = [rng.integers(0, f.order, size=1 << 11, dtype=np.uint64) for _ in range(20)]
for it in range(301):
    ch = p2g.Challenger(f, ctx=ctx)
    b = p2g.PolynomialBatch.from_values(f, cols, 3, False, 2, ctx=ctx, keep_on_device=True)
    proof = p2g.fri_proof(f, [b.merkle_tree], coeffs, None, ch, params, ctx=ctx)
    b.merkle_tree.free()
    if it % 100 == 0: print(it, "free MB", round(free_mb()), "workspace MB", ctx.workspace_bytes() / 1e6, flush=True)
print("ok", proof["pow_witness"])
