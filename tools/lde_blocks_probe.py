"""One GPU, one coset (shard 0 of 8): the LDE of 2^L x 135 coefficient columns into the leaf matrix in one launch pair
against the same work split over column blocks (P2G_FLAG_LEAVES_ONLY per block) -- what the ring exchange does.
usage: lde_blocks_probe.py [log_n]"""
import ctypes
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky2_goldibear_b200 as p2g
from plonky2_goldibear_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
c, r, h, G = 135, 3, 4, 8
n = 1 << L
F = p2g.GOLDILOCKS
ctx = p2g.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.lib.p2g_stream(ctx.handle), device=dev)
with torch.cuda.stream(stream):
    coeffs = torch.randint(0, 2**62, (c, n), dtype=torch.int64, device=dev)
    leaves = torch.empty((n, c), dtype=torch.int64, device=dev)
stream.synchronize()
keep = []


def desc(lo, hi):
    d = _lib.CommitDesc()
    d.field, d.hash, d.repr, d.input_kind = F.field_id, F.hash_id, 0, 1
    d.log_n, d.num_cols, d.rate_bits, d.cap_height = L, hi - lo, r, h
    d.flags = _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT | _lib.FLAG_LEAVES_ONLY
    d.shard_index, d.shard_count = 0, G
    tbl = (ctypes.c_void_p * 1)(coeffs.data_ptr() + lo * n * 8)
    keep.append(tbl)
    d.col_ptrs = ctypes.cast(tbl, ctypes.POINTER(ctypes.c_void_p))
    d.leaves_out, d.leaf_pitch = leaves.data_ptr() + lo * 8, c
    return d


def run(width, reps=5):
    blocks = [(lo, min(lo + width, c)) for lo in range(0, c, width)]
    descs = [desc(lo, hi) for lo, hi in blocks]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(reps + 2):
        ev0.record(stream)
        for d in descs:
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(d)))
        ev1.record(stream)
        stream.synchronize()
        if rep >= 2:
            best = min(best, ev0.elapsed_time(ev1))
    return best, len(blocks)


ref = None
for width in (135, 64, 32, 24, 16, 8):
    ms, nb = run(width)
    chk = int(leaves[::9973].sum().item())
    if ref is None:
        ref = chk
    print("2^%d x %d, one coset: %3d-column blocks (%2d launches pairs): %.3f ms%s" % (L, c, width, nb, ms, "" if chk == ref else "  CHECKSUM DIFFERS"), flush=True)
