"""World-size-2 gloo test (CPU) of the multi-GPU host logic: column-sharded inverse transform, all-gather of the padded
coefficient shares, coset/row/digest/cap slices per rank.  The arithmetic is done by the oracle here (there is no
GPU); what is under test is the partition and the exchange that bench.py runs with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import GL_P, synthetic_columns


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log_n, c, r, h, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import p2o
    from plonky2_goldibear_b200.sharding import plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    sp = plan(rank, world, c, log_n, r, h)
    vals = synthetic_columns(GL_P, np.uint64, c, n)
    # column-sharded inverse transform on this rank's share, zero-padded to cols_per_rank columns
    share = np.zeros((sp.cols_per_rank, n), dtype=np.uint64)
    for j in range(sp.col_begin, sp.col_end):
        share[j - sp.col_begin] = p2o.ifft(p2o.GL, vals[j])
    gathered = [torch.zeros((sp.cols_per_rank, n), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(share.view(np.int64)))
    coeffs = torch.cat(gathered, 0).numpy().view(np.uint64)[:c]
    # every rank now holds all coefficients; it owns cosets [coset_begin, coset_end) = rows [row_begin, row_end)
    full = p2o.commit(p2o.GL, coeffs, p2o.COEFFS, r, h)
    rows = full["leaves"][sp.row_begin:sp.row_end]
    local_digests, local_cap = p2o.merkle_tree(p2o.GL, rows, sp.local_cap_height)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), coeffs=coeffs, rows=rows, digests=local_digests, cap=local_cap,
             plan=np.array([sp.row_begin, sp.row_end, sp.digest_begin, sp.digest_end, sp.cap_begin, sp.cap_end,
                            int(sp.owns_whole_cap_entries)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cap_height", [4, 0])
def test_two_rank_partition_reassembles(tmp_path, oracle, cap_height):
    world, log_n, c, r = 2, 6, 7, 3
    port = _free_port()
    mp.spawn(_worker, args=(world, port, log_n, c, r, cap_height, str(tmp_path)), nprocs=world, join=True)
    vals = synthetic_columns(GL_P, np.uint64, c, 1 << log_n)
    ref = oracle.commit(oracle.GL, vals, oracle.VALUES, r, cap_height)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    for p in parts:
        assert (p["coeffs"] == ref["coeffs"]).all()  # the all-gather reassembled every column on every rank
    assert (np.concatenate([p["rows"] for p in parts]) == ref["leaves"]).all()
    for p in parts:
        rb, re_, db, de, cb, ce, whole = [int(x) for x in p["plan"]]
        assert (ref["digests"][db:de] == p["digests"]).all()  # local digests are a contiguous slice of the global array
        if whole:
            assert (ref["cap"][cb:ce] == p["cap"]).all()
    if cap_height == 0:
        # 2^cap_height < world: the two shard roots are combined into the single cap entry, and they are the pair of
        # digests that sits between the two local blocks in the global layout
        roots = np.concatenate([p["cap"] for p in parts])
        assert (oracle.two_to_one(oracle.GL, roots[0], roots[1]) == ref["cap"][0]).all()
        mid = int(parts[0]["plan"][3])
        assert (ref["digests"][mid:mid + 2] == roots).all()


def test_plan_rejects_bad_worlds():
    from plonky2_goldibear_b200.sharding import plan
    with pytest.raises(ValueError):
        plan(0, 3, 135, 20, 3, 4)
    with pytest.raises(ValueError):
        plan(0, 16, 135, 20, 3, 4)
    sp = plan(7, 8, 135, 20, 3, 4)
    assert (sp.col_begin, sp.col_end, sp.cols_per_rank) == (119, 135, 17)
    assert (sp.coset_begin, sp.coset_end) == (7, 8) and sp.local_cap_height == 1
    assert sp.row_end == 1 << 23 and sp.cap_end == 16


def test_pipelined_exchange_column_groups_cover_every_column_once():
    """bench.py --exchange pipelined: groups are contiguous, ranks own disjoint whole-tile slices, nothing is lost."""
    from plonky2_goldibear_b200.sharding import column_groups
    for c, world, tile in ((135, 8, 8), (135, 4, 8), (135, 2, 8), (256, 8, 16), (256, 2, 16), (5, 8, 8), (64, 8, 8), (1, 1, 8)):
        groups, owned = column_groups(c, world, tile)
        assert groups[0][0] == 0 and groups[-1][1] == c
        assert all(a[1] == b[0] for a, b in zip(groups, groups[1:]))
        seen = []
        for k in range(world):
            assert len(owned[k]) == len(groups)
            for (lo, hi), (glo, ghi) in zip(owned[k], groups):
                assert glo <= lo <= hi <= ghi
                if hi > lo and hi != ghi:
                    assert (hi - lo) % tile == 0  # whole NTT column tiles except at the ragged end
                seen += list(range(lo, hi))
        assert sorted(seen) == list(range(c))
    groups, _ = column_groups(135, 8, 8)
    assert groups == [(0, 64), (64, 128), (128, 135)]


def _group_worker(rank, world, port, log_n, c, r, h, tile, out_dir):
    """bench.py --exchange pipelined on CPU: the columns cross in groups; after the barrier of group s every rank extends
    exactly that column range into its leaf rows (P2G_FLAG_LEAVES_ONLY), and the tree follows when all are in."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import p2o
    from plonky2_goldibear_b200.sharding import column_groups, plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    sp = plan(rank, world, c, log_n, r, h)
    groups, owned = column_groups(c, world, tile)
    gather = np.zeros((c, n), dtype=np.uint64)          # this rank's gather buffer, original column order
    leaves = np.zeros((sp.row_end - sp.row_begin, c), dtype=np.uint64)
    for s, (glo, ghi) in enumerate(groups):
        lo, hi = owned[rank][s]
        mine = synthetic_columns(GL_P, np.uint64, hi - lo, n, col_begin=lo) if hi > lo else np.zeros((0, n), dtype=np.uint64)
        share = np.stack([p2o.ifft(p2o.GL, v) for v in mine]) if hi > lo else mine
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, share))   # the fused epilogue's peer stores + the group's barrier
        for plo, phi, block in parts:
            gather[plo:phi] = block
        # LDE of this group only, written into its column range of the leaf matrix
        ext = p2o.commit(p2o.GL, gather[glo:ghi], p2o.COEFFS, r, 0)["leaves"][sp.row_begin:sp.row_end]
        leaves[:, glo:ghi] = ext
    digests, cap = p2o.merkle_tree(p2o.GL, leaves, sp.local_cap_height)
    np.savez(os.path.join(out_dir, "grank%d.npz" % rank), rows=leaves, digests=digests, cap=cap,
             plan=np.array([sp.digest_begin, sp.digest_end, sp.cap_begin, sp.cap_end]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pipelined_exchange_by_column_groups(tmp_path, oracle):
    world, log_n, c, r, h, tile = 2, 5, 11, 2, 2, 2
    port = _free_port()
    mp.spawn(_group_worker, args=(world, port, log_n, c, r, h, tile, str(tmp_path)), nprocs=world, join=True)
    vals = synthetic_columns(GL_P, np.uint64, c, 1 << log_n)
    ref = oracle.commit(oracle.GL, vals, oracle.VALUES, r, h)
    parts = [np.load(os.path.join(str(tmp_path), "grank%d.npz" % i)) for i in range(world)]
    assert (np.concatenate([p["rows"] for p in parts]) == ref["leaves"]).all()
    for p in parts:
        db, de, cb, ce = [int(x) for x in p["plan"]]
        assert (ref["digests"][db:de] == p["digests"]).all()
        assert (ref["cap"][cb:ce] == p["cap"]).all()


def test_ring_exchange_blocks_and_schedule():
    """bench.py --exchange ring: blocks partition the columns into runs of whole NTT column tiles that differ by at most
    one tile; in step s rank k sends block k to rank (k + s) % G and receives block (k - s) % G, so after G - 1 steps
    every rank holds every block exactly once."""
    from plonky2_goldibear_b200.sharding import ring_blocks
    for c, world, tile in ((135, 8, 8), (135, 4, 8), (135, 2, 8), (256, 8, 16), (256, 2, 16), (5, 8, 8), (64, 8, 8), (1, 1, 8), (17, 4, 8), (135, 8, 1), (135, 4, 1), (5, 8, 1)):
        blocks = ring_blocks(c, world, tile)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == c
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert all(lo % tile == 0 or lo == c for lo, _ in blocks)
        tiles = [-(-(hi - lo) // tile) for lo, hi in blocks]
        assert max(tiles) - min(tiles) <= 1
        for k in range(world):
            got = [k] + [(k - s) % world for s in range(1, world)]   # own block, then one per step
            assert sorted(got) == list(range(world))
            for s in range(1, world):
                assert ((k - s) % world + s) % world == k             # the sender of step s targets this rank
    assert ring_blocks(135, 8, 8) == [(0, 24), (24, 40), (40, 56), (56, 72), (72, 88), (88, 104), (104, 120), (120, 135)]


def _ring_worker(rank, world, port, log_n, c, r, h, tile, out_dir):
    """The ring exchange on CPU: own inverse NTT, then per step one block out / one block in (point to point), the LDE of
    each block written into its column range as soon as it is in, the tree at the end."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    from oracle import p2o
    from plonky2_goldibear_b200.sharding import plan, ring_blocks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << log_n
    sp = plan(rank, world, c, log_n, r, h)
    blocks = ring_blocks(c, world, tile)
    leaves = np.zeros((sp.row_end - sp.row_begin, c), dtype=np.uint64)

    def extend(lo, hi, coeffs):
        if hi > lo:
            leaves[:, lo:hi] = p2o.commit(p2o.GL, coeffs, p2o.COEFFS, r, 0)["leaves"][sp.row_begin:sp.row_end]

    lo, hi = blocks[rank]
    mine = synthetic_columns(GL_P, np.uint64, hi - lo, n, col_begin=lo) if hi > lo else np.zeros((0, n), dtype=np.uint64)
    own = np.stack([p2o.ifft(p2o.GL, v) for v in mine]) if hi > lo else mine
    extend(lo, hi, own)
    for s in range(1, world):
        dst, src = (rank + s) % world, (rank - s) % world
        slo, shi = blocks[src]
        out_t = torch.from_numpy(own.view(np.int64).copy())
        in_t = torch.zeros((shi - slo, n), dtype=torch.int64)
        reqs = []
        if hi > lo:
            reqs.append(dist.isend(out_t, dst))
        if shi > slo:
            reqs.append(dist.irecv(in_t, src))
        for q in reqs:
            q.wait()
        extend(slo, shi, in_t.numpy().view(np.uint64))
    digests, cap = p2o.merkle_tree(p2o.GL, leaves, sp.local_cap_height)
    np.savez(os.path.join(out_dir, "rrank%d.npz" % rank), rows=leaves, digests=digests, cap=cap,
             plan=np.array([sp.digest_begin, sp.digest_end, sp.cap_begin, sp.cap_end]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ring_exchange(tmp_path, oracle):
    world, log_n, c, r, h, tile = 2, 5, 11, 2, 2, 2
    port = _free_port()
    mp.spawn(_ring_worker, args=(world, port, log_n, c, r, h, tile, str(tmp_path)), nprocs=world, join=True)
    vals = synthetic_columns(GL_P, np.uint64, c, 1 << log_n)
    ref = oracle.commit(oracle.GL, vals, oracle.VALUES, r, h)
    parts = [np.load(os.path.join(str(tmp_path), "rrank%d.npz" % i)) for i in range(world)]
    assert (np.concatenate([p["rows"] for p in parts]) == ref["leaves"]).all()
    for p in parts:
        db, de, cb, ce = [int(x) for x in p["plan"]]
        assert (ref["digests"][db:de] == p["digests"]).all()
        assert (ref["cap"][cb:ce] == p["cap"]).all()


def test_final_tile_column_stride_is_bank_conflict_free():
    """csrc/p2g_internal.h tile_col_stride: the column-major tile of the two final NTT kinds (x[w * cs + i], loaded by one
    bulk copy per column) with the kernels' lane map (consecutive lanes take consecutive w, then consecutive units).  An
    8-byte access is served per half-warp, a 4-byte access per warp; the stride must keep every such group on distinct
    banks in all rounds but the last, and every column 16-byte aligned for cp.async.bulk.  (BabyBear tiles of 16 columns
    cannot have both: they are 2-way conflicted)"""
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plonky2_goldibear_b200", "csrc",
                            "p2g_internal.h")).read()
    assert "(128u / elem_bytes) >> lw" in hdr and "16u / elem_bytes" in hdr  # the formula restated below

    def col_stride(elem, lp, lw):
        pad = max((128 // elem) >> lw, 16 // elem)
        return (1 << lp) + pad

    def conflicts(elem, lp, lw, lowbits):
        cs = col_stride(elem, lp, lw)
        assert (cs * elem) % 16 == 0
        group = 16 if elem == 8 else 32          # lanes served together
        words = elem // 4
        worst = 1
        for first in range(0, 256, group):        # a CTA's threads, unit u = thread id
            banks = {}
            for u in range(first, first + group):
                w, t = u & ((1 << lw) - 1), u >> lw
                low, hi = t & ((1 << lowbits) - 1), t >> lowbits
                p0 = (hi << (lowbits + 4)) | low  # a 16-point unit: its first point (the others are + j << lowbits)
                addr_words = (w * cs + p0) * words
                b = (addr_words % 32) // words
                banks[b] = banks.get(b, 0) + 1
            worst = max(worst, max(banks.values()))
        return worst

    for lp, lw in ((10, 3), (11, 2), (9, 3)):                 # Goldilocks: 8 columns up to 2^10 points, 4 at 2^11
        for lowbits in (lp - 4, lp - 7):
            assert conflicts(8, lp, lw, lowbits) == 1, (lp, lw, lowbits)
    for lowbits in (7, 3):                                      # BabyBear 2^11 points x 8 columns (config 3)
        assert conflicts(4, 11, 3, lowbits) == 1
    assert conflicts(4, 10, 4, 6) == 2                          # BabyBear 16 columns: 2-way, as documented
