"""Coset partition of one commit over G GPUs.

Row k = b*n + t of the bit-reversed LDE is the evaluation on the coset with shift g*w_N^{bitrev_r(b)}, so every
block of n consecutive leaves is one size-n coset NTT and owns a contiguous slice of `digests`
(merkle_tree.rs:50-53,135-148).  Rank d computes cosets [d*2^r/G, (d+1)*2^r/G); the inverse NTT is sharded by
columns and followed by one all-gather of the coefficients -- the only data-path collective."""
from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    cols_per_rank: int   # padded share of the inverse NTT (the all-gather needs equal shares)
    col_begin: int
    col_end: int
    coset_begin: int
    coset_end: int
    row_begin: int
    row_end: int
    local_cap_height: int  # cap height of the rank-local tree
    digest_begin: int      # slice of the global `digests` array (valid when 2^cap_height >= world)
    digest_end: int
    cap_begin: int
    cap_end: int
    owns_whole_cap_entries: bool


def plan(rank: int, world: int, num_cols: int, log_n: int, rate_bits: int, cap_height: int) -> ShardPlan:
    if world <= 0 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    lg = world.bit_length() - 1
    if lg > rate_bits:
        raise ValueError("more GPUs than LDE cosets (2^rate_bits)")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    n_total = 1 << (log_n + rate_bits)
    cpr = (num_cols + world - 1) // world
    cosets = (1 << rate_bits) // world
    rows = n_total // world
    whole = cap_height >= lg
    local_h = cap_height - lg if whole else 0
    if whole:
        per = 2 * (rows - (1 << local_h))
        dig = (rank * per, (rank + 1) * per)
        cap = (rank << local_h, (rank + 1) << local_h)
    else:
        # the shard is an aligned block inside one cap subtree: its 2*(rows-1) local digests start at 2*first_leaf
        shards_per_sub = world >> cap_height
        sub, j = divmod(rank, shards_per_sub)
        sub_len = 2 * (shards_per_sub * rows - 1)
        dig = (sub * sub_len + 2 * j * rows, sub * sub_len + 2 * j * rows + 2 * (rows - 1))
        cap = (sub, sub + 1)
    return ShardPlan(rank, world, cpr, min(rank * cpr, num_cols), min((rank + 1) * cpr, num_cols),
                     rank * cosets, (rank + 1) * cosets, rank * rows, (rank + 1) * rows, local_h,
                     dig[0], dig[1], cap[0], cap[1], whole)


def column_groups(num_cols: int, world: int, tile_cols: int):
    """Column groups of the pipelined exchange (bench.py --exchange pipelined): the coefficient columns cross NVLink in
    groups of about half the batch so that the LDE of group s overlaps the inverse NTT + peer stores of group s+1.
    Group s = columns [s*Wg, (s+1)*Wg) with Wg a multiple of world*tile_cols (whole NTT column tiles per rank); inside a
    group rank k owns the k-th slice of Wg/world columns.  Returns (groups, owned): groups[s] = (lo, hi),
    owned[k][s] = (lo, hi) of rank k in group s (possibly empty in a ragged last group)."""
    if world <= 0 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    unit = world * tile_cols
    wg = max(unit, (num_cols // 2) // unit * unit)
    t = wg // world
    groups = [(lo, min(lo + wg, num_cols)) for lo in range(0, num_cols, wg)]
    owned = [[(min(glo + k * t, ghi), min(glo + (k + 1) * t, ghi)) for glo, ghi in groups] for k in range(world)]
    return groups, owned


def ring_blocks(num_cols: int, world: int, tile_cols: int):
    """Column blocks of the ring exchange (bench.py --exchange ring): rank k owns the inverse NTT of block k, a run of whole
    units of tile_cols columns (the last one may be ragged), blocks differ by at most one unit; tile_cols = 1 gives even
    shares to the column, which is what the pass-split LDE (p2g_lde_columns) allows and bench.py uses.  In step
    s = 1 .. world-1 rank k sends its finished block to rank (k + s) % world and extends the block it received from rank
    (k - s) % world into its leaf matrix, so the LDE starts after 1/world of the coefficients has arrived.
    Returns [(lo, hi)] per rank."""
    if world <= 0 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    tiles = (num_cols + tile_cols - 1) // tile_cols
    base, rem = divmod(tiles, world)
    out, t = [], 0
    for k in range(world):
        take = base + (1 if k < rem else 0)
        out.append((min(t * tile_cols, num_cols), min((t + take) * tile_cols, num_cols)))
        t += take
    return out
