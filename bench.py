#!/usr/bin/env python3
"""Benchmark of the commit hot path: PolynomialBatch::from_values (iNTT -> coset LDE -> bit-reversed row-major leaves ->
Poseidon Merkle tree to the cap), metric rows*cols/s (BASELINE.json).

  python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (default workload: BASELINE config 2)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # coset-sharded over N GPUs
  python bench.py --impl reference ...                      # the reference algorithm on the host cores (CPU port)

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in HBM), `e2e` = the same
metric through the C-ABI call with pinned HOST buffers (H2D of the inputs and D2H of every output inside the timed
region).  `roofline` describes the dominant kernel (the leaf sponge); `roofline_ntt` the NTT passes the BASELINE
target speaks about; `cpu_baseline` the CPU restatement timed on this box's cores on a bounded sample.
Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "LDE+Merkle commit rows*cols/s"
UNIT = "rows*cols/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_cores():
    """Threads the CPU arm uses: every core this process may run on.  Asked for explicitly -- torchrun exports
    OMP_NUM_THREADS=1 to its ranks, which would otherwise turn the CPU arm into a single-thread run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, field_name):
    """The reference's CPU algorithm (C++ restatement, OpenMP over columns / subtrees like the reference's rayon) on all
    host cores, on a bounded sample of the workload.  kind = "port": the Rust reference cannot be built (no toolchain)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import p2o
    from util import synthetic_columns
    try:
        p2o.build(native=True)
        native = True
    except Exception as e:  # no compiler on the box: fall back to the portable prebuilt library
        log("native oracle build failed (%s); using the portable build" % e)
        native = False
    fid = p2o.GL if field_name == "gl" else p2o.BB
    order, dtype = p2o.ORDER[fid], p2o.DTYPE[fid]
    log_n = min(args.log_n, args.cpu_sample_log_n)
    p2o.load(native)
    cores = host_cores()
    vals = synthetic_columns(order, dtype, args.cols, 1 << log_n)
    times, stages = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = p2o.commit(fid, vals, p2o.VALUES, args.rate_bits, args.cap_height, threads=cores, want_leaves=True, native=native)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            stages = res["stage_seconds"]
        del res
    t = float(np.mean(times))
    value = (args.cols << log_n) / t
    sample = "2^%d x %d of the 2^%d x %d workload per step (rate_bits=%d, cap_height=%d)" % (
        log_n, args.cols, args.log_n, args.cols, args.rate_bits, args.cap_height)
    # the whole workload once, so that the per-step sample's throughput is checked against the full size it stands for
    full = None
    if log_n < args.log_n and not args.no_full_check:
        es = 8 if field_name == "gl" else 4
        need = args.cols * (1 << args.log_n) * es * (3 + 2 * (1 << args.rate_bits)) * 1.2  # input, coeffs, LDE columns, leaves
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        if avail is not None and avail < need:
            full = {"skipped": "needs ~%.0f GB of host memory, %.0f GB available" % (need / 1e9, avail / 1e9)}
        else:
            try:
                del vals
                big = np.empty((args.cols, 1 << args.log_n), dtype=dtype)
                for cb in range(0, args.cols, 8):
                    ce = min(cb + 8, args.cols)
                    big[cb:ce] = synthetic_columns(order, dtype, ce - cb, 1 << args.log_n, col_begin=cb)
                t0 = time.perf_counter()
                res = p2o.commit(fid, big, p2o.VALUES, args.rate_bits, args.cap_height, threads=cores, want_leaves=True, native=native)
                dt = time.perf_counter() - t0
                full = {"ms": dt * 1e3, "value": (args.cols << args.log_n) / dt, "unit": UNIT, "runs": 1,
                        "ratio_to_sample": ((args.cols << args.log_n) / dt) / value,
                        "stage_ms": {k: v * 1e3 for k, v in zip(("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"), res["stage_seconds"])}}
                del res, big
            except MemoryError as e:
                full = {"skipped": "MemoryError: %s" % e}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64" if field_name == "gl" else "u32", "data": "synthetic",
        "config": workload_config(args, field_name),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "full_size_run": full,
                         "stage_ms": {k: v * 1e3 for k, v in zip(("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"), stages)},
                         "note": "C++ restatement of the reference CPU path (oracle/), g++ -O3 %s -fopenmp, %d threads = every core this "
                                 "process may run on (OMP_NUM_THREADS is ignored: torchrun sets it to 1); the Rust reference cannot be built here" % (
                                     "-march=native" if native else "-march=x86-64-v3", cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, field_name):
    name = "Goldilocks+Poseidon(w12)" if field_name == "gl" else "BabyBear+Poseidon2(w16)"
    return {"workload": "synthetic PolynomialBatch::from_values %s 2^%d rows x %d cols, rate_bits=%d, cap_height=%d" % (
                name, args.log_n, args.cols, args.rate_bits, args.cap_height),
            "field": field_name, "log_rows": args.log_n, "cols": args.cols, "rate_bits": args.rate_bits,
            "cap_height": args.cap_height, "parallelism": "coset-sharded x%d" % args.gpus if args.gpus > 1 else "single GPU",
            "exchange": (args.exchange if args.gpus > 1 else None),
            "exchange_transport": (("copy engines, one block per rank and ring step, under the LDE of the blocks already in" if args.exchange == "ring" else
                                    (os.environ.get("P2G_EXCHANGE") or ("fused peer stores" if args.gpus == 2 else "block push kernel behind the inverse NTT")))
                                   if args.gpus > 1 and args.exchange in ("fused", "pipelined", "ring") else None),
            "l2": "inputs (%.2f GB) and leaves (%.2f GB) are larger than the 126 MB L2; no flush needed" % (
                args.cols * (1 << args.log_n) * (8 if field_name == "gl" else 4) / 1e9,
                args.cols * ((1 << args.rate_bits) << args.log_n) * (8 if field_name == "gl" else 4) / 1e9)}


# ------------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = [None]


def link_ceilings(ctx, d_buf, h_buf, stream, barrier, dev, G, dist, torch):
    """What the host link gives this job at best: every rank copies up to 1 GiB between pinned host memory and HBM at the
    same time (one cudaMemcpyAsync per GPU on the library's stream), aggregate GB/s = total bytes / slowest rank."""
    nbytes = int(min(1 << 30, d_buf.numel() * d_buf.element_size(), h_buf.nbytes))
    out = {}
    for name, kind, dst, src in (("d2h_ceiling_gbs", 1, h_buf.ctypes.data, d_buf.data_ptr()),
                                 ("h2d_ceiling_gbs", 0, d_buf.data_ptr(), h_buf.ctypes.data)):
        best = None
        for _ in range(3):
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            ctx.check(ctx.lib.p2g_memcpy_async(ctx.handle, dst, src, nbytes, kind))
            ev1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if G > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t[0]) if best is None else min(best, float(t[0]))
        out[name] = G * nbytes / (best * 1e-3) / 1e9
    out["ceiling_note"] = "%d concurrent pinned copies of %d MiB, one per GPU, best of 3, max over ranks" % (G, nbytes >> 20)
    return out


def run_ours(args, field_name, check_only=False):
    """check_only: one step of the same code path (same exchange, same buffers, same C-ABI calls) at a size the oracle
    finishes in milliseconds, every output of every rank compared with the oracle; nothing is timed or printed."""
    import torch
    import torch.distributed as dist
    import plonky2_goldibear_b200 as p2g
    from plonky2_goldibear_b200 import _lib
    from util import synthetic_columns

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: park the real stdout and send everything else (NCCL's version banner,
    # library chatter) to stderr until the line is printed
    if _REAL_STDOUT[0] is None:
        sys.stdout.flush()
        _REAL_STDOUT[0] = os.dup(1)
        os.dup2(2, 1)
    real_stdout = _REAL_STDOUT[0]
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with one rank per GPU" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the commit path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    F = p2g.GOLDILOCKS if field_name == "gl" else p2g.BABYBEAR
    tdtype = torch.int64 if field_name == "gl" else torch.int32
    ctx = p2g.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.lib.p2g_stream(ctx.handle), device=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    L, c, r, h = args.log_n, args.cols, args.rate_bits, args.cap_height
    n, N = 1 << L, 1 << (L + r)
    G = world
    if G & (G - 1) or G > (1 << r):
        raise SystemExit("--gpus must be a power of two <= 2^rate_bits")
    lg = G.bit_length() - 1
    cpr = (c + G - 1) // G                      # columns whose inverse NTT this rank owns
    col0, col1 = min(rank * cpr, c), min((rank + 1) * cpr, c)
    rows_local = N // G
    cap_local = max(h - lg, 0)
    d_el = F.num_hash_out_elts
    es = F.elem_bytes

    # ---- synthetic inputs: value[col][row] = SplitMix64(seed ^ (col << 32 | row)) mod p, seed 42
    # (generated a few columns at a time, only this rank's: the whole matrix is 9 GB at 2^23 rows)
    host_in = ctx.pinned_empty((max(col1 - col0, 1), n), F.dtype)
    for cb in range(col0, col1, 8):
        ce = min(cb + 8, col1)
        host_in[cb - col0:ce - col0] = synthetic_columns(F.order, F.dtype, ce - cb, n, col_begin=cb)
    with torch.cuda.stream(stream):
        d_vals = torch.zeros((cpr, n), dtype=tdtype, device=dev)
        d_vals[:col1 - col0].copy_(torch.from_numpy(host_in[:col1 - col0].view(np.int64 if es == 8 else np.int32)))
        d_coeff_slice = torch.zeros((cpr, n), dtype=tdtype, device=dev)
        d_coeffs = torch.zeros((G * cpr, n), dtype=tdtype, device=dev) if G > 1 else d_coeff_slice
        d_leaves = torch.empty((rows_local, c), dtype=tdtype, device=dev)
        d_digests = torch.empty((max(2 * (rows_local - (1 << cap_local)), 1), d_el), dtype=tdtype, device=dev)
        d_cap = torch.empty((1 << cap_local, d_el), dtype=tdtype, device=dev)
    stream.synchronize()

    def make_desc(input_kind, flags):
        desc = _lib.CommitDesc()
        desc.field, desc.hash, desc.repr, desc.input_kind = F.field_id, F.hash_id, 0, input_kind
        desc.log_n, desc.num_cols, desc.rate_bits, desc.cap_height = L, c, r, h
        desc.flags = flags
        desc.shard_index, desc.shard_count = (rank, G) if G > 1 else (0, 0)
        return desc

    keep = []  # keep ctypes arrays alive

    def ptrs(values):
        arr = (ctypes.c_void_p * len(values))(*values)
        keep.append(arr)
        return ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))

    # ---- device-resident step
    dev_desc = make_desc(0 if G == 1 else 1, _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT)
    dev_desc.col_ptrs = ptrs([(d_vals if G == 1 else d_coeffs).data_ptr()])
    dev_desc.coeffs_out = ptrs([d_coeff_slice.data_ptr()])
    dev_desc.leaves_out, dev_desc.leaf_pitch = d_leaves.data_ptr(), c
    dev_desc.digests_out, dev_desc.cap_out = d_digests.data_ptr(), d_cap.data_ptr()
    kernel_s = np.zeros(4)

    # ---- multi-GPU exchange.  Every rank needs all coefficient columns before its coset LDE.  "fused": the inverse NTT's
    # last pass stores its columns straight into every rank's gather buffer over NVLink (CUDA-IPC mapped peer memory,
    # p2g_intt_allgather) and a one-element all-reduce on the same stream is the only collective; two gather buffers
    # alternate so that a fast rank never overwrites coefficients a slow rank is still reading.  "nccl": inverse NTT into
    # a local slice, then ncclAllGather.
    # "pipelined": the same fused kernel, but the columns travel in groups of ~64: a second context (its own stream) runs
    # inverse NTT + peer stores + barrier for group s+1 while the main context extends group s into the leaf matrix
    # (p2g_commit with P2G_FLAG_LEAVES_ONLY on the column range); the tree is built once every column is in
    # (P2G_FLAG_LEAVES_READY).  Measured at 8 GPUs, 2^20 x 135: 21.1 ms (kernel stores) / 23.2 ms (copy engines) against
    # 20.4 ms for the plain fused exchange -- the extra barriers and smaller launches cost more than the overlap hides,
    # so "fused" stays the default.
    # "ring": rank k's inverse NTT stays local; in step s = 1 .. G-1 its finished block goes to rank (k + s) % G through
    # the copy engines (one source and one destination per rank and step, no SMs involved) while the main stream extends
    # the block that arrived in step s-1 (its own block first) into the leaf matrix -- the LDE starts after 1/G of the
    # coefficients has crossed NVLink instead of after all of it.  Blocks are even column shares (sharding.ring_blocks).
    exchange = args.exchange if G > 1 else "none"
    pipelined = exchange in ("pipelined", "pipelined-ce")
    ring = exchange == "ring"
    ce_copies = exchange == "pipelined-ce"  # peers receive the group through the copy engines instead of kernel stores
    if pipelined or ring:
        exchange = "fused"
    gather_ptrs, peer_tables, flag = [], [], None
    if exchange == "fused":
        try:
            nbytes = G * cpr * n * es
            handles = []
            for k in range(2):
                ptr = ctx.lib.p2g_dev_alloc(ctx.handle, nbytes)
                if not ptr:
                    raise RuntimeError("p2g_dev_alloc failed")
                gather_ptrs.append(ptr)
                hbuf = ctypes.create_string_buffer(64)
                ctx.check(ctx.lib.p2g_ipc_export(ctx.handle, ptr, hbuf))
                handles.append(hbuf.raw)
            all_handles = [None] * G
            dist.all_gather_object(all_handles, handles)
            for k in range(2):
                order = [gather_ptrs[k]]
                for rr in range(G):
                    if rr == rank:
                        continue
                    peer = ctypes.c_void_p()
                    ctx.check(ctx.lib.p2g_ipc_import(ctx.handle, all_handles[rr][k], ctypes.byref(peer)))
                    order.append(peer.value)
                tbl = (ctypes.c_void_p * G)(*order)
                keep.append(tbl)
                peer_tables.append(tbl)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
        except Exception as e:  # e.g. IPC not permitted in this container: fall back to the NCCL all-gather
            log("fused exchange unavailable (%s); using ncclAllGather" % e)
            exchange = "nccl"
        ok = torch.tensor([1 if exchange == "fused" else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            exchange = "nccl"
    pipelined = pipelined and exchange == "fused"
    ring = ring and exchange == "fused"
    step_no = [0]

    # ---- pipelined exchange: column groups and who owns what.  Group s = columns [s*Wg, (s+1)*Wg); inside it rank k owns
    # the k-th slice of Wg/G columns (whole NTT tiles), so a group's columns are contiguous in the gather buffer.
    if pipelined or ring:
        from plonky2_goldibear_b200.sharding import column_groups, ring_blocks
        if ring:
            # block k = rank k's columns.  Even shares to the column: the strided passes work column by column, and the last
            # pass runs once over all columns, so a block need not be whole 8- / 16-column tiles; the ring steps are
            # globally synchronised, i.e. every step lasts as long as the largest block's copy
            group_ranges = ring_blocks(c, G, 1)
            own_ranges = [group_ranges[rank]]
        else:
            group_ranges, owned = column_groups(c, G, 8 if field_name == "gl" else 16)
            own_ranges = owned[rank]
        own_total = sum(hi - lo for lo, hi in own_ranges)
        p_host = ctx.pinned_empty((max(own_total, 1), n), F.dtype)
        loc = 0
        for lo, hi in own_ranges:
            if hi > lo:
                p_host[loc:loc + hi - lo] = synthetic_columns(F.order, F.dtype, hi - lo, n, col_begin=lo)
                loc += hi - lo
        ctx2 = p2g.Context(local_rank)  # the exchange side: its own stream and workspace
        stream2 = torch.cuda.ExternalStream(ctx2.lib.p2g_stream(ctx2.handle), device=dev)
        with torch.cuda.stream(stream):
            p_vals = torch.zeros((max(own_total, 1), n), dtype=tdtype, device=dev)
            p_vals[:own_total].copy_(torch.from_numpy(p_host[:own_total].view(np.int64 if es == 8 else np.int32)))
        stream.synchronize()
        ev_g = [torch.cuda.Event() for _ in group_ranges]
        ev_step_done = torch.cuda.Event()
        g_tbls, g_descs = [], []
        for glo, ghi in group_ranges:
            gd = make_desc(1, _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT | _lib.FLAG_LEAVES_ONLY)
            gd.num_cols = ghi - glo
            tbl = (ctypes.c_void_p * 1)()
            keep.append(tbl)
            gd.col_ptrs = ctypes.cast(tbl, ctypes.POINTER(ctypes.c_void_p))
            gd.leaves_out, gd.leaf_pitch = d_leaves.data_ptr() + glo * es, c
            g_tbls.append(tbl)
            g_descs.append(gd)
        tree_desc = make_desc(1, _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT | _lib.FLAG_LEAVES_READY)
        tree_desc.leaves_out, tree_desc.leaf_pitch = d_leaves.data_ptr(), c
        tree_desc.digests_out, tree_desc.cap_out = d_digests.data_ptr(), d_cap.data_ptr()

    def pipelined_step():
        k = step_no[0] & 1
        step_no[0] += 1
        loc = 0
        for s_, (lo, hi) in enumerate(own_ranges):
            if hi > lo and not ce_copies:
                ctx2.check(ctx2.lib.p2g_intt_allgather(ctx2.handle, F.field_id, 0, L, hi - lo, lo,
                                                       p_vals.data_ptr() + loc * n * es, peer_tables[k], G))
                loc += hi - lo
            elif hi > lo:
                ctx2.check(ctx2.lib.p2g_intt_allgather(ctx2.handle, F.field_id, 0, L, hi - lo, lo,
                                                       p_vals.data_ptr() + loc * n * es, peer_tables[k], 1))
                off, nb = lo * n * es, (hi - lo) * n * es
                for pk in range(1, G):
                    ctx2.check(ctx2.lib.p2g_memcpy_async(ctx2.handle, peer_tables[k][pk] + off, gather_ptrs[k] + off, nb, 2))
                loc += hi - lo
            with torch.cuda.stream(stream2):
                dist.all_reduce(flag)  # every rank's stores of this group have landed
                ev_g[s_].record(stream2)
        for s_, (glo, ghi) in enumerate(group_ranges):
            stream.wait_event(ev_g[s_])
            g_tbls[s_][0] = gather_ptrs[k] + glo * n * es
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(g_descs[s_])))
            kernel_s[:] += np.array(list(g_descs[s_].kernel_seconds))
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(tree_desc)))
        kernel_s[:] += np.array(list(tree_desc.kernel_seconds))

    def peer_ptr(k, rr):  # rank rr's gather buffer k as mapped here (peer_tables: own first, then the others in rank order)
        return peer_tables[k][0 if rr == rank else 1 + (rr if rr < rank else rr - 1)]

    ring_events = []

    def ring_step():
        k = step_no[0] & 1
        step_no[0] += 1
        lo, hi = own_ranges[0]
        # the exchange side of a commit starts when the previous commit of this rank is through (no run-ahead into the
        # previous step's hashing: a step is one commit's own schedule, and a peer may only overwrite a gather buffer
        # this rank has finished reading)
        if step_no[0] > 1:
            stream2.wait_event(ev_step_done)
        if hi > lo:
            ctx2.check(ctx2.lib.p2g_intt_allgather(ctx2.handle, F.field_id, 0, L, hi - lo, lo, p_vals.data_ptr(), peer_tables[k], 1))
        ev_g[0].record(stream2)
        off, nb = lo * n * es, (hi - lo) * n * es
        for s_ in range(1, G):
            if nb:
                ctx2.check(ctx2.lib.p2g_memcpy_async(ctx2.handle, peer_ptr(k, (rank + s_) % G) + off, gather_ptrs[k] + off, nb, 2))
            with torch.cuda.stream(stream2):
                dist.all_reduce(flag)  # every rank's copy of this step has landed
                ev_g[s_].record(stream2)
        # the LDE split by pass (p2g_lde_columns): the strided passes of a block as soon as it is in, the last pass -- which
        # writes whole leaf rows -- once over all columns (a split by column range, P2G_FLAG_LEAVES_ONLY per block, makes
        # every block write a narrow strip of every row: 16-column blocks cost the LDE +36..49 %, tools/lde_blocks_probe.py)
        shard = (rank, G)
        for s_ in range(G):
            src = (rank - s_) % G
            stream.wait_event(ev_g[s_])
            glo, ghi = group_ranges[src]
            if ghi > glo:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.check(ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, L, r, shard[0], shard[1], c, glo, ghi - glo,
                                                  gather_ptrs[k] + glo * n * es, None, 0, 0))
                e1.record(stream)
                ring_events.append((e0, e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.check(ctx.lib.p2g_lde_columns(ctx.handle, F.field_id, 0, L, r, shard[0], shard[1], c, 0, c, gather_ptrs[k],
                                          d_leaves.data_ptr(), c, 1))
        e1.record(stream)
        ring_events.append((e0, e1))
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(tree_desc)))
        kernel_s[:] += np.array(list(tree_desc.kernel_seconds))
        ev_step_done.record(stream)

    def exchange_coeffs(d2h_slice=None):
        """inverse NTT of this rank's columns + all-gather; returns the device pointer holding all coefficients"""
        if exchange == "fused":
            k = step_no[0] & 1
            step_no[0] += 1
            ctx.check(ctx.lib.p2g_intt_allgather(ctx.handle, F.field_id, 0, L, cpr, rank * cpr, d_vals.data_ptr(),
                                                 peer_tables[k], G))
            with torch.cuda.stream(stream):
                if d2h_slice is not None:
                    own = gather_ptrs[k] + rank * cpr * n * es
                    ctypes_copy_async(d2h_slice, own, (col1 - col0) * n * es)
                dist.all_reduce(flag)  # cross-GPU barrier ordered on the library's stream
            return gather_ptrs[k]
        ctx.check(ctx.lib.p2g_transform(ctx.handle, F.field_id, 0, 0, L, cpr, 0,
                                        _lib.FLAG_DEVICE_INPUT | _lib.FLAG_DEVICE_OUTPUT,
                                        d_vals.data_ptr(), d_coeff_slice.data_ptr()))
        with torch.cuda.stream(stream):
            if d2h_slice is not None:
                ctypes_copy_async(d2h_slice, d_coeff_slice.data_ptr(), (col1 - col0) * n * es)
            dist.all_gather_into_tensor(d_coeffs, d_coeff_slice)
        stream.synchronize()
        return d_coeffs.data_ptr()

    def ctypes_copy_async(host_arr, dev_ptr, nbytes):
        # D2H of this rank's coefficient slice into pinned memory, on the library's stream
        if nbytes:
            ctx.check(ctx.lib.p2g_memcpy_async(ctx.handle, host_arr.ctypes.data, dev_ptr, nbytes, 1))

    coeff_tbl = (ctypes.c_void_p * 1)()
    keep.append(coeff_tbl)

    def device_step():
        if ring:
            return ring_step()
        if pipelined:
            return pipelined_step()
        if G > 1:
            coeff_tbl[0] = exchange_coeffs()
            dev_desc.col_ptrs = ctypes.cast(coeff_tbl, ctypes.POINTER(ctypes.c_void_p))
        ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(dev_desc)))
        kernel_s[:] += np.array(list(dev_desc.kernel_seconds))

    def barrier():
        if G > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if G > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) * 1e-3, float(t[1]) * 1e-3

    for _ in range(0 if check_only else args.warmup):
        device_step()
    kernel_s[:] = 0
    del ring_events[:]
    launches0 = ctx.kernel_launches() + (ctx2.kernel_launches() if (pipelined or ring) else 0)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not check_only:
        sampler.start()
    dev_total, _ = timed(device_step, 1 if check_only else args.steps)
    clocks = sampler.stop() if rank == 0 and not check_only else None
    if ring:
        kernel_s[1] += sum(a.elapsed_time(b) for a, b in ring_events) * 1e-3  # the LDE passes, without the waits between them
    launches = ctx.kernel_launches() + (ctx2.kernel_launches() if (pipelined or ring) else 0) - launches0
    args.exchange = "ring" if ring else ("pipelined-ce" if ce_copies else "pipelined") if pipelined else (exchange if G > 1 else None)  # what actually ran
    t_step = dev_total / args.steps
    ksec = kernel_s / args.steps
    value = (c << L) / t_step

    # parity guard inside the bench: the cap of the timed run must equal the cap of the e2e (host) run below
    cap_dev = d_cap.cpu().numpy().view(F.dtype).copy()

    ref = None
    if check_only:
        # the oracle's commit of the whole (small) job; this rank's slices of it
        from oracle import p2o
        fid = p2o.GL if field_name == "gl" else p2o.BB
        ref = p2o.commit(fid, synthetic_columns(F.order, F.dtype, c, n), p2o.VALUES, r, h, threads=host_cores())
        if h < lg:
            raise SystemExit("parity check needs cap_height >= log2(gpus)")
        nd, per_cap = len(ref["digests"]) // G, (1 << h) // G
        exp = {"leaves": ref["leaves"][rank * rows_local:(rank + 1) * rows_local],
               "digests": ref["digests"][rank * nd:(rank + 1) * nd],
               "cap": ref["cap"][rank * per_cap:(rank + 1) * per_cap],
               "coeffs": ref["coeffs"][col0:col1]}

        def verdict(what, checks):
            bad = [k for k, (got, want) in checks.items() if got.shape != want.shape or not (got == want).all()]
            okt = torch.tensor([0 if bad else 1], device=dev)
            if G > 1:
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if bad:
                log("PARITY FAILURE (%s, rank %d of %d): %s differ from the oracle" % (what, rank, G, ", ".join(bad)))
            if int(okt) == 0:
                raise SystemExit(3)

        verdict("device-resident step over the %s exchange" % (args.exchange if G > 1 else "single-GPU path"), {
            "leaves": (d_leaves.cpu().numpy().view(F.dtype), exp["leaves"]),
            "digests": (d_digests.cpu().numpy().view(F.dtype)[:nd], exp["digests"]),
            "cap": (cap_dev, exp["cap"])})

    # ---- end-to-end step: host (pinned) buffers in, every output back to the host
    e2e = None
    if not args.no_e2e:
        h_coeffs = ctx.pinned_empty((max(col1 - col0, 1), n), F.dtype)
        h_leaves = ctx.pinned_empty((rows_local, c), F.dtype)
        h_digests = ctx.pinned_empty((max(2 * (rows_local - (1 << cap_local)), 1), d_el), F.dtype)
        h_cap = ctx.pinned_empty((1 << cap_local, d_el), F.dtype)
        if G == 1:
            e_desc = make_desc(0, 0)
            e_desc.col_ptrs = ptrs([host_in[j].ctypes.data for j in range(c)])
            e_desc.coeffs_out = ptrs([h_coeffs[j].ctypes.data for j in range(c)])
        else:
            e_desc = make_desc(1, _lib.FLAG_DEVICE_INPUT)  # coefficients arrive over NVLink, outputs go to the host
            e_desc.col_ptrs = ptrs([d_coeffs.data_ptr()])
        e_desc.leaves_out, e_desc.leaf_pitch = h_leaves.ctypes.data, c
        e_desc.digests_out, e_desc.cap_out = h_digests.ctypes.data, h_cap.ctypes.data
        t_in = torch.from_numpy(host_in[:max(col1 - col0, 1)].view(np.int64 if es == 8 else np.int32))
        t_hc = torch.from_numpy(h_coeffs.view(np.int64 if es == 8 else np.int32))

        def e2e_step():
            if G > 1:
                with torch.cuda.stream(stream):
                    d_vals[:col1 - col0].copy_(t_in[:col1 - col0], non_blocking=True)       # H2D of this rank's columns
                coeff_tbl[0] = exchange_coeffs(d2h_slice=h_coeffs)                          # + D2H of its coefficients
                e_desc.col_ptrs = ctypes.cast(coeff_tbl, ctypes.POINTER(ctypes.c_void_p))
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(e_desc)))

        e2e_step()
        barrier()
        assert (h_cap == cap_dev).all(), "device-resident and end-to-end runs disagree on the Merkle cap"
        if check_only:
            verdict("end-to-end step, host buffers", {
                "leaves": (h_leaves, exp["leaves"]), "digests": (h_digests[:nd], exp["digests"]), "cap": (h_cap, exp["cap"]),
                "coeffs": (h_coeffs[:col1 - col0], exp["coeffs"])})
            if rank == 0:
                log("parity check: 2^%d x %d over %d GPU(s), exchange %s -- device-resident and end-to-end outputs of every "
                    "rank equal the oracle's" % (L, c, G, args.exchange))
            ctx.close()
            return
        ceilings = link_ceilings(ctx, d_leaves, h_leaves, stream, barrier, dev, G, dist, torch)
        _, e2e_wall = timed(e2e_step, args.e2e_steps)
        te = e2e_wall / args.e2e_steps
        h2d = (col1 - col0) * n * es
        d2h = (col1 - col0) * n * es + rows_local * c * es + h_digests.nbytes + h_cap.nbytes
        tb = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if G > 1:
            dist.all_reduce(tb)
        e2e = {"value": (c << L) / te, "unit": UNIT, "h2d_bytes_per_step": int(tb[0]), "d2h_bytes_per_step": int(tb[1]),
               "ms_per_step": te * 1e3, "steps": args.e2e_steps,
               "note": "p2g_commit with pinned host buffers; all of coeffs, leaves, digests, cap copied back"}
        e2e.update(ceilings)
        if ceilings.get("d2h_ceiling_gbs"):
            # what the host link alone allows: every output byte at the measured D2H ceiling (uploads overlap)
            e2e["d2h_floor_ms"] = float(tb[1]) / (ceilings["d2h_ceiling_gbs"] * 1e9) * 1e3
            e2e["frac_of_link_floor"] = e2e["d2h_floor_ms"] / (te * 1e3)

    # ---- the same drop-in call with PAGEABLE buffers: what the unchanged Rust API hands over (Vec<F> from the allocator)
    e2e_pageable = None
    if not args.no_e2e and G == 1 and not check_only:
        p_in = np.array(host_in[:c], copy=True)
        p_coeffs = np.empty((c, n), dtype=F.dtype)
        p_leaves = np.empty((rows_local, c), dtype=F.dtype)
        p_digests = np.empty((max(2 * (rows_local - (1 << cap_local)), 1), d_el), dtype=F.dtype)
        p_cap = np.empty((1 << cap_local, d_el), dtype=F.dtype)
        pg = make_desc(0, 0)
        pg.col_ptrs = ptrs([p_in[j].ctypes.data for j in range(c)])
        pg.coeffs_out = ptrs([p_coeffs[j].ctypes.data for j in range(c)])
        pg.leaves_out, pg.leaf_pitch = p_leaves.ctypes.data, c
        pg.digests_out, pg.cap_out = p_digests.ctypes.data, p_cap.ctypes.data

        def pageable_step():
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(pg)))

        pageable_step()
        assert (p_cap == cap_dev).all() and (p_leaves[-1] == h_leaves[-1]).all(), "pageable run disagrees"
        _, pw = timed(pageable_step, 2)
        e2e_pageable = {"value": (c << L) / (pw / 2), "unit": UNIT, "ms_per_step": pw / 2 * 1e3, "steps": 2,
                        "note": "the same call with pageable host buffers (what an unchanged Vec<F>-based caller hands over): "
                                "every copy goes through the driver's staging buffers, nothing overlaps"}
        del p_in, p_coeffs, p_leaves, p_digests

    # ---- device-resident variant of the same call (N2): inputs from pinned host memory, leaves / digests / coefficients
    # stay in HBM behind a batch handle, the cap comes back, then the prover's access pattern is served by gather kernels
    # (28 FRI queries: one row + one Merkle path each, fri/prover.rs:219-255).  Secondary number; `e2e` above is the
    # full drop-in with every output copied to the host.
    e2e_resident = None
    if not args.no_e2e and G == 1 and not check_only:
        k_desc = make_desc(0, _lib.FLAG_KEEP_ON_DEVICE)
        k_desc.col_ptrs = ptrs([host_in[j].ctypes.data for j in range(c)])
        k_cap = np.empty((1 << h, d_el), dtype=F.dtype)
        k_desc.cap_out = k_cap.ctypes.data
        handle = ctypes.c_void_p()
        k_desc.batch_out = ctypes.pointer(handle)
        q_idx = np.random.default_rng(0).integers(0, N, size=28, dtype=np.uint64)
        q_rows = np.empty((28, c), dtype=F.dtype)
        q_sib = np.empty((28, L + r - h, d_el), dtype=F.dtype)

        def resident_step():
            ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(k_desc)))
            ctx.check(ctx.lib.p2g_batch_get_rows(ctx.handle, handle, q_idx.ctypes.data, 28, q_rows.ctypes.data))
            ctx.check(ctx.lib.p2g_batch_prove(ctx.handle, handle, q_idx.ctypes.data, 28, q_sib.ctypes.data))
            ctx.lib.p2g_batch_free(ctx.handle, handle)

        resident_step()
        assert (k_cap == cap_dev).all(), "device-resident batch disagrees on the Merkle cap"
        _, rw = timed(resident_step, args.e2e_steps)
        tr = rw / args.e2e_steps
        e2e_resident = {"value": (c << L) / tr, "unit": UNIT, "ms_per_step": tr * 1e3,
                        "h2d_bytes_per_step": int(c * n * es), "d2h_bytes_per_step": int(k_cap.nbytes + q_rows.nbytes + q_sib.nbytes),
                        "note": "P2G_FLAG_KEEP_ON_DEVICE: outputs stay in HBM; cap + 28 queried rows and Merkle paths copied back"}

    # ---- the library's own multi-device context (one process drives all GPUs): rank 0 runs it while the others wait on a
    # CPU barrier (an NCCL barrier would park a spinning kernel on the very GPUs being measured)
    lib_multi = None
    if G > 1 and not args.no_library_multi and not check_only:
        cpu_group = dist.new_group(backend="gloo")
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                lib_multi = library_multi_bench(args, field_name, G)
            except SystemExit:
                raise
            except Exception as e:
                lib_multi = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        dist.barrier(group=cpu_group)

    if rank != 0:
        if G > 1:
            dist.destroy_process_group()
        return

    # ---- roofline bookkeeping (algorithmic bytes per kernel)
    peak, peak_src = measured_peak()
    leaves_bytes = rows_local * c * es
    leaf_digest_bytes = rows_local * 32
    t_leaf = ksec[2]
    perms = N * ((c + 7) // 8) + (N - (1 << h)) if c > d_el else N - (1 << h)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "leaf_sponge_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("log_rows") == L and tj.get("cols") == c and tj.get("field") == field_name:
                traffic = tj["dram_bytes_per_launch"] / 1e9
        except Exception:
            pass
    roofline = {"kernel": "leaf_hash_kernel (Poseidon sponge over the leaf rows; integer-pipe bound, see int_pipe)",
                "bound": "hbm", "achieved": (leaves_bytes + leaf_digest_bytes) / t_leaf / 1e9, "peak": peak, "unit": "GB/s",
                "frac": (leaves_bytes + leaf_digest_bytes) / t_leaf / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "ms_per_launch": t_leaf * 1e3,
                "algorithmic_gb_per_launch": (leaves_bytes + leaf_digest_bytes) / 1e9}
    ntt_bytes = es * n * c * (2 + (1 << r)) / G
    t_ntt = ksec[0] + ksec[1]
    roofline_ntt = {"kernels": "ntt_strided_kernel + ntt_final_natural_kernel + ntt_final_leaves_kernel", "bound": "hbm",
                    "achieved": ntt_bytes / t_ntt / 1e9, "peak": peak, "unit": "GB/s", "frac": ntt_bytes / t_ntt / 1e9 / peak,
                    "ms": t_ntt * 1e3, "algorithmic_gb": ntt_bytes / 1e9}
    # integer-pipe view of the dominant kernel: cycles per warp-permutation against the pipe cost model (the costs that
    # reproduce ncu's ALU and fmaheavy utilisation of this kernel: profiles/r2_leaf_sponge_cost_model.json)
    int_pipe = None
    try:
        model = json.load(open(os.path.join(ROOT, "profiles", "r2_leaf_sponge_cost_model.json")))
        key = "goldilocks_poseidon" if field_name == "gl" else "babybear_poseidon2"
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        leaf_perms = rows_local * ((c + 7) // 8)
        measured = t_leaf * sms * 4 * clk * 1e6 / (leaf_perms / 32.0)
        ref = model[key]
        dyn = ref["dynamic_instructions_per_permutation"]
        alu, mul = 2.0 * dyn["ALU pipe"], 4.0 * dyn["IMAD.WIDE/HI"] + 2.0 * dyn["other IMAD"]
        int_pipe = {"kernel": "leaf_hash_kernel", "cycles_per_warp_permutation": measured,
                    "microbench_cycles_per_warp_permutation": ref["measured_cycles_per_warp_permutation"],
                    "alu_pipe_cycles": alu, "multiplier_pipe_cycles": mul,
                    "busier_pipe_share_of_time": max(alu, mul) / measured,
                    "model": "profiles/r2_leaf_sponge_cost_model.json: dynamic warp instructions per permutation (ncu source page) x 2 cycles per "
                             "ALU-pipe instruction, 2 per IMAD of any kind and 4 per IMAD.WIDE/HI on the multiplier pipe (ncu: fmaheavy) -- the "
                             "multiplier pipe is the busiest pipe of the kernel (ncu 88 % Goldilocks, 85 % BabyBear); the pipes run concurrently"}
    except Exception as e:  # the model file is evidence, not a dependency
        int_pipe = {"unavailable": str(e)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64" if field_name == "gl" else "u32", "data": "synthetic",
        "config": workload_config(args, field_name), "clocks": clocks, "e2e": e2e, "e2e_device_resident_outputs": e2e_resident,
        "e2e_pageable": e2e_pageable, "library_multi_device": lib_multi,
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_ntt": roofline_ntt, "int_pipe": int_pipe,
        "stage_ms": {"inverse_ntt": ksec[0] * 1e3, "lde_ntt": ksec[1] * 1e3, "leaf_sponge": ksec[2] * 1e3, "tree_levels": ksec[3] * 1e3},
        "hash": {"permutations_per_step": int(perms), "permutations_per_s": perms / G / (ksec[2] + ksec[3])},
    }
    if G == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, field_name)
    if G == 1 and not args.no_latency:
        try:
            line["latency"] = latency_entries(field_name, ctx, p2g)
        except Exception as e:
            line["latency"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if G == 1 and not args.no_fri:
        try:
            line["fri_commit_phase"] = fri_commit_phase_bench(args, field_name, ctx, p2g)
        except Exception as e:  # a secondary measurement must not cost the headline line
            line["fri_commit_phase"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if G > 1:
        dist.destroy_process_group()


def library_multi_bench(args, field_name, G):
    """The same job through the library's own multi-device context (p2g_init(device_ids, G), ONE process): the inverse-NTT
    exchange, the event barrier and the coset sharding all live behind p2g_commit.  Run by rank 0 after the rank-per-GPU
    measurement (the other ranks wait on a CPU barrier).  Wall clock around the calls -- p2g_commit returns when every
    device has finished -- so launch and join overhead are inside the numbers."""
    import plonky2_goldibear_b200 as p2g
    from plonky2_goldibear_b200 import _lib
    from util import synthetic_columns
    F = p2g.GOLDILOCKS if field_name == "gl" else p2g.BABYBEAR
    L, c, r, h = args.log_n, args.cols, args.rate_bits, args.cap_height
    n, N = 1 << L, 1 << (L + r)
    es, d_el = F.elem_bytes, F.num_hash_out_elts
    ctx = p2g.Context(list(range(G)))
    out = {"devices": G, "how": "one process, p2g_init with %d devices; host wall clock around p2g_commit" % G}
    try:
        # parity first: 2^12 rows through the multi-device context against the oracle
        from oracle import p2o
        fid = p2o.GL if field_name == "gl" else p2o.BB
        small = synthetic_columns(F.order, F.dtype, c, 1 << min(12, L))
        ref = p2o.commit(fid, small, p2o.VALUES, r, h, threads=host_cores())
        b = p2g.PolynomialBatch.from_values(F, list(small), r, False, h, ctx=ctx)
        ok = ((np.stack(b.polynomials) == ref["coeffs"]).all() and (b.merkle_tree.leaves == ref["leaves"]).all() and
              (b.merkle_tree.digests == ref["digests"]).all() and (b.merkle_tree.cap.hashes == ref["cap"]).all())
        if not ok:
            raise SystemExit("PARITY FAILURE: multi-device context disagrees with the oracle at 2^12 x %d" % c)
        out["parity"] = "2^%d x %d commit through the multi-device context equals the oracle's (coefficients, leaves, digests, cap)" % (min(12, L), c)
        del b, ref
        host_in = ctx.pinned_empty((c, n), F.dtype)
        for cb in range(0, c, 8):
            ce = min(cb + 8, c)
            host_in[cb:ce] = synthetic_columns(F.order, F.dtype, ce - cb, n, col_begin=cb)
        keep = []

        def ptrs(values):
            arr = (ctypes.c_void_p * len(values))(*values)
            keep.append(arr)
            return ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))

        def desc_for(flags):
            d = _lib.CommitDesc()
            d.field, d.hash, d.repr, d.input_kind = F.field_id, F.hash_id, 0, 0
            d.log_n, d.num_cols, d.rate_bits, d.cap_height = L, c, r, h
            d.flags = flags
            return d

        cap = np.empty((1 << h, d_el), dtype=F.dtype)
        handle = ctypes.c_void_p()

        def run(desc, steps, free_batch):
            times = []
            for i in range(steps + 2):
                t0 = time.perf_counter()
                ctx.check(ctx.lib.p2g_commit(ctx.handle, ctypes.byref(desc)))
                times.append(time.perf_counter() - t0)
                if free_batch:
                    ctx.lib.p2g_batch_free(ctx.handle, handle)
            return float(np.mean(times[2:]))

        # (a) inputs already in HBM (each device holds its column slice), outputs stay in HBM: the device-resident step
        cpr = (c + G - 1) // G
        dev_ptrs = []
        for e in range(G):
            sub = ctx.lib.p2g_sub_context(ctx.handle, e)
            col0 = min(e * cpr, c)
            ncol = min(cpr, c - col0)
            ptr = ctx.lib.p2g_dev_alloc(sub, max(ncol, 1) * n * es)
            if not ptr:
                raise RuntimeError("p2g_dev_alloc failed on device %d" % e)
            if ncol:
                ctx.check(ctx.lib.p2g_memcpy_async(sub, ptr, host_in[col0].ctypes.data, ncol * n * es, 0))
            dev_ptrs.append((sub, ptr))
        d_res = desc_for(_lib.FLAG_DEVICE_INPUT | _lib.FLAG_KEEP_ON_DEVICE)
        d_res.col_ptrs = ptrs([p_ for _, p_ in dev_ptrs])
        d_res.cap_out, d_res.batch_out = cap.ctypes.data, ctypes.pointer(handle)
        t = run(d_res, args.steps, True)
        out["device_resident"] = {"ms_per_step": t * 1e3, "value": (c << L) / t, "unit": UNIT,
                                  "note": "column slices resident on their devices in, batch resident out, cap to the host"}
        cap_res = cap.copy()
        for sub, ptr in dev_ptrs:
            ctx.lib.p2g_dev_free(sub, ptr)
        # (b) pinned host columns in, outputs resident (the form in which more GPUs can beat the host link)
        d_keep = desc_for(_lib.FLAG_KEEP_ON_DEVICE)
        d_keep.col_ptrs = ptrs([host_in[j].ctypes.data for j in range(c)])
        d_keep.cap_out, d_keep.batch_out = cap.ctypes.data, ctypes.pointer(handle)
        t = run(d_keep, args.e2e_steps, True)
        assert (cap == cap_res).all()
        out["e2e_device_resident_outputs"] = {"ms_per_step": t * 1e3, "value": (c << L) / t, "unit": UNIT,
                                              "h2d_bytes_per_step": int(c * n * es), "d2h_bytes_per_step": int(cap.nbytes)}
        # (c) the full drop-in: every output back in (pinned) host buffers
        if not args.no_e2e:
            h_coeffs = ctx.pinned_empty((c, n), F.dtype)
            h_leaves = ctx.pinned_empty((N, c), F.dtype)
            h_digests = ctx.pinned_empty((2 * (N - (1 << h)), d_el), F.dtype)
            d_all = desc_for(0)
            d_all.col_ptrs = ptrs([host_in[j].ctypes.data for j in range(c)])
            d_all.coeffs_out = ptrs([h_coeffs[j].ctypes.data for j in range(c)])
            d_all.leaves_out, d_all.leaf_pitch = h_leaves.ctypes.data, c
            d_all.digests_out, d_all.cap_out = h_digests.ctypes.data, cap.ctypes.data
            t = run(d_all, args.e2e_steps, False)
            assert (cap == cap_res).all()
            out["e2e"] = {"ms_per_step": t * 1e3, "value": (c << L) / t, "unit": UNIT, "h2d_bytes_per_step": int(c * n * es),
                          "d2h_bytes_per_step": int(h_coeffs.nbytes + h_leaves.nbytes + h_digests.nbytes + cap.nbytes)}
        out["cap0"] = [int(x) for x in cap_res[0]]
    finally:
        ctx.close()
    return out


def cpu_baseline(args, field_name):
    """The oracle (kind "port") timed on this box's host cores on a bounded sample of the same workload."""
    from oracle import p2o
    from util import synthetic_columns
    try:
        p2o.build(native=True)
        native = True
    except Exception:
        native = False
    fid = p2o.GL if field_name == "gl" else p2o.BB
    log_n = min(args.log_n, args.cpu_sample_log_n)
    p2o.load(native)
    cores = host_cores()
    vals = synthetic_columns(p2o.ORDER[fid], p2o.DTYPE[fid], args.cols, 1 << log_n)
    best, stages = None, None
    for _ in range(2):
        t0 = time.perf_counter()
        res = p2o.commit(fid, vals, p2o.VALUES, args.rate_bits, args.cap_height, threads=cores, native=native)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, stages = dt, res["stage_seconds"]
        del res
    return {"value": (args.cols << log_n) / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "2^%d x %d (1/%d of the rows of the workload), best of 2, all %d host threads" % (
                log_n, args.cols, 1 << (args.log_n - log_n), cores),
            "ms": best * 1e3,
            "stage_ms": {k: v * 1e3 for k, v in zip(("IFFT", "FFT + blinding", "transpose LDEs", "build Merkle tree"), stages)}}


def latency_entries(field_name, ctx, p2g):
    """Latency-bound shapes of BASELINE configs 1 and 4: the four oracles of the fibonacci circuit
    (2^4 rows) and of a 2^12-row recursion circuit, and a FRI layer tree of 32-element leaves.  Host wall clock around the
    public call (pageable numpy columns in, every output back on the host), median of 30 after 5 warm-up calls."""
    f = p2g.GOLDILOCKS if field_name == "gl" else p2g.BABYBEAR
    rng = np.random.default_rng(11)
    if field_name == "gl":
        shapes = [(4, 85), (4, 135), (4, 20), (4, 16), (12, 85), (12, 135), (12, 20), (12, 16)]
    else:
        shapes = [(12, 46), (12, 167), (12, 36), (12, 48)]  # narrow BabyBear recursion config
    out = {"how": "host wall clock around PolynomialBatch.from_values / MerkleTree.new, pageable host arrays in and out, "
                  "rate_bits 3, cap_height 4, median of 30 calls after 5 warm-up calls; commit_pinned_reused_ms: the same call with pinned "
                  "input and output buffers allocated once", "commit_ms": {}, "commit_pinned_reused_ms": {}, "merkle_ms": {}}
    reuse = {}
    for log_n, c in shapes:
        cols = [rng.integers(0, f.order, size=1 << log_n, dtype=np.uint64).astype(f.dtype) for _ in range(c)]
        ts = []
        for it in range(35):
            t0 = time.perf_counter()
            p2g.PolynomialBatch.from_values(f, cols, 3, False, 4, ctx=ctx)
            ts.append(time.perf_counter() - t0)
        out["commit_ms"]["2^%d x %d" % (log_n, c)] = float(np.median(ts[5:])) * 1e3
        # the library call alone: pinned buffers allocated once and reused (no host allocation or page faults inside the
        # timed call -- a 2^12 x 135 commit returns 35 MB of leaves, which a fresh allocation above glibc's 32 MB mmap
        # threshold pays for with ~9000 page faults per call)
        pin = ctx.pinned_empty((c, 1 << log_n), f.dtype)
        pin[...] = np.stack(cols)
        ts = []
        for it in range(35):
            t0 = time.perf_counter()
            p2g.PolynomialBatch.from_values(f, list(pin), 3, False, 4, ctx=ctx, pinned_out=True, _reuse=reuse)
            ts.append(time.perf_counter() - t0)
        out["commit_pinned_reused_ms"]["2^%d x %d" % (log_n, c)] = float(np.median(ts[5:])) * 1e3
    d = f.extension_degree
    for log_leaves, leaf_len in ((11, 16 * d), (7, 16 * d)):  # arity-16 layers of a 2^15 codeword (fri/prover.rs:104-108)
        leaves = rng.integers(0, f.order, size=(1 << log_leaves, leaf_len), dtype=np.uint64).astype(f.dtype)
        ts = []
        for it in range(35):
            t0 = time.perf_counter()
            p2g.MerkleTree.new(f, leaves, min(4, log_leaves), ctx=ctx)
            ts.append(time.perf_counter() - t0)
        out["merkle_ms"]["2^%d leaves x %d"% (log_leaves, leaf_len)] = float(np.median(ts[5:])) * 1e3
    return out


def fri_commit_phase_bench(args, field_name, ctx, p2g):
    """fri_committed_trees (fri/prover.rs:83-133) on the codeword of this workload's LDE size, arity 16
    down to a 2^rate_bits..2^(rate_bits+4)-coefficient final polynomial (ConstantArityBits(4, 5)), host arrays in, caps and
    final polynomial out, layer trees left on the device for the query phase.  Timed end to end on the host clock (the
    challenger sits between the device calls); the CPU figure is the oracle's restatement of the same function."""
    from oracle import p2o
    fid = p2o.GL if field_name == "gl" else p2o.BB
    f = p2g.GOLDILOCKS if field_name == "gl" else p2g.BABYBEAR
    d = f.extension_degree
    degree_bits, r, h = args.log_n, args.rate_bits, args.cap_height
    arities = []
    while degree_bits - sum(arities) > 5 and degree_bits + r - sum(arities) - 4 >= h:
        arities.append(4)
    n = 1 << (degree_bits + r)
    rng = np.random.default_rng(7)
    coeffs = ctx.pinned_empty((n, d), f.dtype)
    values = ctx.pinned_empty((n, d), f.dtype)
    coeffs[:] = 0
    coeffs[: 1 << degree_bits] = rng.integers(0, f.order, size=(1 << degree_bits, d), dtype=np.uint64).astype(f.dtype)
    values[:] = rng.integers(0, f.order, size=(n, d), dtype=np.uint64).astype(f.dtype)  # any codeword times the same
    params = p2g.FriParams(p2g.FriConfig(r, h), arities)
    best = None
    for it in range(3):
        ch = p2g.Challenger(f, ctx=ctx)
        t0 = time.perf_counter()
        trees, final = p2g.fri_committed_trees(f, coeffs, values, ch, params, ctx=ctx)
        dt = time.perf_counter() - t0
        for t in trees:
            t.free()
        best = dt if best is None or dt < best else best
    out = {"ms": best * 1e3, "log_n": degree_bits + r, "extension_degree": d, "reduction_arity_bits": arities, "cap_height": h,
           "timing": "host wall clock around fri_committed_trees, pinned host arrays in (%.0f MB H2D), best of 3" % (2 * coeffs.nbytes / 1e6)}
    if not args.no_cpu_baseline:
        cpu_best = None
        for it in range(2):
            ch = p2g.Challenger(f, permute=lambda s: p2o.permute(fid, s))
            t0 = time.perf_counter()
            p2o.fri_committed_trees(fid, coeffs, values, ch, arities, r, h)
            dt = time.perf_counter() - t0
            cpu_best = dt if cpu_best is None or dt < cpu_best else cpu_best
        out["cpu_port_ms"] = cpu_best * 1e3
        out["cpu_port"] = "oracle restatement (fold single-threaded, Merkle trees on all host threads), best of 2"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field", default="gl", choices=["gl", "bb"])
    ap.add_argument("--log-n", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--rate-bits", type=int, default=3)
    ap.add_argument("--cap-height", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="ring", choices=["ring", "fused", "pipelined", "pipelined-ce", "nccl"],
                    help="multi-GPU coefficient exchange: ring (default) -- each rank's finished block of coefficient columns "
                         "goes round the ring through the copy engines while the strided LDE passes of the blocks already in "
                         "run (p2g_lde_columns); fused -- inverse NTT with the all-gather in its epilogue / a push kernel "
                         "behind it, the LDE after a barrier; the same in column groups overlapped with the LDE (pipelined: "
                         "kernel stores, pipelined-ce: copy engines; both measured slower, DESIGN.md section 5); or a "
                         "separate ncclAllGather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fri", action="store_true", help="skip the secondary FRI commit-phase measurement")
    ap.add_argument("--cpu-sample-log-n", type=int, default=16)
    ap.add_argument("--no-full-check", action="store_true",
                    help="--impl reference: skip the single full-size run that checks the per-step sample's throughput")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="N > 1: skip the 2^12-row commit through the same exchange that is compared with the oracle before timing")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-shape latency entries (configs 1 and 4)")
    ap.add_argument("--no-library-multi", action="store_true",
                    help="N > 1: skip the second pass that runs the job through the library's single-process multi-device context")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: timing rules ask for >= 3 warm-up steps; got %d" % args.warmup)
    if args.log_n is None:
        args.log_n = 20 if args.field == "gl" else 22
    if args.cols is None:
        args.cols = 135 if args.field == "gl" else 256
    if args.impl == "reference":
        run_reference(args, args.field)
    else:
        if not args.no_parity_check:
            # before anything is timed: the same path at 2^12 rows, every rank's outputs against the oracle
            import copy
            small = copy.copy(args)
            small.log_n, small.steps, small.warmup, small.e2e_steps, small.no_e2e = min(12, args.log_n), 1, 0, 1, False
            run_ours(small, args.field, check_only=True)
        run_ours(args, args.field)


if __name__ == "__main__":
    main()
