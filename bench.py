#!/usr/bin/env python
"""bench.py — k-mer vectors clustered per second per LSH iteration (BASELINE.json's metric).

A *step* is one pass of the mode-C clustering hot path over one synthetic abundance matrix of the
named shape: Cluster(I=1) (phase 1, threshold 0.95, nested above batch/1000) followed by
Cluster(I=<iters>) (phase 2), i.e. every LSH iteration the reference's mode C would run on it.

  value : sum over all iterations of rows entering the iteration / device time, rows already in HBM
          (state reset between steps by an untimed device-to-device restore)
  e2e   : same metric through the C ABI with HOST buffers: per step the uint16 count matrix is
          copied from pinned host memory, transformed, clustered, and the clusters (centroids +
          member lists) are read back
  roofline     : the dominant kernel family (the in-bucket merge) against the measured HBM peak
  cpu_baseline : the reference's own OpenMP build (oracle/_ref/kmerLSH_ref) on a bounded sample

`--impl reference` times the reference's CPU implementation instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from kmerlsh_b200 import synth  # noqa: E402

# Rows one GPU clusters per step.  C1 and C2 are whole matrices.  C3-C5 do not fit one GPU as one Cluster()
# call and the reference does not run them that way either: it cuts the matrix into batches of 100 M rows
# (app/kmerLSH.cc:285) that are clustered independently (:311-345).  The bench line for those shapes is one
# such batch (D = 64: the reference's 100 M rows; D = 256: 40 M rows, what the windowed merge's scratch
# leaves room for in 180 GB), phase 1 + the -I iterations on that batch's survivors.
BENCH_ROWS = {"C1": 1_000_000, "C2": 50_000_000, "C3": 100_000_000, "C4": 100_000_000, "C5": 40_000_000}

METRIC = "kmer_vectors_clustered_per_sec_per_lsh_iteration"
UNIT = "rows/s"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "kmerLSH_ref")


def b_alg(d, s):
    """Algorithmic bytes per input row per iteration."""
    return 8 * d + 32 + s * (4 * d + 12)


def b_alg_merge(d, s):
    """The merge kernel's share: gather row, write surviving centroid + count/head/tail."""
    return 4 * d + 4 + s * (4 * d + 12)


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        top = sorted(self.samples)[len(self.samples) // 2:]  # under-load half
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_inputs(workload, rows_override, device, seed_offset=0):
    n_full, sa, sb, seed = synth.CONFIGS[workload]
    n = rows_override if rows_override else min(n_full, BENCH_ROWS[workload])
    seed += seed_offset
    if n <= 2_000_000:
        counts, cov = synth.synth_counts(n, sa, sb, seed)
        import torch

        keep = None
        if torch.cuda.is_available():  # pinned for the end-to-end arm; the reference arm needs no GPU
            pinned = torch.empty(counts.shape, dtype=torch.int16, pin_memory=True)
            pinned.numpy().view(np.uint16)[:] = counts
            counts = pinned.numpy().view(np.uint16)
            keep = pinned
    else:
        from kmerlsh_b200.synth_gpu import synth_counts_gpu

        counts, cov = synth_counts_gpu(n, sa, sb, seed, device=device)
        keep = None
    kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    return n, sa, sb, counts, cov, vk, keep


def run_cpu_reference(counts, cov, sa, sb, sample_rows, iters, min_sim, threads, seed=42):
    """Run the reference binary on the first `sample_rows` rows; parse its own --verbose phase
    lines (function/cluster.cc:263, :307, :325).  Returns (rows_iter_per_s, detail)."""
    d = sa + sb
    work = tempfile.mkdtemp(prefix="klsh_cpu_")
    try:
        os.makedirs(os.path.join(work, "tmp"))
        sub = np.ascontiguousarray(counts[:, :sample_rows])
        sub.tofile(os.path.join(work, "kmer_count.bin"))
        subcov = np.log(np.maximum(sub, 1).astype(np.float64)).sum(axis=1)
        open(os.path.join(work, "kmer_count.log"), "w").write(synth.format_log_line(sample_rows, subcov))
        for name, k in (("A.txt", sa), ("B.txt", sb)):
            open(os.path.join(work, name), "w").write("".join("s%d.fq s%d\n" % (i, i) for i in range(k)))
        env = dict(os.environ, KLSH_SEED=str(seed))
        env.pop("OMP_THREAD_LIMIT", None)
        t0 = time.time()
        out = subprocess.run([REF_BIN, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", str(iters),
                              "-N", str(min_sim), "-K", "23", "-T", str(threads), "--verbose"], cwd=work, env=env, check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        wall = time.time() - t0
    finally:
        shutil.rmtree(work, ignore_errors=True)
    rows = sum(int(x) for x in re.findall(r"^Size of profilings : (\d+)", out, flags=re.M))
    secs = sum(float(x) for x in re.findall(r"^(?:hashing|clustering|merging) takes secs:\t([0-9.eE+-]+)", out, flags=re.M))
    return rows / secs, {"rows_iterations": rows, "phase_seconds": secs, "wall_seconds": wall, "dim": d}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/kmerLSH_ref not built (needs /root/reference at build time)"}))
        return
    n, sa, sb, counts, cov, vk, _ = make_inputs(args.workload, args.rows, "cuda:0")
    cores = os.cpu_count() or 1
    sample_rows = min(n, args.cpu_sample_rows)
    iters = args.cpu_sample_iters
    vals = []
    for k in range(args.warmup + args.steps):
        v, detail = run_cpu_reference(counts, cov, sa, sb, sample_rows, iters, args.min_similarity, cores)
        if k >= args.warmup:
            vals.append((v, detail))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d["phase_seconds"] for _, d in vals]) * 1e3)
    sample = "first %d rows of the %s generator x %d samples, phase 1 + I=%d, -T %d (hash+cluster+merge phase seconds)" % (
        sample_rows, args.workload, sa + sb, iters, cores)
    # the config names what THIS arm ran (a bounded sub-sample: the full workload is out of reach of the
    # CPU path), not the GPU arm's full workload
    ref_config = {
        "workload": "%s sub-sample: mode C on the first %d of %d synthetic k-mers x %d samples (%d A + %d B), phase 1 I=1 + phase 2 "
                    "I=%d (full workload: I=%d), N=%.2f" % (args.workload, sample_rows, n, sa + sb, sa, sb, iters, args.iters,
                                                            args.min_similarity),
        "rows": sample_rows, "dim": sa + sb, "iterations": 1 + iters, "min_similarity": args.min_similarity,
        "same_config_as_gpu_arm": bool(sample_rows == n and iters == args.iters),
        "full_workload_rows": n, "full_workload_iterations": 1 + args.iters,
        "parallelism": "%d OpenMP threads (reference binary, -T %d)" % (cores, cores),
    }
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": ref_config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, n, sa, sb, world):
    n_full = synth.CONFIGS[args.workload][0]
    what = "mode C" if n == n_full else "one phase-1 batch (%d of the shape's %d k-mers; the reference clusters such batches independently) of mode C" % (n, n_full)
    return {
        "workload": "%s: %s on synthetic %d k-mers x %d samples (%d A + %d B), phase 1 I=1 + phase 2 I=%d, N=%.2f" % (
            args.workload, what, n, sa + sb, sa, sb, args.iters, args.min_similarity),
        "rows": n, "dim": sa + sb, "iterations": 1 + args.iters, "min_similarity": args.min_similarity,
        "parallelism": "1 GPU" if world == 1 else (
            "%d GPUs, one process each: Cluster() sharded inside libklsh (klsh_mg_cluster: rows replicated, bucket ranges partitioned "
            "over ranks; per iteration one 32-byte ncclAllGather of sizes and one grouped ncclBroadcast per rank carrying survivors + "
            "modified rows + chain writes over NVLink).  The communication-free decomposition (phase-1 batch per GPU + NCCL all-gather "
            "of the survivors) is timed separately in phase1_batch_per_gpu" % world),
        "l2_policy": "inputs larger than L2 (row arena %.1f GB per GPU); no flush" % (n * (sa + sb) * 4 / 1e9),
        "state_reset": "untimed device-to-device restore between steps",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("KLSH_BENCH_WORKLOAD", "C2"), choices=sorted(synth.CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debug only; invalidates the number)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--min-similarity", dest="min_similarity", type=float, default=0.80)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from kmerlsh_b200 import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: the SAME matrix on every rank (the sharded Cluster() replicates the rows and
    # partitions the merge work, so total work is fixed as N grows: strong scaling) ----------------
    n, sa, sb, counts, cov, vk, _keep = make_inputs(args.workload, args.rows, "cuda:%d" % local_rank)
    d = sa + sb
    torch.cuda.empty_cache()
    if world > 1:
        chk = torch.tensor([float(cov.sum()), float(counts[:, :: max(1, n // 4096)].astype(np.float64).sum())], device="cuda", dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise SystemExit("synthetic inputs differ between ranks; the replicated multi-GPU path needs identical rows")

    ctx = Context(local_rank, seed=42)
    p1_thr, p2_thr = 100000, 1000000

    if world > 1:
        # the NCCL communicator lives inside libklsh; torch.distributed only carries the 128-byte id
        from kmerlsh_b200 import nccl_unique_id

        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.mg_init(rank, world, uid[0])

    def one_pass():
        if world == 1:
            st = ctx.cluster(args.min_similarity, 1, p1_thr)
            st += ctx.cluster(args.min_similarity, args.iters, p2_thr)
            return st
        st = ctx.mg_cluster(args.min_similarity, 1, p1_thr)
        st += ctx.mg_cluster(args.min_similarity, args.iters, p2_thr)
        return st

    # ---- device-resident arm ------------------------------------------------------------------------
    ctx.load_counts(counts, vk, 0)
    ctx.snapshot()
    for _ in range(args.warmup):
        ctx.restore()
        ctx.set_seed(42)
        one_pass()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = ctx.launch_count()
    t_steps, rows_steps, all_stats = [], [], []
    for _ in range(args.steps):
        ctx.restore()
        ctx.set_seed(42)
        ctx.sync()
        barrier()
        t0 = time.perf_counter()
        st = one_pass()
        ctx.sync()
        t_steps.append(time.perf_counter() - t0)
        rows_steps.append(sum(s.rows_in for s in st))
        all_stats.append(st)
    barrier()
    launches = ctx.launch_count() - launches0
    dev_ms = [sum(s.ms_total for s in st) for st in all_stats]

    # ---- end-to-end arm (host buffers in, clusters out) ---------------------------------------------
    # caller-owned pinned result buffers, reused across steps (as the pinned input matrix is)
    out_v = torch.empty(n * d, dtype=torch.float32, pin_memory=True).numpy()
    out_o = torch.empty(n + 1, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    out_i = torch.empty(n, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    e2e_t, e2e_rows, d2h_bytes = [], [], 0
    for k in range(1 + args.steps):  # one warm-up
        ctx.set_seed(42)
        barrier()
        t0 = time.perf_counter()
        ctx.load_counts(counts, vk, 0)               # H2D of the uint16 matrix + transform
        st = one_pass()
        values, offs, ids = ctx.get_rows(out=(out_v, out_o, out_i))  # D2H: centroids, counts, heads, member chains
        dt = time.perf_counter() - t0
        if k:
            e2e_t.append(dt)
            e2e_rows.append(sum(s.rows_in for s in st))
            d2h_bytes = values.nbytes + 8 * (len(offs) - 1) + 4 * n
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d_bytes = counts.nbytes + vk.nbytes

    # ---- the first consumer of the result: mode E's statistics step on the clusters the
    # last pass left on the device — AB::WRS per cluster and one label per k-mer id (app/kmerLSH.cc:541-585).
    # Reported beside the headline, not in it.
    stats_block = None
    if world == 1:
        e_ms = []
        for k in range(3):
            t0 = time.perf_counter()
            label, tst = ctx.differential_ids(sa, sb, 0.01, 5, n)
            e_ms.append((time.perf_counter() - t0) * 1e3)
        stats_block = {
            "what": "klsh_differential_ids on the final clusters: two-sample t-test (ALGLIB studentttest2 semantics) per cluster with more "
                    "than 5 members, p <= 0.01, and a group label for each of the %d k-mer ids, D2H of the labels included" % n,
            "ms": min(e_ms[1:]), "clusters": int(tst.rows), "tested": int(tst.tested), "clusters_a": int(tst.rows_a),
            "clusters_b": int(tst.rows_b), "ids_a": int(tst.ids_a), "ids_b": int(tst.ids_b), "margin": int(tst.margin),
            "d2h_bytes": int(label.nbytes),
        }

    # ---- the communication-free decomposition (N > 1): phase 1 is one independent Cluster(I=1) per batch in the
    # reference (app/kmerLSH.cc:311-345), so every GPU clusters ITS OWN batch of this shape concurrently; the
    # survivors are then all-gathered in batch order over NVLink (klsh_mg_gather_rows) — the working set the
    # reference gets by appending to and re-reading its spill file.  Reported beside the headline, not in it.
    p1_block = None
    if world > 1:
        nb, _sa, _sb, bcounts, _bcov, bvk, _k2 = make_inputs(args.workload, args.rows, "cuda:%d" % local_rank, seed_offset=1000 * (rank + 1))
        ctx.load_counts(bcounts, bvk, rank * nb)
        ctx.snapshot()
        p1_t, stp = [], None
        for k in range(1 + max(1, min(args.steps, 3))):
            ctx.restore()
            ctx.set_seed(1000 + rank)
            ctx.sync()
            barrier()
            t0 = time.perf_counter()
            stp = ctx.cluster(args.min_similarity, 1, p1_thr)
            ctx.sync()
            barrier()
            if k:
                p1_t.append(time.perf_counter() - t0)
        my_surv = int(stp[0].rows_out)
        barrier()
        t0 = time.perf_counter()
        ctx.mg_gather_rows()
        ctx.sync()
        barrier()
        t_gather = time.perf_counter() - t0
        gathered = ctx.row_count(False)[0]
        tt = torch.tensor([float(np.mean(p1_t)), t_gather], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        p1_ms, g_ms = (x * 1e3 for x in tt.tolist())
        p1_block = {
            "what": "every GPU clusters its own %d-row batch (phase 1, I=1, threshold 0.95, nested above %d), no communication; then the "
                    "survivors are all-gathered in batch order with NCCL broadcasts" % (nb, p1_thr),
            "rows_per_gpu": nb, "value": world * nb / (p1_ms * 1e-3), "unit": UNIT, "ms_per_step": p1_ms, "scaling": "weak",
            "gather_ms": g_ms, "gathered_rows": int(gathered), "my_survivors": my_surv,
            "gather_GBps_per_gpu": (gathered * (4.0 * d + 12.0) + 4.0 * world * nb) / (g_ms * 1e-3) / 1e9,
        }

    # ---- aggregate over ranks (max time, sum rows) -----------------------------------------------------
    t_total, rows_total = float(sum(t_steps)), float(sum(rows_steps))
    e_total, erows_total = float(sum(e2e_t)), float(sum(e2e_rows))
    if world > 1:
        tt = torch.tensor([t_total, e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rr = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        t_total, e_total = tt.tolist()
        launches = rr.tolist()[0]  # rows are NOT summed: every rank iterates over the same (whole) row set
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (rank 0's steps) -----------------------------------
    peak, peak_src = measured_peaks()
    fam = {"sign": 0.0, "group": 0.0, "merge": 0.0, "compact": 0.0}
    merge_bytes = job_bytes = 0.0
    n_iter = 0
    for st in all_stats:
        for s in st:
            if not s.rows_in:
                continue
            n_iter += 1
            surv = s.rows_out / s.rows_in
            fam["sign"] += s.ms_sign
            fam["group"] += s.ms_group
            fam["merge"] += s.ms_merge
            fam["compact"] += s.ms_compact
            merge_bytes += s.rows_in * b_alg_merge(d, surv)
            job_bytes += s.rows_in * b_alg(d, surv)
    dom = max(fam, key=fam.get)
    dev_total_ms = sum(fam.values())
    if world > 1:  # per-family event times exist only on the single-GPU path; use the step time
        dev_total_ms = t_total * 1e3
        merge_bytes = 0.0
    sign_bytes = sum(s.rows_in * (4.0 * d + 8.0) for st in all_stats for s in st if s.rows_in)
    sign_achieved = sign_bytes / (fam["sign"] * 1e-3) / 1e9 if fam["sign"] > 0 else 0.0
    # DRAM traffic of the merge family for one step of this workload, from the committed ncu capture
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_dram_traffic_C2.json")
    if args.workload == "C2" and not args.rows and world == 1 and os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic = tj["families"]["merge"]["dram_GB"] * 1e9
        traffic_src = "bytes per step over the family's launches, " + tj["source"]
    achieved = merge_bytes / (fam["merge"] * 1e-3) / 1e9 if fam["merge"] > 0 else 0.0
    job_achieved = job_bytes / (dev_total_ms * 1e-3) / 1e9 if dev_total_ms > 0 else 0.0
    # The merge is not bound by bytes: its work is candidates x representatives x D compares (screened on the tensor
    # cores with mma.sync) and a dependent chain of window resolutions.  Counted on the device per window.
    kw = 32 if d <= 32 else (64 if d <= 64 else (d + 31) // 32 * 32)
    pairs = sum(getattr(s_, "screen_pairs", 0) for st in all_stats for s_ in st)
    exact = sum(getattr(s_, "exact_pairs", 0) for st in all_stats for s_ in st)
    clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
    mma_peak = 1324.0 * 148 * clk  # MAC/s: mma.sync.m16n8k16 back to back, measured per SM and clock (tools/microbench/mma_bench.cu)
    compare = None
    if world == 1 and fam["merge"] > 0:
        mac_s = pairs * kw / (fam["merge"] * 1e-3)
        compare = {
            "pairs_per_step": pairs / args.steps, "exact_pairs_per_step": exact / args.steps, "k_width": kw,
            "mac_per_s": mac_s, "mma_peak_mac_per_s": mma_peak, "frac_of_mma_peak": mac_s / mma_peak,
            "sum_largest_bucket_rows_per_step": sum(s_.bucket_max for st in all_stats for s_ in st if s_.rows_in) / args.steps,
            "note": "pairs = window candidates x representatives screened (fp16 mma.sync, k = k_width) + the window's own 64 x 64 block; "
                    "exact_pairs = pairs re-tested with the reference's fp32 chain.  The merge time follows the longest sequential "
                    "chain of windows (the largest bucket of each iteration), not this rate",
        }
    roofline = {
        "bound": "hbm", "kernel": "k_merge_* (in-bucket greedy merge, per iteration)", "compare": compare, "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_src,
        "algorithmic_bytes_per_step": merge_bytes / args.steps, "peak_source": peak_src,
        "share_of_step": (fam["merge"] / dev_total_ms if dev_total_ms else None) if world == 1 else None, "dominant_family": dom,
        "family_ms_per_step": {k: v / args.steps for k, v in fam.items()} if world == 1 else None,
        "streaming": ({"kernel": "k_sign_umma (tcgen05.mma kind::tf32, 3xTF32 projection with rows and sums in tensor memory, + key packing; k_sign_tc_wide above 64 columns)", "achieved": sign_achieved, "unit": "GB/s",
                       "frac": sign_achieved / peak, "bytes_per_row": "4D+8"} if world == 1 else None),
        "job": {"achieved": job_achieved, "frac": job_achieved / (peak * world), "bytes_per_row_iter": "8D+32+s(4D+12)",
                "peak_all_gpus": peak * world},
    }

    if world > 1:
        # per-family event times exist only on the single-GPU path: the block describes the whole sharded step
        roofline = {
            "bound": "hbm", "kernel": "whole step: sign + group on every rank, merge of the rank's bucket range, NCCL exchange",
            "achieved": job_achieved, "peak": peak * world, "unit": "GB/s", "frac": job_achieved / (peak * world), "traffic": None,
            "algorithmic_bytes_per_step": job_bytes / args.steps, "peak_source": peak_src, "bytes_per_row_iter": "8D+32+s(4D+12)",
            "note": "achieved = algorithmic bytes of the job / step time (max over ranks), peak = measured HBM copy peak x GPUs.  Every rank "
                    "signs and sorts all rows (replicated) and the step follows the longest bucket's sequential window chain, which is why "
                    "more GPUs shorten it so little; see roofline.compare of the 1-GPU line for the merge itself",
        }
    cpu = None
    if not args.no_cpu_baseline and os.path.exists(REF_BIN):
        cores = os.cpu_count() or 1
        sample_rows = min(n, args.cpu_sample_rows)
        v, detail = run_cpu_reference(counts, cov, sa, sb, sample_rows, args.cpu_sample_iters, args.min_similarity, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "first %d rows of the same generator x %d samples, phase 1 + I=%d, -T %d, %.1f s of hash+cluster+merge" % (
                   sample_rows, d, args.cpu_sample_iters, cores, detail["phase_seconds"])}
    elif not args.no_cpu_baseline:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref/kmerLSH_ref not present"}

    out = {
        "metric": METRIC, "value": rows_total / t_total, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, n, sa, sb, world),
        "clocks": sampler.summary(),
        "e2e": {"value": erows_total / e_total, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e_total / args.steps * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "device_ms_per_step": float(np.mean(dev_ms)) if world == 1 else None,
        "rows_iterations_per_step": rows_total / args.steps,
        "final_clusters": int(all_stats[-1][-1].rows_out),
        "phase1_ms_per_step": float(np.mean([st[0].ms_total for st in all_stats])) if world == 1 else None,
        "phase1_batch_per_gpu": p1_block,
        "mode_e_statistics": stats_block,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
