"""Dev probe: where the end-to-end time goes (host buffers in, clusters out)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kmerlsh_b200 import Context, synth
from kmerlsh_b200.synth_gpu import synth_counts_gpu
n, sa, sb, seed = synth.CONFIGS["C2"]
counts, cov = synth_counts_gpu(n, sa, sb, seed)
kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
vk = synth.v_kmers_from_cov(cov32, kmap)
ctx = Context(0, seed=42)
for rep in range(2):
    ctx.set_seed(42)
    t0 = time.perf_counter(); ctx.load_counts(counts, vk, 0); t1 = time.perf_counter()
    ctx.cluster(0.8, 1, 100000); ctx.cluster(0.8, 100, 1000000); t2 = time.perf_counter()
    nrows, nids = ctx.row_count(); t3 = time.perf_counter()
    v, o, i = ctx.get_rows(); t4 = time.perf_counter()
    print("rep %d: load %.3f cluster %.3f row_count %.3f get_rows %.3f (rows %d ids %d)" % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, nrows, nids), flush=True)
