"""Dev probe: run mode C (phase 1 I=1, phase 2 I=iters) on a synthetic shape and print per-iteration stats."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from kmerlsh_b200 import Context, synth

import os
ALL = bool(os.environ.get('PROBE_ALL'))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n, sa, sb, seed = synth.CONFIGS[cfg]
if len(sys.argv) > 3: n = int(sys.argv[3])
t = time.time()
if cfg == "C1":
    counts, cov = synth.synth_counts(n, sa, sb, seed)
else:
    from kmerlsh_b200.synth_gpu import synth_counts_gpu
    counts, cov = synth_counts_gpu(n, sa, sb, seed)
print("synth %.1fs" % (time.time() - t), counts.shape, flush=True)
kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
vk = synth.v_kmers_from_cov(cov32, kmap)
ctx = Context(0, seed=42)
t = time.time(); ctx.load_counts(counts, vk, 0); print("load %.3fs rows" % (time.time() - t), ctx.row_count(False)[0], flush=True)
def show(tag, st, wall):
    tot_rows = sum(s.rows_in for s in st); tot_ms = sum(s.ms_total for s in st)
    print("%s: wall %.3fs device %.1f ms rows_in %d -> %.3e rows/s (device), %.3e rows/s (wall)" % (tag, wall, tot_ms, tot_rows, tot_rows / tot_ms * 1e3, tot_rows / wall))
    for k, s in enumerate(st):
        if k < 6 or k % 10 == 9 or ALL:
            print("  it %3d in %9d out %9d H %2d nb %8d bmax %7d nest %d | sign %.3f group %.3f merge %.3f compact %.3f ms" % (
                k + 1, s.rows_in, s.rows_out, s.H, s.buckets, s.bucket_max, s.nested_calls, s.ms_sign, s.ms_group, s.ms_merge, s.ms_compact))
t = time.time(); st = ctx.cluster(0.80, 1, 100000); show("phase1", st, time.time() - t)
t = time.time(); st = ctx.cluster(0.80, iters, 1000000); show("phase2", st, time.time() - t)
print("final rows", ctx.row_count(False)[0], "launches", ctx.launch_count())
