"""One-off large parity check (not a pytest: the oracle needs minutes of CPU at this size):
phase 1 + I iterations on N rows of a named shape, CUDA path vs the C oracle, bit for bit.

  python tools/parity_large.py C2 5000000 10            both sides in one process (needs a GPU)
  python tools/parity_large.py C2 5000000 10 gpu  F     CUDA side only, md5 digests written to F (JSON)
  python tools/parity_large.py C2 5000000 10 oracle F   oracle side only (no GPU needed)
The split form lets the oracle's minutes run on a box without a GPU; compare the two JSON files.
"""
import hashlib, json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from kmerlsh_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
part = sys.argv[4] if len(sys.argv) > 4 else "both"
out_path = sys.argv[5] if len(sys.argv) > 5 else None
_, sa, sb, seed = synth.CONFIGS[cfg]
counts, cov = synth.synth_counts(n, sa, sb, seed)
kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
vk = synth.v_kmers_from_cov(cov32, kmap)


def digest(res):
    return {"clusters": int(len(res[1]) - 1), "ids": int(len(res[2])),
            "values_md5": hashlib.md5(np.ascontiguousarray(res[0]).view(np.uint8)).hexdigest(),
            "offsets_md5": hashlib.md5(np.ascontiguousarray(res[1]).astype(np.uint64).view(np.uint8)).hexdigest(),
            "ids_md5": hashlib.md5(np.ascontiguousarray(res[2]).astype(np.uint64).view(np.uint8)).hexdigest(),
            "config": [cfg, n, iters], "counts_md5": hashlib.md5(counts.view(np.uint8)).hexdigest()}


got = want = None
if part in ("both", "gpu"):
    from kmerlsh_b200 import Context

    t = time.time()
    ctx = Context(0, seed=42)
    ctx.load_counts(counts, vk, 0)
    s1 = ctx.cluster(0.80, 1, n // 1000)
    s2 = ctx.cluster(0.80, iters, 1000000)
    got = ctx.get_rows()
    print("gpu %.1fs: %d -> %d -> %d clusters; bmax %s nested %d eps rows %d" % (
        time.time() - t, s1[0].rows_in, s1[0].rows_out, s2[-1].rows_out, max(x.bucket_max for x in s1 + s2),
        sum(x.nested_calls for x in s1 + s2), sum(x.eps_margin_rows for x in s1 + s2)), flush=True)
if part in ("both", "oracle"):
    from oracle_lib import Oracle

    t = time.time()
    o = Oracle()
    values, ids = o.convert_counts(counts, vk, 0)
    rows = o.rows(values, np.arange(len(ids) + 1, dtype=np.uint64), ids)
    planes = o.planes(42)
    rows.cluster(0.80, 1, n // 1000, planes)
    rows.cluster(0.80, iters, 1000000, planes)
    want = rows.export()
    print("oracle %.1fs" % (time.time() - t), flush=True)
if part == "both":
    ok = got[0].tobytes() == want[0].tobytes() and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    print("IDENTICAL" if ok else "MISMATCH", len(want[1]) - 1, "clusters,", len(want[2]), "ids")
    sys.exit(0 if ok else 1)
d = digest(got if part == "gpu" else want)
print(json.dumps(d))
if out_path:
    json.dump(d, open(out_path, "w"))
