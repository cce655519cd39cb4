"""Mint the golden fixtures from the REAL reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run here, in the build container (`python tests/golden/make_golden.py`);
the outputs (*.npz, golden.json) are committed and travel to the GPU box, the reference does not.

The reference ships no tests or golden vectors, so these known-answer tests
are outputs of the reference's own functions called through oracle/ref_harness.cc, plus md5s of a
full seeded `kmerLSH -M C` run of the reference binary (T=1, OMP_THREAD_LIMIT=1).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

os.environ["OMP_THREAD_LIMIT"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from oracle_lib import REF_BIN, RefLib  # noqa: E402

from kmerlsh_b200 import synth  # noqa: E402


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    r = RefLib()
    meta = {}
    rng = np.random.default_rng(20261018)

    # (1) hyperplanes + keys: seeded tables and LSH::random_projection on C1-law rows, all 19 bits
    r.reseed(42)
    tables = {"t19x20": r.table(19, 20), "t5x7": r.table(5, 7), "t25x32": r.table(25, 32), "t3x1": r.table(3, 1)}
    meta["planes_master_draws"] = int(r.draws())
    counts, cov = synth.synth_counts(6000, 10, 10, 20261018)
    kmap, cov32 = synth.parse_log_line(synth.format_log_line(6000, cov), 20)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    rows20, ids20 = r.convert_counts(counts, vk, 0)
    keys = r.sign(rows20, tables["t19x20"])
    np.savez_compressed(os.path.join(HERE, "planes_keys.npz"), counts=counts, vk=vk, rows=rows20, ids=ids20, keys=keys, **tables)

    # (2) cosine on adversarial pairs
    a = rng.standard_normal((64, 20)).astype(np.float32)
    b = (a + np.float32(0.3) * rng.standard_normal((64, 20))).astype(np.float32)
    a[0] = 0.0                      # zero vector -> NaN
    b[1] = a[1]                     # identical
    b[2] = -a[2]                    # opposite
    a[3] *= np.float32(1e-20)       # tiny magnitudes
    b[4] *= np.float32(1e18)        # huge magnitudes
    dist = np.array([r.cosine_distance(a[i], b[i]) for i in range(64)], dtype=np.float32)
    # (3) SetConsensus with member counts 1, 2^24, 2^24+1 (int -> float rounding)
    cc = [(1, 1), (1, 7), (3, 2), (1 << 24, 1), ((1 << 24) + 1, 1), (1, (1 << 24) + 1), (123456, 654321)]
    cons = np.stack([r.consensus(a[10 + i], c1, b[10 + i], c2) for i, (c1, c2) in enumerate(cc)])
    np.savez_compressed(os.path.join(HERE, "scalar_kats.npz"), a=a, b=b, dist=dist, cons_counts=np.array(cc, dtype=np.int64), cons=cons)

    # (4) p_cluster on hand-built buckets: swap-remove order, chain merges, merged rep not re-compared
    base = rng.standard_normal((1, 12)).astype(np.float32)
    buckets = {
        "identical": np.repeat(base, 9, axis=0),
        "noisy": (base + np.float32(0.25) * rng.standard_normal((80, 12))).astype(np.float32),
        "two_groups": np.concatenate([base + np.float32(0.2) * rng.standard_normal((15, 12)),
                                      -base + np.float32(0.2) * rng.standard_normal((15, 12))]).astype(np.float32)[rng.permutation(30)],
        "random": rng.standard_normal((40, 12)).astype(np.float32),
    }
    pc = {}
    for name, v in buckets.items():
        for thr in (0.95, 0.8, 0.5):
            ov, oo, oi = r.p_cluster(v, thr)
            key = "%s_%d" % (name, int(thr * 100))
            pc[key + "_in"] = v
            pc[key + "_values"] = ov
            pc[key + "_offs"] = oo
            pc[key + "_ids"] = oi
    np.savez_compressed(os.path.join(HERE, "p_cluster.npz"), **pc)

    # (6) convertHTMat: LUT for all 65536 counts (D=1, v_kmers=0) and the keep-filter edge at D=20
    allc = np.arange(65536, dtype=np.uint16).reshape(1, -1)
    lv, li = r.convert_counts(allc, np.zeros(1, np.float32), 0)
    edge = np.zeros((20, 4), dtype=np.uint16)
    edge[0, 1] = 2                  # total 2 == 0.1*20 -> dropped
    edge[0, 2] = 3                  # total 3 -> kept
    edge[:, 3] = 1                  # total 20 -> kept
    ev, ei = r.convert_counts(edge, np.linspace(0.1, 2.0, 20).astype(np.float32), 77)
    np.savez_compressed(os.path.join(HERE, "convert.npz"), lut_values=lv[:, 0], lut_ids=li, edge_counts=edge, edge_values=ev, edge_ids=ei)

    # (8)+(9) Cluster: multi-iteration, nested buckets (RNG consumption order), N=1
    cl = {}
    counts8, cov8 = synth.synth_counts(12000, 4, 4, 7)
    kmap8, c32 = synth.parse_log_line(synth.format_log_line(12000, cov8), 8)
    vk8 = synth.v_kmers_from_cov(c32, kmap8)
    rows8, ids8 = r.convert_counts(counts8, vk8, 0)
    cl["counts"] = counts8
    cl["vk"] = vk8
    for tag, (seed, iters, minsim, thr) in {"plain": (7, 8, 0.85, 100000), "nested": (7, 3, 0.85, 50), "one_iter": (11, 1, 0.8, 12)}.items():
        r.reseed(seed)
        ov, oo, oi = r.cluster(rows8, minsim, iters, thr)
        cl[tag + "_params"] = np.array([seed, iters, thr], dtype=np.int64)
        cl[tag + "_minsim"] = np.float32(minsim)
        cl[tag + "_values"] = ov
        cl[tag + "_offs"] = oo
        cl[tag + "_ids"] = oi
        cl[tag + "_draws"] = np.int64(r.draws())
    r.reseed(5)
    ov, oo, oi = r.cluster(rows8[:1], 0.8, 3, 1000)
    cl["single_values"], cl["single_offs"], cl["single_ids"] = ov, oo, oi
    r.reseed(13)
    ov, oo, oi = r.nested_cluster(rows8[:5000], 0.9)
    cl["nestedfn_values"], cl["nestedfn_offs"], cl["nestedfn_ids"] = ov, oo, oi
    np.savez_compressed(os.path.join(HERE, "cluster.npz"), **cl)

    # (7)+(5) full mode C with the reference BINARY: C1 (1M x 20, I=100, N=0.80) and a small case
    for tag, (n, sa, sb, gseed, iters, minsim, kseed) in {
        "modec_small": (40000, 4, 4, 7, 6, 0.85, 7),
        "modec_C1": (1000000, 10, 10, 20261018, 100, 0.80, 42),
    }.items():
        work = tempfile.mkdtemp(prefix="klsh_golden_")
        synth.write_mode_c_inputs(work, n, sa, sb, gseed)
        env = dict(os.environ, KLSH_SEED=str(kseed), OMP_THREAD_LIMIT="1")
        out = subprocess.run([REF_BIN, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", str(iters),
                              "-N", str(minsim), "-K", "23", "-T", "1", "--verbose"], cwd=work, env=env, check=True,
                             stdout=subprocess.PIPE, text=True).stdout
        thr_lines = [ln.split("\t")[2].split(" ")[0] for ln in out.splitlines() if ln.startswith("Iteration:")]
        sizes = [int(ln.split(":")[1]) for ln in out.splitlines() if ln.startswith("Size of profilings")]
        f = os.path.join(work, "clustering_result.txt")
        meta[tag] = {
            "n": n, "sa": sa, "sb": sb, "gen_seed": gseed, "iters": iters, "min_similarity": minsim, "klsh_seed": kseed,
            "bin_md5": md5(f), "clust_md5": md5(f + ".clust"), "bin_bytes": os.path.getsize(f),
            "clust_bytes": os.path.getsize(f + ".clust"),
            "tmp_bin_md5": md5(os.path.join(work, "tmp", "0.bin")), "tmp_clust_md5": md5(os.path.join(work, "tmp", "0.bin.clust")),
            "thresholds_printed": thr_lines, "rows_in": sizes,
            "kmer_count_bin_md5": md5(os.path.join(work, "kmer_count.bin")),
            "kmer_count_log": open(os.path.join(work, "kmer_count.log")).read(),
        }
        shutil.rmtree(work)
    with open(os.path.join(HERE, "golden.json"), "w") as fjs:
        json.dump(meta, fjs, indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
