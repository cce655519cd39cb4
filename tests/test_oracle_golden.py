"""CPU: the oracle (C restatement) against the golden vectors minted from the real reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from kmerlsh_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_planes_match_reference_generator(oracle):
    g = load("planes_keys.npz")
    p = oracle.planes(42)
    for name, (h, d) in (("t19x20", (19, 20)), ("t5x7", (5, 7)), ("t25x32", (25, 32)), ("t3x1", (3, 1))):
        assert p.table(h, d).tobytes() == g[name].tobytes(), name
    assert p.draws() == json.load(open(os.path.join(G, "golden.json")))["planes_master_draws"]


def test_transform_and_keys(oracle):
    g = load("planes_keys.npz")
    values, ids = oracle.convert_counts(g["counts"], g["vk"], 0)
    assert values.tobytes() == g["rows"].tobytes()
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(oracle.sign(values, g["t19x20"]), g["keys"])


def test_synth_generator_is_stable():
    """The committed counts are the generator's output (numpy stream unchanged)."""
    g = load("planes_keys.npz")
    counts, _ = synth.synth_counts(6000, 10, 10, 20261018)
    assert np.array_equal(counts, g["counts"])


def test_cosine_kats(oracle):
    g = load("scalar_kats.npz")
    got = np.array([oracle.cosine_distance(g["a"][i], g["b"][i]) for i in range(len(g["a"]))], dtype=np.float32)
    assert got.tobytes() == g["dist"].tobytes()
    assert np.isnan(got[0])            # zero vector never merges
    assert got[1] <= np.float32(1e-6)  # identical vectors


def test_consensus_kats(oracle):
    g = load("scalar_kats.npz")
    for i, (c1, c2) in enumerate(g["cons_counts"]):
        got = oracle.consensus(g["a"][10 + i], int(c1), g["b"][10 + i], int(c2))
        assert got.tobytes() == g["cons"][i].tobytes(), (c1, c2)


def test_p_cluster_kats(oracle):
    g = load("p_cluster.npz")
    keys = sorted(k[:-3] for k in g.files if k.endswith("_in"))
    assert len(keys) == 12
    for key in keys:
        thr = int(key.rsplit("_", 1)[1]) / 100.0
        rows = oracle.rows(g[key + "_in"])
        rows.p_cluster(thr)
        v, o, i = rows.export()
        assert v.tobytes() == g[key + "_values"].tobytes(), key
        assert np.array_equal(o, g[key + "_offs"]) and np.array_equal(i, g[key + "_ids"]), key


def test_convert_lut_and_filter_edge(oracle):
    g = load("convert.npz")
    lut = oracle.log_lut()
    assert np.array_equal(g["lut_ids"], np.arange(1, 65536, dtype=np.uint64))  # count 0 is dropped at D=1
    assert lut[1:].tobytes() == g["lut_values"].tobytes()
    v, i = oracle.convert_counts(g["edge_counts"], np.linspace(0.1, 2.0, 20).astype(np.float32), 77)
    assert np.array_equal(i, g["edge_ids"]) and v.tobytes() == g["edge_values"].tobytes()
    assert list(i) == [79, 80]  # total 2 == 0.1*20 dropped, 3 and 20 kept


@pytest.mark.parametrize("tag", ["plain", "nested", "one_iter"])
def test_cluster_kats(oracle, tag):
    g = load("cluster.npz")
    values, ids = oracle.convert_counts(g["counts"], g["vk"], 0)
    seed, iters, thr = (int(x) for x in g[tag + "_params"])
    planes = oracle.planes(seed)
    rows = oracle.rows(values)
    rows.cluster(float(g[tag + "_minsim"]), iters, thr, planes)
    v, o, i = rows.export()
    assert v.tobytes() == g[tag + "_values"].tobytes()
    assert np.array_equal(o, g[tag + "_offs"]) and np.array_equal(i, g[tag + "_ids"])
    assert planes.draws() == int(g[tag + "_draws"])  # nested tables consumed in the reference's order


def test_cluster_single_row_and_nested_fn(oracle):
    g = load("cluster.npz")
    values, _ = oracle.convert_counts(g["counts"], g["vk"], 0)
    rows = oracle.rows(values[:1])
    rows.cluster(0.8, 3, 1000, oracle.planes(5))
    v, o, i = rows.export()
    assert v.tobytes() == g["single_values"].tobytes() and np.array_equal(i, g["single_ids"])
    rows = oracle.rows(values[:5000])
    rows.nested_cluster(0.9, oracle.planes(13))
    v, o, i = rows.export()
    assert v.tobytes() == g["nestedfn_values"].tobytes()
    assert np.array_equal(o, g["nestedfn_offs"]) and np.array_equal(i, g["nestedfn_ids"])


def test_threshold_recurrence(oracle):
    """fp32 recurrence threshold -= step (function/cluster.cc:330), against the values the reference
    binary printed, and its closed-form neighbours at 100/500 steps."""
    meta = json.load(open(os.path.join(G, "golden.json")))["modec_C1"]
    printed = meta["thresholds_printed"]  # phase-1 call (1 line) then the 100 phase-2 iterations
    assert printed[0] == "0.95"
    for k in range(100):
        assert "%g" % oracle.threshold_after(0.80, 100, k) == printed[1 + k], k
    for minsim, iters in ((0.80, 100), (0.90, 500)):
        t = np.float32(0.95)
        step = np.float32((np.float32(0.95) - np.float32(minsim)) / np.float32(iters))
        for _ in range(iters):
            t = np.float32(t - step)
        assert oracle.threshold_after(minsim, iters, iters) == t
        assert abs(float(t) - minsim) < 1e-4


@pytest.mark.parametrize("tag", ["modec_small", "modec_C1"])
def test_mode_c_end_to_end(oracle, tag, tmp_path):
    """Whole mode C (transform -> phase 1 -> tmp spill -> reload -> -I iterations -> >5 filter):
    byte-identical files to the reference binary's seeded T=1 run."""
    m = json.load(open(os.path.join(G, "golden.json")))[tag]
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, m["n"], m["sa"], m["sb"], m["gen_seed"])
    if md5(os.path.join(work, "kmer_count.bin")) != m["kmer_count_bin_md5"]:
        pytest.skip("numpy generator stream differs from the one the golden run used")
    assert open(os.path.join(work, "kmer_count.log")).read() == m["kmer_count_log"]
    out = os.path.join(work, "clustering_result.txt")
    stats = oracle.mode_c(work, m["sa"] + m["sb"], m["min_similarity"], m["iters"], out, m["klsh_seed"])
    assert [s.rows_in for s in stats] == m["rows_in"][1:]
    assert os.path.getsize(out) == m["bin_bytes"]
    assert md5(out) == m["bin_md5"]
    assert md5(out + ".clust") == m["clust_md5"]
    assert md5(os.path.join(work, "tmp", "0.bin")) == m["tmp_bin_md5"]
    assert md5(os.path.join(work, "tmp", "0.bin.clust")) == m["tmp_clust_md5"]


def test_mode_c_two_batches(oracle, tmp_path):
    """Reduced batch size: independent phase-1 batches appended to one tmp file, then the re-batch
    loop (similarity -= 0.001, 5 iterations per batch) while survivors exceed the batch size."""
    work = str(tmp_path)
    counts, cov = synth.write_mode_c_inputs(work, 30000, 3, 3, 99)
    out = os.path.join(work, "r.txt")
    oracle.mode_c(work, 6, 0.85, 4, out, 17, batch_thresh=8000, phase2_bucket_threshold=1000000)
    # same thing by hand with the function-level API
    kmap, cov32 = synth.parse_log_line(open(os.path.join(work, "kmer_count.log")).read(), 6)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    planes = oracle.planes(17)
    parts = []
    for b in range(0, 30000, 8000):
        v, ids = oracle.convert_counts(counts[:, b:b + 8000], vk, b)
        rows = oracle.rows(v, np.arange(len(ids) + 1, dtype=np.uint64), ids)
        rows.cluster(0.85, 1, 8, planes)
        parts.append(rows.export())
    total = sum(len(p[1]) - 1 for p in parts)
    sim = np.float32(0.85)
    while total > 8000:
        sim = np.float32(np.float64(sim) - 0.001)
        v = np.concatenate([p[0] for p in parts])
        offs = [np.uint64(0)]
        for p in parts:
            offs.extend(list(p[1][1:] + offs[-1]))
        offs = np.array(offs, dtype=np.uint64)
        ids = np.concatenate([p[2] for p in parts])
        parts = []
        for b in range(0, total, 8000):
            e = min(b + 8000, total)
            o = offs[b:e + 1] - offs[b]
            rows = oracle.rows(v[b:e], o, ids[int(offs[b]):int(offs[e])])
            rows.cluster(float(sim), 5, 8, planes)
            parts.append(rows.export())
        total = sum(len(p[1]) - 1 for p in parts)
    v = np.concatenate([p[0] for p in parts])
    offs = [np.uint64(0)]
    for p in parts:
        offs.extend(list(p[1][1:] + offs[-1]))
    rows = oracle.rows(v, np.array(offs, dtype=np.uint64), np.concatenate([p[2] for p in parts]))
    rows.cluster(0.85, 4, 1000000, planes)
    rows.save(os.path.join(work, "hand.txt"), True, 5)
    assert md5(out) == md5(os.path.join(work, "hand.txt"))
    assert md5(out + ".clust") == md5(os.path.join(work, "hand.txt.clust"))
