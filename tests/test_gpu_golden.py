"""GPU: the CUDA path against the golden vectors minted from the real reference, and the mode-C
command line against the reference binary's seeded T=1 output (byte-identical files)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from kmerlsh_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def load(name):
    return np.load(os.path.join(G, name))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_planes_and_keys(gpu):
    g = load("planes_keys.npz")
    gpu.set_seed(42)
    for name, (h, d) in (("t19x20", (19, 20)), ("t5x7", (5, 7)), ("t25x32", (25, 32)), ("t3x1", (3, 1))):
        assert gpu.draw_table(h, d).tobytes() == g[name].tobytes(), name
    gpu.load_counts(g["counts"], g["vk"], 0)
    v, o, i = gpu.get_rows()
    assert v.tobytes() == g["rows"].tobytes() and np.array_equal(i, g["ids"])
    assert np.array_equal(gpu.sign(g["rows"], g["t19x20"]), g["keys"].astype(np.uint64))


def test_convert_lut_all_counts(gpu):
    g = load("convert.npz")
    gpu.load_counts(np.arange(65536, dtype=np.uint16).reshape(1, -1), np.zeros(1, np.float32), 0)
    v, o, i = gpu.get_rows()
    assert np.array_equal(i, g["lut_ids"]) and v[:, 0].tobytes() == g["lut_values"].tobytes()
    gpu.load_counts(g["edge_counts"], np.linspace(0.1, 2.0, 20).astype(np.float32), 77)
    v, o, i = gpu.get_rows()
    assert np.array_equal(i, g["edge_ids"]) and v.tobytes() == g["edge_values"].tobytes()


def test_p_cluster_kats(gpu):
    g = load("p_cluster.npz")
    for key in sorted(k[:-3] for k in g.files if k.endswith("_in")):
        thr = int(key.rsplit("_", 1)[1]) / 100.0
        gpu.set_rows(g[key + "_in"])
        gpu.p_cluster(thr)
        v, o, i = gpu.get_rows()
        assert v.tobytes() == g[key + "_values"].tobytes(), key
        assert np.array_equal(o, g[key + "_offs"]) and np.array_equal(i, g[key + "_ids"]), key


def test_consensus_counts_beyond_2_24(gpu):
    """Member counts 2^24 and 2^24+1: the int->float conversion must round like
    cvtsi2ss.  Through p_cluster with that many REAL member ids (the merge kernels' consensus) ..."""
    g = load("scalar_kats.npz")
    for k, (c1, c2) in enumerate(g["cons_counts"]):
        c1, c2 = int(c1), int(c2)
        a, b = g["a"][10 + k], g["b"][10 + k]
        # two rows that certainly merge at threshold -1: current=row1 (c1 ids), candidate=row0 (c2 ids)
        values = np.stack([b, a])
        offs = np.array([0, c2, c1 + c2], dtype=np.uint64)
        gpu.set_rows(values, offs, np.arange(c1 + c2, dtype=np.uint64))
        gpu.p_cluster(-1.0)
        v, o, i = gpu.get_rows()
        assert len(v) == 1 and v[0].tobytes() == g["cons"][k].tobytes(), (c1, c2)
        assert len(i) == c1 + c2 and np.array_equal(i[:c1], np.arange(c2, c1 + c2, dtype=np.uint64)), (c1, c2)
        assert np.array_equal(i[c1:], np.arange(c2, dtype=np.uint64)), (c1, c2)


def test_set_consensus_function(gpu):
    """... and through the function-level entry klsh_set_consensus (AB::SetConsensus, funcAB.cc:49-71)."""
    g = load("scalar_kats.npz")
    for k, (c1, c2) in enumerate(g["cons_counts"]):
        got = gpu.set_consensus(g["a"][10 + k], int(c1), g["b"][10 + k], int(c2))
        assert got.tobytes() == g["cons"][k].tobytes(), (c1, c2)


def test_cosine_distance_function(gpu):
    """klsh_cosine_distance = Distance::cosine (distance.cc:27-38) on the reference's own outputs for
    64 adversarial pairs: zero vector (NaN), identical, opposite, 1e-20 and 1e18 magnitudes."""
    g = load("scalar_kats.npz")
    got = gpu.cosine_distance(g["a"], g["b"])
    assert got.view(np.uint32).tolist() == g["dist"].view(np.uint32).tolist() or all(
        (x == y) or (np.isnan(x) and np.isnan(y)) for x, y in zip(got.tolist(), g["dist"].tolist()))
    assert np.isnan(got[0]) and got[1] <= 1e-6


@pytest.mark.parametrize("tag", ["plain", "nested", "one_iter"])
def test_cluster_kats(gpu, tag):
    g = load("cluster.npz")
    seed, iters, thr = (int(x) for x in g[tag + "_params"])
    gpu.set_seed(seed)
    gpu.load_counts(g["counts"], g["vk"], 0)
    gpu.cluster(float(g[tag + "_minsim"]), iters, thr)
    v, o, i = gpu.get_rows()
    assert v.tobytes() == g[tag + "_values"].tobytes()
    assert np.array_equal(o, g[tag + "_offs"]) and np.array_equal(i, g[tag + "_ids"])


@pytest.mark.parametrize("tag,extra", [("modec_small", []), ("modec_C1", []), ("modec_small", ["--reload-tmp"]),
                                       ("modec_small", ["--no-tmp-files"])])
def test_mode_c_cli_matches_reference_binary(tag, extra, tmp_path):
    """kmerLSH_b200 -M C --only ... in a directory laid out like the reference's: the output files
    and the phase-1 spill are byte-identical to the seeded T=1 reference run — with the phase-1
    survivors kept on the device (default for a single batch), with the reference's re-read of the
    spill files (--reload-tmp), and without writing the spill at all (--no-tmp-files)."""
    m = json.load(open(os.path.join(G, "golden.json")))[tag]
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, m["n"], m["sa"], m["sb"], m["gen_seed"])
    if md5(os.path.join(work, "kmer_count.bin")) != m["kmer_count_bin_md5"]:
        pytest.skip("numpy generator stream differs from the one the golden run used")
    exe = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")
    subprocess.run([exe, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", str(m["iters"]),
                    "-N", str(m["min_similarity"]), "-K", "23", "-T", "1", "--seed=%d" % m["klsh_seed"]] + extra, cwd=work,
                   check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(work, "clustering_result.txt")
    if "--no-tmp-files" in extra:
        assert not os.path.exists(os.path.join(work, "tmp", "0.bin"))
    else:
        assert md5(os.path.join(work, "tmp", "0.bin")) == m["tmp_bin_md5"]
        assert md5(os.path.join(work, "tmp", "0.bin.clust")) == m["tmp_clust_md5"]
    assert md5(out) == m["bin_md5"]
    assert md5(out + ".clust") == m["clust_md5"]


@pytest.mark.parametrize("tag", ["modec_small", "modec_C1"])
def test_reference_program_with_cluster_shim(tag, tmp_path):
    """The drop-in proof for function/cluster.h:42: oracle/_ref/kmerLSH_shim is the reference's OWN
    app/kmerLSH.cc and objects, linked with oracle/cluster_b200.cc in place of
    function/cluster.o, so every Cluster() call of mode C (app/kmerLSH.cc:323, :490) runs on libklsh
    with the reference's own LSH::generateHashTable behind the plane callback.  Its output files must be
    byte-identical to the unmodified reference binary's (the golden md5s)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "kmerLSH_shim")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/kmerLSH_shim not built (needs /root/reference at build time)")
    m = json.load(open(os.path.join(G, "golden.json")))[tag]
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, m["n"], m["sa"], m["sb"], m["gen_seed"])
    if md5(os.path.join(work, "kmer_count.bin")) != m["kmer_count_bin_md5"]:
        pytest.skip("numpy generator stream differs from the one the golden run used")
    env = dict(os.environ, KLSH_SEED=str(m["klsh_seed"]), OMP_THREAD_LIMIT="1")
    subprocess.run([exe, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", str(m["iters"]),
                    "-N", str(m["min_similarity"]), "-K", "23", "-T", "1"], cwd=work, env=env, check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(work, "clustering_result.txt")
    assert md5(os.path.join(work, "tmp", "0.bin")) == m["tmp_bin_md5"]
    assert md5(os.path.join(work, "tmp", "0.bin.clust")) == m["tmp_clust_md5"]
    assert md5(out) == m["bin_md5"]
    assert md5(out + ".clust") == m["clust_md5"]


@pytest.mark.parametrize("extra", [[], ["--resident"], ["--resident", "--no-tmp-files"], ["--binary-tmp-ids"]])
def test_mode_c_cli_two_batches(oracle, tmp_path, extra):
    """--batch smaller than the input: independent phase-1 batches appended to tmp/0.bin, the
    re-batch loop (similarity -= 0.001, 5 iterations per batch) while survivors exceed the batch
    size, then the -I iterations — byte-identical to the oracle's mode C with the same batch size
    (the reference hard-codes 100 M rows per batch, app/kmerLSH.cc:285).  With --resident the survivors
    of every batch are appended to a device-resident stash (klsh_stash_rows) instead of being re-read from
    the spill files; with --no-tmp-files the spill is only written if a re-batch round needs it; with
    --binary-tmp-ids the spill files' member lists are binary (<n>.bin.clust.bin) and the result is the same text."""
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, 30000, 3, 3, 99)
    out = os.path.join(work, "oracle_result.txt")
    os.makedirs(os.path.join(work, "otmp"))
    oracle.mode_c(work, 6, 0.85, 4, out, 17, batch_thresh=8000, tmp_dir=os.path.join(work, "otmp") + "/")
    exe = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")
    subprocess.run([exe, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", "4", "-N", "0.85",
                    "-T", "1", "--seed=17", "--batch=8000"] + extra, cwd=work, check=True, stdout=subprocess.DEVNULL)
    res = os.path.join(work, "clustering_result.txt")
    assert md5(res) == md5(out)
    assert md5(res + ".clust") == md5(out + ".clust")
    if "--binary-tmp-ids" in extra:
        spills = os.listdir(os.path.join(work, "tmp"))
        assert any(f.endswith(".clust.bin") for f in spills) and not any(f.endswith(".clust") for f in spills), spills


def test_binary_member_lists_round_trip(oracle, gpu, tmp_path):
    """klsh_set_id_format(1): <F>.clust.bin holds per cluster a uint64 count and the ids; saving (with append and
    with ignore_small) and reading back (whole file and a window of lines) give the rows the text format gives."""
    from helpers import assert_rows_equal, synth_rows

    _, _, values, _ = synth_rows(oracle, 5000, 4, 4, 12)
    gpu.set_id_format(0)
    gpu.set_rows(values)
    gpu.set_seed(3)
    gpu.cluster(0.85, 4, 100000)
    want = gpu.get_rows()
    t, b = str(tmp_path / "t.bin"), str(tmp_path / "b.bin")
    gpu.save(t, True, 0)
    gpu.save(t, False, 2)            # append: clusters of more than 2 members once more
    try:
        gpu.set_id_format(1)
        gpu.save(b, True, 0)
        gpu.save(b, False, 2)
        assert md5(t) == md5(b)      # the centroid file does not depend on the id format
        assert not os.path.exists(b + ".clust")
        raw = np.fromfile(b + ".clust.bin", dtype=np.uint64)
        assert raw[0] == want[1][1] - want[1][0] and np.array_equal(raw[1:1 + int(raw[0])], want[2][: int(raw[0])])
        gpu.load_cluster_file(b, 8)
        got_b = gpu.get_rows()
        gpu.load_cluster_file(b, 8, 7, 40)
        win_b = gpu.get_rows()
    finally:
        gpu.set_id_format(0)
    gpu.load_cluster_file(t, 8)
    assert_rows_equal(got_b, gpu.get_rows(), "binary member lists, whole file")
    gpu.load_cluster_file(t, 8, 7, 40)
    assert_rows_equal(win_b, gpu.get_rows(), "binary member lists, lines 7..46")
    assert len(got_b[1]) - 1 > len(want[1]) - 1


def test_mode_c_cli_resident_batches_without_rebatch(oracle, tmp_path):
    """Several phase-1 batches whose survivors fit one batch: with --resident --no-tmp-files no spill file is
    ever written and the -I iterations start from the device-resident stash — same output as the oracle."""
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, 20000, 3, 3, 5)
    out = os.path.join(work, "oracle_result.txt")
    os.makedirs(os.path.join(work, "otmp"))
    oracle.mode_c(work, 6, 0.85, 4, out, 23, batch_thresh=14000, tmp_dir=os.path.join(work, "otmp") + "/")
    exe = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")
    subprocess.run([exe, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", "4", "-N", "0.85",
                    "-T", "1", "--seed=23", "--batch=14000", "--resident", "--no-tmp-files"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    res = os.path.join(work, "clustering_result.txt")
    assert md5(res) == md5(out) and md5(res + ".clust") == md5(out + ".clust")
    assert not os.path.exists(os.path.join(work, "tmp", "0.bin"))


@pytest.mark.parametrize("gpus", [2, 3])
def test_mode_c_cli_several_workers(oracle, tmp_path, gpus):
    """--gpus=N: phase-1 batches and re-batch rounds run round-robin on N worker contexts (sharing the
    physical GPUs when there are fewer), the hash tables still come from ONE seeded stream in the
    reference's call order (a context draws only at its call's turn and hands the stream on through the
    draws-done callback), spills are appended in batch order.  The output must be byte-identical to the
    oracle's sequential mode C, i.e. to --gpus=1."""
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, 60000, 3, 3, 123)
    out = os.path.join(work, "oracle_result.txt")
    os.makedirs(os.path.join(work, "otmp"))
    oracle.mode_c(work, 6, 0.85, 4, out, 29, batch_thresh=9000, tmp_dir=os.path.join(work, "otmp") + "/")
    exe = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")
    stats = os.path.join(work, "stats.json")
    subprocess.run([exe, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "C", "--only", "-I", "4", "-N", "0.85",
                    "-T", "1", "--seed=29", "--batch=9000", "--gpus=%d" % gpus, "--stats-json=" + stats], cwd=work, check=True,
                   stdout=subprocess.DEVNULL)
    res = os.path.join(work, "clustering_result.txt")
    assert md5(res) == md5(out)
    assert md5(res + ".clust") == md5(out + ".clust")
    recs = [json.loads(ln) for ln in open(stats)]
    assert {r["phase"] for r in recs} >= {"phase1"} and all(r["rows_out"] <= r["rows_in"] for r in recs)
    assert len({r["gpu"] for r in recs if r["phase"] == "phase1"}) == gpus
