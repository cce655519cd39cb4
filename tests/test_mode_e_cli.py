"""GPU: `kmerLSH_b200 -M E` (statistics + read extraction, kmerlsh_b200/host/modee.cc) against the reference binary's
seeded run: the extracted-read files are byte-identical (md5s in tests/golden/modee.json, minted by
tests/golden/make_golden_modee.py from oracle/_ref/kmerLSH_ref -M C --only followed by -M E --only).  Inputs are
regenerated here by the same seeded generator (mode-C matrix, kmer_set.hex, one FASTQ file per sample — one of them
gzip-compressed, one spanning two 65 536-read parts in `modee_parts`)."""
import hashlib
import json
import os
import subprocess

import pytest

from kmerlsh_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "kmerlsh_b200", "kmerLSH_b200")


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def prepare(work, c):
    synth.write_mode_c_inputs(work, c["n"], c["sa"], c["sb"], c["gen_seed"])
    _, files = synth.write_mode_e_inputs(work, c["n"], c["sa"], c["sb"], c["k"], c["gen_seed"] + 1000, c["reads"], c["big"])
    for f, want in c["inputs_md5"].items():
        if md5(os.path.join(work, f)) != want:
            pytest.skip("numpy generator stream differs from the one the golden run used (%s)" % f)
    return files


def args(c, mode, only=True):
    a = ["-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", mode, "-I", str(c["iters"]), "-N", str(c["minsim"]), "-K", str(c["k"]),
         "-S", str(c["S"]), "-P", str(c["P"]), "-V", str(c["V"]), "-T", "1", "--seed=%d" % c["klsh_seed"]]
    return a + (["--only"] if only else [])


def check_outputs(work, c):
    assert md5(os.path.join(work, "clustering_result.txt.clust")) == c["clust_md5"]
    for name, want in c["outputs"].items():
        path = os.path.join(work, name)
        assert os.path.exists(path), name
        assert os.path.getsize(path) == want["bytes"], name
        assert md5(path) == want["md5"], name


@pytest.mark.parametrize("tag", ["modee_small", "modee_parts"])
def test_mode_e_cli_matches_reference_binary(tag, tmp_path):
    c = json.load(open(os.path.join(G, "modee.json")))[tag]
    work = str(tmp_path)
    prepare(work, c)
    subprocess.run([EXE] + args(c, "C"), cwd=work, check=True, stdout=subprocess.DEVNULL)
    subprocess.run([EXE] + args(c, "E"), cwd=work, check=True, stdout=subprocess.DEVNULL)
    check_outputs(work, c)


def test_mode_c_without_only_runs_the_extraction_too(tmp_path):
    """`-M C` without --only is clustering followed by extraction in the reference (app/kmerLSH.cc:260-266)."""
    c = json.load(open(os.path.join(G, "modee.json")))["modee_small"]
    work = str(tmp_path)
    prepare(work, c)
    out = subprocess.run([EXE] + args(c, "C", only=False) + ["--verbose"], cwd=work, check=True, stdout=subprocess.PIPE, text=True).stdout
    check_outputs(work, c)
    assert "# of differential kmers in group A" in out and "writing to oa_A_0.fq" in out


def test_mode_e_cli_errors(tmp_path):
    work = str(tmp_path)
    synth.write_mode_c_inputs(work, 2000, 2, 2, 5)
    r = subprocess.run([EXE, "-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", "E", "--only"], cwd=work, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "klsh_load_cluster_file failed" in r.stderr   # no clustering result yet
    r = subprocess.run([EXE, "-a", "A.txt", "-b", "B.txt", "-M", "E", "-K", "40"], cwd=work, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "kmer_size" in r.stderr
