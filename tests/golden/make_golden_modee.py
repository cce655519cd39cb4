"""Mint the mode-E goldens (md5 of the extracted-read files) from the REAL reference binary (oracle/_ref/kmerLSH_ref):
`-M C --only` on synthetic inputs, then `-M E --only` on kmer_set.hex and FASTQ files of kmerlsh_b200/synth.py's
generator.  Run in the build container: `python tests/golden/make_golden_modee.py`; writes tests/golden/modee.json.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import REF_BIN  # noqa: E402

from kmerlsh_b200 import synth  # noqa: E402

CASES = {
    # n, sa, sb, generator seed, mode-C iterations, N, KLSH_SEED, k, -S, -P, -V, reads per file, reads of the big file
    "modee_small": dict(n=40000, sa=4, sb=4, gen_seed=7, iters=6, minsim=0.85, klsh_seed=7, k=23, S=1, P=0.2, V=0.02, reads=1500, big=0),
    "modee_parts": dict(n=30000, sa=3, sb=3, gen_seed=11, iters=5, minsim=0.85, klsh_seed=3, k=23, S=1, P=0.2, V=0.012, reads=500, big=70000),
}


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def prepare(work, c):
    synth.write_mode_c_inputs(work, c["n"], c["sa"], c["sb"], c["gen_seed"])
    _, files = synth.write_mode_e_inputs(work, c["n"], c["sa"], c["sb"], c["k"], c["gen_seed"] + 1000, c["reads"], c["big"])
    return files


def mode_args(c, mode):
    return ["-a", "A.txt", "-b", "B.txt", "-o", "oa", "-p", "ob", "-M", mode, "--only", "-I", str(c["iters"]), "-N", str(c["minsim"]),
            "-K", str(c["k"]), "-S", str(c["S"]), "-P", str(c["P"]), "-V", str(c["V"]), "-T", "1"]


def main():
    out = {}
    for tag, c in CASES.items():
        work = tempfile.mkdtemp(prefix="klsh_modee_")
        files = prepare(work, c)
        env = dict(os.environ, KLSH_SEED=str(c["klsh_seed"]), OMP_THREAD_LIMIT="1")
        subprocess.run([REF_BIN] + mode_args(c, "C"), cwd=work, env=env, check=True, stdout=subprocess.DEVNULL)
        subprocess.run([REF_BIN] + mode_args(c, "E"), cwd=work, env=env, check=True, stdout=subprocess.DEVNULL)
        rec = dict(c)
        rec["inputs_md5"] = {f: md5(os.path.join(work, f)) for f in ["kmer_count.bin", "kmer_set.hex"] + [x for x in files if x != "A_1.fq"]}
        rec["clust_md5"] = md5(os.path.join(work, "clustering_result.txt.clust"))
        rec["outputs"] = {}
        for f in files:
            o = ("oa_" if f.startswith("A_") else "ob_") + f
            p = os.path.join(work, o)
            rec["outputs"][o] = {"md5": md5(p), "bytes": os.path.getsize(p), "reads": open(p, "rb").read().count(b"\n+\n")}
        out[tag] = rec
        print(tag, {k: (v["reads"], v["bytes"]) for k, v in rec["outputs"].items()})
    with open(os.path.join(HERE, "modee.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
