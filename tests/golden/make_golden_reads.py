"""Mint tests/golden/reads.npz — known answers for the read-extraction votes — from the REAL
reference: Kmer (kmer/Kmer.cc) and IOFQ::CheckRead (io/ioFastQ.cc:5-75) called through oracle/ref_harness.cc.
Kmer::set_k works once per process, so every k is minted by its own child process.
Run in the build container: `python tests/golden/make_golden_reads.py`.
"""
import os
import subprocess
import sys

os.environ["OMP_THREAD_LIMIT"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

KS = (23, 31, 32, 20, 5)
VOTES = (0.0, 0.1, 0.3, 0.5, 0.75)


def make_case(k, seed):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    odd = np.frombuffer(b"Nacgtn.", dtype=np.uint8)

    def rand_read(length, p_odd=0.02):
        a = alpha[rng.integers(0, 4, length)].copy()
        m = rng.random(length) < p_odd
        a[m] = odd[rng.integers(0, len(odd), int(m.sum()))]
        return a.tobytes()

    genome = rand_read(20000, 0.0)
    reads = []
    for i in range(1500):
        length = int(rng.integers(k + 5, 300))
        if i % 3 == 0:   # from the "genome", forward or reverse complement, with a few substitutions
            p = int(rng.integers(0, len(genome) - length))
            a = np.frombuffer(genome[p:p + length], dtype=np.uint8).copy()
            if i % 2:
                comp = {65: 84, 67: 71, 71: 67, 84: 65}
                a = np.array([comp[x] for x in a[::-1]], dtype=np.uint8)
            m = rng.random(length) < 0.03
            a[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            reads.append(a.tobytes())
        else:
            reads.append(rand_read(length))
    reads[5] = b"\0" + reads[5][1:]          # "abnormal read entry skipped"
    reads[7] = reads[7][: k + 9].ljust(k + 9, b"A")    # one short of the minimum length
    reads[8] = reads[8][: k + 10].ljust(k + 10, b"C")  # exactly the minimum
    reads[9] = b"T" * 120                     # poly-T: its twin is poly-A, the all-ones record when k = 32
    reads[10] = b"A" * 120
    reads[11] = b""
    seq = b"".join(reads)
    offs = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.uint64)
    return genome, seq, offs


def child(k, out_path):
    from oracle_lib import RefLib

    r = RefLib()
    genome, seq, offs = make_case(k, 1000 + k)
    # the differential k-mer set: canonical forms of every third k-mer of the genome, poly-T and a few random ones
    strings = [genome[j:j + k] for j in range(0, len(genome) - k + 1, 3)] + [b"T" * k, b"ACGT" * 8]
    kms, reps = [], []
    for s in strings:
        km, rep = r.kmer_rep(s[:k].ljust(k, b"A"), k)
        kms.append(km)
        reps.append(rep)
    kms, reps = np.array(kms, dtype=np.uint8), np.array(reps, dtype=np.uint8)
    out = {"k": np.int32(k), "seq": np.frombuffer(seq, dtype=np.uint8), "offs": offs, "strings": np.array([s[:k].ljust(k, b"A") for s in strings]),
           "km": kms, "rep": reps}
    for v in VOTES:
        out["rec_%g" % v] = r.check_reads(reps, k, seq, offs, v)
    np.savez_compressed(out_path, **out)
    print("k", k, "reads", len(offs) - 1, "set", len(reps), {v: int(out["rec_%g" % v].sum()) for v in VOTES})


def main():
    if len(sys.argv) == 3:
        child(int(sys.argv[1]), sys.argv[2])
        return
    merged = {}
    for k in KS:
        tmp = os.path.join(HERE, "_reads_k%d.npz" % k)
        subprocess.run([sys.executable, os.path.abspath(__file__), str(k), tmp], check=True, stdout=sys.stdout)
        g = np.load(tmp)
        for name in g.files:
            merged["k%d_%s" % (k, name)] = g[name]
        os.remove(tmp)
    np.savez_compressed(os.path.join(HERE, "reads.npz"), **merged)


if __name__ == "__main__":
    main()
