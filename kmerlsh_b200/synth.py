"""Synthetic k-mer abundance inputs in the reference's on-disk layout.

The reference ships no sample data; this is the generator the survey
verified: latent genomes with log-normal abundance, Poisson counts,
written as the sample-major uint16 matrix that `ReadHT` consumes (reference io/ioHT.cc:59-81)
plus the companion `kmer_count.log` line (reference io/ioHT.cc:171,185).
"""
from __future__ import annotations

import os

import numpy as np

# (N, samples in group A, samples in group B, generator seed) of BASELINE.json's configs
CONFIGS = {
    "C1": (1_000_000, 10, 10, 20261018),
    "C2": (50_000_000, 16, 16, 2),
    "C3": (500_000_000, 32, 32, 3),
    "C4": (1_000_000_000, 32, 32, 4),
    "C5": (200_000_000, 128, 128, 5),
}


def synth_counts(n: int, sa: int, sb: int, seed: int):
    """Return (counts[S][n] uint16 sample-major, coverage[S] float64).

    Draw order is fixed: base, samp, diff, genome id, then Poisson per sample ascending.
    """
    s = sa + sb
    rng = np.random.default_rng(seed)
    g_num = max(8, n // 2000)
    base = rng.lognormal(2.0, 1.0, size=(g_num, 1))
    samp = rng.lognormal(0.0, 0.6, size=(g_num, s))
    diff = rng.random(g_num) < 0.2
    fold = np.ones((g_num, s))
    fold[diff, sa:] *= 4.0
    lam = base * samp * fold
    g = rng.integers(0, g_num, size=n)
    counts = np.empty((s, n), dtype=np.uint16)
    cov = np.empty(s, dtype=np.float64)
    for j in range(s):
        c = np.minimum(rng.poisson(lam[g, j]), 65535).astype(np.uint16)
        cov[j] = np.log(np.maximum(c, 1).astype(np.float64)).sum()
        counts[j] = c
    return counts, cov


def format_log_line(n: int, cov) -> str:
    """`kmap_size\\tcov_0\\t...` with %f, as the reference writes it."""
    return str(n) + "".join("\t%f" % float(c) for c in cov) + "\n"


def write_mode_c_inputs(out_dir: str, n: int, sa: int, sb: int, seed: int):
    """Write kmer_count.bin / kmer_count.log / A.txt / B.txt / tmp/ under out_dir."""
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(os.path.join(out_dir, "tmp"), exist_ok=True)
    counts, cov = synth_counts(n, sa, sb, seed)
    counts.tofile(os.path.join(out_dir, "kmer_count.bin"))
    with open(os.path.join(out_dir, "kmer_count.log"), "w") as f:
        f.write(format_log_line(n, cov))
    for name, k in (("A.txt", sa), ("B.txt", sb)):
        with open(os.path.join(out_dir, name), "w") as f:
            for i in range(k):
                f.write("%s_%d.fq %s_%d\n" % (name[0], i, name[0], i))
    return counts, cov


def parse_log_line(line: str, d: int):
    """kmap_size and float32 coverage exactly as `ss >> float` reads them
    (reference app/kmerLSH.cc:473-481)."""
    tok = line.split()
    kmap_size = int(tok[0])
    cov = np.array([np.float32(t) for t in tok[1 : 1 + d]], dtype=np.float32)
    return kmap_size, cov


def v_kmers_from_cov(cov32, kmap_size: int):
    """float(cov_j) / kmap_size in fp32 (reference app/kmerLSH.cc:480)."""
    return (np.asarray(cov32, dtype=np.float32) / np.float32(kmap_size)).astype(np.float32)
