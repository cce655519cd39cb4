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


def kmer_records(strings, k: int):
    """Canonical 8-byte records of k-mers given as an [n][k] uint8 array of 'ACGT' characters, in the reference's
    Kmer layout (kmer/Kmer.cc:131-150: base i in byte i/4 at bit 2*(i%4), A=0 C=1 G=2 T=3) and canonical form
    (`rep = (km < twin) ? km : twin`, operator< = memcmp over the 8 bytes, :98-100, :160-185)."""
    strings = np.asarray(strings, dtype=np.uint8).reshape(-1, k)
    code = np.zeros(256, dtype=np.uint64)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    c = code[strings]
    shifts = (2 * np.arange(k, dtype=np.uint64))[None, :]
    fw = np.bitwise_or.reduce(c << shifts, axis=1)
    rc = np.bitwise_or.reduce((np.uint64(3) - c[:, ::-1]) << shifts, axis=1)
    fw_b = fw.astype("<u8").view(np.uint8).reshape(-1, 8)
    rc_b = rc.astype("<u8").view(np.uint8).reshape(-1, 8)
    less = fw.astype("<u8").byteswap() < rc.astype("<u8").byteswap()  # memcmp order = big-endian integer order
    return np.where(less[:, None], fw_b, rc_b).astype(np.uint8)


def write_mode_e_inputs(out_dir: str, n: int, sa: int, sb: int, k: int, seed: int, reads_per_file: int, big_file_reads: int = 0,
                        gz_file: bool = True):
    """Mode-E inputs beside the mode-C ones (call after write_mode_c_inputs with the same n, sa, sb): kmer_set.hex with
    n distinct random canonical k-mers (record i = k-mer id i) and one FASTQ file per line of A.txt / B.txt
    (A_0.fq ...).  Reads are runs of k-mers of the file (so that some carry differential k-mers), forward or reverse
    complemented, with a few substitutions, characters outside ACGT, header comments and '@'/'+' inside quality strings;
    the second file of group A is gzip-compressed (same name: zlib detects it); `big_file_reads` > 65 536 makes the last
    file of group B span more than one part of the reader."""
    import gzip

    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    # distinct canonical k-mers
    strings = alpha[rng.integers(0, 4, (int(n * 1.05) + 16, k))]
    recs = kmer_records(strings, k)
    _, first = np.unique(recs.view("<u8").reshape(-1), return_index=True)
    keep = np.sort(first)[:n]
    assert len(keep) == n
    strings, recs = strings[keep], recs[keep]
    recs.tofile(os.path.join(out_dir, "kmer_set.hex"))
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    odd = np.frombuffer(b"Nacgtn", dtype=np.uint8)
    qual_alpha = np.frombuffer(b"IIIIHG#5@+!~", dtype=np.uint8)

    def make_reads(count):
        out = []
        for i in range(count):
            if i % 4 == 3:
                length = int(rng.integers(k + 5, 200))
                seq = alpha[rng.integers(0, 4, length)].copy()
            else:
                picks = rng.integers(0, n, int(rng.integers(2, 7)))
                seq = strings[picks].reshape(-1).copy()
                if i % 2:
                    seq = comp[seq[::-1]]
                m = rng.random(len(seq)) < 0.01
                seq[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            m = rng.random(len(seq)) < 0.004
            seq[m] = odd[rng.integers(0, len(odd), int(m.sum()))]
            qual = qual_alpha[rng.integers(0, len(qual_alpha), len(seq))]
            head = b"@r%d" % i + (b" len=%d extra" % len(seq) if i % 5 == 0 else b"")
            out.append(head + b"\n" + seq.tobytes() + b"\n+\n" + qual.tobytes() + b"\n")
        return b"".join(out)

    files = ["A_%d.fq" % i for i in range(sa)] + ["B_%d.fq" % i for i in range(sb)]
    for idx, name in enumerate(files):
        count = big_file_reads if (big_file_reads and idx == len(files) - 1) else reads_per_file
        data = make_reads(count)
        path = os.path.join(out_dir, name)
        if gz_file and idx == 1:
            with gzip.GzipFile(path, "wb", mtime=0) as f:
                f.write(data)
        else:
            with open(path, "wb") as f:
                f.write(data)
    return recs, files


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
