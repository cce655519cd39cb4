"""Read-extraction votes (IOFQ::CheckRead io/ioFastQ.cc:5-75 over Kmer, kmer/Kmer.cc).

CPU: the oracle's restatement against tests/golden/reads.npz (minted from the reference's own Kmer and CheckRead by
tests/golden/make_golden_reads.py, one process per k) and, where oracle/_ref is present, against the reference live.
GPU: klsh_kmer_set_load / klsh_check_reads against the oracle and the golden file.  Integer and byte work plus one
float division: everything is compared bit for bit.
"""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KS = (23, 31, 32, 20, 5)
VOTES = (0.0, 0.1, 0.3, 0.5, 0.75)


def load():
    return np.load(os.path.join(G, "reads.npz"))


def random_reads(rng, n, k, lo=None, hi=300, p_odd=0.02):
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    odd = np.frombuffer(b"Nacgtn.", dtype=np.uint8)
    reads = []
    for _ in range(n):
        length = int(rng.integers(lo if lo is not None else max(1, k - 3), hi))
        a = alpha[rng.integers(0, 4, length)].copy()
        m = rng.random(length) < p_odd
        a[m] = odd[rng.integers(0, len(odd), int(m.sum()))]
        reads.append(a.tobytes())
    return reads


def pack(reads):
    return b"".join(reads), np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.uint64)


def kmer_set_of(oracle, reads, k, every=2):
    recs = [oracle.kmer_rep(x[j:j + k], k)[1] for x in reads for j in range(0, max(0, len(x) - k + 1), every)]
    return np.array(recs, dtype=np.uint8).reshape(-1, 8)


# ---------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("k", KS)
def test_oracle_kmer_bytes_and_votes_golden(oracle, k):
    g = load()
    pre = "k%d_" % k
    for s, km, rep in zip(g[pre + "strings"][::7], g[pre + "km"][::7], g[pre + "rep"][::7]):
        got = oracle.kmer_rep(bytes(s), k)
        assert np.array_equal(got[0], km) and np.array_equal(got[1], rep), s
    seq, offs = g[pre + "seq"].tobytes(), g[pre + "offs"]
    for v in VOTES:
        rec, votes = oracle.check_reads(g[pre + "rep"], k, seq, offs, v)
        assert np.array_equal(rec, g[pre + "rec_%g" % v]), (k, v)


def test_oracle_votes_against_reference_live(oracle, reflib):
    """k = 23 (the reference's -K default); Kmer::set_k can be called once per process, hence one k here."""
    try:
        reflib.kmer_rep(b"A" * 23, 23)
    except RuntimeError:
        pytest.skip("Kmer::k already set to another value in this process")
    rng = np.random.default_rng(77)
    reads = random_reads(rng, 800, 23)
    reads[3] = b"\0" + reads[3][1:]
    seq, offs = pack(reads)
    kmers = kmer_set_of(oracle, reads[::3], 23)
    for v in (0.0, 0.2, 0.45, 0.5):
        assert np.array_equal(oracle.check_reads(kmers, 23, seq, offs, v)[0], reflib.check_reads(kmers, 23, seq, offs, v)), v


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("k", KS)
def test_gpu_votes_golden(oracle, gpu, k):
    g = load()
    pre = "k%d_" % k
    seq, offs = g[pre + "seq"].tobytes(), g[pre + "offs"]
    gpu.kmer_set_load(g[pre + "rep"])
    for v in VOTES:
        rec, votes = gpu.check_reads(k, seq, offs, v)
        assert np.array_equal(rec, g[pre + "rec_%g" % v]), (k, v)
        assert np.array_equal(votes, oracle.check_reads(g[pre + "rep"], k, seq, offs, v)[1]), (k, v)


@pytest.mark.gpu
@pytest.mark.parametrize("k,n,every", [(23, 20000, 2), (32, 5000, 1), (31, 5000, 5), (1, 300, 1), (4, 2000, 3), (16, 3000, 2), (17, 3000, 2)])
def test_gpu_votes_random(oracle, gpu, k, n, every):
    """Random reads (2 % characters outside ACGT, a NUL-leading read, reads around the k+10 minimum, empty reads),
    the set drawn from a third of them: flags and vote counts against the oracle for several thresholds."""
    rng = np.random.default_rng(1000 * k + n)
    reads = random_reads(rng, n, k)
    reads[1] = b"\0" + reads[1][1:]
    reads[2] = b""
    reads[3] = (reads[3] * 40)[: k + 9]
    reads[4] = (reads[4] * 40)[: k + 10]
    reads[6] = b"T" * 100
    seq, offs = pack(reads)
    kmers = kmer_set_of(oracle, reads[::3] + [b"T" * 64], k, every)
    gpu.kmer_set_load(kmers)
    for v in (0.0, 0.05, 0.3, 0.5, 0.999, 1.0):
        rec, votes = gpu.check_reads(k, seq, offs, v)
        o_rec, o_votes = oracle.check_reads(kmers, k, seq, offs, v)
        assert np.array_equal(votes, o_votes), (k, v)
        assert np.array_equal(rec, o_rec), (k, v)
    # a window of the same buffer (offsets that do not start at 0) and an empty set
    rec, votes = gpu.check_reads(k, seq, offs[100:200], 0.1)
    o_rec, o_votes = oracle.check_reads(kmers, k, seq, offs[100:200], 0.1)
    assert np.array_equal(rec, o_rec) and np.array_equal(votes, o_votes)
    gpu.kmer_set_load(np.zeros((0, 8), dtype=np.uint8))
    rec, votes = gpu.check_reads(k, seq, offs, 0.0)
    assert not rec.any() and not votes.any()


@pytest.mark.gpu
def test_gpu_mode_e_chain_statistics_to_reads(oracle):
    """The statistics step feeding the read votes: clusters -> t-test -> labels -> k-mer records -> set -> reads."""
    from helpers import synth_rows
    from kmerlsh_b200 import Context

    n, sa, sb, k = 30000, 10, 10, 23
    counts, vk, values, ids = synth_rows(oracle, n, sa, sb, 321)
    rng = np.random.default_rng(4)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    kmer_strings = [alpha[rng.integers(0, 4, k)].tobytes() for _ in range(n)]
    hexfile = np.array([oracle.kmer_rep(s, k)[1] for s in kmer_strings], dtype=np.uint8)  # kmer_set.hex: record i = k-mer id i
    with Context(0, seed=5) as ctx:
        ctx.load_counts(counts, vk, 0)
        ctx.cluster(0.8, 8, 100000)
        label, st = ctx.differential_ids(sa, sb, 0.05, 3, n)
        a, b = ctx.select_kmers(hexfile, label)
        assert len(a) + len(b) > 0
        picks = np.flatnonzero(label)[: 400]
        reads = [b"".join(kmer_strings[i] for i in rng.choice(picks, 5)) for _ in range(300)] + random_reads(rng, 300, k, lo=60)
        seq, offs = pack(reads)
        for recs in (a, b):
            ctx.kmer_set_load(recs)
            rec, votes = ctx.check_reads(k, seq, offs, 0.02)
            o_rec, o_votes = oracle.check_reads(recs, k, seq, offs, 0.02)
            assert np.array_equal(rec, o_rec) and np.array_equal(votes, o_votes)
        assert o_rec.any() or len(b) == 0


@pytest.mark.gpu
def test_gpu_check_reads_argument_errors(oracle):
    from kmerlsh_b200 import Context, KlshError

    with Context(0) as ctx:
        with pytest.raises(KlshError):
            ctx.check_reads(23, b"ACGT" * 20, [0, 80], 0.1)          # no set loaded
        with pytest.raises(KlshError):
            ctx.kmer_set_load(np.zeros((4, 16), dtype=np.uint8))    # MAX_K = 64 records are not supported
        ctx.kmer_set_load(np.zeros((1, 8), dtype=np.uint8))
        with pytest.raises(KlshError):
            ctx.check_reads(33, b"ACGT" * 20, [0, 80], 0.1)
        with pytest.raises(KlshError):
            ctx.check_reads(23, b"ACGT" * 20, [40, 0], 0.1)
        rec, votes = ctx.check_reads(23, b"A" * 80, [0, 80], 0.5)    # poly-A is the all-zero record
        assert rec[0] == 1 and votes[0] == 58
