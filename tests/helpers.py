"""Shared input builders for the parity tests."""
import numpy as np

from kmerlsh_b200 import synth


def synth_rows(oracle, n, sa, sb, seed):
    """Transformed rows of the synthetic generator via the oracle."""
    counts, cov = synth.synth_counts(n, sa, sb, seed)
    kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    values, ids = oracle.convert_counts(counts, vk, 0)
    return counts, vk, values, ids


def assert_rows_equal(a, b, what=""):
    av, ao, ai = a
    bv, bo, bi = b
    assert len(ao) == len(bo), "%s: row count %d != %d" % (what, len(ao) - 1, len(bo) - 1)
    assert np.array_equal(ao, bo), what + ": id offsets differ"
    assert np.array_equal(ai, bi), what + ": id lists differ"
    assert av.tobytes() == bv.tobytes(), what + ": centroid bits differ"
