"""Mode-E statistics (AB::WRS function/funcAB.cc:73-109, the loop and the k-mer join of
app/kmerLSH.cc:541-585).

CPU: the oracle's restatement against tests/golden/ttest.npz (minted from the reference's own ALGLIB and AB::WRS by
tests/golden/make_golden_ttest.py) and, where oracle/_ref is present, against the reference live.
GPU: klsh_ttest / klsh_differential_ids / klsh_select_kmers against the oracle and the golden file.

Tolerance: the test statistic is restated operation by operation; the Student distribution's lower tail for
t < -2 is ALGLIB's Cephes incomplete beta in the reference and a continued fraction here and in the oracle, so tail
probabilities are compared to RTOL = 1e-9 relative (ATOL 1e-290: ALGLIB flushes tails below 1e-300 to 0), and
decisions must be identical for every cluster whose tails are not within RTOL of the threshold.
"""
import os

import numpy as np
import pytest

from helpers import synth_rows

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-9, 1e-290


def load():
    return np.load(os.path.join(G, "ttest.npz"))


def close(a, b, other=None):
    """a against the expected b.  `other` = the expected complementary tail: righttail is computed as 1 - lefttail
    (statistics.cpp:12614), so it inherits lefttail's ABSOLUTE error, RTOL * lefttail."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    tol = ATOL + RTOL * np.abs(b)
    if other is not None:
        tol = tol + RTOL * np.abs(np.asarray(other, dtype=np.float64))
    return bool(np.all(both_nan | (np.abs(a - b) <= tol)))


def one_row_per_case(g):
    """The KAT rows as clusters: every case becomes a row set of its own width class."""
    by_shape = {}
    for k in range(len(g["kat_rows"])):
        n1, n2 = (int(x) for x in g["kat_n"][k])
        by_shape.setdefault((n1, n2), []).append(k)
    return by_shape


# ---------------------------------------------------------------------------------------------- CPU
def test_oracle_ttest2_kats(oracle):
    g = load()
    for k in range(len(g["kat_rows"])):
        n1, n2 = (int(x) for x in g["kat_n"][k])
        row = g["kat_rows"][k].astype(np.float64)
        got = oracle.ttest2(row[:n1], row[n1:n1 + n2])
        want = g["kat_tails"][k]
        assert close(got[1], want[1]) and close(got[2], want[2], want[1]) and close(got[0], want[0], 2 * want[1]), (k, n1, n2, got, want)


def test_oracle_wrs_and_join_kats(oracle):
    g = load()
    for tag in "abcd":
        pthr, sthr = g["wrs_%s_params" % tag]
        group, left, right = oracle.wrs_rows(g["wrs_values"], g["wrs_offs"], 10, 10, float(pthr), int(sthr))
        assert np.array_equal(group, g["wrs_%s_group" % tag]), tag
        label = oracle.differential_ids(group, g["wrs_offs"], g["wrs_ids"], 30000)
        assert np.array_equal(label, g["wrs_%s_label" % tag]), tag
    group, _, _ = oracle.wrs_rows(g["dup_values"], g["dup_offs"], 4, 4, 0.01, 1)
    assert np.array_equal(group, g["dup_group"])
    assert np.array_equal(oracle.differential_ids(group, g["dup_offs"], g["dup_ids"], 20), g["dup_label"])


def test_oracle_select_kmers(oracle):
    rng = np.random.default_rng(3)
    rec = rng.integers(0, 256, (5000, 8), dtype=np.uint8)
    label = rng.choice(np.array([0, 0, 0, 1, 2], dtype=np.uint8), 5000)
    a, b = oracle.select_kmers(rec, label)
    assert np.array_equal(a, rec[label == 1]) and np.array_equal(b, rec[label == 2])


def test_oracle_statistic_is_the_references(oracle, reflib):
    """Live against ALGLIB (oracle/_ref): random halves, constant halves, both distribution branches."""
    rng = np.random.default_rng(11)
    for trial in range(3000):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        if trial % 7 == 0:
            n, m = int(rng.integers(100, 400)), int(rng.integers(100, 400))
        x = rng.normal(0, 1, n).astype(np.float32).astype(np.float64)
        y = (rng.normal(0, 1, m) + rng.choice([0, 0.1, 0.5, 1, 2, 5, 20])).astype(np.float32).astype(np.float64)
        if trial % 11 == 0:
            x[:] = x[0]
        if trial % 13 == 0:
            y[:] = y[0]
        if trial % 2:
            x, y = y, x
        got, want = oracle.ttest2(x, y), reflib.ttest2(x, y)
        assert close(got[1], want[1]) and close(got[2], want[2], want[1]), (trial, n, m)


# ---------------------------------------------------------------------------------------------- GPU
def _check_against_oracle(oracle, gpu, values, offs, ids, n1, n2, pthr, sthr, n_kmers, what):
    o_group, o_left, o_right = oracle.wrs_rows(values, offs, n1, n2, pthr, sthr)
    gpu.set_rows(values, offs, ids)
    group, left, right, st = gpu.ttest(n1, n2, pthr, sthr)
    assert close(left, o_left) and close(right, o_right, o_left), what
    assert st.margin == 0, what  # no decision inside the tolerance band, so decisions must be identical
    assert np.array_equal(group, o_group), what
    members = np.diff(offs.astype(np.int64))
    # `ids.size() > size_thresh` compares size_t with int: a negative threshold becomes huge and nothing is tested
    assert st.rows == len(values) and st.tested == (0 if sthr < 0 else int(np.sum(members > sthr)))
    assert (st.rows_a, st.rows_b) == (int(np.sum(group == 1)), int(np.sum(group == 2)))
    assert (st.ids_a, st.ids_b) == (int(members[group == 1].sum()), int(members[group == 2].sum()))
    label, st2 = gpu.differential_ids(n1, n2, pthr, sthr, n_kmers)
    assert np.array_equal(label, oracle.differential_ids(o_group, offs, ids, n_kmers)), what
    assert (st2.rows_a, st2.rows_b, st2.tested) == (st.rows_a, st.rows_b, st.tested)
    return group, label


@pytest.mark.gpu
def test_gpu_ttest_kats(gpu):
    """studentttest2 known answers of the reference's ALGLIB, one cluster per case."""
    g = load()
    for (n1, n2), ks in sorted(one_row_per_case(g).items()):
        rows = np.ascontiguousarray(g["kat_rows"][ks][:, : n1 + n2])
        gpu.set_rows(rows)
        group, left, right, st = gpu.ttest(n1, n2, 0.01, 0)
        want = g["kat_tails"][ks]
        assert close(left, want[:, 1]) and close(right, want[:, 2], want[:, 1]), (n1, n2)
        exp = np.where(want[:, 1] <= np.float32(0.01), 2, np.where(want[:, 2] <= np.float32(0.01), 1, 0))
        near = (np.abs(want[:, 1] - np.float32(0.01)) <= 1e-8) | (np.abs(want[:, 2] - np.float32(0.01)) <= 1e-8)
        assert np.array_equal(group[~near], exp[~near]), (n1, n2)
        assert st.tested == len(ks)


@pytest.mark.gpu
def test_gpu_wrs_golden(oracle, gpu):
    """AB::WRS of the reference over real clusters: row groups and the id sets, four threshold settings, plus ids
    shared by both groups (first set wins) and ids beyond kmap_size."""
    g = load()
    for tag in "abcd":
        pthr, sthr = g["wrs_%s_params" % tag]
        group, label = _check_against_oracle(oracle, gpu, g["wrs_values"], g["wrs_offs"], g["wrs_ids"], 10, 10, float(pthr),
                                             int(sthr), 30000, tag)
        assert np.array_equal(group, g["wrs_%s_group" % tag]) and np.array_equal(label, g["wrs_%s_label" % tag]), tag
    group, label = _check_against_oracle(oracle, gpu, g["dup_values"], g["dup_offs"], g["dup_ids"], 4, 4, 0.01, 1, 20, "dup")
    assert np.array_equal(group, g["dup_group"]) and np.array_equal(label, g["dup_label"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,sa,sb,iters,pthr,sthr", [(60000, 10, 10, 10, 0.01, 5), (40000, 16, 16, 6, 0.05, 0), (20000, 32, 32, 5, 0.01, 2),
                                                     (8000, 100, 100, 3, 0.01, 1), (30000, 3, 9, 8, 0.2, 3)])
def test_gpu_mode_e_after_clustering(oracle, n, sa, sb, iters, pthr, sthr):
    """The whole mode-E statistics step on the hot path's own output: cluster on the GPU (implicit member ids, the
    chains stay on the device), test every cluster, label every k-mer id, select the k-mer records — against the
    oracle run on the exported clusters."""
    from kmerlsh_b200 import Context

    counts, vk, values, ids = synth_rows(oracle, n, sa, sb, 300 + n % 97)
    with Context(0, seed=5) as ctx:
        ctx.load_counts(counts, vk, 0)
        ctx.cluster(0.8, iters, 100000)
        cv, co, ci = ctx.get_rows()
        o_group, o_left, o_right = oracle.wrs_rows(cv, co, sa, sb, pthr, sthr)
        group, left, right, st = ctx.ttest(sa, sb, pthr, sthr)
        assert close(left, o_left) and close(right, o_right, o_left)
        assert st.margin == 0 and np.array_equal(group, o_group)
        label, _ = ctx.differential_ids(sa, sb, pthr, sthr, n)
        o_label = oracle.differential_ids(o_group, co, ci, n)
        assert np.array_equal(label, o_label)
        assert label.max(initial=0) > 0  # the case does select something
        rec = np.random.default_rng(n).integers(0, 256, (n, 8), dtype=np.uint8)
        a, b = ctx.select_kmers(rec, label)
        oa, ob = oracle.select_kmers(rec, o_label)
        assert np.array_equal(a, oa) and np.array_equal(b, ob)
        # the clustering result read back from disk (what the reference's mode E does, ReadClusterAll): same answers
        import tempfile

        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "clustering_result.txt")
            ctx.save(path, True, 0)
            ctx.load_cluster_file(path, sa + sb)
            group2, _, _, _ = ctx.ttest(sa, sb, pthr, sthr)
            label2, _ = ctx.differential_ids(sa, sb, pthr, sthr, n)
            assert np.array_equal(group2, group) and np.array_equal(label2, label)


@pytest.mark.gpu
def test_gpu_select_kmers_sizes(oracle, gpu):
    """Two-way compaction: empty input, record widths other than 8 bytes, more than one 8 M-record pass."""
    a, b = gpu.select_kmers(np.zeros((0, 8), dtype=np.uint8), np.zeros(0, dtype=np.uint8))
    assert len(a) == 0 and len(b) == 0
    rng = np.random.default_rng(9)
    for n, rb in ((1, 8), (1025, 8), (70001, 16), (3000, 5), ((8 << 20) + 12345, 8)):
        rec = rng.integers(0, 256, (n, rb), dtype=np.uint8)
        label = rng.choice(np.array([0, 0, 1, 2], dtype=np.uint8), n)
        a, b = gpu.select_kmers(rec, label)
        assert np.array_equal(a, rec[label == 1]) and np.array_equal(b, rec[label == 2]), (n, rb)


@pytest.mark.gpu
def test_gpu_ttest_argument_errors(gpu):
    from kmerlsh_b200 import KlshError

    gpu.set_rows(np.ones((4, 8), dtype=np.float32))
    with pytest.raises(KlshError):
        gpu.ttest(5, 4, 0.01, 0)   # 9 samples, 8 values per row
    with pytest.raises(KlshError):
        gpu.ttest(-1, 4, 0.01, 0)
    group, left, right, st = gpu.ttest(0, 8, 0.01, 0)  # an empty group: ALGLIB answers 1.0 for every tail
    assert np.all(left == 1.0) and np.all(right == 1.0) and not group.any()
