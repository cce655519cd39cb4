"""Mint tests/golden/ttest.npz — known answers for the mode-E statistics — from the REAL
reference: alglib::studentttest2 and AB::WRS (function/funcAB.cc:73-109) called through oracle/ref_harness.cc in
oracle/_ref/libklsh_ref.so.  Run in the build container: `python tests/golden/make_golden_ttest.py`.
"""
import os
import sys

os.environ["OMP_THREAD_LIMIT"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from helpers import synth_rows  # noqa: E402
from oracle_lib import Oracle, RefLib  # noqa: E402


def kat_rows(rng):
    """(row, n1, n2) cases for studentttest2: both branches of the distribution (t >= -2 series, t < -2 incomplete
    beta), odd and even degrees of freedom, zero variance, one-sample halves, wide rows, NaN."""
    cases = []
    for n1, n2 in ((10, 10), (16, 16), (5, 8), (32, 32), (3, 4), (2, 1), (1, 1), (128, 128), (200, 56), (1, 7), (7, 1)):
        for shift in (0.0, 0.05, 0.3, 1.0, 2.5, 6.0, 40.0, -0.3, -1.0, -2.5, -6.0, -40.0):
            for _ in range(3):
                row = rng.normal(0.0, 1.0, n1 + n2)
                row[n1:] += shift
                cases.append((row.astype(np.float32), n1, n2))
    for n1, n2 in ((10, 10), (4, 9)):
        row = np.full(n1 + n2, 1.5, dtype=np.float32)
        cases.append((row.copy(), n1, n2))            # s == 0, equal means
        row[n1:] = 2.5
        cases.append((row.copy(), n1, n2))            # s == 0, x < y
        row[n1:] = -2.5
        cases.append((row.copy(), n1, n2))            # s == 0, x > y
        row = rng.normal(0, 1, n1 + n2).astype(np.float32)
        row[:n1] = row[0]
        cases.append((row.copy(), n1, n2))            # constant first half only
        row = rng.normal(0, 1, n1 + n2).astype(np.float32)
        row[3] = np.nan
        cases.append((row.copy(), n1, n2))            # NaN: every comparison is false
        row = (rng.normal(0, 1, n1 + n2) * 1e-30).astype(np.float32)
        cases.append((row.copy(), n1, n2))            # tiny magnitudes
        row = (rng.normal(0, 1, n1 + n2) * 1e30).astype(np.float32)
        cases.append((row.copy(), n1, n2))            # huge magnitudes (squares stay finite in double)
    return cases


def main():
    r = RefLib()
    o = Oracle()
    rng = np.random.default_rng(20261019)
    out = {}

    cases = kat_rows(rng)
    width = max(len(c[0]) for c in cases)
    rows = np.zeros((len(cases), width), dtype=np.float32)
    n12 = np.zeros((len(cases), 2), dtype=np.int32)
    tails = np.zeros((len(cases), 3), dtype=np.float64)
    for k, (row, n1, n2) in enumerate(cases):
        rows[k, : len(row)] = row
        n12[k] = (n1, n2)
        tails[k] = r.ttest2(row[:n1].astype(np.float64), row[n1:].astype(np.float64))
    out["kat_rows"], out["kat_n"], out["kat_tails"] = rows, n12, tails

    # AB::WRS over real clusters: C1-law rows (10 + 10 samples) clustered by the oracle, then the reference's loop
    _, _, values, _ = synth_rows(o, 30000, 10, 10, 20261018)
    rs = o.rows(values)
    rs.cluster(0.80, 12, 100000, o.planes(42))
    cv, co, ci = rs.export()
    n_kmers = 30000
    for tag, (pthr, sthr) in {"a": (0.01, 5), "b": (0.2, 0), "c": (0.5, -1), "d": (1e-6, 50)}.items():
        group, label = r.wrs(cv, co, ci, 10, 10, pthr, sthr, n_kmers)
        out["wrs_%s_params" % tag] = np.array([pthr, sthr], dtype=np.float64)
        out["wrs_%s_group" % tag] = group
        out["wrs_%s_label" % tag] = label
    out["wrs_values"], out["wrs_offs"], out["wrs_ids"] = cv, co, ci

    # ids shared by clusters of both groups (cannot happen in a clustering result, but the join defines it:
    # the first set wins, app/kmerLSH.cc:571-576) and ids beyond kmap_size (never asked about)
    dv = np.zeros((6, 8), dtype=np.float32)
    dv[:, :4] = rng.normal(0, 0.05, (6, 4))
    dv[:, 4:] = rng.normal(0, 0.05, (6, 4))
    dv[0, 4:] += 5.0   # x < y: left tail small -> second set
    dv[1, :4] += 5.0   # x > y: right tail small -> first set
    dv[2, 4:] += 5.0
    dv[3, :4] += 5.0
    doffs = np.array([0, 3, 6, 8, 11, 13, 14], dtype=np.uint64)
    dids = np.array([1, 2, 3, 3, 4, 5, 5, 99, 6, 1, 98, 7, 8, 9], dtype=np.uint64)
    group, label = r.wrs(dv, doffs, dids, 4, 4, 0.01, 1, 20)
    out["dup_values"], out["dup_offs"], out["dup_ids"], out["dup_group"], out["dup_label"] = dv, doffs, dids, group, label

    np.savez_compressed(os.path.join(HERE, "ttest.npz"), **out)
    print("kat cases", len(cases), "wrs clusters", len(co) - 1, "groups a:", np.bincount(out["wrs_a_group"], minlength=3),
          "labels a:", np.bincount(out["wrs_a_label"], minlength=3), "dup group", group, "dup label", label)


if __name__ == "__main__":
    main()
