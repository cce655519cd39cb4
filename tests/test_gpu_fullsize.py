"""Full-size checks (BASELINE.json configs[1]: 50 M k-mers x 32 samples, phase 1 + I=100) through
properties that do not need the oracle to replay the run (it would take hours on one core):

* conservation: every k-mer that passes convertHTMat's keep filter (io/ioMatrix.cc:366-377) ends up
  in exactly one cluster, member counts add up, offsets are strictly increasing;
* the iteration chain is consistent (rows_out[k] == rows_in[k+1]) and the thresholds follow the
  reference's fp32 recurrence (function/cluster.cc:190-192, :330);
* singleton clusters carry their transformed row bit for bit (checked against the oracle's
  convertHTMat on a sample), and sampled centroids lie inside the per-dimension envelope of their
  members' rows (AB::SetConsensus is a convex combination, funcAB.cc:49-71);
* the run is deterministic: a second run from the same seed gives byte-identical clusters;
* the signing kernel at full launch geometry agrees with the oracle on every row of a 4 M-row slice.

KLSH_FULLSIZE_ROWS shrinks the shape for a quick look (default: the full 50 M).
"""
import hashlib
import os

import numpy as np
import pytest

from kmerlsh_b200 import Context, synth
from kmerlsh_b200.distributed import float32_threshold_schedule

pytestmark = pytest.mark.gpu

ROWS = int(os.environ.get("KLSH_FULLSIZE_ROWS", "50000000"))
ITERS, MIN_SIM = 100, 0.80


@pytest.fixture(scope="module")
def c2():
    from kmerlsh_b200.synth_gpu import synth_counts_gpu

    _, sa, sb, seed = synth.CONFIGS["C2"]
    counts, cov = synth_counts_gpu(ROWS, sa, sb, seed)
    kmap, cov32 = synth.parse_log_line(synth.format_log_line(ROWS, cov), sa + sb)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    return ROWS, sa + sb, counts, vk


def _run(ctx, counts, vk):
    ctx.set_seed(42)
    ctx.load_counts(counts, vk, 0)
    kept = ctx.row_count(False)[0]
    st = ctx.cluster(MIN_SIM, 1, 100000)
    st += ctx.cluster(MIN_SIM, ITERS, 1000000)
    values, offs, ids = ctx.get_rows()
    return kept, st, values, offs, ids


def _digest(values, offs, ids):
    h = hashlib.md5()
    for a in (values, offs, ids):
        h.update(np.ascontiguousarray(a).view(np.uint8))
    return h.hexdigest()


def test_full_size_mode_c_properties(c2, oracle):
    n, d, counts, vk = c2
    ctx = Context(0, seed=42)
    kept, st, values, offs, ids = _run(ctx, counts, vk)

    # conservation
    total = np.zeros(n, dtype=np.uint32)
    for j in range(d):
        total += counts[j]
    expect_ids = np.flatnonzero(total.astype(np.float64) > 0.1 * float(d)).astype(np.uint64)
    assert kept == len(expect_ids)
    sizes = np.diff(offs.astype(np.int64))
    assert offs[0] == 0 and sizes.min() >= 1
    assert int(offs[-1]) == kept == len(ids)
    assert np.array_equal(np.sort(ids), expect_ids)

    # iteration chain and thresholds
    phase2 = st[1:]
    assert st[0].rows_in == kept and st[0].rows_out == phase2[0].rows_in
    for a, b in zip(phase2, phase2[1:]):
        assert a.rows_out == b.rows_in and a.rows_out <= a.rows_in
    assert phase2[-1].rows_out == len(offs) - 1 == values.shape[0]
    for s, thr in zip(phase2, float32_threshold_schedule(MIN_SIM, ITERS)):
        assert np.float32(s.threshold) == thr
    assert np.isfinite(values).all()

    # singletons are untouched transformed rows
    rng = np.random.default_rng(7)
    single = np.flatnonzero(sizes == 1)
    assert len(single) > 0
    pick = np.sort(rng.choice(single, size=min(2000, len(single)), replace=False))
    cols = ids[offs[pick].astype(np.int64)].astype(np.int64)
    rows, rid = oracle.convert_counts(np.ascontiguousarray(counts[:, cols]), vk, 0)
    assert len(rid) == len(cols)
    assert np.array_equal(rows.view(np.uint32), values[pick].view(np.uint32))

    # centroids inside the envelope of their members
    multi = np.flatnonzero((sizes >= 2) & (sizes <= 500))
    for c in rng.choice(multi, size=min(300, len(multi)), replace=False):
        mem = ids[int(offs[c]):int(offs[c + 1])].astype(np.int64)
        rows, _ = oracle.convert_counts(np.ascontiguousarray(counts[:, mem]), vk, 0)
        lo, hi = rows.min(axis=0), rows.max(axis=0)
        tol = 1e-4 * np.maximum(1.0, np.maximum(np.abs(lo), np.abs(hi)))
        assert np.all(values[c] >= lo - tol) and np.all(values[c] <= hi + tol)

    # determinism
    first = _digest(values, offs, ids)
    _, st2, v2, o2, i2 = _run(ctx, counts, vk)
    assert _digest(v2, o2, i2) == first
    assert [s.rows_out for s in st2] == [s.rows_out for s in st]


def test_full_size_sign_slice(c2, oracle):
    """Every key of a 4 M-row slice, signed in one launch with the run's first table (H = 25)."""
    n, d, counts, vk = c2
    m = min(n, 4_000_000)
    rows, _ = oracle.convert_counts(np.ascontiguousarray(counts[:, :m]), vk, 0)
    ctx = Context(0, seed=42)
    h = int(np.floor(np.log2(n)))
    table = ctx.draw_table(h, d)
    got = ctx.sign(rows, table)
    want = oracle.sign(rows, table)
    assert np.array_equal(got.astype(np.uint64), want.astype(np.uint64))
