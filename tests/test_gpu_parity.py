"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Bit-exact for keys, assignments, id order AND centroid floats (stricter than the 1e-5 the
north star allows)."""
import numpy as np
import pytest

from helpers import assert_rows_equal, synth_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,H", [(1, 0), (1, 3), (7, 5), (20, 19), (32, 25), (64, 29), (256, 27), (12, 32),
                                 (65, 9), (100, 20), (130, 31), (200, 24)])
def test_sign_keys(gpu, oracle, D, H):
    rng = np.random.default_rng(100 + D + H)
    n = 5000 if D < 100 else 1500
    rows = rng.standard_normal((n, D)).astype(np.float32)
    rows[::17] = 0.0                      # zero rows: every sum is +0 -> bit 1
    rows[1::29] *= np.float32(1e-30)      # denormal products
    table = oracle.planes(5).table(H, D) if H else np.zeros((0, D), np.float32)
    want = oracle.sign(rows, table).astype(np.uint64)
    got = gpu.sign(rows, table)
    assert np.array_equal(got, want)


def test_sign_nan_rows(gpu, oracle):
    rows = np.zeros((64, 8), np.float32)
    rows[3, 2] = np.nan                   # NaN sum -> bit 0 (sum >= 0 is false)
    rows[5, 0] = np.inf
    table = oracle.planes(9).table(6, 8)
    assert np.array_equal(gpu.sign(rows, table), oracle.sign(rows, table).astype(np.uint64))


@pytest.mark.parametrize("n,sa,sb", [(4096, 2, 2), (50000, 4, 4), (30011, 10, 10)])
def test_convert_counts(gpu, oracle, n, sa, sb):
    counts, vk, values, ids = synth_rows(oracle, n, sa, sb, 11)
    counts = counts.copy()
    counts[:, 5] = 0                      # dropped: total 0
    counts[:, 6] = 0
    counts[0, 6] = 65535                  # kept; LUT's last entry
    d = sa + sb
    counts[:, 7] = 0
    counts[0, 7] = int(0.1 * d)           # total == floor(0.1*D): dropped unless > 0.1*D
    counts[:, 8] = 0
    counts[0, 8] = int(0.1 * d) + 1       # kept
    values, ids = oracle.convert_counts(counts, vk, 1000)
    gpu.load_counts(counts, vk, 1000)
    gv, go, gi = gpu.get_rows()
    assert np.array_equal(gi, ids)
    assert gv.tobytes() == values.tobytes()
    assert np.array_equal(go, np.arange(len(ids) + 1, dtype=np.uint64))


def _bucket_cases():
    rng = np.random.default_rng(3)
    base = rng.standard_normal((1, 16)).astype(np.float32)
    cases = {}
    cases["all_identical"] = np.repeat(base, 40, axis=0)
    cases["two"] = np.concatenate([base, base * np.float32(1.0000001)])
    cases["none_merge"] = rng.standard_normal((37, 16)).astype(np.float32)
    x = base + np.float32(0.25) * rng.standard_normal((300, 16)).astype(np.float32)
    cases["noisy_cluster_300"] = x
    y = np.concatenate([base + np.float32(0.2) * rng.standard_normal((20, 16)).astype(np.float32),
                        -base + np.float32(0.2) * rng.standard_normal((20, 16)).astype(np.float32)])
    cases["two_groups_interleaved"] = y[rng.permutation(40)]
    z = rng.standard_normal((9, 16)).astype(np.float32)
    z[4] = 0.0                            # zero vector: cosine is NaN, never merges
    z[7] = z[1]
    cases["zero_row"] = z
    cases["chain_2000"] = (base + np.float32(0.3) * rng.standard_normal((2000, 16)).astype(np.float32))
    return cases


@pytest.mark.parametrize("name", sorted(_bucket_cases()))
@pytest.mark.parametrize("thr", [0.95, 0.8, 0.5])
def test_p_cluster(gpu, oracle, name, thr):
    values = _bucket_cases()[name]
    rows = oracle.rows(values)
    rows.p_cluster(thr)
    gpu.set_rows(values)
    gpu.p_cluster(thr)
    assert_rows_equal(gpu.get_rows(), rows.export(), name)


def test_p_cluster_prior_members(gpu, oracle):
    """Rows that already carry id lists (phase-2 input): count-weighted consensus and id order."""
    rng = np.random.default_rng(8)
    values = (rng.standard_normal((1, 20)) + 0.2 * rng.standard_normal((60, 20))).astype(np.float32)
    sizes = rng.integers(1, 9, size=60)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ids = rng.permutation(int(offs[-1])).astype(np.uint64) + 1000
    rows = oracle.rows(values, offs, ids)
    rows.p_cluster(0.9)
    gpu.set_rows(values, offs, ids)
    gpu.p_cluster(0.9)
    assert_rows_equal(gpu.get_rows(), rows.export())


@pytest.mark.parametrize("n,sa,sb,iters,minsim,thr,seed", [
    (1, 2, 2, 3, 0.8, 1000, 1),            # N=1 -> H=0, one bucket
    (2, 2, 2, 2, 0.8, 1000, 1),
    (3000, 3, 3, 4, 0.85, 100000, 2),
    (40000, 4, 4, 8, 0.85, 100000, 7),
    (40000, 4, 4, 3, 0.85, 50, 7),         # nested buckets (fresh tables consumed in bucket order)
    (30000, 6, 6, 8, 0.80, 100000, 3),
    (120000, 10, 10, 5, 0.80, 1000, 4),    # nested + large buckets
    (60000, 16, 16, 6, 0.90, 100000, 5),
    (20000, 32, 32, 4, 0.80, 100000, 6),
])
def test_cluster(gpu, oracle, n, sa, sb, iters, minsim, thr, seed):
    _, _, values, ids = synth_rows(oracle, n, sa, sb, 40 + seed)
    offs = np.arange(len(ids) + 1, dtype=np.uint64)
    rows = oracle.rows(values, offs, ids)
    planes = oracle.planes(seed)
    ost = rows.cluster(minsim, iters, thr, planes)
    gpu.set_seed(seed)
    gpu.set_rows(values, offs, ids)
    gst = gpu.cluster(minsim, iters, thr)
    for k in range(iters):
        if ost[k].rows_in == 0:
            break
        assert (gst[k].rows_in, gst[k].rows_out, gst[k].H) == (ost[k].rows_in, ost[k].rows_out, ost[k].H), k
        assert gst[k].nested_calls == ost[k].nested_calls
        assert np.float32(gst[k].threshold) == np.float32(ost[k].threshold)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_cluster_wide_rows(gpu, oracle):
    """D=256 (config 5's shape): multi-chunk rows in every kernel."""
    _, _, values, ids = synth_rows(oracle, 6000, 128, 128, 77)
    rows = oracle.rows(values)
    rows.cluster(0.8, 3, 100000, oracle.planes(5))
    gpu.set_seed(5)
    gpu.set_rows(values)
    gpu.cluster(0.8, 3, 100000)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_cluster_rows_beyond_the_windowed_merge(gpu, oracle):
    """D = 300: the windowed merge's shared-memory window no longer fits (about 768 bytes per float of row
    width), so the library falls back to the block-per-bucket merge kernel on its own — same results."""
    rng = np.random.default_rng(31)
    base = rng.standard_normal((40, 300)).astype(np.float32)
    values = (base[rng.integers(0, 40, size=2500)] + np.float32(0.25) * rng.standard_normal((2500, 300))).astype(np.float32)
    rows = oracle.rows(values)
    rows.cluster(0.8, 3, 100000, oracle.planes(8))
    gpu.set_seed(8)
    gpu.set_rows(values)
    gpu.cluster(0.8, 3, 100000)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_nested_cluster(gpu, oracle):
    _, _, values, ids = synth_rows(oracle, 20000, 4, 4, 21)
    rows = oracle.rows(values)
    rows.nested_cluster(0.9, oracle.planes(13))
    gpu.set_seed(13)
    gpu.set_rows(values)
    gpu.nested_cluster(0.9)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_two_phase_from_counts(gpu, oracle):
    """load_counts -> Cluster(I=1) -> Cluster(I=6) on resident rows == oracle's two calls."""
    counts, vk, values, ids = synth_rows(oracle, 50000, 5, 5, 31)
    planes = oracle.planes(99)
    rows = oracle.rows(values, np.arange(len(ids) + 1, dtype=np.uint64), ids)
    rows.cluster(0.8, 1, 100, planes)
    rows.cluster(0.8, 6, 1000, planes)
    gpu.set_seed(99)
    gpu.load_counts(counts, vk, 0)
    gpu.cluster(0.8, 1, 100)
    gpu.cluster(0.8, 6, 1000)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_snapshot_restore(gpu, oracle):
    _, _, values, ids = synth_rows(oracle, 20000, 4, 4, 5)
    gpu.set_rows(values)
    gpu.snapshot()
    gpu.set_seed(1)
    gpu.cluster(0.8, 3, 100000)
    a = gpu.get_rows()
    gpu.restore()
    gpu.set_seed(1)
    gpu.cluster(0.8, 3, 100000)
    assert_rows_equal(gpu.get_rows(), a)


def test_plane_callback(gpu, oracle):
    """Caller-supplied tables (what a reference-side binding does with its own generator)."""
    _, _, values, ids = synth_rows(oracle, 15000, 4, 4, 6)
    src = oracle.planes(321)
    rows = oracle.rows(values)
    rows.cluster(0.85, 4, 100000, oracle.planes(321))
    gpu.set_plane_source(lambda H, D: src.table(H, D))
    gpu.set_rows(values)
    gpu.cluster(0.85, 4, 100000)
    gpu.set_seed(0)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_save_files(gpu, oracle, tmp_path):
    _, _, values, ids = synth_rows(oracle, 30000, 4, 4, 9)
    rows = oracle.rows(values)
    rows.cluster(0.8, 6, 100000, oracle.planes(2))
    rows.save(str(tmp_path / "o.bin"), True, 5)
    gpu.set_seed(2)
    gpu.set_rows(values)
    gpu.cluster(0.8, 6, 100000)
    gpu.save(str(tmp_path / "g.bin"), True, 5)
    assert (tmp_path / "o.bin").read_bytes() == (tmp_path / "g.bin").read_bytes()
    assert (tmp_path / "o.bin.clust").read_bytes() == (tmp_path / "g.bin.clust").read_bytes()
    # reload through the reference's reader restatement and through ours
    gpu.load_cluster_file(str(tmp_path / "g.bin"), 8)
    back = oracle.read_cluster(str(tmp_path / "o.bin"), 8)
    assert_rows_equal(gpu.get_rows(), back.export())


@pytest.mark.parametrize("cta_max,cluster_max,cluster2_max,csize,direct_min", [
    (8, 64, 200, 8, 8192), (1, 1, 1, 8, 8192), (2, 1000000, 1000000, 4, 8192), (16, 100, 150, 16, 8192), (4, 4, 1000000, 8, 8192),
    (1000000, 1000000, 1000000, 8, 8192),
    # buckets of >= direct_min rows take the direct pipeline (cluster teams on the second stream, concurrent
    # with the single-CTA stage), with its own escalation to large clusters and the grid
    (8, 64, 200, 8, 40), (1000000, 20, 60, 8, 100), (4, 1000000, 1000000, 16, 33), (2, 2, 2, 8, 300)])
def test_merge_team_variants(oracle, monkeypatch, cta_max, cluster_max, cluster2_max, csize, direct_min):
    """The windowed merge with the escalation thresholds (representatives per team) forced low, so
    that small test buckets travel CTA -> cluster -> large cluster -> cooperative grid, against the oracle.
    Variants with an odd cta_max also switch the speculative resolver to its warp-parallel scan.  In the direct
    pipeline the cluster teams' opt-in screen pool is on from 48 representatives for direct_min 40 and 100 (chunk
    claims, stale board entries), off (the default) for direct_min 300, so that pipeline's cooperative-grid stage runs too."""
    from kmerlsh_b200 import Context

    monkeypatch.setenv("KLSH_PAR_SCAN", "1" if cta_max % 2 else "0")
    if direct_min in (40, 100):
        monkeypatch.setenv("KLSH_CPOOL", "1")
        monkeypatch.setenv("KLSH_CPOOL_MIN", "48")
    monkeypatch.setenv("KLSH_DIRECT_MIN", str(direct_min))
    monkeypatch.setenv("KLSH_CTA_MAX", str(cta_max))
    monkeypatch.setenv("KLSH_CLUSTER_MAX", str(cluster_max))
    monkeypatch.setenv("KLSH_CLUSTER2_MAX", str(cluster2_max))
    monkeypatch.setenv("KLSH_CLUSTER_SIZE", str(csize))
    cases = [(40000, 4, 4, 6, 0.85, 100000, 7), (120000, 10, 10, 4, 0.80, 1000, 4), (30000, 16, 16, 5, 0.9, 100000, 5),
             (8000, 40, 40, 3, 0.8, 100000, 9), (20000, 24, 24, 4, 0.85, 100000, 13)]
    with Context(0) as ctx:
        for n, sa, sb, iters, minsim, thr, seed in cases:
            _, _, values, ids = synth_rows(oracle, n, sa, sb, 60 + seed)
            rows = oracle.rows(values)
            rows.cluster(minsim, iters, thr, oracle.planes(seed))
            ctx.set_seed(seed)
            ctx.set_rows(values)
            ctx.cluster(minsim, iters, thr)
            assert_rows_equal(ctx.get_rows(), rows.export(), "n=%d D=%d" % (n, sa + sb))
        for name, values in sorted(_bucket_cases().items()):
            for thr in (0.95, 0.5):
                rows = oracle.rows(values)
                rows.p_cluster(thr)
                ctx.set_rows(values)
                ctx.p_cluster(thr)
                assert_rows_equal(ctx.get_rows(), rows.export(), name)


def test_merge_v1_kernel(oracle, monkeypatch):
    from kmerlsh_b200 import Context

    monkeypatch.setenv("KLSH_MERGE_V1", "1")
    _, _, values, ids = synth_rows(oracle, 60000, 6, 6, 71)
    rows = oracle.rows(values)
    rows.cluster(0.8, 5, 500, oracle.planes(3))
    with Context(0, seed=3) as ctx:
        ctx.set_rows(values)
        ctx.cluster(0.8, 5, 500)
        assert_rows_equal(ctx.get_rows(), rows.export())


@pytest.mark.parametrize("world,n,sa,sb,iters,nest", [(2, 60000, 4, 4, 5, 100000), (3, 60000, 4, 4, 3, 60), (4, 150000, 10, 10, 4, 1000)])
def test_sharded_cluster_in_process(oracle, world, n, sa, sb, iters, nest):
    """The multi-GPU protocol (replicated rows, partitioned bucket ranges, exchanged update logs) with
    `world` contexts stepped in lockstep on one GPU: every replica must end bit-identical to the
    oracle's single-process Cluster."""
    import torch

    from kmerlsh_b200 import Context
    from kmerlsh_b200 import distributed as kd

    _, _, values, ids = synth_rows(oracle, n, sa, sb, 91)
    rows = oracle.rows(values)
    ost = rows.cluster(0.82, iters, nest, oracle.planes(17))
    want = rows.export()
    ctxs = [Context(0, seed=17) for _ in range(world)]
    try:
        for c in ctxs:
            c.set_rows(values)
        stats = []
        kd.run_in_process([kd.TorchBackend(c, torch.device("cuda:0")) for c in ctxs], 0.82, iters, nest, stats)
        assert [s["rows_out"] for s in stats] == [s.rows_out for s in ost]
        assert all(s["my_buckets"] < s["buckets"] for s in stats if s["buckets"] > world)
        for c in ctxs:
            assert_rows_equal(c.get_rows(), want)
    finally:
        for c in ctxs:
            c.close()


def test_prefilter_boundaries(gpu, oracle):
    """The tensor-core / FMA prefilters may only drop pairs the exact test would drop: thresholds set
    exactly on (and one ulp either side of) similarities that occur in the bucket, rows scaled to
    tiny and huge magnitudes (fp16 copies are taken of unit-norm rows), a zero row and a NaN row."""
    rng = np.random.default_rng(123)
    base = rng.standard_normal((1, 24)).astype(np.float32)
    x = (base + np.float32(0.35) * rng.standard_normal((400, 24))).astype(np.float32)
    sims = np.array([np.float32(1) - oracle.cosine_distance(x[i], x[0]) for i in range(1, 60)], dtype=np.float32)
    thrs = []
    for s_ in np.sort(sims)[[5, 20, 40]]:
        thrs += [float(np.nextafter(s_, np.float32(-1))), float(s_), float(np.nextafter(s_, np.float32(2)))]
    scaled = x.copy()
    scaled[::3] *= np.float32(1e-18)
    scaled[1::3] *= np.float32(1e15)
    odd = x.copy()
    odd[7] = 0.0
    odd[11, 3] = np.nan
    for values in (x, scaled, odd):
        for thr in thrs[:6] if values is not x else thrs:
            rows = oracle.rows(values)
            rows.p_cluster(thr)
            gpu.set_rows(values)
            gpu.p_cluster(thr)
            assert_rows_equal(gpu.get_rows(), rows.export(), "thr=%r" % thr)


def test_degenerate_inputs(gpu, oracle):
    """Empty row set, everything dropped by the keep filter, a single column."""
    gpu.set_rows(np.zeros((0, 5), np.float32), np.zeros(1, np.uint64), np.zeros(0, np.uint64))
    st = gpu.cluster(0.8, 3, 1000)
    assert st[0].rows_in == 0 and gpu.row_count() == (0, 0)
    v, o, i = gpu.get_rows()
    assert v.shape == (0, 5) and list(o) == [0] and len(i) == 0
    gpu.load_counts(np.zeros((4, 100), np.uint16), np.ones(4, np.float32), 0)
    assert gpu.row_count() == (0, 0)
    gpu.cluster(0.8, 2, 1000)
    rng = np.random.default_rng(2)
    vals = rng.standard_normal((500, 1)).astype(np.float32)     # D = 1: cosine is +-1
    rows = oracle.rows(vals)
    rows.cluster(0.8, 3, 1000, oracle.planes(3))
    gpu.set_seed(3)
    gpu.set_rows(vals)
    gpu.cluster(0.8, 3, 1000)
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_eps_margin_rows_are_reported(gpu, oracle):
    """Rows that needed the exact re-evaluation of a plane are counted (and still signed exactly)."""
    rng = np.random.default_rng(5)
    vals = rng.standard_normal((20000, 16)).astype(np.float32)
    table = oracle.planes(1).table(14, 16)
    # rows built to be (numerically) orthogonal to plane 0: their sum lands inside the eps margin
    w = table[0].astype(np.float64)
    ortho = vals[:200].astype(np.float64)
    ortho -= np.outer(ortho @ w / (w @ w), w)
    vals[:200] = ortho.astype(np.float32)
    assert np.array_equal(gpu.sign(vals, table), oracle.sign(vals, table).astype(np.uint64))
    src = oracle.planes(1)
    rows = oracle.rows(vals)
    rows.cluster(0.9, 1, 100000, oracle.planes(1))
    gpu.set_plane_source(lambda H, D: src.table(H, D))
    gpu.set_rows(vals)
    st = gpu.cluster(0.9, 1, 100000)
    gpu.set_seed(0)
    assert st[0].eps_margin_rows >= 100   # most of the 200 constructed rows
    assert st[0].eps_margin_rows < 2000
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_plane_stream_tell_seek_and_done_callback(gpu, oracle):
    """klsh_plane_tell / klsh_plane_seek position the seeded hyperplane stream; the draws-done callback fires
    once per cluster() call, after the call's last table has been drawn."""
    _, _, values, ids = synth_rows(oracle, 20000, 4, 4, 15)
    gpu.set_seed(77)
    t1 = gpu.draw_table(9, 8)
    seed, drawn = gpu.plane_tell()
    assert (seed, drawn) == (77, 9)
    t2 = gpu.draw_table(5, 8)
    gpu.plane_seek(77, 9)
    assert gpu.draw_table(5, 8).tobytes() == t2.tobytes()
    gpu.plane_seek(77, 0)
    assert gpu.draw_table(9, 8).tobytes() == t1.tobytes()
    fired = []
    gpu.set_draws_done_callback(lambda: fired.append(gpu.plane_tell()[1]))
    gpu.set_seed(5)
    gpu.set_rows(values)
    st = gpu.cluster(0.85, 3, 60)      # nested buckets draw tables too
    end = gpu.plane_tell()[1]
    gpu.set_draws_done_callback(None)
    assert len(fired) == 1 and fired[0] == end and end >= sum(s.H for s in st)
    rows = oracle.rows(values)
    rows.cluster(0.85, 3, 60, oracle.planes(5))
    assert_rows_equal(gpu.get_rows(), rows.export())


def test_stash_rows_equals_file_round_trip(gpu, oracle, tmp_path):
    """klsh_stash_rows / klsh_unstash_rows: survivors of two batches concatenated on the device — same rows, order
    and member lists as appending both to a spill file and reading it back (klsh_save + klsh_load_cluster_file),
    with implicit ids (contiguous klsh_load_counts batches) and with explicit ids (klsh_set_rows batches)."""
    counts, vk, values, ids = synth_rows(oracle, 40000, 4, 4, 12)
    spill = str(tmp_path / "spill.bin")
    for explicit in (False, True):
        for b, (lo, hi) in enumerate(((0, 25000), (25000, 40000))):
            gpu.set_seed(100 + b)
            if explicit:
                sel = (ids >= lo) & (ids < hi)
                gpu.set_rows(values[sel], np.arange(sel.sum() + 1, dtype=np.uint64), ids[sel] * 3 + 7)
            else:
                gpu.load_counts(np.ascontiguousarray(counts[:, lo:hi]), vk, lo)
            gpu.cluster(0.8, 1, 40)
            gpu.save(spill, b == 0, 0)
            gpu.stash_rows()
        assert gpu.stash_count() > 0
        gpu.unstash_rows()
        assert gpu.stash_count() == 0
        got = gpu.get_rows()
        gpu.load_cluster_file(spill, 8)
        assert_rows_equal(got, gpu.get_rows(), "explicit ids" if explicit else "implicit ids")
        gpu.set_seed(9)
        gpu.cluster(0.8, 3, 1000)      # and the stash is a working row set
        a = gpu.get_rows()
        assert len(a[1]) - 1 < len(got[1]) - 1


@pytest.mark.parametrize("env", [{"KLSH_NO_SPEC": "1"}, {"KLSH_PAR_SCAN": "1"}, {"KLSH_SCAN": "0"}, {"KLSH_SCAN": "2"}, {},
                                 {"KLSH_DIRECT_MIN": "2000"},
                                 {"KLSH_CPOOL": "1", "KLSH_CPOOL_MIN": "64", "KLSH_DIRECT_MIN": "2000", "KLSH_CTA_MAX": "300"},
                                 {"KLSH_CPOOL": "1", "KLSH_CPOOL_MIN": "64", "KLSH_CPOOL_HELPERS": "64", "KLSH_DIRECT_MIN": "2000"},
                                 {"KLSH_POOL": "1", "KLSH_POOL_MIN": "64"}])
def test_window_resolution_modes(oracle, monkeypatch, env):
    """However a window is resolved and whoever screens it, the clusters are the oracle's bit for bit: the
    sequential loop only; the speculative resolver with each of its scans (scalar replay, closed-form order,
    lean replay; default: per team kind); the direct pipeline's cluster teams from 2000 rows on without the screen
    pool (the default) and with the opt-in pool from 64 representatives on (chunks claimed by the cluster's own CTAs;
    or also by idle teams staying on as helpers); single-CTA teams with the pool for every bucket."""
    from kmerlsh_b200 import Context

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cases = [(150000, 8, 8, 5, 0.85, 100000, 3), (60000, 16, 16, 4, 0.9, 100000, 5), (90000, 4, 4, 6, 0.8, 60, 7),
             (50000, 24, 24, 4, 0.85, 100000, 11)]
    with Context(0) as ctx:
        for n, sa, sb, iters, minsim, thr, seed in cases:
            _, _, values, ids = synth_rows(oracle, n, sa, sb, 80 + seed)
            rows = oracle.rows(values)
            rows.cluster(minsim, iters, thr, oracle.planes(seed))
            ctx.set_seed(seed)
            ctx.set_rows(values)
            ctx.cluster(minsim, iters, thr)
            assert_rows_equal(ctx.get_rows(), rows.export(), "n=%d D=%d %r" % (n, sa + sb, env))
