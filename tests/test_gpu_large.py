"""GPU: parity at sizes the oracle needs minutes for (against COMMITTED oracle digests), an
adversarial test of the signing kernel's error budget, and the multi-GPU path over real NCCL."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from kmerlsh_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIGESTS = json.load(open(os.path.join(ROOT, "tests", "golden", "large_digests.json")))


def _md5(a):
    return hashlib.md5(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_large_parity_against_oracle_digests(gpu, key):
    """Phase 1 (nested buckets) + I iterations on millions of rows of the C2 / C3 / C5 shapes.  The
    expected md5s of centroids, offsets and member ids were produced by the C oracle
    (`tools/parity_large.py <cfg> <n> <I> oracle`, minutes of CPU) and are committed in
    tests/golden/large_digests.json; the CUDA path must reproduce them bit for bit."""
    want = DIGESTS[key]
    cfg, n, iters = want["config"]
    _, sa, sb, seed = synth.CONFIGS[cfg]
    counts, cov = synth.synth_counts(n, sa, sb, seed)
    if _md5(counts) != want["counts_md5"]:
        pytest.skip("numpy generator stream differs from the one the oracle digest was made with")
    kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), sa + sb)
    vk = synth.v_kmers_from_cov(cov32, kmap)
    gpu.set_seed(42)
    gpu.load_counts(counts, vk, 0)
    gpu.cluster(0.80, 1, n // 1000)
    gpu.cluster(0.80, iters, 1000000)
    values, offs, ids = gpu.get_rows()
    assert (len(offs) - 1, len(ids)) == (want["clusters"], want["ids"])
    assert _md5(values) == want["values_md5"]
    assert _md5(offs.astype(np.uint64)) == want["offsets_md5"]
    assert _md5(ids.astype(np.uint64)) == want["ids_md5"]


@pytest.mark.parametrize("D,H,n", [(32, 25, 300000), (20, 19, 100000), (64, 29, 150000), (200, 24, 40000)])
def test_sign_adversarial_cancellation(gpu, oracle, D, H, n):
    """The tensor-core projection is trusted only outside eps = c * 2^-24 * |w| * |x|.  Rows built so
    that ONE plane's sum is a catastrophic cancellation inside a single 8-wide k-step: two terms of
    +-M (M up to 2^20) whose difference, plus many tiny terms, lands between 0 and a few eps on either
    side of zero — the region where a too-small budget would flip a key bit.  Every bit must equal
    the reference chain's (oracle.sign)."""
    rng = np.random.default_rng(1000 + D)
    table = oracle.planes(77).table(H, D)
    rows = np.empty((n, D), np.float32)
    c_eps = (D + 32 + 30 * ((D + 7) // 8)) * 2.0 ** -24
    for r0 in range(0, n, 20000):
        m = min(20000, n - r0)
        p = (np.arange(r0, r0 + m) % H)
        w = table[p].astype(np.float64)                                  # the plane each row attacks
        ks = rng.integers(0, (D + 7) // 8, size=m)                       # the k-step holding the pair
        lo = ks * 8
        width = np.minimum(8, D - lo)
        a = lo + rng.integers(0, 8, size=m) % width
        b = lo + (a - lo + 1 + rng.integers(0, 7, size=m) % np.maximum(width - 1, 1)) % width
        single = width < 2
        M = 2.0 ** rng.choice([0, 6, 12, 20], size=m)
        x = rng.standard_normal((m, D)) * 1e-3                           # the tiny terms
        idx = np.arange(m)
        x[idx, a] = M / w[idx, a]
        x[idx, b] = np.where(single, x[idx, b], -M / w[idx, b])
        x32 = x.astype(np.float32).astype(np.float64)
        # steer the exact sum to alpha * eps with alpha in [-4, 4] by nudging one tiny element
        eps = c_eps * np.linalg.norm(w, axis=1) * np.linalg.norm(x32, axis=1)
        target = rng.uniform(-4, 4, size=m) * eps
        s = (w * x32).sum(axis=1)
        free = (a + 2) % D
        clash = (free == a) | (free == b)
        free = np.where(clash, (free + 2) % D, free)
        x32[idx, free] += (target - s) / w[idx, free]
        rows[r0:r0 + m] = x32.astype(np.float32)
    rows[::1001] *= np.float32(1e30)        # overflowing sums (inf - inf = NaN -> bit 0)
    want = oracle.sign(rows, table).astype(np.uint64)
    got = gpu.sign(rows, table)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, "%d rows differ, first %r" % (bad.size, bad[:5])


def test_sharded_cluster_over_nccl(tmp_path):
    """Real NCCL: the multi-GPU path under torchrun on 2 GPUs must end with the same clusters, bit
    for bit, as one GPU (skipped when fewer than 2 GPUs are visible)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "mg.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "mg_check.py"), "1500000", "12", str(out)]
    subprocess.run(cmd, check=True, cwd=ROOT, timeout=600)
    res = json.load(open(out))
    assert res["identical"] and res["world"] == 2
