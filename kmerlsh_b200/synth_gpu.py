"""Device-side synthetic generator for the large benchmark shapes (C2-C5).

Same law as synth.synth_counts but drawn with torch's CUDA generators,
so the exact stream differs from numpy's: these shapes are throughput workloads, parity is pinned
on C1 and on slices checked against the oracle.  torch is used here only to
make synthetic input; it is not on the measured path.
"""
from __future__ import annotations

import numpy as np


def synth_counts_gpu(n: int, sa: int, sb: int, seed: int, device="cuda:0", pin=True):
    """Return (counts[S][n] uint16 host array (pinned if requested), coverage[S] float64)."""
    import torch

    s = sa + sb
    g_num = max(8, n // 2000)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    base = torch.exp(2.0 + 1.0 * torch.randn(g_num, 1, generator=gen, device=device, dtype=torch.float64))
    samp = torch.exp(0.6 * torch.randn(g_num, s, generator=gen, device=device, dtype=torch.float64))
    diff = torch.rand(g_num, generator=gen, device=device) < 0.2
    fold = torch.ones(g_num, s, device=device, dtype=torch.float64)
    fold[diff, sa:] *= 4.0
    lam = (base * samp * fold).to(torch.float32)
    g = torch.randint(0, g_num, (n,), generator=gen, device=device)
    host = torch.empty((s, n), dtype=torch.int16, pin_memory=pin)
    cov = np.empty(s, dtype=np.float64)
    for j in range(s):
        c = torch.poisson(lam[g, j], generator=gen).clamp_(max=65535.0)
        cov[j] = float(torch.log(c.clamp(min=1.0).to(torch.float64)).sum().item())
        host[j].copy_(c.to(torch.int32).to(torch.int16))  # two's complement bits == uint16 bits
    return host.numpy().view(np.uint16), cov
