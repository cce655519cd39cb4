"""Cluster() sharded over several GPUs of one node: replicated row state, partitioned merge work.

Every rank holds the same rows and the same hyperplane source.  Per LSH iteration each rank signs
and groups all rows (no communication), merges only its own contiguous range of buckets, and the
ranks all-gather three things: the survivors of each range, the rows each range modified (values +
member metadata) and the member-chain pointer writes.  Contiguous bucket ranges in rank order keep
the reference's canonical row order, so the result is identical to the single-GPU `klsh_cluster`
(and hence to the reference's T=1 run).  See DESIGN.md section 7.

The protocol is written once, as a generator that yields at every collective:
`sharded_cluster_steps(backend, rank, world, ...)` yields a list of local arrays and receives the
list (over ranks) of the gathered arrays.  Drivers:
  * `run_with_torch_distributed` — one process per GPU, NCCL all-gather on device tensors;
  * `run_in_process` — several ranks stepped in lockstep inside one process (tests on one GPU).
The backend is anything with the `mg_*` methods of `kmerlsh_b200.api.Context` wrapped by
`TorchBackend` below (tests drive the same protocol over gloo with a CPU test double).
"""
from __future__ import annotations

import numpy as np


def float32_threshold_schedule(min_similarity: float, iterations: int):
    """threshold_k of Cluster(): fp32 recurrence (reference function/cluster.cc:190-192, :330)."""
    max_similarity = np.float32(0.95)
    step = np.float32((max_similarity - np.float32(min_similarity)) / np.float32(iterations))
    thr = max_similarity
    for _ in range(iterations):
        yield thr
        thr = np.float32(thr - step)


def sharded_cluster_steps(backend, rank: int, world: int, min_similarity: float, iterations: int,
                          bucket_size_threshold: int, stats: list | None = None):
    """Generator implementing Cluster() for one rank.  Yields `[counts, surv, mod_rows, mod_vals,
    mod_meta, chain_slots, chain_vals]` (backend arrays) at each iteration's exchange point and expects
    to be sent the list over ranks of those lists."""
    for it, thr in enumerate(float32_threshold_schedule(min_similarity, iterations)):
        n, H, nb = backend.mg_pass_begin()
        if n == 0:
            break
        splits = backend.mg_plan(world)
        b_lo, b_hi = splits[rank], splits[rank + 1]
        n_surv, n_mod, n_chain = backend.mg_merge(b_lo, b_hi, float(thr), bucket_size_threshold)
        local = backend.mg_export(n_surv, n_mod, n_chain)
        gathered = yield local
        total = 0
        alive_parts = []
        for r in range(world):
            counts, surv, mod_rows, mod_vals, mod_meta, slots, vals = gathered[r]
            ns, nm, nc = (int(x) for x in backend.to_host(counts))
            if r != rank:
                backend.mg_apply(mod_rows, mod_vals, mod_meta, nm, slots, vals, nc)
            alive_parts.append((surv, ns))
            total += ns
        backend.mg_set_alive(alive_parts, total)
        if stats is not None:
            stats.append({"iteration": it + 1, "rows_in": n, "rows_out": total, "H": H, "buckets": nb,
                          "threshold": float(thr), "my_buckets": b_hi - b_lo, "my_survivors": n_surv,
                          "my_modified_rows": n_mod})


class TorchBackend:
    """kmerlsh_b200.Context + torch CUDA tensors as exchange buffers."""

    def __init__(self, ctx, device):
        import torch

        self.torch = torch
        self.ctx = ctx
        self.device = device

    def mg_pass_begin(self):
        return self.ctx.mg_pass_begin()

    def mg_plan(self, world):
        return self.ctx.mg_plan(world)

    def mg_merge(self, b_lo, b_hi, thr, nest):
        return self.ctx.mg_merge(b_lo, b_hi, thr, nest)

    def mg_export(self, n_surv, n_mod, n_chain):
        t = self.torch
        dev = self.device
        stride = self.ctx.row_stride()  # known once rows are loaded
        counts = t.tensor([n_surv, n_mod, n_chain], dtype=t.int64, device=dev)
        surv = t.empty(max(n_surv, 1), dtype=t.int32, device=dev)
        mod_rows = t.empty(max(n_mod, 1), dtype=t.int32, device=dev)
        mod_vals = t.empty((max(n_mod, 1), stride), dtype=t.float32, device=dev)
        mod_meta = t.empty((max(n_mod, 1), 3), dtype=t.int32, device=dev)
        slots = t.empty(max(n_chain, 1), dtype=t.int32, device=dev)
        vals = t.empty(max(n_chain, 1), dtype=t.int32, device=dev)
        self.ctx.mg_export(surv.data_ptr(), mod_rows.data_ptr(), mod_vals.data_ptr(), mod_meta.data_ptr(),
                           slots.data_ptr(), vals.data_ptr())
        return [counts, surv, mod_rows, mod_vals, mod_meta, slots, vals]

    def to_host(self, counts):
        return counts.cpu().tolist()

    def mg_apply(self, mod_rows, mod_vals, mod_meta, nm, slots, vals, nc):
        self.torch.cuda.current_stream().synchronize()
        self.ctx.mg_apply(mod_rows.data_ptr(), mod_vals.data_ptr(), mod_meta.data_ptr(), nm, slots.data_ptr(), vals.data_ptr(), nc)

    def mg_set_alive(self, parts, total):
        t = self.torch
        alive = t.cat([s[:n] for s, n in parts]) if total else t.empty(1, dtype=t.int32, device=self.device)
        t.cuda.current_stream().synchronize()
        self.ctx.mg_set_alive(alive.data_ptr(), total)


def _pad_gather(torch, dist, tensors, world):
    """all_gather of variable-length tensors: lengths first, then padded payloads."""
    out = [[None] * len(tensors) for _ in range(world)]
    lens = torch.tensor([t.shape[0] for t in tensors], dtype=torch.int64, device=tensors[0].device)
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    all_lens = torch.stack(all_lens).cpu()
    for k, t in enumerate(tensors):
        m = int(all_lens[:, k].max())
        buf = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        recv = [torch.empty_like(buf) for _ in range(world)]
        if buf.numel():
            dist.all_gather(recv, buf)
        for r in range(world):
            out[r][k] = recv[r][: int(all_lens[r, k])]
    return out


def run_with_torch_distributed(backend, min_similarity, iterations, bucket_size_threshold, stats=None):
    """Drive the protocol with torch.distributed (NCCL on GPUs, gloo on CPU test doubles)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    gen = sharded_cluster_steps(backend, rank, world, min_similarity, iterations, bucket_size_threshold, stats)
    try:
        local = next(gen)
        while True:
            gathered = _pad_gather(torch, dist, local, world)
            local = gen.send(gathered)
    except StopIteration:
        pass


def run_in_process(backends, min_similarity, iterations, bucket_size_threshold, stats=None):
    """Several ranks stepped in lockstep inside one process (one GPU, several contexts)."""
    world = len(backends)
    gens = [sharded_cluster_steps(b, r, world, min_similarity, iterations, bucket_size_threshold,
                                  stats if r == 0 else None) for r, b in enumerate(backends)]
    locals_ = []
    alive = [True] * world
    for g in gens:
        try:
            locals_.append(next(g))
        except StopIteration:
            return
    while True:
        nxt = []
        for r, g in enumerate(gens):
            try:
                nxt.append(g.send(locals_))
            except StopIteration:
                alive[r] = False
        if not all(alive):
            return
        locals_ = nxt
