"""torchrun --nproc-per-node N tools/mg_check.py [rows] [iters] [out.json]

Multi-GPU parity over real NCCL, everything through the C ABI (the NCCL calls are inside libklsh):
  1. Cluster sharded over N GPUs (klsh_mg_cluster) vs klsh_cluster on one GPU, same rows;
  2. phase 1 batch-per-GPU (each rank clusters its own slice of the count matrix, I=1), the survivors
     all-gathered in rank order over NVLink (klsh_mg_gather_rows), then the sharded -I iterations, vs one
     GPU running the batches one after the other and clustering the concatenation.
torch.distributed is used only to hand the NCCL unique id to the other processes.
"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import torch
import torch.distributed as dist

from kmerlsh_b200 import Context, nccl_unique_id, synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
out_path = sys.argv[3] if len(sys.argv) > 3 else None

uid = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
counts, cov = synth.synth_counts(n, 10, 10, 20261018)
kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), 20)
vk = synth.v_kmers_from_cov(cov32, kmap)


def same(a, b):
    return a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


ctx = Context(lr, seed=42)
ctx.mg_init(rank, world, uid[0])

# ---- 1. one Cluster call sharded over the ranks ------------------------------------------------------
ctx.load_counts(counts, vk, 0)
dist.barrier()
t0 = time.time()
ctx.mg_cluster(0.80, iters, 1000000)
ctx.sync()
t_sharded = time.time() - t0
got1 = ctx.get_rows()

# ---- 2. phase 1 batch per GPU + all-gather of the survivors + sharded phase 2 -------------------------
per = (n + world - 1) // world
lo, hi = rank * per, min(n, (rank + 1) * per)
ctx.set_seed(1000 + rank)                       # every batch its own hyperplane stream
ctx.load_counts(np.ascontiguousarray(counts[:, lo:hi]), vk, lo)
ctx.cluster(0.80, 1, max(1, per // 1000))
ctx.mg_gather_rows()
ctx.set_seed(42)
ctx.mg_cluster(0.80, iters, 1000000)
got2 = ctx.get_rows()

ok1 = ok2 = True
if rank == 0:
    ref = Context(lr, seed=42)
    ref.load_counts(counts, vk, 0)
    t0 = time.time()
    ref.cluster(0.80, iters, 1000000)
    ref.sync()
    t_single = time.time() - t0
    ok1 = same(got1, ref.get_rows())
    parts = []
    for b in range(world):
        blo, bhi = b * per, min(n, (b + 1) * per)
        ref.set_seed(1000 + b)
        ref.load_counts(np.ascontiguousarray(counts[:, blo:bhi]), vk, blo)
        ref.cluster(0.80, 1, max(1, per // 1000))
        parts.append(ref.get_rows())
    values = np.concatenate([p[0] for p in parts])
    sizes = np.concatenate([np.diff(p[1].astype(np.int64)) for p in parts])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ids = np.concatenate([p[2] for p in parts])
    ref.set_seed(42)
    ref.set_rows(values, offs, ids)
    ref.cluster(0.80, iters, 1000000)
    ok2 = same(got2, ref.get_rows())
    print("world %d rows %d iters %d: sharded %.3fs single %.3fs | sharded identical=%s, batch-per-GPU + gather identical=%s, clusters=%d" % (
        world, n, iters, t_sharded, t_single, ok1, ok2, len(got2[1]) - 1), flush=True)
    if out_path:
        json.dump({"world": world, "rows": n, "iters": iters, "identical": bool(ok1 and ok2), "sharded_identical": bool(ok1),
                   "gather_identical": bool(ok2), "sharded_s": t_sharded, "single_s": t_single}, open(out_path, "w"))
flags = [ok1 and ok2]
dist.broadcast_object_list(flags, src=0)
ctx.mg_finalize()
dist.destroy_process_group()
sys.exit(0 if flags[0] else 1)
