"""torchrun --nproc-per-node N tools/mg_check.py: sharded Cluster over N GPUs (NCCL) vs one GPU."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, torch.distributed as dist
from kmerlsh_b200 import Context, synth, distributed as kd

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
counts, cov = synth.synth_counts(n, 10, 10, 20261018)
kmap, cov32 = synth.parse_log_line(synth.format_log_line(n, cov), 20)
vk = synth.v_kmers_from_cov(cov32, kmap)
ctx = Context(lr, seed=42)
ctx.load_counts(counts, vk, 0)
be = kd.TorchBackend(ctx, torch.device("cuda", lr))
stats = []
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
kd.run_with_torch_distributed(be, 0.80, iters, 1000000, stats)
torch.cuda.synchronize(); dist.barrier(); t1 = time.time() - t0
got = ctx.get_rows()
if rank == 0:
    ref = Context(lr, seed=42)
    ref.load_counts(counts, vk, 0)
    t0 = time.time(); st = ref.cluster(0.80, iters, 1000000); t2 = time.time() - t0
    want = ref.get_rows()
    ok = got[0].tobytes() == want[0].tobytes() and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    print("world %d rows %d iters %d: sharded %.3fs single %.3fs identical=%s final=%d" % (world, n, iters, t1, t2, ok, len(want[1]) - 1))
    assert ok
dist.destroy_process_group()
