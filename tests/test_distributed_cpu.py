"""CPU, world_size 2 over gloo: the sharded-Cluster protocol (kmerlsh_b200/distributed.py) driven
with a CPU test double of the klsh_mg_* building blocks must reproduce the single-process oracle
exactly — partition planning, update exchange, replay and canonical order are what is under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, values, seed, minsim, iters, nest, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from fake_mg import FakeMgBackend
    from oracle_lib import Oracle

    from kmerlsh_b200 import distributed as kd

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    be = FakeMgBackend(Oracle(), values, seed)
    stats = []
    kd.run_with_torch_distributed(be, minsim, iters, nest, stats)
    v, o, i = be.get_rows()
    np.savez(out_path % rank, v=v, o=o, i=i, rows_out=np.array([s["rows_out"] for s in stats]),
             mine=np.array([s["my_survivors"] for s in stats]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d,iters,nest,world", [(2500, 6, 3, 100000, 2), (3000, 8, 2, 40, 2), (1500, 4, 3, 100000, 3)])
def test_sharded_protocol_matches_oracle(oracle, tmp_path, n, d, iters, nest, world):
    from helpers import synth_rows

    _, _, values, ids = synth_rows(oracle, n, d // 2, d - d // 2, 5)
    seed, minsim = 11, 0.85
    rows = oracle.rows(values)
    st = rows.cluster(minsim, iters, nest, oracle.planes(seed))
    want = rows.export()
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(world, _free_port(), values, seed, minsim, iters, nest, out), nprocs=world, join=True)
    got = [np.load(out % r) for r in range(world)]
    for r in range(world):  # every replica ends identical to the single-process result
        assert got[r]["v"].tobytes() == want[0].tobytes(), r
        assert np.array_equal(got[r]["o"], want[1]) and np.array_equal(got[r]["i"], want[2]), r
    assert list(got[0]["rows_out"]) == [s.rows_out for s in st[:len(got[0]["rows_out"])]]
    # the work really was partitioned: every rank produced part of the survivors in the first iteration
    firsts = [int(g["mine"][0]) for g in got]
    assert sum(firsts) == st[0].rows_out and all(f > 0 for f in firsts)


def test_threshold_schedule_matches_oracle(oracle):
    from kmerlsh_b200.distributed import float32_threshold_schedule

    for minsim, iters in ((0.8, 100), (0.9, 500), (0.85, 7)):
        sched = list(float32_threshold_schedule(minsim, iters))
        assert len(sched) == iters
        for k in (0, 1, iters // 2, iters - 1):
            assert sched[k] == oracle.threshold_after(minsim, iters, k)
