"""N>1 host logic on CPU: two gloo ranks shard volumes, produce (fake) per-volume HSP lists and
gather them on rank 0 in database order.  The GPU kernels are not involved (no GPU here)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gblastn_b200 import shard, abi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_vol, seqs_per_vol = 5, 7
        mine = shard.assign_volumes(n_vol, world)[rank]
        parts = []
        for v in mine:
            rng = np.random.default_rng(100 + v)
            n = int(rng.integers(0, 6))
            h = np.zeros(n, dtype=abi.HSP_DTYPE)
            h["oid"] = np.sort(rng.integers(0, seqs_per_vol, size=n))
            h["score"] = 1000 * v + np.arange(n)          # encodes (volume, list position)
            parts.append(shard.globalize_oids(h, v * seqs_per_vol))
        local = np.concatenate(parts) if parts else np.zeros(0, dtype=abi.HSP_DTYPE)
        allh = shard.gather_hsps(local, dist)
        if rank == 0:
            q.put((allh["oid"].tolist(), allh["score"].tolist()))
        else:
            assert allh is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_matches_sequential_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    oids, scores = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: a single sequential pass over volumes 0..4
    exp_oid, exp_score = [], []
    for v in range(5):
        rng = np.random.default_rng(100 + v)
        n = int(rng.integers(0, 6))
        o = np.sort(rng.integers(0, 7, size=n)) + 7 * v
        exp_oid += o.tolist()
        exp_score += (1000 * v + np.arange(n)).tolist()
    assert oids == exp_oid
    assert scores == exp_score


def test_assign_volumes_round_robin():
    sys.path.insert(0, ROOT)
    from gblastn_b200 import shard
    assert shard.assign_volumes(8, 8) == [[i] for i in range(8)]
    assert shard.assign_volumes(5, 2) == [[0, 2, 4], [1, 3]]
    assert shard.assign_volumes(0, 4) == [[], [], [], []]
