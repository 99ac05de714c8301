"""Multi-rank host logic on CPU (gloo, world_size 2), no solver involved: every rank asks the library for its own
contiguous model range (hdb_shard_range, the rule hdb_multi_* and bench.py use), the ranges tile the batch exactly
for even and ragged sizes, the inputs a rank builds for its range (bench.build_inputs: a pure function of the global
model index) are the corresponding slice of the single-rank batch bit for bit, and the only collectives of the path
- a summed failure count and a max-over-ranks time - behave as bench.py expects.  The sharded GPU solve itself is
tests/test_multi_gpu.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bays, totals, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from hd_b200 import _lib
    for total in totals:
        first, count = _lib.shard_range(total, world, rank)
        # every rank learns every range: they must tile [0, total) in rank order, sizes differing by at most one
        mine = torch.tensor([first, count], dtype=torch.int64)
        ranges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ranges, mine)
        pos = 0
        for f, c in (r.tolist() for r in ranges):
            assert f == pos and c >= 0
            pos += c
        assert pos == total
        counts = [int(r[1]) for r in ranges]
        assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
        sec, loads = bench.build_inputs(n_bays, first, count, 3)
        np.save(os.path.join(out_dir, f"sec_{total}_{rank}.npy"), sec)
        np.save(os.path.join(out_dir, f"loads_{total}_{rank}.npy"), loads)
    bad = torch.tensor([float(rank)], dtype=torch.float64)      # rank 1 reports one failed model
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(bad, op=dist.ReduceOp.SUM)  # the only collectives of the path (bench.py): failure count ...
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # ... and the max-over-ranks time
    assert bad.item() == world * (world - 1) / 2 and t.item() == world
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_tiles_the_batch_and_reproduces_its_inputs(tmp_path):
    import bench
    world, n_bays, totals = 2, 4, (6, 7, 1, 0)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_bays, totals, str(tmp_path)), nprocs=world, join=True)
    for total in totals:
        sec, loads = bench.build_inputs(n_bays, 0, total, 3)
        got_sec = np.concatenate([np.load(tmp_path / f"sec_{total}_{r}.npy") for r in range(world)])
        got_loads = np.concatenate([np.load(tmp_path / f"loads_{total}_{r}.npy") for r in range(world)])
        assert np.array_equal(got_sec, np.asarray(sec)) and np.array_equal(got_loads, np.asarray(loads))


def test_shard_range_rule():
    from hd_b200 import _lib
    for total in (0, 1, 5, 70, 8192, 65536):
        for n in (1, 2, 3, 8):
            pos = 0
            for d in range(n):
                first, count = _lib.shard_range(total, n, d)
                assert first == pos and count in (total // n, total // n + 1)
                pos += count
            assert pos == total
    with __import__("pytest").raises(_lib.HdbError):
        _lib.shard_range(4, 2, 2)
