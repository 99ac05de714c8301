"""Multi-rank logic on CPU (gloo, world_size 2): the batch is sharded by contiguous model ranges with no
data-path collective; inputs are a pure function of the global model index, so any sharding reproduces the
single-rank batch; only a reduced status summary crosses ranks (the oracle stands in for the GPU solve)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hd_b200 import frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bays, per_rank, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import hd_oracle as oracle
    fm = oracle.flatten_model(frames.uniform_frame(n_bays))
    sec = frames.uniform_frame_sections(n_bays, rank * per_rank, per_rank)
    loads = frames.uniform_frame_loads(n_bays, rank * per_rank, per_rank, n_cases=3)
    status, disp, _, react = oracle.static_batch(fm, sec, loads, n_threads=1, want=("disp", "react"))
    np.save(os.path.join(out_dir, f"disp{rank}.npy"), disp)
    bad = torch.tensor([float(np.abs(status).sum())], dtype=torch.float64)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(bad, op=dist.ReduceOp.SUM)  # the only collective of the path
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # max-over-ranks timing as in bench.py
    assert bad.item() == 0 and t.item() == world
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_batch(tmp_path, oracle):
    world, per_rank, n_bays = 2, 3, 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_bays, per_rank, str(tmp_path)), nprocs=world, join=True)
    fm = oracle.flatten_model(frames.uniform_frame(n_bays))
    sec = frames.uniform_frame_sections(n_bays, 0, world * per_rank)
    loads = frames.uniform_frame_loads(n_bays, 0, world * per_rank, n_cases=3)
    _, disp, _, _ = oracle.static_batch(fm, sec, loads, n_threads=1, want=("disp",))
    got = np.concatenate([np.load(tmp_path / f"disp{r}.npy") for r in range(world)])
    assert np.array_equal(got, disp)
