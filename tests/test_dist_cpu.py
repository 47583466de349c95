"""world_size-2 gloo test of the multi-GPU plumbing on CPU: instance sharding + the statistics gather."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_ilqr_mujoco_b200.sharding import gather_instance_stats, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    ids = np.arange(lo, hi)
    cost, iters, status = ids * 1.5, ids % 7, (ids % 5 == 0).astype(np.int64)
    c, i, s = gather_instance_stats(cost, iters, status)
    if rank == 0:
        q.put((c.tolist(), i.tolist(), s.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    total, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r == 0 else 0, total, q)) for r in range(world)]
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    c, i, s = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = np.arange(total)
    assert c == (ids * 1.5).tolist() and i == (ids % 7).tolist() and s == (ids % 5 == 0).astype(int).tolist()
