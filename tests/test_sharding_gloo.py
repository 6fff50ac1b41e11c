"""CPU suite, part 3: the N>1 launch path (one process per GPU, pair-index sharding, max-over-ranks
timing) exercised with world_size 2 on the gloo backend.  The data path has no collective; only
the barrier and the timing reduction go through torch.distributed -- exactly what bench.py does."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import oracle as O
    from interpolatory_b200.sharding import pairs_of_rank, frames_needed
    from interpolatory_b200.synth import make_sequence
    frames = make_sequence(40, 56, n_pairs + 1, seed=9)
    mine = pairs_of_rank(n_pairs, rank, world)
    need = frames_needed(mine)
    digests = {}
    for k in mine:                      # the oracle stands in for the device path on this CPU box
        assert k in need and k + 1 in need
        f_, b_ = O.fs(8, 4, frames[k], frames[k + 1]), O.fs(8, 4, frames[k + 1], frames[k])
        digests[k] = int(O.bidir_helper(frames[k], frames[k + 1], f_, b_, np.float32(0.5)).astype(np.int64).sum())
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)      # pretend elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        q.put((float(t[0]), gathered))
    dist.destroy_process_group()


def test_two_rank_pair_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_pairs = 2, 5
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_max, gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 2.0                                   # max over ranks
    merged = {}
    for d in gathered:
        assert not (set(d) & set(merged))                  # no pair done twice
        merged.update(d)
    assert sorted(merged) == list(range(n_pairs))          # every pair done once
    # and the sharded result equals the single-process result
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as O
    from interpolatory_b200.synth import make_sequence
    frames = make_sequence(40, 56, n_pairs + 1, seed=9)
    for k in range(n_pairs):
        f_, b_ = O.fs(8, 4, frames[k], frames[k + 1]), O.fs(8, 4, frames[k + 1], frames[k])
        assert merged[k] == int(O.bidir_helper(frames[k], frames[k + 1], f_, b_, np.float32(0.5)).astype(np.int64).sum())
