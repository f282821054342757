"""world_size-2 gloo run of the multi-GPU host logic (sharding, stat reduction, rollout gather) on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nascargymnasium_b200 import distributed as D


def test_shard_range_partitions():
    for n in (1, 7, 8, 4096, 65536, 65537):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_track_ids_cover_all_tracks_sorted():
    ids = np.concatenate([D.shard_track_ids(65536, 8, r, 8) for r in range(8)])
    assert np.bincount(ids, minlength=8).tolist() == [8192] * 8
    one = D.shard_track_ids(65536, 8, 3, 8)
    assert np.all(np.diff(one) >= 0) and len(one) == 8192


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(9, rank, world)            # 5 + 4 envs: uneven shards
    stats = D.reduce_stats({"car_steps": (hi - lo) * 10, "episodes": rank + 1, "return_sum": -1.5 * (rank + 1)})
    tmax = D.max_over_ranks(1.0 + rank)
    local = torch.arange(lo, hi, dtype=torch.float32).repeat(3, 1).unsqueeze(-1) + 100.0 * torch.arange(3).view(3, 1, 1)
    full = D.gather_rollout(local, dst=0)
    if rank == 0:
        out.put((stats, tmax, full.numpy()))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, tmax, full = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert stats == {"car_steps": 90.0, "episodes": 3.0, "return_sum": -4.5}
    assert tmax == 2.0
    assert full.shape == (3, 9, 1)
    assert np.array_equal(full[0, :, 0], np.arange(9, dtype=np.float32))
    assert np.array_equal(full[2, :, 0], np.arange(9, dtype=np.float32) + 200.0)
