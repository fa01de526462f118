"""N>1 host logic on CPU: POI sharding, image broadcast and result gather with torch.distributed
(gloo, world_size 2 and 3).  The per-rank compute is stood in by the CPU oracle here (the CUDA
engine needs a GPU); what is under test is opencorr_b200/distributed.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opencorr_b200 import distributed as obd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 50000, 500001):
        for world in (1, 2, 3, 8):
            spans = [obd.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == obd.all_shard_sizes(n, world)
    with pytest.raises(ValueError):
        obd.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_poi, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opencorr_b200 import make_poi2d, synth
    from oracle.oracle import Oracle2D
    try:
        # rank 0 owns the inputs; the others start from zeros and must receive them
        if rank == 0:
            ref, tar = synth.speckle_pair_2d(160, 144)
            xy = synth.grid_2d(30, 30, n_poi, 1, 2, 1)[:n_poi]
            t_ref, t_tar = torch.from_numpy(ref), torch.from_numpy(tar)
            allq = torch.from_numpy(make_poi2d(xy))
        else:
            t_ref, t_tar = torch.zeros(144, 160), torch.zeros(144, 160)
            allq = None
        obd.broadcast_images(t_ref, t_tar, src=0)
        shard = obd.scatter_pois(allq, n_poi, 25, "cpu", src=0)
        lo, hi = obd.shard_bounds(n_poi, world, rank)
        assert shard.shape == (hi - lo, 25)
        q = shard.numpy().copy()
        o = Oracle2D(t_ref.numpy(), t_tar.numpy(), threads=1)
        o.fftcc2d(q, 10, 10)
        o.icgn2d1(q, 10, 10, 0.001, 10)
        q[:, 19] = rank  # 'feature' column carries the rank that processed the record
        gathered = obd.gather_pois(torch.from_numpy(q), n_poi, dst=0)
        if rank == 0:
            np.save(out_path, gathered.numpy())
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_poi", [(2, 9), (3, 10), (2, 1)])
def test_scatter_compute_gather_matches_single_process(tmp_path, world, n_poi):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_poi, out), nprocs=world, join=True)
    got = np.load(out)
    from opencorr_b200 import make_poi2d, synth
    from oracle.oracle import Oracle2D
    ref, tar = synth.speckle_pair_2d(160, 144)
    xy = synth.grid_2d(30, 30, n_poi, 1, 2, 1)[:n_poi]
    q = make_poi2d(xy)
    o = Oracle2D(ref, tar, threads=1)
    o.fftcc2d(q, 10, 10)
    o.icgn2d1(q, 10, 10, 0.001, 10)
    owner = np.concatenate([np.full(obd.shard_bounds(n_poi, world, r)[1] - obd.shard_bounds(n_poi, world, r)[0], r) for r in range(world)])
    assert np.array_equal(got[:, 19], owner.astype(np.float32))
    got[:, 19] = 0
    assert np.array_equal(got, q)  # same records, same order as one process over the whole queue
