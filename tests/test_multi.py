"""N > 1 path on CPU (gloo, world_size 2): every rank renders its interleaved column bands with the
product's shard mapping (host-compiled integrator, tests/native/hostcheck.cu), the per-tile sample
sums are combined with ONE sum collective at tonemap time, and the result equals the single-rank
image bit for bit (each pixel is non-zero on exactly one rank, so the fp32 sum is exact)."""
import os
import socket
import sys

import numpy as np
import pytest

import common

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from raytracingpbr_b200 import scenes
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, objs, cam, _ = scenes.cornell_box_shortest(72, 24, max_bounces=8, seed=3)
    part = common.hostcheck_pathtrace(cfg, cam, objs, 3, rank=rank, nranks=world, band=8)
    own = ((np.arange(72) // 8) % world) == rank
    assert (part[~own] == 0).all() and (part[own][..., 3] == 3).all()
    t = torch.from_numpy(part)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)          # the NCCL tile reduce of rtpbr_reduce_tiles, on gloo
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_sharding_and_tile_reduce(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    from raytracingpbr_b200 import scenes
    cfg, objs, cam, _ = scenes.cornell_box_shortest(72, 24, max_bounces=8, seed=3)
    full = common.hostcheck_pathtrace(cfg, cam, objs, 3)
    got = np.load(os.path.join(str(tmp_path), "reduced.npy"))
    assert np.array_equal(got, full)
