"""CPU, world_size 2, gloo: the N>1 host logic (pair sharding, arena broadcast, pose gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vista_slam_b200.dist import pair_shard


def test_pair_shard_partitions_exactly():
    for total in (0, 1, 7, 8, 16, 255, 256):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = pair_shard(total, r, world)
                assert 0 <= lo <= hi <= total
                cover += list(range(lo, hi))
            assert cover == list(range(total))
            sizes = [pair_shard(total, r, world)[1] - pair_shard(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vista_slam_b200.dist import broadcast_arena, gather_poses, pair_shard as ps
    # weight arena: rank 0 holds the packed bytes, the others receive them
    g = torch.Generator().manual_seed(7)
    ref = torch.randint(0, 256, (4097,), dtype=torch.uint8, generator=g)
    arena = ref.clone() if rank == 0 else torch.zeros_like(ref)
    broadcast_arena(arena, src=0)
    ok_arena = bool(torch.equal(arena, ref))
    # every rank "runs" its contiguous shard of a global pair batch and the poses are gathered in global order
    lo, hi = ps(total, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float32)
    pose = torch.eye(4).repeat(hi - lo, 1, 1)
    pose[:, 0, 3] = idx
    conf = idx / 100.0
    gp, gc = gather_poses(pose, conf, total)
    ok_gather = gp.shape == (total, 4, 4) and bool(torch.equal(gp[:, 0, 3], torch.arange(total, dtype=torch.float32)))
    ok_gather = ok_gather and bool(torch.allclose(gc, torch.arange(total, dtype=torch.float32) / 100.0))
    ret[rank] = (ok_arena, ok_gather)
    dist.destroy_process_group()


def test_world2_gloo_broadcast_and_gather():
    world, total = 2, 7  # ragged: 4 + 3 pairs
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    assert dict(ret) == {0: (True, True), 1: (True, True)}
