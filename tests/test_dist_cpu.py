"""world_size-2 gloo tests of the multi-GPU host logic (frame sharding, reference-frame broadcast)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kvazaar_b200 import dist as kd


def test_shard_frames_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in kd.shard_frames(37, r, world))
        assert seen == list(range(37))
        assert all(kd.owner_of(i, world) == r for r in range(world) for i in kd.shard_frames(37, r, world))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = 5
        mine = kd.shard_frames(frames, rank, world)
        ok = True
        for f in range(frames):
            buf = torch.full((64 * 64 * 3 // 2,), f + 1, dtype=torch.uint8) if f in mine else torch.zeros(64 * 64 * 3 // 2, dtype=torch.uint8)
            kd.broadcast_reference_frame(buf, f, world)
            ok &= bool((buf == f + 1).all())
        sizes = kd.gather_result_sizes(100 + rank)
        ok &= sizes == [100 + r for r in range(world)]
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_broadcast_reference_frame_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
    assert np.all([out[0], out[1]])
