"""Host logic of the image-batch sharding (balf_b200/sharding.py) on CPU: world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from balf_b200 import sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    xy = torch.randint(0, 640, (5, 17, 2), generator=g, dtype=torch.int32)
    sc = torch.rand(5, 17, generator=g)
    cnt = torch.randint(0, 18, (5,), generator=g, dtype=torch.int32)
    a, b, c = sharding.unpack_records(sharding.pack_records(xy, sc, cnt))
    assert torch.equal(a, xy) and torch.equal(b, sc) and torch.equal(c, cnt)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total, k = 6, 9
        g = torch.Generator().manual_seed(42)                     # every rank builds the same global batch
        xy = torch.randint(0, 640, (total, k, 2), generator=g, dtype=torch.int32)
        sc = torch.rand(total, k, generator=g)
        cnt = torch.randint(0, k + 1, (total,), generator=g, dtype=torch.int32)
        s, e = sharding.shard_range(total, rank, world)
        gx, gs, gc = sharding.gather_keypoints(xy[s:e].contiguous(), sc[s:e].contiguous(), cnt[s:e].contiguous())
        ok = torch.equal(gx, xy) and torch.equal(gs, sc) and torch.equal(gc, cnt)
        ids = torch.arange(rank * 100, rank * 100 + 2 * 4 * 2, dtype=torch.int32).reshape(2, 4, 2)
        gi, gn = sharding.gather_matches(ids, torch.tensor([3, rank], dtype=torch.int32))
        ok = ok and gi.shape == (2 * world, 4, 2) and gn.tolist() == [3, 0, 3, 1] and int(gi[2, 0, 0]) == 100
        np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([int(ok)]))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int(np.load(tmp_path / ("ok%d.npy" % r))[0]) == 1


def test_gather_is_identity_without_process_group():
    xy = torch.zeros(2, 3, 2, dtype=torch.int32)
    out = sharding.gather_keypoints(xy, torch.zeros(2, 3), torch.zeros(2, dtype=torch.int32))
    assert out[0] is xy
