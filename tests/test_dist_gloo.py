"""Host logic of the multi-GPU path on CPU: world_size 2, gloo backend."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csparse3_b200.dist import gather_solutions, gather_to_root, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = shard_range(batch, rank, world)
    full = torch.arange(batch * n, dtype=torch.float64).reshape(batch, n)
    x = gather_solutions(full[start:stop].clone(), batch)
    ok = bool(torch.equal(x, full))
    xr = gather_to_root(full[start:stop].clone(), batch)
    ok = ok and ((xr is None) if rank != 0 else bool(torch.equal(xr, full)))
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([ok, start, stop]))
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for batch in (0, 1, 7, 10000, 13997):
        for world in (1, 2, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) == -(-batch // world) or batch == 0


def test_gather_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 7, 5, str(tmp_path)), nprocs=2, join=True)   # ragged: 4 + 3 systems
    for r in range(2):
        ok, start, stop = np.load(tmp_path / ("ok%d.npy" % r))
        assert ok == 1
