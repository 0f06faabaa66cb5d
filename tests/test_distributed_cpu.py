"""world_size-2 gloo test of the multi-GPU host logic on CPU: the per-rank sample buffers, gathered
rank-major as ncclAllGather lays them out, de-interleave to exactly one value per pixel using the
reference's StaticWorkDistribution (host ABI functions), independent of the number of ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pixel_value(x, y):
    return np.float32(x * 1000 + y)


def _worker(rank, world, port, w, h, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from optix_raytracer_b200 import host
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = host.wd_num_samples(w, h, world)
    local = np.zeros((n, 4), np.float32)
    for s in range(n):
        x, y = host.wd_sample_pixel(w, h, world, rank, s)
        if x < w and y < h:
            local[s] = (_pixel_value(x, y), rank, s, 1.0)
    gathered = [torch.zeros((n, 4)) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(local))
    if rank == 0:
        img = np.full((h, w, 4), -1, np.float32)
        cover = np.zeros((h, w), np.int32)
        for g in range(world):
            buf = gathered[g].numpy()
            for s in range(n):
                x, y = host.wd_sample_pixel(w, h, world, g, s)
                if x < w and y < h:
                    img[y, x] = buf[s]
                    cover[y, x] += 1
        q.put((img, cover))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(64, 24), (100, 37)])
def test_gather_and_deinterleave_two_ranks(w, h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, q)) for r in range(2)]
    for p in procs:
        p.start()
    img, cover = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.all(cover == 1), "every pixel must be owned by exactly one (rank, sample)"
    ys, xs = np.mgrid[0:h, 0:w]
    assert np.array_equal(img[..., 0], (xs * 1000 + ys).astype(np.float32))
    # ownership follows the 8-wide column interleave with per-strip-row rotation
    owner = img[..., 1].astype(int)
    assert np.array_equal(owner, ((xs // 8) + (2 - (ys // 4) % 2)) % 2)
