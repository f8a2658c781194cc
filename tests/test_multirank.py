"""N > 1 host logic on CPU: world_size-2 gloo run of the tile sharding + frame gather (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from core_b200.sharding import assemble_frame, gather_frame, tile_pixels


def test_tiles_partition_the_frame():
    for pixels, world in (((40, 24), 3), ((128, 128), 8), ((17, 33), 2)):
        parts = [tile_pixels(pixels, r, world) for r in range(world)]
        allpix = np.sort(np.concatenate(parts))
        assert np.array_equal(allpix, np.arange(pixels[0] * pixels[1]))
        sizes = [p.size for p in parts]
        assert max(sizes) - min(sizes) <= 16 * 16 * (1 + (pixels[1] + 15) // 16)
    rows = [np.repeat(tile_pixels((40, 24), r, 3)[:, None], 5, axis=1).astype(np.float32) for r in range(3)]
    frame = assemble_frame(rows, (40, 24))
    assert np.array_equal(frame[:, :, 0].ravel(), np.arange(960, dtype=np.float32))


def _worker(rank, world, port, pixels, bins, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pix = tile_pixels(pixels, rank, world)
    rows = torch.from_numpy(pix.astype(np.float32))[:, None] * torch.arange(1, bins + 1, dtype=torch.float32)[None, :]
    frame = gather_frame(rows, pixels, dst=0)
    if rank == 0:
        expect = torch.arange(pixels[0] * pixels[1], dtype=torch.float32)[:, None] * torch.arange(1, bins + 1, dtype=torch.float32)[None, :]
        out.put(bool(torch.equal(frame.reshape(-1, bins), expect)))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, (48, 32), 7, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _shared_worker(rank, world, port, pixels, bins, out):
    # the N > 1 end-to-end layout: every rank writes its tiles to their pixels of ONE image-ordered frame in shared host memory
    from core_b200.sharding import open_shared_frame, unlink_shared_frame
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name = "cb2_test_frame_%d" % port
    n = pixels[0] * pixels[1]
    if rank == 0:
        open_shared_frame(name, n, bins, create=True)
    dist.barrier()
    frame, mm = open_shared_frame(name, n, bins, create=False)
    dist.barrier()
    if rank == 0:
        unlink_shared_frame(name)
    pix = tile_pixels(pixels, rank, world)                       # destination rows of this rank's rays (cb2_emission_render_rows)
    frame[pix] = pix[:, None].astype(np.float32) * np.arange(1, bins + 1, dtype=np.float32)[None, :]
    dist.barrier()
    if rank == 0:
        expect = np.arange(n, dtype=np.float32)[:, None] * np.arange(1, bins + 1, dtype=np.float32)[None, :]
        out.put(bool(np.array_equal(frame, expect)) and not os.path.exists("/dev/shm/" + name))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_image_ordered_shared_frame():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_shared_worker, args=(r, 2, port, (40, 56), 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
