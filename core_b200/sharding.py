"""Image-tile sharding of a camera frame over ranks (one process per GPU).

Rays are independent: 16x16-pixel tiles are dealt round-robin to the ranks, scene tables are replicated and
there is no data-path collective.  The only collective is the gather of the finished tiles to rank 0
(`gather_frame`, NCCL on GPUs / gloo in the CPU tests).
"""
import numpy as np

TILE = 16


def tile_pixels(pixels, rank, world, tile=TILE):
    """Pixel indices (ix * ny + iy) of the tiles dealt round-robin to `rank`, tile-major order."""
    nx, ny = (pixels, pixels) if np.isscalar(pixels) else pixels
    tx, ty = (nx + tile - 1) // tile, (ny + tile - 1) // tile
    tiles = np.arange(tx * ty)[rank::world]
    ox, oy = np.meshgrid(np.arange(tile), np.arange(tile), indexing="ij")
    ix = (tiles // ty)[:, None] * tile + ox.ravel()[None, :]
    iy = (tiles % ty)[:, None] * tile + oy.ravel()[None, :]
    ok = (ix < nx) & (iy < ny)
    return (ix * ny + iy)[ok]


def assemble_frame(parts, pixels, tile=TILE):
    """Inverse of the sharding: parts[r] = rows [n_pixels_r, bins] of rank r -> frame [nx, ny, bins]."""
    nx, ny = (pixels, pixels) if np.isscalar(pixels) else pixels
    world = len(parts)
    bins = parts[0].shape[1]
    frame = np.zeros((nx * ny, bins), dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        frame[tile_pixels((nx, ny), r, world, tile)] = p
    return frame.reshape(nx, ny, bins)


def gather_frame(local_rows, pixels, dst=0):
    """Gather every rank's rows to `dst` with torch.distributed (the path's only collective); returns the assembled
    frame on `dst` (as a torch tensor on the rows' device) and None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [tile_pixels(pixels, r, world).size for r in range(world)]
    bins = local_rows.shape[1]
    bufs = [torch.empty((c, bins), dtype=local_rows.dtype, device=local_rows.device) for c in counts] if rank == dst else None
    dist.gather(local_rows.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    nx, ny = (pixels, pixels) if np.isscalar(pixels) else pixels
    frame = torch.empty((nx * ny, bins), dtype=local_rows.dtype, device=local_rows.device)
    for r, b in enumerate(bufs):
        frame[torch.from_numpy(tile_pixels((nx, ny), r, world)).to(local_rows.device)] = b
    return frame.reshape(nx, ny, bins)


def open_shared_frame(name, rows, bins, create):
    """The [rows, bins] float32 frame in shared host memory (/dev/shm/<name>) every rank of a node writes its tiles into — the
    image-ordered frame of the N > 1 end-to-end path.  ``create``: allocate the file (one rank, before the others open it;
    posix_fallocate fails here, not at the first touch, when /dev/shm is too small).  Returns (numpy frame, mmap object)."""
    import mmap
    import os
    path = os.path.join("/dev/shm", name)
    nbytes = int(rows) * int(bins) * 4
    if create:
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o600)
        try:
            os.posix_fallocate(fd, 0, nbytes)
        except OSError:
            os.close(fd)
            os.unlink(path)
            raise
        os.close(fd)
    fd = os.open(path, os.O_RDWR)
    try:
        mm = mmap.mmap(fd, nbytes)
    finally:
        os.close(fd)
    return np.frombuffer(mm, dtype=np.float32).reshape(int(rows), int(bins)), mm


def unlink_shared_frame(name):
    """Remove the file; existing mappings keep the memory alive."""
    import os
    try:
        os.unlink(os.path.join("/dev/shm", name))
    except OSError:
        pass
