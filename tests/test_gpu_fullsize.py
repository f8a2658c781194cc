"""BASELINE config C3 at its FULL size (1024 x 1024 rays, 2048 bins, 8 lines + Bremsstrahlung) on the CUDA path, checked through
properties that do not need the oracle to finish a 4e9-sample frame: the exact sample count, linearity under scale / accumulate,
invariance of every ray's spectrum to how the frame is partitioned, and the oracle on a seeded handful of the frame's rays."""
import os
import sys

import numpy as np
import pytest

import core_b200 as cb
from core_b200.engine import EmissionScene
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the scene and camera of the benchmark)

pytestmark = pytest.mark.gpu
PIXELS, BINS = 1024, 2048


def test_c3_full_frame_properties():
    import torch
    plasma, flat = bench.build_scene(BINS)
    cam = cb.PinholeCamera((PIXELS, PIXELS), fov=45, transform=cb.look_at(bench.CAMERA_POS, bench.CAMERA_TARGET))
    sx, sy = cb.stratified_offsets(4)[5]
    pin = cb.DevicePinhole(cam, plasma.geometry, to_world=plasma.geometry_to_world())
    rays = pin.rays(sx, sy)
    scene = EmissionScene(flat)
    dev = pin.device
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    frame = torch.zeros((PIXELS * PIXELS, BINS), dtype=torch.float32, device=dev)
    scene.render_device(rays, frame, scale=1.0, accumulate=False, stats=stats)
    torch.cuda.synchronize()

    # (1) the trapezium marcher evaluates intervals + 1 samples per chord, intervals = max(min_samples - 1, ceil(L / step)):
    #     exact integer agreement with the count made from the chords themselves
    length = (rays.seg_t1[:rays.n_segments] - rays.seg_t0[:rays.n_segments]).cpu().numpy()
    d = flat.desc
    intervals = np.maximum(d.min_samples - 1, np.ceil(length / d.step)).astype(np.int64)
    assert int(stats[0].item()) == int((intervals + 1).sum()) > 3_000_000_000
    assert bool(torch.isfinite(frame).all()) and float(frame.min()) >= 0.0 and float(frame.max()) > 0.0

    # (2) linearity: two accumulated passes == one pass at scale 2, up to the float32 rounding of the extra additions (the line
    #     and the continuum parts are added to the frame by two kernels, so the association differs)
    twice = torch.zeros_like(frame)
    scene.render_device(rays, twice, scale=2.0, accumulate=False)
    frame2 = frame.clone()
    scene.render_device(rays, frame2, scale=1.0, accumulate=True)
    torch.cuda.synchronize()
    err = (frame2 - twice).abs_()
    assert bool((err <= 4e-7 * twice.abs() + 1e-30).all())
    del frame2, twice, err

    # (3) a ray's spectrum does not depend on which other rays share its launch: the 16 x 16-pixel tiles of rank 1 of 3,
    #     rendered alone, reproduce their rows of the full frame bit for bit
    idx = bench.rank_pixels(PIXELS, 1, 3)
    part_rays = cb.DevicePinhole(cam, plasma.geometry, to_world=plasma.geometry_to_world(), pixel_index=idx).rays(sx, sy)
    part = torch.zeros((idx.size, BINS), dtype=torch.float32, device=dev)
    scene.render_device(part_rays, part, scale=1.0, accumulate=False)
    torch.cuda.synchronize()
    assert torch.equal(part, frame[torch.as_tensor(idx, device=dev)])

    # (4) the oracle on a seeded handful of the frame's rays (the acceptance rule of SURVEY 8(d))
    pick = np.sort(np.random.default_rng(7).choice(PIXELS * PIXELS, size=6, replace=False))
    host_rays = bench.make_rays(plasma, PIXELS, pick, 5)
    ref, _ = oracle.emission_render(flat, host_rays)
    got = frame[torch.as_tensor(pick, device=dev)].double().cpu().numpy()
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
    assert ref.max() > 0 and np.all(np.abs(got - ref) <= tol + 6e-8 * np.abs(ref))       # + fp32 storage of the frame
    scene.close()
