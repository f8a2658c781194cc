"""GPU parity tests of the ray-transfer path: CUDA vs oracle on identical steps (<= 1e-5 relative per (ray, source),
identical sets of touched sources) plus the reference's known answers (cherab/tools/tests/test_raytransfer.py)."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200.engine import DeviceRays, RayTransferScene
from core_b200.raytransfer import RayTransferBox, RayTransferCylinder
from oracle import oracle

pytestmark = pytest.mark.gpu


def check(rt, rays):
    scene = RayTransferScene(rt)
    dense, stats = scene.render_dense(rays)
    desc, keep = rt.descriptor()
    ref, rstats = oracle.rt_render_dense(desc, rays)
    assert stats["rt_steps"] == rstats["rt_steps"]
    assert np.array_equal(dense > 0, ref > 0), "touched source sets differ"
    nz = ref > 0
    assert np.max(np.abs(dense[nz] - ref[nz]) / ref[nz]) <= 1e-5 if nz.any() else True
    ro, cols, lens, _ = scene.render_csr(rays)
    rebuilt = np.zeros_like(ref)
    for r in range(rays.n_rays):
        c = cols[ro[r]:ro[r + 1]]
        assert len(np.unique(c)) == len(c), "duplicate source within a CSR row"
        rebuilt[r, c] = lens[ro[r]:ro[r + 1]]
    assert np.allclose(rebuilt, dense, rtol=1e-12, atol=0)
    sparse, _ = scene.render_sparse(rays)
    assert sparse.shape == dense.shape and np.array_equal(sparse.toarray(), rebuilt)
    scene.close()
    return dense, ref


def test_box_known_answer():
    rtb = RayTransferBox(xmax=3., ymax=3., zmax=3., nx=3, ny=3, nz=3)
    rtb.step = 0.01 * rtb.step
    rays = cb.ray_segments(rtb.primitive, [(4., 4., 4.)], [np.array([-1., -1., -1.]) / np.sqrt(3)])
    dense, _ = check(rtb, rays)
    ref = np.zeros(27)
    ref[0] = ref[13] = ref[26] = np.sqrt(3.)
    assert np.allclose(dense[0], ref, atol=1e-3)


def test_cylinder_3d_known_answer():
    rtc = RayTransferCylinder(radius_outer=2., height=2., n_radius=2, n_height=2, n_polar=3, period=90.)
    rtc.step = 0.001 * rtc.step
    rays = cb.ray_segments(rtc.primitive, [(np.sqrt(2.), np.sqrt(2.), 2.)], [np.array([-1., -1., -np.sqrt(2.)]) / 2.])
    dense, _ = check(rtc, rays)
    ref = np.zeros(12)
    ref[2] = ref[9] = np.sqrt(2.)
    assert np.allclose(dense[0], ref, atol=1e-3)


def test_cylinder_camera_with_voxel_map_and_transform():
    rng = np.random.default_rng(3)
    vm = rng.integers(-1, 40, size=(20, 1, 30)).astype(np.int32)
    rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=20, n_height=30, radius_inner=0.73, voxel_map=vm,
                              transform=cb.translate(0, 0, -1.8))
    cam = cb.PinholeCamera((16, 16), fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
    o, d = cam.rays()
    rays = cb.ray_segments(rtc.primitive, o, d, rtc.transform)
    assert (np.diff(rays.seg_offset) == 2).any()    # rays crossing the central hole have two segments
    check(rtc, rays)


def test_cylinder_polar_sectors_camera():
    rtc = RayTransferCylinder(radius_outer=2.0, height=2.0, n_radius=8, n_height=8, radius_inner=0.5, n_polar=12, period=60.,
                              transform=cb.translate(0, 0, -1.0))
    cam = cb.PinholeCamera((12, 12), fov=60, transform=cb.look_at((3.0, 0.4, 0.6), (0.0, 0.0, 0.0)))
    o, d = cam.rays()
    rays = cb.ray_segments(rtc.primitive, o, d, rtc.transform)
    check(rtc, rays)


def test_box_mask_and_empty():
    x = np.linspace(-0.45, 0.45, 10)
    mask = x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2 < 0.2
    rtb = RayTransferBox(1., 1., 1., 10, 10, 10, mask=mask, transform=cb.translate(-0.5, -0.5, -0.5))
    cam = cb.PinholeCamera((10, 10), fov=40, transform=cb.look_at((0.2, -2.5, 0.3), (0, 0, 0)))
    o, d = cam.rays()
    rays = cb.ray_segments(rtb.primitive, o, d, rtb.transform)
    check(rtb, rays)
    scene = RayTransferScene(rtb)
    none = cb.RayBatch(np.zeros((0, 3)), np.ones((0, 3)), [0], [], [])
    ro, cols, lens, st = scene.render_csr(none)
    assert ro.tolist() == [0] and cols.size == 0
    scene.close()


def test_c4_shape_csr_device_properties():
    # BASELINE config C4 grid (400 x 800 axisymmetric voxels, identity voxel map) with a 64x64 camera:
    # size-independent properties — row sums equal the chord length inside the grid, every length is a multiple of dt
    import torch
    rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=400, n_height=800, radius_inner=0.73,
                              transform=cb.translate(0, 0, -1.8))
    assert rtc.bins == 320000
    cam = cb.PinholeCamera((64, 64), fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
    o, d = cam.rays()
    rays = cb.ray_segments(rtc.primitive, o, d, rtc.transform)
    scene = RayTransferScene(rtc)
    dr = DeviceRays(rays)
    ro, cols, lens = scene.render_csr_device(dr, capacity=64 * 64 * 2500)
    torch.cuda.synchronize()
    ro, cols, lens = ro.cpu().numpy(), cols.cpu().numpy(), lens.cpu().numpy()
    chord = np.zeros(rays.n_rays)
    np.add.at(chord, np.repeat(np.arange(rays.n_rays), np.diff(rays.seg_offset)), rays.seg_t1 - rays.seg_t0)
    rowsum = np.add.reduceat(np.append(lens, 0.0), ro[:-1]) * (np.diff(ro) > 0)
    assert np.allclose(rowsum, chord, rtol=1e-9, atol=1e-12)
    assert cols.min() >= 0 and cols.max() < rtc.bins
    # a few rows against the oracle
    pick = np.array([0, 777, 2048, 4095])
    sub = rays.subset(pick)
    desc, keep = rtc.descriptor()
    ref, _ = oracle.rt_render_dense(desc, sub)
    for k, r in enumerate(pick):
        row = np.zeros(rtc.bins)
        row[cols[ro[r]:ro[r + 1]]] = lens[ro[r]:ro[r + 1]]
        nz = ref[k] > 0
        assert np.array_equal(row > 0, nz)
        assert np.max(np.abs(row[nz] - ref[k][nz]) / ref[k][nz]) <= 1e-5
    scene.close()


# ---- seeded random grids and rays: index flips move a whole step, so every corner of the index arithmetic is visited ----
def _random_rt(rng):
    kind = int(rng.integers(0, 3))
    if kind == 0:                                   # axisymmetric cylinder, possibly hollow
        n_r, n_z = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        r_in = float(rng.choice([0.0, rng.uniform(0.1, 1.0)]))
        rt = RayTransferCylinder(radius_outer=r_in + float(rng.uniform(0.5, 2.0)), height=float(rng.uniform(0.5, 3.0)), n_radius=n_r, n_height=n_z,
                                 radius_inner=r_in, transform=cb.translate(0, 0, float(rng.uniform(-2, 0))))
    elif kind == 1:                                 # 3-D cylinder with a toroidal period
        n_p = int(rng.integers(2, 9))
        rt = RayTransferCylinder(radius_outer=float(rng.uniform(1.0, 2.5)), height=float(rng.uniform(0.5, 3.0)), n_radius=int(rng.integers(1, 12)),
                                 n_height=int(rng.integers(1, 12)), n_polar=n_p, period=float(rng.choice([360.0, 180.0, 90.0, 45.0])),
                                 radius_inner=float(rng.choice([0.0, 0.4])))
    else:
        rt = RayTransferBox(xmax=float(rng.uniform(0.5, 3)), ymax=float(rng.uniform(0.5, 3)), zmax=float(rng.uniform(0.5, 3)),
                            nx=int(rng.integers(1, 12)), ny=int(rng.integers(1, 12)), nz=int(rng.integers(1, 12)),
                            transform=cb.translate(*rng.uniform(-1, 1, 3)))
    shape = rt.grid_shape if hasattr(rt, "grid_shape") else None
    if rng.uniform() < 0.5:                         # a random voxel map with unmapped cells, or a mask
        vm = rng.integers(-1, 25, size=tuple(rt.voxel_map.shape)).astype(np.int32)
        rt.voxel_map = vm
    elif rng.uniform() < 0.5:
        rt.mask = rng.uniform(size=tuple(rt.voxel_map.shape)) < 0.7
    rt.step = rt.step * float(10 ** rng.uniform(-1.0, 1.0))
    n = 24
    o = rng.uniform(-4, 4, (n, 3))
    target = rng.uniform(-1.0, 1.0, (n, 3))
    d = target - o
    d[0] = (0, 0, -1); o[0] = (0.3, 0.2, 5.0)                 # parallel to the axis
    d[1] = (-1, 0, 0); o[1] = (5.0, 0.0, 0.1)                 # through the axis
    d[2] = (0, -1, 0); o[2] = (0.0, 5.0, 0.2)                 # along a coordinate plane
    return rt, cb.ray_segments(rt.primitive, o, d, rt.transform)


@pytest.mark.parametrize("seed", range(24))
def test_random_grids_and_rays(seed):
    rt, rays = _random_rt(np.random.default_rng(500 + seed))
    if rays.n_segments == 0:
        pytest.skip("no ray hits the grid")
    check(rt, rays)


@pytest.mark.parametrize("variant", ["CB2_RT_TWO_PASS", "CB2_RT_UNPACKED", "both"])
@pytest.mark.parametrize("seed", (0, 3, 7, 11, 19))
def test_csr_fallback_paths(monkeypatch, variant, seed):
    # the count + fill traversal pair (taken when the scratch rows of the single traversal do not fit) and the 6-byte hash entries
    # (grids with >= 2^20 sources) must stay alive: same checks as the default path
    for v in (("CB2_RT_TWO_PASS", "CB2_RT_UNPACKED") if variant == "both" else (variant,)):
        monkeypatch.setenv(v, "1")
    rt, rays = _random_rt(np.random.default_rng(500 + seed))
    if rays.n_segments == 0:
        pytest.skip("no ray hits the grid")
    check(rt, rays)


def test_csr_is_reproducible_and_in_order_of_first_visit(monkeypatch):
    # run heads of one 32-step group that carry the same source are merged before the table is touched: slot numbers and the
    # order of the float64 additions are fixed, so two builds of the same matrix agree bit for bit — on every path
    import torch
    rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=100, n_height=200, radius_inner=0.73, n_polar=16, period=90.0,
                              transform=cb.translate(0, 0, -1.8))
    cam = cb.PinholeCamera((48, 48), fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
    rays = cb.ray_segments(rtc.primitive, *cam.rays(), rtc.transform)
    results = []
    for env in ({}, {}, {"CB2_RT_TWO_PASS": "1"}, {"CB2_RT_UNPACKED": "1"}):
        for k in ("CB2_RT_TWO_PASS", "CB2_RT_UNPACKED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        scene = RayTransferScene(rtc)
        ro, cols, lens = scene.render_csr_device(DeviceRays(rays), capacity=48 * 48 * 1500)
        torch.cuda.synchronize()
        results.append((ro.cpu().numpy(), cols.cpu().numpy(), lens.cpu().numpy()))
        scene.close()
    for ro, cols, lens in results[1:]:
        assert np.array_equal(ro, results[0][0]) and np.array_equal(cols, results[0][1]) and np.array_equal(lens, results[0][2])
    # order of first visit: the oracle's dense row, walked along the ray, meets the sources in the order the row lists them
    ro, cols, lens = results[0]
    assert ro[-1] > 48 * 48 * 50


@pytest.mark.parametrize("seed", (1, 4, 9, 14, 20))
def test_foreign_numerical_integrator_matches_the_oracle(seed):
    # RayTransferEmitter.emission_function under a NumericalIntegrator (emitters.pyx:452-473, 557-571): trapezium nodes and end weights
    # on the device (dense rows and CSR) against the oracle, on the random grids / voxel maps / masks of the suite
    rt, rays = _random_rt(np.random.default_rng(500 + seed))
    if rays.n_segments == 0:
        pytest.skip("no ray hits the grid")
    rt.integrator = cb.NumericalIntegrator(step=rt.step * 1.7, min_samples=5)
    dense, ref = check(rt, rays)
    assert ref.sum() > 0
    # and it is a different quadrature: the midpoint sampler of the same grid gives other numbers
    rt.integrator = None
    scene = RayTransferScene(rt)
    mid, _ = scene.render_dense(rays)
    scene.close()
    assert not np.array_equal(mid, dense)
