"""Pin the oracle's ray-transfer sampler against cherab/tools/tests/test_raytransfer.py: voxel_map/mask semantics
(:33-89,126-164), the integration known answers (:109-118 cylinder 3-D, :166-175 box) and the 2-D cylinder case
(:91-107, whose right-hand side in the reference is a CSG ToroidalVoxelGrid = exact chord lengths per voxel)."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200.raytransfer import RayTransferBox, RayTransferCylinder, RayTransferPipeline2D
from oracle import oracle


def _render(rt, origin, direction):
    rays = cb.ray_segments(rt.primitive, [origin], [direction], rt.transform)
    desc, keep = rt.descriptor()
    out, stats = oracle.rt_render_dense(desc, rays)
    return out[0], stats


def test_mask_2d():
    rtc = RayTransferCylinder(radius_outer=8., height=10., n_radius=4, n_height=10, radius_inner=4.)
    mask = np.zeros((4, 10), dtype=bool)
    mask[:, 3:6] = True
    rtc.mask = mask[:, None, :]
    ref = -1 * np.ones((4, 10), dtype=np.int32)
    ref[:, 3:6] = np.arange(12, dtype=int).reshape((4, 3))
    assert np.all(ref == rtc.voxel_map[:, 0, :]) and rtc.bins == 12


def test_voxel_map_3d():
    rtc = RayTransferCylinder(radius_outer=8., height=10., n_radius=4, n_height=10, radius_inner=4., n_polar=10, period=10.)
    voxel_map = -1 * np.ones((4, 10, 10), dtype=np.int32)
    voxel_map[1, 3:5, 3:5] = 0
    voxel_map[2, 5:7, 5:7] = 7
    rtc.voxel_map = voxel_map
    assert rtc.bins == 8 and rtc.mask.sum() == 8
    inv = rtc.invert_voxel_map()
    assert np.all(np.array(inv[0]) == np.array((np.array([1, 1, 1, 1]), np.array([3, 3, 4, 4]), np.array([3, 4, 3, 4]))))


def test_bad_arguments():
    import pytest
    with pytest.raises(ValueError):
        RayTransferCylinder(radius_outer=8., height=10., n_radius=4, n_height=10, period=7.)
    rtb = RayTransferBox(xmax=1., ymax=1., zmax=1., nx=2, ny=2, nz=2)
    with pytest.raises(ValueError):
        rtb.step = 0
    with pytest.raises(ValueError):
        rtb.mask = np.ones((3, 3, 3))


def test_box_integration():
    rtb = RayTransferBox(xmax=3., ymax=3., zmax=3., nx=3, ny=3, nz=3)
    rtb.step = 0.01 * rtb.step
    got, stats = _render(rtb, (4., 4., 4.), np.array([-1., -1., -1.]) / np.sqrt(3))
    ref = np.zeros(rtb.bins)
    ref[0] = ref[13] = ref[26] = np.sqrt(3.)
    assert np.allclose(ref, got, atol=0.001)
    assert stats["rt_steps"] > 1000


def test_cylinder_integration_3d():
    rtc = RayTransferCylinder(radius_outer=2., height=2., n_radius=2, n_height=2, n_polar=3, period=90.)
    rtc.step = 0.001 * rtc.step
    got, _ = _render(rtc, (np.sqrt(2.), np.sqrt(2.), 2.), np.array([-1., -1., -np.sqrt(2.)]) / 2.)
    ref = np.zeros(rtc.bins)
    ref[2] = ref[9] = np.sqrt(2.)
    assert np.allclose(ref, got, atol=0.001)


def _exact_chords_2d(origin, direction, r_edges, z_edges):
    """Exact chord length of a ray in each (ir, iz) annular cell (what the CSG ToroidalVoxelGrid of the reference
    test measures): split the ray at every radial / axial boundary crossing and bin the pieces."""
    o, d = np.asarray(origin, float), np.asarray(direction, float)
    ts = [0.0, 20.0]
    a = d[0] ** 2 + d[1] ** 2
    b = 2 * (o[0] * d[0] + o[1] * d[1])
    for r in r_edges:
        c = o[0] ** 2 + o[1] ** 2 - r * r
        disc = b * b - 4 * a * c
        if disc > 0:
            ts += [(-b - np.sqrt(disc)) / (2 * a), (-b + np.sqrt(disc)) / (2 * a)]
    for z in z_edges:
        ts.append((z - o[2]) / d[2])
    ts = np.sort([t for t in ts if 0 <= t <= 20])
    out = np.zeros((len(r_edges) - 1, len(z_edges) - 1))
    for t0, t1 in zip(ts[:-1], ts[1:]):
        p = o + 0.5 * (t0 + t1) * d
        r = np.hypot(p[0], p[1])
        ir = np.searchsorted(r_edges, r) - 1
        iz = np.searchsorted(z_edges, p[2]) - 1
        if 0 <= ir < out.shape[0] and 0 <= iz < out.shape[1]:
            out[ir, iz] += t1 - t0
    return out


def test_cylinder_integration_2d():
    rtc = RayTransferCylinder(radius_outer=4., height=2., n_radius=2, n_height=2, radius_inner=2.)
    rtc.step = 0.001 * rtc.step
    origin, direction = (4., 1., 2.), np.array([-4., -1., -2.]) / np.sqrt(21.)
    got, _ = _render(rtc, origin, direction)
    ref = _exact_chords_2d(origin, direction, [2., 3., 4.], [0., 1., 2.]).reshape(-1)   # source = ir * n_z + iz
    assert np.allclose(ref, got, atol=0.001)
    assert got.sum() > 1.0


def test_transform_and_pipeline_2d():
    rtb = RayTransferBox(xmax=1., ymax=1., zmax=1., nx=4, ny=4, nz=4, transform=cb.translate(-0.5, -0.5, -0.5))
    cam = cb.PinholeCamera((6, 5), fov=30, transform=cb.look_at((0, -3, 0.1), (0, 0, 0)))
    o, d = cam.rays()
    rays = cb.ray_segments(rtb.primitive, o, d, rtb.transform)
    desc, keep = rtb.descriptor()
    out, _ = oracle.rt_render_dense(desc, rays)
    pipe = RayTransferPipeline2D(kind="radiance")
    pipe.initialise((6, 5), 1, 600., 601., rtb.bins, 1, True)
    for k in range(30):
        pipe.update(k // 5, k % 5, 0, (out[k], 0))
    pipe.finalise()
    assert pipe.matrix.shape == (6, 5, 64)
    # total path length per ray == chord through the unit box (midpoint sampling sums dt over the whole chord)
    chord = np.zeros(30)
    for r in range(30):
        s = slice(rays.seg_offset[r], rays.seg_offset[r + 1])
        chord[r] = (rays.seg_t1[s] - rays.seg_t0[s]).sum()
    assert np.allclose(pipe.matrix.reshape(30, -1).sum(axis=1), chord, rtol=1e-9, atol=1e-12)
    import pytest
    with pytest.raises(ValueError):
        RayTransferPipeline2D(kind="flux")


# ---- pipelines and pixel processors: cherab/tools/tests/test_raytransfer.py:245-393 ----
def test_pipelines_initialise_and_kind():
    from core_b200.raytransfer import RayTransferPipeline0D, RayTransferPipeline1D, Spectrum
    nbins, pixels, samples, sensitivity, value = 10, 20, 1, 2.0, 1.0
    spectrum = Spectrum(1.0, 2.0, nbins)
    spectrum.samples[:] = value
    p0 = RayTransferPipeline0D("test_pipeline_0D", kind="power")
    p0.initialise(0, 0, nbins, 0, 0)
    assert p0.matrix.shape == (nbins,) and p0.name == "test_pipeline_0D" and p0.kind == "power"
    with pytest.raises(ValueError):
        RayTransferPipeline0D("test_pipeline_0D", "blah")
    p1 = RayTransferPipeline1D("test_pipeline_1D", kind="radiance")
    p1.initialise(pixels, samples, 0, 0, nbins, 1, 0)
    assert p1.matrix.shape == (pixels, nbins) and p1.kind == "radiance" and p1._samples == samples
    p2 = RayTransferPipeline2D("test_pipeline_2D", kind="power")
    p2.initialise((pixels, pixels), samples, 0, 0, nbins, 1, 0)
    assert p2.matrix.shape == (pixels, pixels, nbins)
    for pipe, args in ((p0, (0,)), (p1, (0, 0)), (p2, (0, 0, 0))):
        pipe.kind = "power"
        proc = pipe.pixel_processor(*args)
        proc.add_sample(spectrum, sensitivity)
        assert np.all(proc.pack_results()[0] == sensitivity * value)        # multiplied by the sensitivity
        pipe.kind = "radiance"
        proc = pipe.pixel_processor(*args)
        proc.add_sample(spectrum, sensitivity)
        assert np.all(proc.pack_results()[0] == value)                      # not multiplied
    # 0-D accumulates over update() calls and normalises at finalise (pipelines.py:111-119)
    p0.initialise(0, 0, nbins, 0, 0)
    p0.update(0, (np.full(nbins, 3.0), 0), 2)
    p0.update(0, (np.full(nbins, 5.0), 0), 2)
    p0.finalise()
    assert np.all(p0.matrix == 2.0)


def test_foreign_numerical_integrator_over_the_emitter():
    # emitters.pyx:452-473, 557-571: RayTransferEmitter.emission_function puts unit emissivity into the sample's cell, so a
    # NumericalIntegrator [raysect] over it gives the trapezium-rule path lengths: along the x axis of a 3 x 3 x 3 box of unit cells,
    # step 0.25 -> 12 intervals, nodes at 0, 0.25 .. 3.0; the node on a cell face belongs to the upper cell (truncation), the end
    # nodes weigh h / 2
    rtb = RayTransferBox(xmax=3.0, ymax=3.0, zmax=3.0, nx=3, ny=3, nz=3)
    rtb.integrator = cb.NumericalIntegrator(step=0.25, min_samples=2)
    rays = cb.ray_segments(rtb.primitive, np.array([[-1.0, 0.5, 0.5]]), np.array([[1.0, 0.0, 0.0]]))
    desc, keep = rtb.descriptor()
    assert desc.integrator == 1 and desc.step == 0.25
    row, st = oracle.rt_render_dense(desc, rays)
    assert st["rt_steps"] == 13
    h = (rays.seg_t1[0] - rays.seg_t0[0]) / 12
    # marching from the far end (x = 3 - eps) to the near end: cell 2 holds nodes 0..3 (+ half of node 0), cell 1 nodes 4..7 ...
    got = row[0][[rtb.voxel_map[i, 0, 0] for i in range(3)]]
    assert abs(got.sum() - 12 * h) < 1e-12
    np.testing.assert_allclose(sorted(got), sorted([3.5 * h, 4 * h, 4.5 * h]), rtol=1e-9)
    # the emitter's own integrator on the same ray: midpoints of 29 steps (the eps-shrunk chord is a hair under 3 m): 9 or 10 per cell
    rtb.integrator = None
    desc, keep = rtb.descriptor()
    row2, _ = oracle.rt_render_dense(desc, rays)
    np.testing.assert_allclose(row2[0][[rtb.voxel_map[i, 0, 0] for i in range(3)]], 1.0, atol=0.11)
    with pytest.raises(TypeError):
        rtb.integrator = "trapezium"
        rtb.descriptor()


def test_row_sums_are_chord_lengths_under_both_integrators():
    # size-independent property: with every cell mapped, a ray's path lengths add up to its chord through the grid — exactly (to
    # rounding) under the trapezium rule (weights h/2, h, ..., h/2 sum to L) and under the midpoint rule (n dt = L)
    rng = np.random.default_rng(42)
    rtc = RayTransferCylinder(radius_outer=2.0, height=3.0, n_radius=13, n_height=17, radius_inner=0.5, n_polar=5, period=72.0,
                              transform=cb.translate(0, 0, -1.5))
    o = rng.uniform(-4, 4, (40, 3))
    d = rng.uniform(-1, 1, (40, 3)) - o
    rays = cb.ray_segments(rtc.primitive, o, d, rtc.transform)
    assert rays.n_segments > 10
    chord = np.zeros(rays.n_rays)
    np.add.at(chord, np.repeat(np.arange(rays.n_rays), np.diff(rays.seg_offset)), rays.seg_t1 - rays.seg_t0)
    for integrator in (None, cb.NumericalIntegrator(step=0.013, min_samples=7)):
        rtc.integrator = integrator
        desc, keep = rtc.descriptor()
        rows, _ = oracle.rt_render_dense(desc, rays)
        np.testing.assert_allclose(rows.sum(axis=1), chord, rtol=1e-12, atol=1e-12)
