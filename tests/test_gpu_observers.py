"""Device-side observer front-end (cb2_pinhole_rays_device, core_b200/observers.py) against the host mirror of the same
geometry and, for whole frames, against the oracle on the host-generated rays."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200 import generomak
from core_b200.engine import EmissionScene, RayTransferScene
from oracle import oracle

pytestmark = pytest.mark.gpu


def _compare_rays(dev_batch, host_batch):
    assert np.array_equal(dev_batch.seg_offset, host_batch.seg_offset)
    np.testing.assert_allclose(dev_batch.origin, host_batch.origin, rtol=0, atol=1e-15)
    np.testing.assert_allclose(dev_batch.direction, host_batch.direction, rtol=0, atol=4e-16)
    np.testing.assert_allclose(dev_batch.seg_t0, host_batch.seg_t0, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dev_batch.seg_t1, host_batch.seg_t1, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("case", ["hollow", "solid", "sphere", "box", "inside"])
def test_pinhole_rays_and_chords_match_the_host_mirror(case):
    cam = cb.PinholeCamera((37, 29), fov=52.0, transform=cb.look_at((3.6, 0.4, 1.3), (0.2, -0.1, -0.2)))
    to_world = None
    if case == "hollow":
        prim = cb.HollowCylinder(0.73, 2.41, -1.8, 1.55)          # rays see 0, 1 or 2 chords
    elif case == "solid":
        prim = cb.HollowCylinder(0.0, 1.2, 0.0, 2.0, transform=cb.translate(0.1, 0.2, -1.0))
    elif case == "sphere":
        prim = cb.Sphere(1.25, transform=cb.translate(0.3, 0.0, 0.1))
    elif case == "box":
        prim = cb.Box((-1.0, -0.5, -0.7), (0.8, 0.9, 0.6))
        to_world = cb.look_at((0.2, 0.1, 0.0), (1.0, 1.0, 0.3))     # a rotated box
    else:
        prim = cb.HollowCylinder(0.73, 2.41, -1.8, 1.55)
        cam = cb.PinholeCamera((16, 16), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))   # camera inside the vessel
    pin = cb.DevicePinhole(cam, prim, to_world=to_world)
    for sx, sy in ((0.5, 0.5), (0.125, 0.875)):
        dev = pin.rays(sx, sy).to_host()
        host = cb.ray_segments(prim, *cam.rays(sx, sy), to_world)
        assert host.n_segments > 0
        _compare_rays(dev, host)
    # a pixel subset (what a rank's tiles are)
    idx = np.arange(0, cam.pixels[0] * cam.pixels[1], 7)
    dev = cb.DevicePinhole(cam, prim, to_world=to_world, pixel_index=idx).rays().to_host()
    _compare_rays(dev, cb.ray_segments(prim, *cam.rays(pixel_index=idx), to_world))


def test_observe_frame_mean_over_pixel_samples():
    # the frame loop on the device (4 stratified samples per pixel) against the oracle on the host-generated rays
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.01)
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 128)
    cam = cb.PinholeCamera((6, 5), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    pin = cb.DevicePinhole(cam, plasma.geometry, to_world=plasma.geometry_to_world())
    scene = EmissionScene(flat)
    import torch
    spectral, total = cb.SpectralRadiancePipeline2D(), cb.RadiancePipeline2D()
    frame = cb.observe(scene, pin, pixel_samples_side=2, dtype=torch.float64, pipelines=[spectral, total]).cpu().numpy()
    scene.close()
    assert spectral.frame.mean.shape == (6, 5, 128) and spectral.frame.samples == 4 and spectral.bins == 128
    assert np.array_equal(spectral.frame.mean.reshape(30, 128), frame)
    np.testing.assert_allclose(total.frame.mean, frame.reshape(6, 5, 128).sum(axis=2) * (10.0 / 128), rtol=1e-12)
    ref = np.zeros_like(frame)
    for sx, sy in cb.stratified_offsets(2):
        rays = cb.ray_segments(plasma.geometry, *cam.rays(sx, sy), plasma.geometry_to_world())
        ref += oracle.emission_render(flat, rays)[0] / 4.0
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
    assert ref.max() > 0 and np.all(np.abs(frame - ref) <= tol)


def test_device_rays_feed_the_ray_transfer_kernel():
    import torch
    from core_b200.raytransfer import RayTransferCylinder
    rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=20, n_height=30, radius_inner=0.73, transform=cb.translate(0, 0, -1.8))
    prim = cb.HollowCylinder(0.73, 2.41, -1.8, 1.55)
    cam = cb.PinholeCamera((12, 12), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    pin = cb.DevicePinhole(cam, prim)
    scene = RayTransferScene(rtc)
    ro, co, le = scene.render_csr_device(pin.rays(), capacity=144 * 64)
    h_ro, h_co, h_le, _ = scene.render_csr(cb.ray_segments(prim, *cam.rays()))
    scene.close()
    assert np.array_equal(ro.cpu().numpy(), h_ro) and np.array_equal(co.cpu().numpy(), h_co)
    np.testing.assert_allclose(le.cpu().numpy(), h_le, rtol=1e-9)


def test_fibre_group_observes_in_one_render():
    # demos/observers/groups.py:57-76 shape: five fibres at the pinhole position looking into the divertor, one render call
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.005)
    flat = cb.flatten_scene(plasma, 655.5, 656.9, 256)
    group = cb.FibreOpticGroup(name="Divertor Fibre Optic Array")
    for i, ang in enumerate((-63.8, -66.5, -69.2, -71.9, -74.6)):
        target = (2.3 - np.cos(np.deg2rad(ang)), 0.0, 1.25 + np.sin(np.deg2rad(ang)))
        group.add_observer(cb.FibreOptic(name=str(i + 1), transform=cb.look_at((2.3, 0.0, 1.25), target, up=(0, 1, 0))))
    group.acceptance_angle, group.radius, group.pixel_samples = 1.4, 0.001, 24
    scene = EmissionScene(flat)
    spectra = group.observe(scene, plasma.geometry, plasma.geometry_to_world())
    scene.close()
    o, d, w, owner = group.gather_rays()
    ref_rays, _ = oracle.emission_render(flat, cb.ray_segments(plasma.geometry, o, d, plasma.geometry_to_world()))
    for i in range(5):
        sel = owner == i
        ref = (ref_rays[sel] * w[sel, None]).sum(axis=0) / w[sel].sum()
        assert ref.max() > 0
        assert np.all(np.abs(spectra[i] - ref) <= 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max())
        assert group.observers[i].spectrum is spectra[i] or np.array_equal(group.observers[i].spectrum, spectra[i])
        ob = group.observers[i]
        power = (ref_rays[sel] * w[sel, None]).mean(axis=0) * ob.solid_angle * ob.collection_area      # etendue-weighted: W / nm
        assert np.all(np.abs(group.power_spectra[i] - power) <= 1e-4 * np.abs(power) + 1e-9 * np.abs(power).max())


def _fibre_fan(n_fibres=7, samples=24):
    group = cb.FibreOpticGroup(name="fan", transform=cb.translate(0.0, 0.02, -0.01))
    for i, ang in enumerate(np.linspace(-60.0, -80.0, n_fibres)):
        target = (2.3 - np.cos(np.deg2rad(ang)), 0.05 * i, 1.25 + np.sin(np.deg2rad(ang)))
        group.add_observer(cb.FibreOptic(name=str(i), transform=cb.look_at((2.3, 0.0, 1.25), target, up=(0, 1, 0)),
                                         acceptance_angle=1.0 + 0.3 * i, radius=0.001 * (1 + i), pixel_samples=samples + i))
    return group


def test_group_rays_on_the_device_match_the_host_bundles():
    # cb2_observer0d_rays_device against FibreOptic.rays / SightLine.rays + geometry.ray_segments (cos / sin of the device's libm: 1e-14)
    import ctypes as C
    import torch
    from core_b200 import _abi
    from core_b200.observers import DeviceRayBuffer, _primitive_desc
    prim = cb.HollowCylinder(0.73, 2.41, -1.8, 1.55)
    sight = cb.SightLineGroup([cb.SightLine(transform=cb.look_at((2.3, 0.1 * k, 1.25), (1.0, 0.5, -0.5 + 0.1 * k)), sensitivity=1.0 + k) for k in range(5)])
    for group in (_fibre_fan(), sight):
        o, d, w, owner = group.gather_rays()
        host = cb.ray_segments(prim, o, d)
        arr, offs, etendue = group._descs()
        assert offs[-1] == o.shape[0] and np.array_equal(np.searchsorted(offs, np.arange(offs[-1]), side="right") - 1, owner)
        lib = _abi.load_library()
        buf = DeviceRayBuffer(int(offs[-1]), "cuda:0")
        weight = torch.empty(int(offs[-1]), dtype=torch.float64, device="cuda:0")
        rs = buf.as_struct()
        pd = _primitive_desc(prim)
        _abi.check(lib, lib.cb2_observer0d_rays_device(arr, len(group.observers), C.byref(pd), C.byref(rs), C.c_void_p(weight.data_ptr()), None))
        buf.n_rays, buf.n_segments = int(rs.n_rays), int(rs.n_segments)
        dev = buf.to_host()
        assert np.array_equal(dev.seg_offset, host.seg_offset) and host.n_segments > 0
        np.testing.assert_allclose(dev.origin, host.origin, rtol=0, atol=1e-14)
        np.testing.assert_allclose(dev.direction, host.direction, rtol=0, atol=1e-14)
        np.testing.assert_allclose(dev.seg_t0, host.seg_t0, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(dev.seg_t1, host.seg_t1, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(weight.cpu().numpy(), w, rtol=0, atol=1e-15)


def test_group_observe_on_the_device_matches_the_host_ray_path():
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.005)
    flat = cb.flatten_scene(plasma, 655.5, 656.9, 200)
    scene = EmissionScene(flat)
    group = _fibre_fan()
    group.connect_pipelines([cb.SpectralRadiancePipeline0D, cb.SpectralPowerPipeline0D, cb.RadiancePipeline0D, cb.PowerPipeline0D])
    spectra = group.observe(scene, plasma.geometry, plasma.geometry_to_world())
    ref_rad, ref_pow = group.observe_host_rays(scene, plasma.geometry, plasma.geometry_to_world())
    assert spectra.shape == (7, 200) and ref_rad.max() > 0
    # the results packed into every observer's pipelines (group/base.py:383-436)
    for i, ob in enumerate(group.observers):
        rad, pw, tot, totp = ob.pipelines
        assert (rad.min_wavelength, rad.max_wavelength, rad.bins) == (655.5, 656.9, 200) and rad.samples.samples == ob.pixel_samples
        np.testing.assert_allclose(rad.wavelengths, 655.5 + (np.arange(200) + 0.5) * 0.007, rtol=1e-12)
        assert np.array_equal(rad.samples.mean, spectra[i]) and np.array_equal(pw.samples.mean, group.power_spectra[i])
        np.testing.assert_allclose(tot.value.mean, spectra[i].sum() * 0.007, rtol=1e-12)
        np.testing.assert_allclose(totp.value.mean, group.power_spectra[i].sum() * 0.007, rtol=1e-12)
    np.testing.assert_allclose(spectra, ref_rad, rtol=1e-9, atol=1e-12 * ref_rad.max())
    np.testing.assert_allclose(group.power_spectra, ref_pow, rtol=1e-9, atol=1e-12 * ref_pow.max())
    # sight lines: one ray each, power = radiance x sensitivity
    sight = cb.SightLineGroup([cb.SightLine(transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.1 * k, -0.9)), sensitivity=2.0 + k) for k in range(4)])
    rad = sight.observe(scene, plasma.geometry, plasma.geometry_to_world())
    o, d, w, owner = sight.gather_rays()
    ref, _ = oracle.emission_render(flat, cb.ray_segments(plasma.geometry, o, d, plasma.geometry_to_world()))
    scene.close()
    assert np.all(np.abs(rad - ref) <= 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)) and ref.max() > 0
    np.testing.assert_allclose(sight.power_spectra, rad * np.array(sight.sensitivity)[:, None], rtol=1e-14)
    with pytest.raises(ValueError):
        cb.FibreOpticGroup().observe(scene, plasma.geometry)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("order", ["tiles", "random", "ragged"])
def test_render_rows_puts_every_ray_on_its_row_of_the_frame(order, pinned):
    # cb2_emission_render_rows: a rank's tile-ordered rays land on their pixels of the image-ordered host frame (strided D2H per tile);
    # any other row list must work too (single rows, runs of different lengths)
    from core_b200.sharding import tile_pixels
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.01)
    flat = cb.flatten_scene(plasma, 655.0, 657.5, 96)
    nx = ny = 40
    cam = cb.PinholeCamera((nx, ny), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    if order == "tiles":
        pix = tile_pixels((nx, ny), 1, 3)               # rank 1 of 3: 16 x 16 tiles, partial tiles at the frame's edge
    elif order == "random":
        pix = np.random.default_rng(3).permutation(nx * ny)[:500]
    else:
        pix = np.concatenate([np.arange(100, 137), np.arange(10, 12), [5], np.arange(400, 464), np.arange(300, 364), np.arange(200, 264), [1599]])
    rays = cb.ray_segments(plasma.geometry, *cam.rays(pixel_index=pix), plasma.geometry_to_world())
    scene = EmissionScene(flat)
    monkey_batch = 128                                   # several batches, so the overlapped per-batch copies are exercised
    import os
    os.environ["CB2_BATCH_RAYS"] = str(monkey_batch)
    try:
        plain, _ = scene.render(rays, dtype=np.float32)
        if pinned:                                       # a page-locked frame is written by the scatter kernel through its device mapping,
            import torch                                 # a pageable one by strided copies
            keep = torch.full((nx * ny, 96), -1.0, dtype=torch.float32).pin_memory()
            frame = keep.numpy()
        else:
            frame = np.full((nx * ny, 96), -1.0, dtype=np.float32)
        out, st = scene.render(rays, out=frame, rows=pix)
    finally:
        del os.environ["CB2_BATCH_RAYS"]
    scene.close()
    assert out is frame and plain.max() > 0
    assert np.array_equal(frame[pix], plain)
    untouched = np.ones(nx * ny, dtype=bool)
    untouched[pix] = False
    assert np.all(frame[untouched] == -1.0)
    with pytest.raises(ValueError):
        scene2 = EmissionScene(flat)
        try:
            scene2.render(rays, out=frame, rows=pix + nx * ny)
        finally:
            scene2.close()
