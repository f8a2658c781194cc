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
    frame = cb.observe(scene, pin, pixel_samples_side=2, dtype=torch.float64).cpu().numpy()
    scene.close()
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
