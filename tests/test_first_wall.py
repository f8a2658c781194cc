"""First-wall occlusion: the host side (component placement, reference first_wall.py:11-184) here, the BVH
first-hit kernel against the brute-force float64 oracle on the GPU."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200 import first_wall as fw
from oracle import oracle


def test_component_table_follows_the_reference():
    # first_wall.py:11-117: ten 32-fold components 11.25 degrees apart from -90, the inner limiter in 10 rows from z = -0.561 m,
    # the outer limiter 8-fold from -45 degrees
    assert len(fw.FIRST_WALL_COMPONENT) == 11
    for name, c in fw.FIRST_WALL_COMPONENT.items():
        if name == "OuterWallLimiter":
            assert (c["initial_toroidal_shift"], c["toroidal_step"], c["toroidal_instances"]) == (-45, 45, 8)
        else:
            assert (c["initial_toroidal_shift"], c["toroidal_step"], c["toroidal_instances"]) == (-90, 11.25, 32)
    inner = fw.FIRST_WALL_COMPONENT["InnerWallLimiter"]
    assert inner["vertical_instances"] == 10 and inner["initial_vertical_shift"] == -0.561 and inner["vertical_step"] == 187e-3


def test_instance_transforms():
    t = fw.instance_transforms(toroidal_step=90, toroidal_instances=4, initial_toroidal_shift=-90, vertical_step=0.5, vertical_instances=2,
                               initial_vertical_shift=-1.0)
    assert sorted(t) == sorted("{:d}, {:d}".format(a, b) for a in range(4) for b in range(2))
    p = np.array([1.0, 0.0, 0.0, 1.0])
    np.testing.assert_allclose(t["0, 0"] @ p, [0.0, -1.0, -1.0, 1.0], atol=1e-15)      # rotate_z(-90) then translate(0, 0, -1)
    np.testing.assert_allclose(t["1, 1"] @ p, [1.0, 0.0, -0.5, 1.0], atol=1e-15)
    np.testing.assert_allclose(t["2, 0"] @ p, [0.0, 1.0, -1.0, 1.0], atol=1e-15)


def test_wall_meshes_load_and_are_toroidally_periodic():
    parts = fw.load_first_wall()
    assert "OuterWallLimiter" not in parts and len(parts) == 10             # the OBJ is missing from the reference checkout
    baffle = parts["TopBaffle"]
    m = baffle.shape[0] // 32
    a = np.deg2rad(11.25)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    np.testing.assert_allclose(baffle[m:2 * m], baffle[:m] @ rot.T, atol=1e-12)
    r = np.hypot(baffle[..., 0], baffle[..., 1])
    assert 0.5 < r.min() and r.max() < 2.5 and baffle[..., 2].min() > 0.5


def test_oracle_wall_hit_known_answers():
    # one triangle in the plane x = 1: a ray along +x from the origin hits at t = 1; |direction| = 2 halves t; a ray that starts on the
    # far side, points away or passes outside the triangle misses
    tri = np.array([[[1.0, -1.0, -1.0], [1.0, 2.0, -1.0], [1.0, -1.0, 2.0]]])
    o = np.array([[0, 0, 0], [0, 0, 0], [2, 0, 0], [0, 0, 0], [0, 1.9, 1.9]], dtype=float)
    d = np.array([[1, 0, 0], [2, 0, 0], [1, 0, 0], [-1, 0, 0], [1, 0, 0]], dtype=float)
    t = oracle.wall_hit(tri, o, d)
    np.testing.assert_allclose(t[:2], [1.0, 0.5], rtol=1e-15)
    assert np.all(np.isinf(t[2:]))


def _camera_rays(n=48):
    cam = cb.PinholeCamera((n, n), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    return cam, cam.rays()


@pytest.mark.gpu
def test_first_hit_matches_the_brute_force_oracle():
    wall = cb.FirstWall()
    assert wall.n_triangles == 2877632
    cam, (o, d) = _camera_rays(40)
    rng = np.random.default_rng(5)
    # camera rays plus rays from random points inside the vessel in random directions
    r, phi, z = rng.uniform(1.0, 1.9, 400), rng.uniform(0, 2 * np.pi, 400), rng.uniform(-0.9, 0.9, 400)
    o2 = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    d2 = rng.normal(size=(400, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    o, d = np.concatenate([o, o2]), np.concatenate([d, d2])
    t = wall.hit(o, d)
    ref = oracle.wall_hit(wall.triangles, o, d)
    hit = np.isfinite(ref)
    assert 0.1 < hit.mean() and np.array_equal(hit, np.isfinite(t))
    # the exact test is the same float64 expression on both sides; the BVH only prunes
    np.testing.assert_allclose(t[hit], ref[hit], rtol=1e-12, atol=1e-12)
    # a direction of another length rescales t
    t2 = wall.hit(o[:64], 2.5 * d[:64])
    h = np.isfinite(t[:64])
    np.testing.assert_allclose(t2[h] * 2.5, t[:64][h], rtol=1e-12)
    wall.close()


@pytest.mark.gpu
def test_clip_device_cuts_segments_at_the_wall_and_the_frame_follows():
    import torch
    from core_b200 import generomak
    from core_b200.engine import EmissionScene
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.005)
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 64)
    cam = cb.PinholeCamera((24, 24), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    pin = cb.DevicePinhole(cam, plasma.geometry, to_world=plasma.geometry_to_world())
    wall = cb.FirstWall()
    free = pin.rays().to_host()
    buf = pin.rays()
    t_dev = torch.empty(pin.n_rays, dtype=torch.float64, device=pin.device)
    wall.clip_device(buf, hit_out=t_dev)
    clipped = buf.to_host()
    t_ref = oracle.wall_hit(wall.triangles, free.origin, free.direction)
    np.testing.assert_allclose(t_dev.cpu().numpy(), t_ref, rtol=1e-12)
    per_seg = np.repeat(t_ref, np.diff(free.seg_offset))
    assert np.array_equal(clipped.seg_offset, free.seg_offset) and np.array_equal(clipped.seg_t0, free.seg_t0)
    np.testing.assert_allclose(clipped.seg_t1, np.maximum(free.seg_t0, np.minimum(free.seg_t1, per_seg)), rtol=1e-12, atol=1e-12)
    assert (clipped.seg_t1 < free.seg_t1).mean() > 0.1                     # the wall does cut chords of this view
    host_clip = wall.clip(free)
    np.testing.assert_allclose(host_clip.seg_t1, clipped.seg_t1, rtol=1e-12, atol=1e-12)
    # the frame rendered behind the wall is the oracle's integral over the clipped chords
    scene = EmissionScene(flat)
    frame = cb.observe(scene, pin, wall=wall, dtype=torch.float64).cpu().numpy()
    scene.close()
    wall.close()
    ref = oracle.emission_render(flat, clipped)[0]
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
    assert ref.max() > 0 and np.all(np.abs(frame - ref) <= tol)
    open_frame = oracle.emission_render(flat, free)[0]
    assert np.abs(open_frame - ref).max() > 1e-3 * ref.max()                # and it is not the frame of the open vessel
