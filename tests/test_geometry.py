"""Bounding primitives of the host mirror (core_b200/geometry.py) against brute-force membership tests, and the beam's choice of
bounding volume (cherab/core/beam/node.pyx:505-554)."""
import numpy as np

import core_b200 as cb
from core_b200.geometry import TruncatedCone, ray_segments


def _chord_lengths(rays):
    return np.array([(rays.seg_t1[a:b] - rays.seg_t0[a:b]).sum() for a, b in zip(rays.seg_offset[:-1], rays.seg_offset[1:])])


def _brute(inside, o, d, t_max=8.0, n=8001):
    ts = np.linspace(0.0, t_max, n)
    p = o[:, None, :] + ts[None, :, None] * d[:, None, :]
    return inside(p).sum(axis=1) * (ts[1] - ts[0])


def _random_rays(seed, n):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-2, 2, (n, 3))
    o[:, 2] = rng.uniform(-1, 4, n)
    d = rng.normal(size=(n, 3))
    d[:20] = [0, 0, 1]                      # along the axis (a < 0 branch), across it, through the axis
    d[20:40] = [1, 0, 0]
    o[40:60, :2] = 0
    return o, d / np.linalg.norm(d, axis=1, keepdims=True)


def test_truncated_cone_chords():
    rs, re, length = 0.1, 0.6, 3.0
    da = rs * length / (re - rs)
    k = re / (length + da)
    o, d = _random_rays(1, 600)
    got = _chord_lengths(ray_segments(TruncatedCone(rs, re, length), o, d))
    ref = _brute(lambda p: (p[..., 2] >= 0) & (p[..., 2] <= length) & (np.hypot(p[..., 0], p[..., 1]) <= k * (p[..., 2] + da)), o, d)
    assert (got > 0).sum() > 30 and np.max(np.abs(got - ref)) < 3e-3


def test_hollow_cylinder_chords():
    o, d = _random_rays(2, 600)
    got = _chord_lengths(ray_segments(cb.HollowCylinder(0.4, 1.1, 0.0, 3.0), o, d))
    r = lambda p: np.hypot(p[..., 0], p[..., 1])
    ref = _brute(lambda p: (p[..., 2] >= 0) & (p[..., 2] <= 3.0) & (r(p) <= 1.1) & (r(p) >= 0.4), o, d)
    assert (got > 0).sum() > 100 and np.max(np.abs(got - ref)) < 3e-3


def test_beam_bounding_volume_follows_the_reference_rule():
    beam = cb.Beam()
    beam.attenuator = cb.SingleRayAttenuator(clamp_sigma=5.0)
    beam.sigma, beam.length = 0.05, 3.0
    beam.divergence_x = beam.divergence_y = 0.0
    g = beam.geometry                                   # no divergence: Cylinder(num_sigma * sigma, length)   node.pyx:524-525
    assert isinstance(g, cb.HollowCylinder) and g.r_outer == 0.25 and (g.z_min, g.z_max) == (0.0, 3.0)
    beam.divergence_x = 0.1                             # a cone would save < 10 % of the volume: same cylinder   node.pyx:543-545
    g = beam.geometry
    assert isinstance(g, cb.HollowCylinder) and g.r_outer == 0.25
    beam.divergence_y = 2.0                             # cone from 5 sigma to 5 sqrt(sigma^2 + (L tan 2deg)^2)      node.pyx:527-554
    g = beam.geometry
    assert isinstance(g, TruncatedCone) and g.radius_start == 0.25
    assert abs(g.radius_end - 5.0 * np.sqrt(0.05 ** 2 + (3.0 * np.tan(np.deg2rad(2.0))) ** 2)) < 1e-15
