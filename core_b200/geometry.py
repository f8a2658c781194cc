"""Ray generation and ray/primitive clipping: the caller side of the hot path (SURVEY 8(f2) front-end, minimal).

The reference obtains ray segments from Raysect's tracer (World.hit -> primitive entry/exit pairs -> evaluate_volume
per segment).  Here the same entry/exit intervals are computed analytically, in float64
numpy, for the primitives the BASELINE configs use, and handed to the kernels as ``RayBatch`` segments.
"""
import numpy as np

from .flatten import RayBatch, affine_inverse


def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def look_at(position, target, up=(0.0, 0.0, 1.0)):
    """camera -> world matrix with the camera looking along its +z axis at ``target`` and +y up (raysect convention)."""
    position, target, up = (np.asarray(v, dtype=np.float64) for v in (position, target, up))
    zc = target - position
    zc /= np.linalg.norm(zc)
    xc = np.cross(up, zc)
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = xc, yc, zc, position
    return m


def _interval_intersect(a0, a1, b0, b1):
    return np.maximum(a0, b0), np.minimum(a1, b1)


def _cyl_interval(o, d, radius):
    """t-interval where an infinite z-axis cylinder of ``radius`` contains o + t d (empty: t0 > t1)."""
    a = d[:, 0] ** 2 + d[:, 1] ** 2
    b = 2.0 * (o[:, 0] * d[:, 0] + o[:, 1] * d[:, 1])
    c = o[:, 0] ** 2 + o[:, 1] ** 2 - radius * radius
    disc = b * b - 4.0 * a * c
    t0 = np.full(o.shape[0], np.inf)
    t1 = np.full(o.shape[0], -np.inf)
    par = a < 1e-300
    ok = (~par) & (disc > 0)
    sq = np.sqrt(np.where(ok, disc, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = np.where(ok, (-b - sq) / (2 * a), t0)
        t1 = np.where(ok, (-b + sq) / (2 * a), t1)
    inside_par = par & (c < 0)
    t0 = np.where(inside_par, -np.inf, t0)
    t1 = np.where(inside_par, np.inf, t1)
    return t0, t1


def _slab_interval(o, d, axis, lo, hi):
    with np.errstate(divide="ignore", invalid="ignore"):
        ta = (lo - o[:, axis]) / d[:, axis]
        tb = (hi - o[:, axis]) / d[:, axis]
    t0, t1 = np.minimum(ta, tb), np.maximum(ta, tb)
    par = d[:, axis] == 0
    inside = (o[:, axis] >= lo) & (o[:, axis] <= hi)      # a ray lying in a face plane counts as inside (test_beamcxline.py)
    t0 = np.where(par, np.where(inside, -np.inf, np.inf), t0)
    t1 = np.where(par, np.where(inside, np.inf, -np.inf), t1)
    return t0, t1


class Primitive:
    """Base: ``intervals(o, d)`` returns a list of (t0, t1) array pairs in primitive-local space."""
    transform = None  # local -> world

    def intervals(self, o, d):
        raise NotImplementedError


class HollowCylinder(Primitive):
    """Subtract(Cylinder(r_outer, height), Cylinder(r_inner, ...)) with base at z=z_min (plasma.py:673-681;
    raytransfer.py:198).  r_inner=0 gives a solid Cylinder."""

    def __init__(self, r_inner, r_outer, z_min, z_max, transform=None):
        self.r_inner, self.r_outer, self.z_min, self.z_max = float(r_inner), float(r_outer), float(z_min), float(z_max)
        self.transform = transform

    def intervals(self, o, d):
        a0, a1 = _cyl_interval(o, d, self.r_outer)
        s0, s1 = _slab_interval(o, d, 2, self.z_min, self.z_max)
        a0, a1 = _interval_intersect(a0, a1, s0, s1)
        a0 = np.maximum(a0, 0.0)
        if self.r_inner <= 0:
            return [(a0, a1)]
        b0, b1 = _cyl_interval(o, d, self.r_inner)
        # [a0,a1] minus [b0,b1]
        hit = b1 > b0
        first = (a0, np.where(hit, np.minimum(a1, b0), a1))
        second = (np.where(hit, np.maximum(a0, b1), np.inf), a1)
        return [first, second]


class TruncatedCone(Primitive):
    """Intersect(Cone(radius_end, cone_height, translate(0, 0, length) * rotate_x(180)), Cylinder(1.01 radius_end, 1.01 length)):
    the bounding volume of a diverging beam (cherab/core/beam/node.pyx:527-554) — radius_start at z = 0 growing linearly to
    radius_end at z = length."""

    def __init__(self, radius_start, radius_end, length, transform=None):
        if not radius_end > radius_start > 0 or not length > 0:
            raise ValueError("TruncatedCone needs 0 < radius_start < radius_end and a positive length")
        self.radius_start, self.radius_end, self.length = float(radius_start), float(radius_end), float(length)
        self.transform = transform

    def intervals(self, o, d):
        da = self.radius_start * self.length / (self.radius_end - self.radius_start)      # apex at z = -da
        k2 = (self.radius_end / (self.length + da)) ** 2
        zo = o[:, 2] + da
        a = d[:, 0] ** 2 + d[:, 1] ** 2 - k2 * d[:, 2] ** 2
        b = 2.0 * (o[:, 0] * d[:, 0] + o[:, 1] * d[:, 1] - k2 * zo * d[:, 2])
        c = o[:, 0] ** 2 + o[:, 1] ** 2 - k2 * zo * zo
        s0, s1 = _slab_interval(o, d, 2, 0.0, self.length)
        s0 = np.maximum(s0, 0.0)
        disc = b * b - 4.0 * a * c
        lin = np.abs(a) < 1e-300
        with np.errstate(divide="ignore", invalid="ignore"):
            sq = np.sqrt(np.maximum(disc, 0.0))
            r0 = np.where(a > 0, (-b - sq) / (2 * a), (-b + sq) / (2 * a))              # smaller root
            r1 = np.where(a > 0, (-b + sq) / (2 * a), (-b - sq) / (2 * a))              # larger root
            tl = -c / b
        miss = np.full(o.shape[0], np.inf), np.full(o.shape[0], -np.inf)
        # a > 0: inside between the roots; a < 0: inside outside the roots (the lower nappe lies below the slab); a = 0: half line
        has = disc > 0
        f0 = np.where(lin, np.where(b > 0, s0, np.maximum(s0, tl)), np.where(a > 0, np.where(has, np.maximum(s0, r0), miss[0]), s0))
        f1 = np.where(lin, np.where(b > 0, np.minimum(s1, tl), s1), np.where(a > 0, np.where(has, np.minimum(s1, r1), miss[1]),
                                                                           np.where(has, np.minimum(s1, r0), s1)))
        g0 = np.where((~lin) & (a < 0) & has, np.maximum(s0, r1), miss[0])
        g1 = np.where((~lin) & (a < 0) & has, s1, miss[1])
        lin_out = lin & (np.abs(b) < 1e-300) & (c > 0)
        f0 = np.where(lin_out, miss[0], f0)
        return [(f0, f1), (g0, g1)]


class Sphere(Primitive):
    def __init__(self, radius, transform=None):
        self.radius, self.transform = float(radius), transform

    def intervals(self, o, d):
        b = 2.0 * np.einsum("ij,ij->i", o, d)
        c = np.einsum("ij,ij->i", o, o) - self.radius ** 2
        disc = b * b - 4.0 * c
        ok = disc > 0
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t0 = np.where(ok, (-b - sq) / 2, np.inf)
        t1 = np.where(ok, (-b + sq) / 2, -np.inf)
        return [(np.maximum(t0, 0.0), t1)]


class Box(Primitive):
    def __init__(self, lower, upper, transform=None):
        self.lower, self.upper, self.transform = np.asarray(lower, float), np.asarray(upper, float), transform

    def intervals(self, o, d):
        t0 = np.zeros(o.shape[0])
        t1 = np.full(o.shape[0], np.inf)
        for ax in range(3):
            s0, s1 = _slab_interval(o, d, ax, self.lower[ax], self.upper[ax])
            t0, t1 = _interval_intersect(t0, t1, s0, s1)
        return [(t0, t1)]


def ray_segments(primitive, origins, directions, to_world=None):
    """Clip rays (world space) against ``primitive`` -> RayBatch with 0..k segments per ray, ordered along the ray.
    ``to_world`` (4x4, primitive-local -> world) overrides ``primitive.transform``; pass ``plasma.geometry_to_world()``
    for a plasma's geometry."""
    o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    if to_world is None:
        to_world = primitive.transform
    if to_world is not None:
        w2l = affine_inverse(to_world)
        ol = o @ w2l[:3, :3].T + w2l[:3, 3]
        dl = d @ w2l[:3, :3].T
    else:
        ol, dl = o, d
    ivs = primitive.intervals(ol, dl)
    valid = [(t1 > t0) & np.isfinite(t0) & np.isfinite(t1) for t0, t1 in ivs]
    counts = np.sum(valid, axis=0).astype(np.int64)
    seg_offset = np.concatenate([[0], np.cumsum(counts)])
    n = o.shape[0]
    seg_t0 = np.empty(seg_offset[-1])
    seg_t1 = np.empty(seg_offset[-1])
    cursor = seg_offset[:-1].copy()
    for (t0, t1), v in zip(ivs, valid):
        idx = cursor[v]
        seg_t0[idx] = t0[v]
        seg_t1[idx] = t1[v]
        cursor[v] += 1
    assert n == counts.size
    return RayBatch(o, d, seg_offset, seg_t0, seg_t1)


class PinholeCamera:
    """raysect PinholeCamera geometry: image plane at z=1 in camera space, width 2 tan(fov/2),
    pixel (ix, iy) looks through (w/2 - delta (ix + sx), h/2 - delta (iy + sy), 1); sx, sy in [0,1) is the
    sub-pixel sample position (0.5 = pixel centre)."""

    def __init__(self, pixels, fov=45.0, transform=None):
        self.pixels, self.fov = (int(pixels[0]), int(pixels[1])), float(fov)
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)

    def rays(self, sub_x=0.5, sub_y=0.5, pixel_index=None):
        """Return (origins, directions) for every pixel in row-major (ix, iy) order, or for ``pixel_index`` only."""
        nx, ny = self.pixels
        width = 2.0 * np.tan(np.pi / 180.0 * 0.5 * self.fov)
        delta = width / nx
        start_x, start_y = 0.5 * width, 0.5 * delta * ny
        if pixel_index is None:
            ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
            ix, iy = ix.ravel(), iy.ravel()
        else:
            pixel_index = np.asarray(pixel_index, dtype=np.int64)
            ix, iy = pixel_index // ny, pixel_index % ny
        x = start_x - delta * (ix + sub_x)
        y = start_y - delta * (iy + sub_y)
        dc = np.stack([x, y, np.ones_like(x)], axis=1)
        dc /= np.linalg.norm(dc, axis=1, keepdims=True)
        m = self.transform
        d = dc @ m[:3, :3].T
        o = np.broadcast_to(m[:3, 3], d.shape).copy()
        return o, d


def stratified_offsets(n_side=4):
    """Deterministic n_side x n_side stratified sub-pixel offsets (16 samples/pixel)."""
    k = (np.arange(n_side) + 0.5) / n_side
    return [(float(a), float(b)) for a in k for b in k]
