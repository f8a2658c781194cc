"""First-wall occlusion: the Generomak wall components as one triangle soup behind a BVH on the device.

The reference adds the wall to the Raysect scene graph (cherab/generomak/machine/first_wall.py:11-184: every component is a
Wavefront mesh instanced toroidally — 32 copies 11.25 degrees apart — and, for the inner limiter, in 10 vertical rows) and the
tracer ends a ray at the first opaque hit, so the plasma's volume integral stops there.  Here the same instances are flattened to
world-space triangles, a bounding-volume hierarchy is built once on the host, and ``FirstWall.clip_device`` shortens / empties the
ray segments beyond each ray's first wall hit on the device, between ray generation (``DevicePinhole.rays``) and the marcher.
Only the occluding role of the wall is modelled (an absorbing surface): no reflected light.

The component meshes come from ``core_b200/data/generomak_first_wall.npz`` (generated from the reference's OBJ files by
tools/make_data_tables.py; ``OuterWallLimiter.obj`` is not in the reference checkout and is skipped).  No CPU fallback:
``FirstWall.hit`` / ``clip`` run on the GPU; the brute-force float64 check lives in ``oracle/``.
"""
import ctypes as C
import os

import numpy as np

from . import _abi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "generomak_first_wall.npz")

# (component, toroidal shift of the first instance [deg], toroidal step [deg], toroidal instances, first vertical shift [m],
#  vertical step [m], vertical instances) — first_wall.py:11-117
FIRST_WALL_COMPONENT = {
    "InnerWallLimiter": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32, initial_vertical_shift=-0.561,
                             vertical_step=187e-3, vertical_instances=10),
    "InnerDivertorBaffle": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "BottomBaffle": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "BottomDivertorFloor": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "BottomOuterVerticalTarget": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "BottomInnerVerticalTarget": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "TopBaffle": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "TopDivertorFloor": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "TopInnerVerticalTarget": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "TopOuterVerticalTarget": dict(initial_toroidal_shift=-90, toroidal_step=11.25, toroidal_instances=32),
    "OuterWallLimiter": dict(initial_toroidal_shift=-45, toroidal_step=45, toroidal_instances=8),
}


def instance_transforms(toroidal_step=0, toroidal_instances=1, initial_toroidal_shift=0, vertical_step=0, vertical_instances=1,
                        initial_vertical_shift=0):
    """{"tor, vert": 4x4 matrix} = translate(0, 0, z0 + vert dz) * rotate_z(phi0 + tor dphi), first_wall.py:148-160."""
    out = {}
    for tor in range(toroidal_instances):
        for vert in range(vertical_instances):
            a = np.deg2rad(initial_toroidal_shift + tor * toroidal_step)
            m = np.eye(4)
            m[0, 0], m[0, 1], m[1, 0], m[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
            m[2, 3] = initial_vertical_shift + vert * vertical_step
            out["{:d}, {:d}".format(tor, vert)] = m
    return out


def load_component_group(vertices, triangles, **placement):
    """World-space triangles [instances * m, 3, 3] of one component group (first_wall.py:120-162)."""
    base = np.asarray(vertices, dtype=np.float64)[np.asarray(triangles, dtype=np.int64)]          # [m, 3, 3]
    parts = []
    for m in instance_transforms(**placement).values():
        parts.append(base @ m[:3, :3].T + m[:3, 3])
    return np.concatenate(parts, axis=0)


def load_first_wall(mesh_file=None, components=None):
    """{component name: world-space triangles} for the Generomak first wall (first_wall.py:165-184).  Components whose mesh is not
    in the data file (OuterWallLimiter: missing from the reference checkout) are left out."""
    data = np.load(mesh_file or _DATA)
    out = {}
    for name, placement in FIRST_WALL_COMPONENT.items():
        if components is not None and name not in components:
            continue
        if name + "_vertices" not in data.files:
            continue
        out[name] = load_component_group(data[name + "_vertices"], data[name + "_triangles"], **placement)
    return out


class FirstWall:
    """Opaque triangle geometry on one GPU: BVH built at construction (cb2_wall_create), first-hit queries and segment clipping."""

    def __init__(self, triangles=None, device=0):
        self._lib = _abi.load_library()
        if triangles is None:
            triangles = np.concatenate(list(load_first_wall().values()), axis=0)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 3, 3)
        if self.triangles.shape[0] < 1:
            raise ValueError("a wall needs at least one triangle")
        self.device = int(device)
        d = _abi.WallDesc()
        d.abi_version = _abi.ABI_VERSION
        d.n_triangles = self.triangles.shape[0]
        d.vertices = self.triangles.ctypes.data_as(_abi.c_double_p)
        self._h = C.c_void_p()
        _abi.check(self._lib, self._lib.cb2_wall_create(C.byref(d), self.device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cb2_wall_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def n_triangles(self):
        return int(self.triangles.shape[0])

    def hit(self, origin, direction):
        """Distance along each ray (units of |direction|) to its first wall hit, +inf where it misses.  Host arrays [n, 3]."""
        o = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origin and direction must have the same shape")
        t = np.empty(o.shape[0], dtype=np.float64)
        _abi.check(self._lib, self._lib.cb2_wall_hit(self._h, o.ctypes.data_as(_abi.c_double_p), d.ctypes.data_as(_abi.c_double_p),
                                                     o.shape[0], t.ctypes.data_as(_abi.c_double_p)))
        return t

    def clip(self, rays):
        """RayBatch with every segment cut at the ray's first wall hit (segments behind it become empty).  Host buffers."""
        from .flatten import RayBatch
        t = self.hit(rays.origin, rays.direction)
        per_seg = np.repeat(t, np.diff(rays.seg_offset))
        t0 = rays.seg_t0.copy()
        t1 = np.maximum(np.minimum(rays.seg_t1, per_seg), t0)
        return RayBatch(rays.origin, rays.direction, rays.seg_offset, t0, t1)

    def clip_device(self, dev_rays, hit_out=None):
        """In place on a DeviceRays / DeviceRayBuffer: seg_t1 = max(seg_t0, min(seg_t1, t_hit)) on the current stream.
        ``hit_out``: optional torch float64 tensor [n_rays] that receives the hit distances."""
        import torch
        rs = dev_rays.as_struct()
        stream = torch.cuda.current_stream(dev_rays.device).cuda_stream
        _abi.check(self._lib, self._lib.cb2_wall_clip_device(self._h, C.byref(rs), C.c_void_p(hit_out.data_ptr()) if hit_out is not None else None,
                                                             C.c_void_p(stream)))
        return dev_rays
