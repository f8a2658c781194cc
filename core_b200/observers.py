"""Observer front-end on the device.

``DevicePinhole`` generates the rays of a ``PinholeCamera`` and their chords through a bounding primitive directly in
device memory (cb2_pinhole_rays_device), and ``observe`` is the frame loop raysect's ``Observer2D.observe()`` runs pixel by
pixel in Python: every pixel sample is one device render accumulated into the frame with weight 1 / pixel_samples — the mean
spectral radiance per pixel that ``SpectralRadiancePipeline2D`` reports — so the frame crosses PCIe once, when the caller
asks for it.  The host mirror of the same geometry (``geometry.PinholeCamera.rays`` + ``geometry.ray_segments``) stays the
path for explicit ray lists.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _abi
from .flatten import affine_inverse
from .geometry import Box, HollowCylinder, Sphere, stratified_offsets


def _primitive_desc(primitive, to_world=None):
    d = _abi.PrimitiveDesc()
    if isinstance(primitive, HollowCylinder):
        d.kind = _abi.PRIM_HOLLOW_CYLINDER
        vals = (primitive.r_inner, primitive.r_outer, primitive.z_min, primitive.z_max, 0.0, 0.0)
    elif isinstance(primitive, Sphere):
        d.kind = _abi.PRIM_SPHERE
        vals = (primitive.radius, 0.0, 0.0, 0.0, 0.0, 0.0)
    elif isinstance(primitive, Box):
        d.kind = _abi.PRIM_BOX
        vals = tuple(primitive.lower) + tuple(primitive.upper)
    else:
        raise TypeError("Unsupported primitive for device ray generation: %r" % (primitive,))
    for k, v in enumerate(vals):
        d.p[k] = float(v)
    if to_world is None:
        to_world = primitive.transform
    w2l = np.eye(4) if to_world is None else affine_inverse(np.asarray(to_world, dtype=np.float64))
    for i in range(3):
        for j in range(4):
            d.world_to_local[4 * i + j] = w2l[i, j]
    return d


class DeviceRayBuffer:
    """Device-resident cb2_rays (torch tensors), duck-compatible with engine.DeviceRays."""

    def __init__(self, n, device):
        import torch
        self.device = torch.device(device)
        self.n_rays, self.n_segments = int(n), 0
        f64 = dict(dtype=torch.float64, device=self.device)
        self.origin = torch.empty((n, 3), **f64)
        self.direction = torch.empty((n, 3), **f64)
        self.seg_offset = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        self.seg_t0 = torch.empty(2 * n, **f64)
        self.seg_t1 = torch.empty(2 * n, **f64)
        self.nbytes = 0                                   # nothing crosses PCIe

    def as_struct(self):
        r = _abi.Rays()
        r.n_rays, r.n_segments = self.n_rays, self.n_segments
        r.origin = C.cast(C.c_void_p(self.origin.data_ptr()), _abi.c_double_p)
        r.direction = C.cast(C.c_void_p(self.direction.data_ptr()), _abi.c_double_p)
        r.seg_offset = C.cast(C.c_void_p(self.seg_offset.data_ptr()), _abi.c_int64_p)
        r.seg_t0 = C.cast(C.c_void_p(self.seg_t0.data_ptr()), _abi.c_double_p)
        r.seg_t1 = C.cast(C.c_void_p(self.seg_t1.data_ptr()), _abi.c_double_p)
        return r

    def to_host(self):
        """RayBatch copy (tests, debugging)."""
        from .flatten import RayBatch
        n = self.n_segments
        return RayBatch(self.origin.cpu().numpy(), self.direction.cpu().numpy(), self.seg_offset.cpu().numpy(),
                        self.seg_t0[:n].cpu().numpy(), self.seg_t1[:n].cpu().numpy())


class DevicePinhole:
    """A geometry.PinholeCamera bound to a bounding primitive on one GPU."""

    def __init__(self, camera, primitive, to_world=None, device=0, pixel_index=None):
        import torch
        self._lib = _abi.load_library()
        self.camera = camera
        self.device = torch.device("cuda", int(device))
        self._prim = _primitive_desc(primitive, to_world)
        self._cam = _abi.PinholeDesc()
        self._cam.nx, self._cam.ny = camera.pixels
        self._cam.width = 2.0 * np.tan(np.pi / 180.0 * 0.5 * camera.fov)
        m = np.asarray(camera.transform, dtype=np.float64)
        for i in range(3):
            for j in range(4):
                self._cam.to_world[4 * i + j] = m[i, j]
        self.pixel_index = None
        if pixel_index is not None:
            self.pixel_index = torch.as_tensor(np.ascontiguousarray(pixel_index, dtype=np.int64)).to(self.device)
        self.n_rays = int(self.pixel_index.numel()) if self.pixel_index is not None else camera.pixels[0] * camera.pixels[1]

    def rays(self, sub_x=0.5, sub_y=0.5, out=None):
        """Rays through the sub-pixel position (sub_x, sub_y) of every (listed) pixel -> DeviceRayBuffer (reused if given)."""
        import torch
        buf = out if out is not None else DeviceRayBuffer(self.n_rays, self.device)
        rs = buf.as_struct()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _abi.check(self._lib, self._lib.cb2_pinhole_rays_device(
                C.byref(self._cam), C.byref(self._prim), C.c_void_p(self.pixel_index.data_ptr()) if self.pixel_index is not None else None,
                self.n_rays, float(sub_x), float(sub_y), C.byref(rs), C.c_void_p(stream)))
        buf.n_rays, buf.n_segments = int(rs.n_rays), int(rs.n_segments)
        return buf


def observe(scene, pinhole, pixel_samples_side=1, frame=None, dtype=None):
    """Mean spectral radiance per pixel over pixel_samples_side^2 stratified sub-pixel samples, on the device.

    ``scene``: engine.EmissionScene or engine.PlasmaRenderer; ``pinhole``: DevicePinhole.  Returns a torch tensor
    [n_pixels, bins] (float32 unless ``dtype``/``frame`` say otherwise) — ``SpectralRadiancePipeline2D.frame.mean`` of the
    reference flow, W / (m^2 sr nm)."""
    import torch
    bins = scene.scene.bins if hasattr(scene, "scene") else scene.bins
    if frame is None:
        frame = torch.zeros((pinhole.n_rays, bins), dtype=dtype or torch.float32, device=pinhole.device)
    else:
        frame.zero_()
    offsets = stratified_offsets(pixel_samples_side)
    buf = DeviceRayBuffer(pinhole.n_rays, pinhole.device)
    for sx, sy in offsets:
        rays = pinhole.rays(sx, sy, out=buf)
        scene.render_device(rays, frame, scale=1.0 / len(offsets), accumulate=True)
    return frame
