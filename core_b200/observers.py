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


def observe(scene, pinhole, pixel_samples_side=1, frame=None, dtype=None, wall=None, pipelines=()):
    """Mean spectral radiance per pixel over pixel_samples_side^2 stratified sub-pixel samples, on the device.

    ``scene``: engine.EmissionScene or engine.PlasmaRenderer; ``pinhole``: DevicePinhole.  Returns a torch tensor
    [n_pixels, bins] (float32 unless ``dtype``/``frame`` say otherwise) — ``SpectralRadiancePipeline2D.frame.mean`` of the
    reference flow, W / (m^2 sr nm).  ``wall``: optional first_wall.FirstWall — every ray's chords end at its first wall hit
    (what the opaque wall meshes do in the reference's scene graph, generomak/machine/first_wall.py:120-184).  ``pipelines``:
    SpectralRadiancePipeline2D / RadiancePipeline2D objects that receive the frame on the host."""
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
        if wall is not None:
            wall.clip_device(rays)
        scene.render_device(rays, frame, scale=1.0 / len(offsets), accumulate=True)
    if pipelines:
        # result packing into 2-D pipelines (one D2H of the frame): only for a camera that renders all of its pixels
        if pinhole.pixel_index is not None:
            raise ValueError("pipelines need the whole frame: the pinhole was built with a pixel subset")
        nx, ny = pinhole.camera.pixels
        host = frame.cpu().numpy().reshape(nx, ny, bins)
        grid = (scene.scene if hasattr(scene, "scene") else scene).flat.desc.grid
        for p in pipelines:
            p._fill_frame(grid.min_wavelength, grid.max_wavelength, host, len(offsets))
    return frame


# ------------------------------------------------------------------------------------------------------------------
# 0-D pipelines: what an observer's result is packed into [raysect SpectralRadiancePipeline0D / SpectralPowerPipeline0D /
# RadiancePipeline0D / PowerPipeline0D], as the groups forward them (cherab/tools/observers/group/base.py:130-135, 383-436)
# ------------------------------------------------------------------------------------------------------------------
class _Stats:
    """The part of raysect's StatsArray1D / StatsBin a consumer reads: mean, variance, samples (rays), errors()."""

    def __init__(self, mean, samples):
        self.mean = mean
        self.variance = np.zeros_like(mean)          # deterministic ray sets: no Monte-Carlo variance estimate
        self.samples = samples

    def errors(self):
        return np.zeros_like(self.mean)


class _Pipeline0D:
    _POWER = False

    def __init__(self, name=None, accumulate=True, display_progress=False):
        self.name = name or type(self).__name__
        self.accumulate, self.display_progress = accumulate, display_progress
        self.min_wavelength = self.max_wavelength = self.delta_wavelength = None
        self.bins = 0
        self.wavelengths = None

    def _set_grid(self, min_wavelength, max_wavelength, bins):
        self.min_wavelength, self.max_wavelength, self.bins = float(min_wavelength), float(max_wavelength), int(bins)
        self.delta_wavelength = (self.max_wavelength - self.min_wavelength) / self.bins
        self.wavelengths = self.min_wavelength + (np.arange(self.bins) + 0.5) * self.delta_wavelength


class SpectralRadiancePipeline0D(_Pipeline0D):
    """``samples.mean[bins]``: mean spectral radiance over the observer's rays, W / (m^2 sr nm)."""

    def _fill(self, min_wavelength, max_wavelength, radiance, power, n_rays):
        self._set_grid(min_wavelength, max_wavelength, radiance.size)
        self.samples = _Stats((power if self._POWER else radiance).copy(), int(n_rays))


class SpectralPowerPipeline0D(SpectralRadiancePipeline0D):
    """``samples.mean[bins]``: spectral power collected by the observer (etendue-weighted), W / nm."""
    _POWER = True


class RadiancePipeline0D(_Pipeline0D):
    """``value.mean``: the spectral radiance integrated over the spectral range, W / (m^2 sr)."""

    def _fill(self, min_wavelength, max_wavelength, radiance, power, n_rays):
        self._set_grid(min_wavelength, max_wavelength, radiance.size)
        self.value = _Stats(np.float64((power if self._POWER else radiance).sum() * self.delta_wavelength), int(n_rays))


class PowerPipeline0D(RadiancePipeline0D):
    """``value.mean``: the collected power integrated over the spectral range, W."""
    _POWER = True


class SpectralRadiancePipeline2D(_Pipeline0D):
    """``frame.mean[nx, ny, bins]``: mean spectral radiance per pixel over the pixel samples, W / (m^2 sr nm) [raysect]."""

    def _fill_frame(self, min_wavelength, max_wavelength, frame, pixel_samples):
        self._set_grid(min_wavelength, max_wavelength, frame.shape[-1])
        self.frame = _Stats(frame, int(pixel_samples))


class RadiancePipeline2D(_Pipeline0D):
    """``frame.mean[nx, ny]``: radiance per pixel integrated over the spectral range, W / (m^2 sr) [raysect]."""

    def _fill_frame(self, min_wavelength, max_wavelength, frame, pixel_samples):
        self._set_grid(min_wavelength, max_wavelength, frame.shape[-1])
        self.frame = _Stats(frame.sum(axis=-1) * self.delta_wavelength, int(pixel_samples))


# ------------------------------------------------------------------------------------------------------------------
# 0-D observers and their groups (cherab/tools/observers/group/{base,fibreoptic,sightline}.py)
# ------------------------------------------------------------------------------------------------------------------
class SightLine:
    """raysect SightLine: one ray from the observer's origin along its +z axis; ``sensitivity`` scales power pipelines
    (cherab/tools/observers/group/sightline.py:73-95)."""

    def __init__(self, transform=None, name="", sensitivity=1.0):
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        self.name, self.sensitivity = name, float(sensitivity)
        self.spectrum = None
        self.pipelines = []

    def rays(self):
        m = self.transform
        return m[:3, 3][None, :].copy(), m[:3, 2][None, :].copy(), np.ones(1)


class FibreOptic:
    """raysect FibreOptic: rays start on the fibre tip (disc of ``radius``) and leave within ``acceptance_angle`` degrees of the
    +z axis.  raysect draws both at random; here the ``pixel_samples`` rays are deterministic (sunflower points on the disc,
    Fibonacci points uniform in solid angle on the cone cap — SURVEY 8(d) C5), each weighted by cos(theta), the projected
    tip area it sees, so the weighted mean is the etendue-averaged spectral radiance."""

    def __init__(self, transform=None, name="", acceptance_angle=5.0, radius=0.001, pixel_samples=64):
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        self.name = name
        self.acceptance_angle, self.radius, self.pixel_samples = float(acceptance_angle), float(radius), int(pixel_samples)
        self.spectrum = None
        self.pipelines = []

    @property
    def solid_angle(self):
        return 2.0 * np.pi * (1.0 - np.cos(np.deg2rad(self.acceptance_angle)))

    @property
    def collection_area(self):
        return np.pi * self.radius ** 2

    def rays(self):
        if not 0.0 < self.acceptance_angle <= 90.0:
            raise ValueError("Acceptance angle must be in the range (0, 90] degrees.")
        if self.radius <= 0 or self.pixel_samples < 1:
            raise ValueError("The fibre radius and the number of pixel samples must be positive.")
        n = self.pixel_samples
        k = np.arange(n) + 0.5
        golden = np.pi * (3.0 - np.sqrt(5.0))
        cos_max = np.cos(np.deg2rad(self.acceptance_angle))
        cos_t = 1.0 - (1.0 - cos_max) * k / n                   # uniform in solid angle on the cap
        sin_t = np.sqrt(np.maximum(0.0, 1.0 - cos_t * cos_t))
        phi = golden * k
        dl = np.stack([sin_t * np.cos(phi), sin_t * np.sin(phi), cos_t], axis=1)
        rr = self.radius * np.sqrt(k / n)                       # uniform in area on the disc, decorrelated from the directions
        psi = golden * k * 7.0 + 1.0
        ol = np.stack([rr * np.cos(psi), rr * np.sin(psi), np.zeros(n)], axis=1)
        m = self.transform
        return ol @ m[:3, :3].T + m[:3, 3], dl @ m[:3, :3].T, cos_t


class _Observer0DGroup:
    """Observer0DGroup (group/base.py:60-436): a set of 0-D observers rendered together.  ``observe`` clips every observer's rays
    against the bounding primitive, renders ALL of them in one device call and reduces them per observer (weighted mean spectral
    radiance -> ``observer.spectrum`` and ``self.spectra[i]``)."""
    _OBSERVER_TYPE = object

    def __init__(self, observers=(), name="", transform=None):
        self.name = name
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        self._observers = ()
        self.spectra = None
        for o in observers:
            self.add_observer(o)

    @property
    def observers(self):
        return self._observers

    def add_observer(self, observer):
        if not isinstance(observer, self._OBSERVER_TYPE):
            raise ValueError("Can only add {} objects".format(self._OBSERVER_TYPE))
        self._observers = self._observers + (observer,)

    @property
    def names(self):
        return [o.name for o in self._observers]

    @names.setter
    def names(self, value):
        if not isinstance(value, (list, tuple)):
            raise TypeError("The names attribute must be a list or tuple.")
        if len(value) != len(self._observers):
            raise ValueError("The length of 'names' ({}) mismatches the number of observers ({}).".format(len(value), len(self._observers)))
        for o, v in zip(self._observers, value):
            o.name = v

    def _broadcast(self, attr, value):
        if isinstance(value, (list, tuple, np.ndarray)):
            if len(value) != len(self._observers):
                raise ValueError("The length of '{}' ({}) mismatches the number of observers ({}).".format(attr, len(value), len(self._observers)))
            for o, v in zip(self._observers, value):
                setattr(o, attr, v)
        else:
            for o in self._observers:
                setattr(o, attr, value)

    @property
    def pipelines(self):
        """The pipelines of every observer (a list of lists, group/base.py:383-396)."""
        return [ob.pipelines for ob in self._observers]

    @pipelines.setter
    def pipelines(self, pipelist):
        if len(pipelist) == len(self._observers):
            for ob, p in zip(self._observers, pipelist):
                ob.pipelines = p
        else:
            raise ValueError('Length of pipelines list do not match number of observers in the group.')

    def connect_pipelines(self, pipeline_classes, keywords_list=None, suppress_display_progress=True):
        """A new set of the given pipeline classes on every observer (group/base.py:398-436)."""
        if keywords_list is None:
            keywords_list = [dict() for _ in pipeline_classes]
        if len(pipeline_classes) != len(keywords_list):
            raise ValueError('The number of given pipeline classes does not match the number of dicts in keyword list.')
        for ob in self._observers:
            pipelines = []
            for cls, kwargs in zip(pipeline_classes, keywords_list):
                p = cls(**kwargs)
                if suppress_display_progress:
                    try:
                        p.display_progress = False
                    except AttributeError:
                        pass
                pipelines.append(p)
            ob.pipelines = pipelines

    def gather_rays(self):
        """(origins, directions, weights, owner index) of every observer's rays, in world space."""
        os_, ds_, ws_, owner = [], [], [], []
        g = self.transform
        for i, ob in enumerate(self._observers):
            o, d, w = ob.rays()
            os_.append(o @ g[:3, :3].T + g[:3, 3])
            ds_.append(d @ g[:3, :3].T)
            ws_.append(w)
            owner.append(np.full(o.shape[0], i))
        return np.concatenate(os_), np.concatenate(ds_), np.concatenate(ws_), np.concatenate(owner)

    def _descs(self):
        """cb2_observer0d array of the group (group transform applied) and the per-observer ray offsets / etendues."""
        n = len(self._observers)
        arr = (_abi.Observer0DDesc * n)()
        offs = np.zeros(n + 1, dtype=np.int64)
        etendue = np.ones(n, dtype=np.float64)
        g = self.transform
        for i, ob in enumerate(self._observers):
            m = g @ ob.transform
            for r in range(3):
                for c in range(4):
                    arr[i].to_world[4 * r + c] = m[r, c]
            if isinstance(ob, FibreOptic):
                if not 0.0 < ob.acceptance_angle <= 90.0:
                    raise ValueError("Acceptance angle must be in the range (0, 90] degrees.")
                if ob.radius <= 0 or ob.pixel_samples < 1:
                    raise ValueError("The fibre radius and the number of pixel samples must be positive.")
                arr[i].radius, arr[i].acceptance_angle, arr[i].samples = ob.radius, ob.acceptance_angle, ob.pixel_samples
                etendue[i] = ob.solid_angle * ob.collection_area
            else:
                arr[i].radius, arr[i].acceptance_angle, arr[i].samples = 0.0, 0.0, 1
                etendue[i] = ob.sensitivity
            offs[i + 1] = offs[i] + arr[i].samples
        return arr, offs, etendue

    def observe(self, scene, primitive, to_world=None, wall=None, device=0):
        """The whole group in one pass on the device: ray generation (cb2_observer0d_rays_device), optional first-wall clip, one
        render of all rays, per-observer reduction (cb2_observer0d_reduce_device); only the [n_observers, bins] results cross PCIe.
        ``scene``: engine.EmissionScene / PlasmaRenderer; ``wall``: optional first_wall.FirstWall that ends every ray at its first
        hit.  Returns spectra[n_observers, bins] (mean spectral radiance, W / (m^2 sr nm)); ``self.power_spectra`` holds the
        spectral power [W / nm] a SpectralPowerPipeline0D reports: the mean of L cos(theta) over the uniform solid-angle samples
        times the etendue (solid angle x collection area); a sight line scales by its sensitivity."""
        import torch
        if not self._observers:
            raise ValueError("The group has no observers.")
        lib = _abi.load_library()
        arr, offs, etendue = self._descs()
        n_obs, n_rays = len(self._observers), int(offs[-1])
        dev = torch.device("cuda", int(device))
        bins = scene.scene.bins if hasattr(scene, "scene") else scene.bins
        prim = _primitive_desc(primitive, to_world)
        buf = DeviceRayBuffer(n_rays, dev)
        weight = torch.empty(n_rays, dtype=torch.float64, device=dev)
        rs = buf.as_struct()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _abi.check(lib, lib.cb2_observer0d_rays_device(arr, n_obs, C.byref(prim), C.byref(rs), C.c_void_p(weight.data_ptr()), C.c_void_p(stream)))
            buf.n_rays, buf.n_segments = int(rs.n_rays), int(rs.n_segments)
            if wall is not None:
                wall.clip_device(buf)
            per_ray = torch.zeros((n_rays, bins), dtype=torch.float64, device=dev)
            stats = torch.zeros(8, dtype=torch.int64, device=dev)
            scene.render_device(buf, per_ray, stats=stats)
            out = torch.empty((2, n_obs, bins), dtype=torch.float64, device=dev)
            _abi.check(lib, lib.cb2_observer0d_reduce_device(C.c_void_p(per_ray.data_ptr()), 1, C.c_void_p(weight.data_ptr()),
                                                             offs.ctypes.data_as(_abi.c_int64_p), etendue.ctypes.data_as(_abi.c_double_p),
                                                             n_obs, bins, C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                                             C.c_void_p(stream)))
            host = out.cpu().numpy()
            ood = int(stats[5].item())
        if ood > 0:
            raise ValueError("The specified value is outside of the range of the supplied data and/or extrapolation range: %d table "
                             "lookups of this render left their tables." % ood)
        self.spectra, self.power_spectra = host[0], host[1]
        self.last_rays = buf
        grid = (scene.scene if hasattr(scene, "scene") else scene).flat.desc.grid
        for i, (ob, s, pw) in enumerate(zip(self._observers, self.spectra, self.power_spectra)):
            ob.spectrum, ob.power_spectrum = s, pw
            for p in ob.pipelines:                      # result packing into the observer's pipelines
                p._fill(grid.min_wavelength, grid.max_wavelength, s, pw, offs[i + 1] - offs[i])
        return self.spectra

    def observe_host_rays(self, scene, primitive, to_world=None, wall=None):
        """The same observation from host-generated rays (gather_rays + geometry.ray_segments) through the host-buffer render call and a
        numpy reduction — the explicit-ray-list path; tests compare the device pass with it."""
        from .geometry import ray_segments
        if not self._observers:
            raise ValueError("The group has no observers.")
        o, d, w, owner = self.gather_rays()
        rays = ray_segments(primitive, o, d, to_world)
        if wall is not None:
            rays = wall.clip(rays)
        per_ray, _ = scene.render(rays)
        n = len(self._observers)
        num = np.zeros((n, per_ray.shape[1]))
        den = np.zeros(n)
        np.add.at(num, owner, per_ray * w[:, None])
        np.add.at(den, owner, w)
        counts = np.bincount(owner, minlength=n).astype(np.float64)
        etendue = np.array([getattr(ob, "solid_angle", 1.0) * getattr(ob, "collection_area", 1.0) * getattr(ob, "sensitivity", 1.0)
                            for ob in self._observers])
        return num / den[:, None], num / counts[:, None] * etendue[:, None]


class SightLineGroup(_Observer0DGroup):
    _OBSERVER_TYPE = SightLine

    @property
    def sensitivity(self):
        return [o.sensitivity for o in self._observers]

    @sensitivity.setter
    def sensitivity(self, value):
        self._broadcast("sensitivity", value)


class FibreOpticGroup(_Observer0DGroup):
    _OBSERVER_TYPE = FibreOptic

    @property
    def acceptance_angle(self):
        return [o.acceptance_angle for o in self._observers]

    @acceptance_angle.setter
    def acceptance_angle(self, value):
        self._broadcast("acceptance_angle", value)

    @property
    def radius(self):
        return [o.radius for o in self._observers]

    @radius.setter
    def radius(self, value):
        self._broadcast("radius", value)

    @property
    def pixel_samples(self):
        return [o.pixel_samples for o in self._observers]

    @pixel_samples.setter
    def pixel_samples(self, value):
        self._broadcast("pixel_samples", value)
