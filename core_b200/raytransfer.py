"""Ray-transfer (geometry matrix) objects and pipelines: host-side mirror of cherab.tools.raytransfer.

RayTransferCylinder / RayTransferBox follow cherab/tools/raytransfer/raytransfer.py:129-268 (same arguments, default
step, eps-shrunk primitive, voxel_map/mask semantics of emitters.pyx:227-337); RayTransferPipeline0D/1D/2D follow
pipelines.py:73-240 (kind 'power'/'radiance', matrix = sum(samples * sensitivity) / pixel_samples).  The path-length
sampling itself (emitters.pyx:88-224) runs in the CUDA library; dense matrices are only produced on request, the
native output is CSR (the dense 512x512x320000 matrix of config C4 would be 671 TB).
"""
import ctypes as C

import numpy as np

from . import _abi
from .flatten import affine_inverse
from .geometry import Box, HollowCylinder


class RayTransferObject:
    """raytransfer.py:30-126."""

    def __init__(self, grid_shape, grid_steps, step, voxel_map=None, mask=None, transform=None, min_samples=2):
        if len(grid_shape) != 3:
            raise ValueError("Attribute 'grid_shape' must contain 3 elements.")
        if len(grid_steps) != 3:
            raise ValueError("Attribute 'grid_steps' must contain 3 elements.")
        for i in grid_shape:
            if i < 1:
                raise ValueError('Number of grid cells must be > 0.')
        for s in grid_steps:
            if s <= 0:
                raise ValueError('Grid steps must be > 0.')
        self.grid_shape = tuple(int(i) for i in grid_shape)
        self.grid_steps = tuple(float(s) for s in grid_steps)
        self.step = step
        self.min_samples = min_samples
        self.transform = transform
        # None: the emitter's own RayTransferIntegrator(step, min_samples) (emitters.pyx:88-224).  A plasma.NumericalIntegrator here is
        # the "foreign integrator" case of the reference — Raysect's trapezium rule over RayTransferEmitter.emission_function, unit
        # emissivity in the cell of every sample (emitters.pyx:452-473, 557-571) — with that integrator's step and min_samples
        self.integrator = None
        if voxel_map is None:
            self.mask = mask
        else:
            self.voxel_map = voxel_map

    @property
    def step(self):
        return self._step

    @step.setter
    def step(self, value):
        if value <= 0:
            raise ValueError("Numerical integration step size can not be less than or equal to zero.")
        self._step = float(value)

    @property
    def min_samples(self):
        return self._min_samples

    @min_samples.setter
    def min_samples(self, value):
        if value < 2:
            raise ValueError("At least two samples are required to perform the numerical integration.")
        self._min_samples = int(value)

    @property
    def bins(self):
        return self._bins

    @property
    def voxel_map(self):
        return self._voxel_map

    @voxel_map.setter
    def voxel_map(self, value):
        value = np.asarray(value)
        if value.shape != self.grid_shape:
            raise ValueError('Voxel_map array must be of shape: %s.' % (' '.join(['%d' % i for i in self.grid_shape])))
        self._voxel_map = np.ascontiguousarray(value.astype(np.int32))
        self._bins = int(self._voxel_map.max()) + 1

    @property
    def mask(self):
        return self._voxel_map > -1

    @mask.setter
    def mask(self, value):
        if value is not None:
            value = np.asarray(value)
            if value.shape != self.grid_shape:
                raise ValueError('Mask array must be of shape: %s.' % (' '.join(['%d' % i for i in self.grid_shape])))
            value = value.astype(bool)
        else:
            value = np.ones(self.grid_shape, dtype=bool)
        voxel_map = -1 * np.ones(value.shape, dtype=np.int32)
        voxel_map[value] = np.arange(value.sum(), dtype=np.int32)
        self._voxel_map = np.ascontiguousarray(voxel_map)
        self._bins = int(self._voxel_map.max()) + 1

    def invert_voxel_map(self):
        return [np.where(self._voxel_map == i) for i in range(self._bins)]

    def emission_function(self, point, direction, spectrum, *_unused):
        """RayTransferEmitter.emission_function (emitters.pyx:452-473, 557-571) for one point in the object's local space: unit
        emissivity in the light source the point's cell maps to — ``spectrum`` (array of ``bins`` samples) is added to and returned.
        The protocol method foreign integrators call; whole rays go through the device (``integrator`` attribute)."""
        x, y, z = (float(c) for c in point)
        if self.kind == _abi.RT_CYLINDRICAL:
            i2 = int(z / self.grid_steps[2])
            i0 = int((np.sqrt(x * x + y * y) - getattr(self, "rmin", 0.0)) / self.grid_steps[0])
            if self.grid_shape[1] == 1:
                i1 = 0
            else:
                phi = ((180.0 / np.pi) * np.arctan2(y, x) + 360.0) % getattr(self, "period", 360.0)
                i1 = int(phi / self.grid_steps[1])
        else:
            i0, i1, i2 = int(x / self.grid_steps[0]), int(y / self.grid_steps[1]), int(z / self.grid_steps[2])
        isource = self._voxel_map[i0, i1, i2]           # (IndexError outside the grid, as in the reference)
        if isource >= 0:
            spectrum[isource] += 1.0
        return spectrum

    def descriptor(self):
        """(cb2_rt_desc, keepalive) in the layout of include/cherab_b200.h."""
        d = _abi.RTDesc()
        d.abi_version = _abi.ABI_VERSION
        d.kind = self.kind
        for i in range(3):
            d.grid_shape[i] = self.grid_shape[i]
            d.grid_steps[i] = self.grid_steps[i]
        d.min_samples = self._min_samples
        d.rmin = getattr(self, "rmin", 0.0)
        d.period = getattr(self, "period", 360.0)
        d.step = self._step
        d.integrator = 0
        if self.integrator is not None:
            from .plasma import NumericalIntegrator
            if not isinstance(self.integrator, NumericalIntegrator):
                raise TypeError("integrator must be None (the ray-transfer integrator) or a NumericalIntegrator, not %r" % (self.integrator,))
            d.integrator = 1
            d.step = float(self.integrator.step)
            d.min_samples = max(2, int(self.integrator.min_samples))
        w2l = np.eye(4) if self.transform is None else affine_inverse(self.transform)
        for i in range(3):
            for j in range(4):
                d.world_to_local[4 * i + j] = w2l[i, j]
        vm = self._voxel_map
        d.voxel_map = vm.ctypes.data_as(_abi.c_int32_p)
        d.bins = self._bins
        return d, [vm]


class RayTransferCylinder(RayTransferObject):
    """raytransfer.py:129-200."""
    kind = _abi.RT_CYLINDRICAL

    def __init__(self, radius_outer, height, n_radius, n_height, radius_inner=0, n_polar=1, period=360., step=None,
                 voxel_map=None, mask=None, transform=None):
        num_sectors = 360. / period
        if abs(round(num_sectors) - num_sectors) > 1.e-3:
            raise ValueError("The period %.3f is not a multiple of 360." % period)
        dr = (radius_outer - radius_inner) / n_radius
        dz = height / n_height
        dphi = period / n_polar
        eps_r, eps_z = 1.e-5 * dr, 1.e-5 * dz
        self.rmin, self.period = float(radius_inner), float(period)
        step = step or 0.1 * min(dr, dz)
        super().__init__((n_radius, n_polar, n_height), (dr, dphi, dz), step, voxel_map, mask, transform)
        # Subtract(Cylinder(radius_outer - eps_r, height - eps_z), Cylinder(radius_inner + eps_r, height - eps_z))
        self.primitive = HollowCylinder(radius_inner + eps_r, radius_outer - eps_r, 0.0, height - eps_z)

    @property
    def dr(self):
        return self.grid_steps[0]

    @property
    def dphi(self):
        return self.grid_steps[1]

    @property
    def dz(self):
        return self.grid_steps[2]


class RayTransferBox(RayTransferObject):
    """raytransfer.py:203-268."""
    kind = _abi.RT_CARTESIAN

    def __init__(self, xmax, ymax, zmax, nx, ny, nz, step=None, voxel_map=None, mask=None, transform=None):
        dx, dy, dz = xmax / nx, ymax / ny, zmax / nz
        step = step or 0.1 * min(dx, dy, dz)
        super().__init__((nx, ny, nz), (dx, dy, dz), step, voxel_map, mask, transform)
        self.primitive = Box((0, 0, 0), (xmax - 1.e-5 * dx, ymax - 1.e-5 * dy, zmax - 1.e-5 * dz))


# ------------------------------------------------------------------------------------------------------------------
# pipelines (pipelines.py:28-240)
# ------------------------------------------------------------------------------------------------------------------
class RayTransferPipelineBase:
    def __init__(self, name=None, kind='power'):
        self.name = name
        self._matrix = None
        self._samples = 0
        self._bins = 0
        self.kind = kind

    @property
    def kind(self):
        return self._kind

    @kind.setter
    def kind(self, value):
        _kind = value.lower()
        if _kind in ('power', 'radiance'):
            self._kind = _kind
        else:
            raise ValueError("The kind property must be 'power' or 'radiance'.")

    @property
    def matrix(self):
        return self._matrix


class Spectrum:
    """The slice of raysect's Spectrum the pixel processors read: ``samples[bins]`` on [min_wavelength, max_wavelength]."""

    def __init__(self, min_wavelength, max_wavelength, bins):
        self.min_wavelength, self.max_wavelength, self.bins = float(min_wavelength), float(max_wavelength), int(bins)
        self.samples = np.zeros(self.bins)


class RayTransferPixelProcessorBase:
    """pipelines.py:213-222: accumulates the ray-transfer matrix row of one pixel."""

    def __init__(self, bins):
        self._matrix = np.zeros(bins)

    def pack_results(self):
        return (self._matrix, 0)


class RadianceRayTransferPixelProcessor(RayTransferPixelProcessorBase):
    """pipelines.py:225-231: path lengths in [m], the detector sensitivity is ignored."""

    def add_sample(self, spectrum, sensitivity):
        self._matrix += spectrum.samples


class PowerRayTransferPixelProcessor(RayTransferPixelProcessorBase):
    """pipelines.py:234-240: [m^3 sr], every sample multiplied by the detector sensitivity."""

    def add_sample(self, spectrum, sensitivity):
        self._matrix += spectrum.samples * sensitivity


class RayTransferPipeline0D(RayTransferPipelineBase):
    def __init__(self, name='RayTransferPipeline0D', kind='power'):
        super().__init__(name, kind)

    def initialise(self, min_wavelength, max_wavelength, spectral_bins, spectral_slices, quiet):
        self._samples = 0
        self._bins = spectral_bins
        self._matrix = np.zeros(spectral_bins)

    def pixel_processor(self, slice_id):
        return (PowerRayTransferPixelProcessor if self._kind == 'power' else RadianceRayTransferPixelProcessor)(self._bins)

    def update(self, slice_id, packed_result, pixel_samples):
        self._samples += pixel_samples
        self._matrix += packed_result[0]

    def finalise(self):
        self._matrix /= self._samples


class RayTransferPipeline1D(RayTransferPipelineBase):
    def __init__(self, name='RayTransferPipeline1D', kind='power'):
        super().__init__(name, kind)
        self._pixels = None

    def initialise(self, pixels, pixel_samples, min_wavelength, max_wavelength, spectral_bins, spectral_slices, quiet):
        self._pixels, self._samples, self._bins = pixels, pixel_samples, spectral_bins
        self._matrix = np.zeros((pixels, spectral_bins))

    def pixel_processor(self, pixel, slice_id):
        return (PowerRayTransferPixelProcessor if self._kind == 'power' else RadianceRayTransferPixelProcessor)(self._bins)

    def update(self, pixel, slice_id, packed_result):
        self._matrix[pixel] = packed_result[0] / self._samples

    def finalise(self):
        pass


class RayTransferPipeline2D(RayTransferPipelineBase):
    def __init__(self, name='RayTransferPipeline2D', kind='power'):
        super().__init__(name, kind)
        self._pixels = None

    def initialise(self, pixels, pixel_samples, min_wavelength, max_wavelength, spectral_bins, spectral_slices, quiet):
        self._pixels, self._samples, self._bins = pixels, pixel_samples, spectral_bins
        self._matrix = np.zeros((pixels[0], pixels[1], spectral_bins))

    def pixel_processor(self, x, y, slice_id):
        return (PowerRayTransferPixelProcessor if self._kind == 'power' else RadianceRayTransferPixelProcessor)(self._bins)

    def update(self, x, y, slice_id, packed_result):
        self._matrix[x, y] = packed_result[0] / self._samples

    def finalise(self):
        pass


# The pipelines and pixel processors above are protocol classes (initialise / pixel_processor / update / finalise, pack_results):
# where Cherab and Raysect are importable the reference's own classes replace them, so frames produced here feed the very objects a
# Cherab user holds (cherab/tools/raytransfer/pipelines.py:28-240, pixelprocessors.pyx).
try:
    from cherab.tools.raytransfer.pipelines import (RayTransferPipeline0D, RayTransferPipeline1D,     # noqa: F401,F811
                                                    RayTransferPipeline2D)
    from cherab.tools.raytransfer.pixelprocessors import (PowerRayTransferPixelProcessor,             # noqa: F401,F811
                                                          RadianceRayTransferPixelProcessor)
except Exception:
    pass
