"""Scene flattening: Plasma + models + AtomicData -> the flat ``cb2_scene_desc`` of include/cherab_b200.h.

Does once per scene what the reference does lazily per model on the first ``emission()`` call
(``_populate_cache``: impact_excitation.pyx:102-128, recombination.pyx:102-131, bremsstrahlung.pyx:210-234) and what
``Plasma._configure_geometry`` wires up (plasma/node.pyx:511-543): resolves target species, rate tables, wavelengths,
line-shape parameters and the plasma<-world transform.  Unknown model / field classes raise TypeError — there is no
CPU fallback to hand them to.
"""
import ctypes as C

import numpy as np

from . import _abi
from .atomic import ConstantRate, RateTable, RateTable3D
from .models import Bremsstrahlung, ThermalCXLine, TotalRadiatedPower, _LineModel
from .plasma import AxisymBlend, AxisymBlendVector, EFITMagneticField, _as_scalar_field, _as_vector_field


def affine_inverse(m):
    m = np.asarray(m, dtype=np.float64)
    if m.shape == (3, 4):
        m = np.vstack([m, [0, 0, 0, 1]])
    return np.linalg.inv(m)


class FlatScene:
    """Owns the ctypes descriptor and keeps every numpy array it points at alive."""

    def __init__(self):
        self.desc = _abi.SceneDesc()
        self.keep = []
        self.n_gaussian_components = 0


def _fill_rate(r, rate, keep):
    if rate is None:
        # a Null rate (OpenADAS(missing_rates_return_null=True), openadas/rates/pec.pyx:80-90): always zero
        r.n_ne = r.n_te = 0
        r.constant = 0.0
        r.extrapolate = 1
    elif isinstance(rate, ConstantRate):
        r.n_ne = r.n_te = 0
        r.constant = rate.value
        r.extrapolate = 1
    elif isinstance(rate, RateTable):
        r.n_ne, r.n_te = rate.ne.size, rate.te.size
        keep.extend([rate.ne, rate.te, rate.rate])
        r.ne = rate.ne.ctypes.data_as(_abi.c_double_p)
        r.te = rate.te.ctypes.data_as(_abi.c_double_p)
        r.rate = rate.rate.ctypes.data_as(_abi.c_double_p)
        r.extrapolate = 1 if rate.extrapolate else 0
    else:
        raise TypeError("Unsupported rate object %r (expected RateTable or ConstantRate)" % (rate,))


def _fill_rate3(r, rate, keep):
    if isinstance(rate, ConstantRate):
        r.n_ne = r.n_te = r.n_td = 0
        r.constant = rate.value
        r.extrapolate = 1
    elif rate is None:                                   # NullThermalCXPEC (pec.pyx:197-205)
        r.n_ne = r.n_te = r.n_td = 0
        r.constant = 0.0
        r.extrapolate = 1
    elif isinstance(rate, RateTable3D):
        r.n_ne, r.n_te, r.n_td = rate.ne.size, rate.te.size, rate.td.size
        keep.extend([rate.ne, rate.te, rate.td, rate.rate])
        r.ne = rate.ne.ctypes.data_as(_abi.c_double_p)
        r.te = rate.te.ctypes.data_as(_abi.c_double_p)
        r.td = rate.td.ctypes.data_as(_abi.c_double_p)
        r.rate = rate.rate.ctypes.data_as(_abi.c_double_p)
        r.extrapolate = 1 if rate.extrapolate else 0
    else:
        raise TypeError("Unsupported thermal CX rate object %r" % (rate,))


def flatten_scene(plasma, min_wavelength, max_wavelength, bins, quad_rtol=1e-5, quad_min_order=1, quad_max_order=50,
                  brems_quadrature=0):
    """Flatten ``plasma`` (with its models) for a Spectrum(min_wavelength, max_wavelength, bins)."""
    if plasma.electron_distribution is None:
        raise RuntimeError("The plasma must have a defined electron distribution.")
    if plasma.atomic_data is None and any(isinstance(m, (_LineModel, Bremsstrahlung)) for m in plasma.models):
        raise RuntimeError("The plasma must have an atomic data source to be used with an emission model.")
    if bins < 1 or not max_wavelength > min_wavelength:
        raise ValueError("Invalid spectral range.")
    fs = FlatScene()
    d, keep = fs.desc, fs.keep
    d.abi_version = _abi.ABI_VERSION
    d.step, d.min_samples = plasma.integrator.step, plasma.integrator.min_samples
    d.grid.min_wavelength, d.grid.max_wavelength, d.grid.bins = float(min_wavelength), float(max_wavelength), int(bins)
    d.quad_rtol, d.quad_min_order, d.quad_max_order = quad_rtol, quad_min_order, quad_max_order
    d.brems_quadrature = brems_quadrature

    # world -> plasma space: the composition of Raysect's world_to_primitive with PlasmaMaterial's local_to_plasma
    # (material.pyx:55-57) is the inverse of the Plasma node's own transform, whatever the geometry_transform is
    w2p = np.eye(4) if plasma.transform is None else affine_inverse(plasma.transform)
    for i in range(3):
        for j in range(4):
            d.world_to_plasma[4 * i + j] = w2p[i, j]

    uses_axisym = False
    _as_scalar_field(plasma.electron_distribution.density)._fill(d.electron_density, keep)
    _as_scalar_field(plasma.electron_distribution.temperature)._fill(d.electron_temperature, keep)
    uses_axisym |= isinstance(plasma.electron_distribution.density, AxisymBlend) or isinstance(plasma.electron_distribution.temperature, AxisymBlend)

    species = list(plasma.composition)
    sp_arr = (_abi.SpeciesDesc * max(1, len(species)))()
    for i, s in enumerate(species):
        sp_arr[i].charge = s.charge
        sp_arr[i].atomic_weight = s.element.atomic_weight
        dist = s.distribution
        _as_scalar_field(dist.density)._fill(sp_arr[i].density, keep)
        _as_scalar_field(dist.temperature)._fill(sp_arr[i].temperature, keep)
        _as_vector_field(dist.velocity)._fill(sp_arr[i].velocity, keep)
        uses_axisym |= isinstance(dist.density, AxisymBlend) or isinstance(dist.temperature, AxisymBlend) \
            or isinstance(dist.velocity, AxisymBlendVector)
    keep.append(sp_arr)
    d.n_species = len(species)
    d.species = C.cast(sp_arr, C.POINTER(_abi.SpeciesDesc))

    if isinstance(plasma.b_field, EFITMagneticField):
        d.b_field_kind = 1
        uses_axisym = True
    else:
        b = _as_vector_field(plasma.b_field)
        if not hasattr(b, "v"):
            raise TypeError("Unsupported b_field for the B200 path: %r" % (plasma.b_field,))
        d.b_field_kind = 0
        d.b_field[0], d.b_field[1], d.b_field[2] = b.v

    if uses_axisym:
        if plasma.axisym is None:
            raise RuntimeError("AxisymBlend fields need plasma.axisym (an AxisymContext).")
        ax = _abi.Axisym()
        plasma.axisym._fill(ax, keep)
        keep.append(ax)
        d.axisym = C.pointer(ax)

    models = list(plasma.models)
    mo_arr = (_abi.ModelDesc * max(1, len(models)))()
    need_gaunt = None
    for i, m in enumerate(models):
        mo = mo_arr[i]
        mo.kind = m.kind if m.kind is not None else -1
        if isinstance(m, _LineModel):
            atomic = m.atomic_data or plasma.atomic_data
            index, rate, wavelength, shape = m.populate(plasma, atomic)
            mo.species = index
            mo.wavelength = wavelength
            mo.atomic_weight = m.line.element.atomic_weight
            if isinstance(m, ThermalCXLine):
                donors = m.donors(plasma, atomic)
                ext = _abi.ModelExt()
                ext.n_donors = len(donors)
                idx = np.ascontiguousarray([i for i, _ in donors], dtype=np.int32)
                rates = (_abi.Rate3D * max(1, len(donors)))()
                for k, (_, r3) in enumerate(donors):
                    _fill_rate3(rates[k], r3, keep)
                ext.donor_species = idx.ctypes.data_as(_abi.c_int32_p)
                ext.donor_rates = C.cast(rates, C.POINTER(_abi.Rate3D))
                keep.extend([idx, rates, ext])
                mo.ext = C.pointer(ext)
                mo.pec.n_ne = mo.pec.n_te = 0
                mo.pec.constant = 0.0
            else:
                _fill_rate(mo.pec, rate, keep)
            shape._fill(mo.shape, keep)
            keep.append(shape)
        elif isinstance(m, TotalRadiatedPower):
            atomic = m.atomic_data or plasma.atomic_data
            i_line, i_recom, hyd, plt, prb, prc = m.populate(plasma, atomic)
            ext = _abi.ModelExt()
            ext.line_rad_species, ext.recom_species = i_line, i_recom
            hidx = np.ascontiguousarray(hyd, dtype=np.int32)
            ext.n_hydrogen = hidx.size
            ext.hydrogen_species = hidx.ctypes.data_as(_abi.c_int32_p)
            for name, rate in (("plt", plt), ("prb", prb), ("prc", prc)):
                setattr(ext, "has_" + name, 0 if rate is None else 1)
                if rate is not None:
                    _fill_rate(getattr(ext, name), rate, keep)
            keep.extend([hidx, ext])
            mo.species = -1
            mo.ext = C.pointer(ext)
        elif isinstance(m, Bremsstrahlung):
            mo.species = -1
            need_gaunt = m.gaunt_factor or (m.atomic_data or plasma.atomic_data).free_free_gaunt_factor()
        else:
            raise TypeError("Unsupported PlasmaModel for the B200 path: %r" % (m,))
    keep.append(mo_arr)
    d.n_models = len(models)
    d.models = C.cast(mo_arr, C.POINTER(_abi.ModelDesc))

    if need_gaunt is not None:
        u, g2, gff = (np.ascontiguousarray(a, dtype=np.float64) for a in need_gaunt)
        keep.extend([u, g2, gff])
        d.gaunt.n_u, d.gaunt.n_gamma2 = u.size, g2.size
        d.gaunt.u = u.ctypes.data_as(_abi.c_double_p)
        d.gaunt.gamma2 = g2.ctypes.data_as(_abi.c_double_p)
        d.gaunt.gaunt = gff.ctypes.data_as(_abi.c_double_p)
    return fs


class RayBatch:
    """Ray segments in the layout of ``cb2_rays`` (what Raysect's tracer hands to VolumeIntegrator.integrate)."""

    def __init__(self, origin, direction, seg_offset, seg_t0, seg_t1):
        self.origin = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
        self.direction = np.ascontiguousarray(d / np.linalg.norm(d, axis=1, keepdims=True))
        self.seg_offset = np.ascontiguousarray(seg_offset, dtype=np.int64)
        self.seg_t0 = np.ascontiguousarray(seg_t0, dtype=np.float64)
        self.seg_t1 = np.ascontiguousarray(seg_t1, dtype=np.float64)
        if self.seg_offset.size != self.origin.shape[0] + 1 or self.seg_offset[-1] != self.seg_t0.size:
            raise ValueError("seg_offset must have n_rays+1 entries ending at n_segments")

    @property
    def n_rays(self):
        return self.origin.shape[0]

    @property
    def n_segments(self):
        return self.seg_t0.size

    def pin(self):
        """Move the arrays into page-locked host memory (torch's pinned allocator): the host-buffer render calls then upload them at
        PCIe speed instead of through the driver's staging buffer.  Returns self."""
        import torch
        self._pinned = []
        for name in ("origin", "direction", "seg_offset", "seg_t0", "seg_t1"):
            t = torch.from_numpy(getattr(self, name)).pin_memory()
            self._pinned.append(t)
            setattr(self, name, t.numpy())
        return self

    def as_struct(self):
        r = _abi.Rays()
        r.n_rays, r.n_segments = self.n_rays, self.n_segments
        r.origin = self.origin.ctypes.data_as(_abi.c_double_p)
        r.direction = self.direction.ctypes.data_as(_abi.c_double_p)
        r.seg_offset = self.seg_offset.ctypes.data_as(_abi.c_int64_p)
        r.seg_t0 = self.seg_t0.ctypes.data_as(_abi.c_double_p)
        r.seg_t1 = self.seg_t1.ctypes.data_as(_abi.c_double_p)
        return r

    def subset(self, index):
        """Rays ``index`` (array of ray ids) as a new batch."""
        index = np.asarray(index, dtype=np.int64)
        counts = self.seg_offset[index + 1] - self.seg_offset[index]
        offs = np.concatenate([[0], np.cumsum(counts)])
        sel = np.concatenate([np.arange(self.seg_offset[i], self.seg_offset[i + 1]) for i in index]) if index.size else np.zeros(0, np.int64)
        sel = sel.astype(np.int64)
        return RayBatch(self.origin[index], self.direction[index], offs, self.seg_t0[sel], self.seg_t1[sel])

    def sample_count(self, step, min_samples):
        """Sum over segments of intervals+1 with intervals = max(min_samples-1, ceil(L/step)) (the metric's unit)."""
        length = self.seg_t1 - self.seg_t0
        length = length[length > 0]
        iv = np.maximum(min_samples - 1, np.ceil(length / step)).astype(np.int64)
        return int((iv + 1).sum())
