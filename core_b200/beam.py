"""Neutral-beam objects of the emission path, host-side mirror with the reference's names and arguments:
Beam (cherab/core/beam/node.pyx:100-583), SingleRayAttenuator (cherab/core/model/attenuator/singleray.pyx:36-346),
BeamCXLine (cherab/core/model/beam/charge_exchange.pyx:36-374) and the rate tables of cherab/openadas/rates/{beam,cx}.pyx.
The arithmetic lives in the CUDA library; these classes carry what the reference resolves lazily (stopping rates for every
plasma species, the CX effective emission coefficient of the ground-state donor) and flatten it into ``cb2_beam_desc``.

Not yet on the device path: excited donor metastables (BeamPopulationRate) and BeamEmissionLine / the MSE multiplet.
"""
import ctypes as C

import numpy as np

from . import _abi
from .atomic import ConstantRate, Line
from .geometry import HollowCylinder, TruncatedCone
from .models import GaussianLine, LineShapeModel, PlasmaModel
from .notify import Notifier
from .plasma import NumericalIntegrator


class BeamStoppingTable:
    """BeamStoppingRate data dict (openadas/rates/beam.pyx:40-103): e [eV/amu], n [m^-3], t [eV], sen [N x M], st [K], sref.
    ``extrapolate``: 'linear' (2-D part) / 'quadratic' (1-D parts) beyond the tables (beam.pyx:73-84); without it a lookup
    outside is the reference's ValueError (counted on the device, raised by EmissionScene.render)."""

    def __init__(self, e, n, t, sen, st, sref, extrapolate=False):
        self.extrapolate = bool(extrapolate)
        self.e, self.n, self.t = (np.ascontiguousarray(a, dtype=np.float64) for a in (e, n, t))
        self.sen = np.ascontiguousarray(sen, dtype=np.float64)
        self.st = np.ascontiguousarray(st, dtype=np.float64)
        self.sref = float(sref)
        if self.sen.shape != (self.e.size, self.n.size) or self.st.shape != (self.t.size,):
            raise ValueError("sen must have shape (len(e), len(n)) and st shape (len(t),)")


class BeamCXTable:
    """BeamCXPEC data dict (openadas/rates/cx.pyx:66-103): eb, ti, ni, z, b grids with qeb, qti, qni, qz, qb and qref.
    ``extrapolate``: 'quadratic' in log10 E, 'nearest' for the four factors (cx.pyx:96-102)."""

    def __init__(self, donor_metastable, eb, ti, ni, z, b, qeb, qti, qni, qz, qb, qref, extrapolate=False):
        self.extrapolate = bool(extrapolate)
        self.donor_metastable = int(donor_metastable)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.eb, self.ti, self.ni, self.z, self.b = f(eb), f(ti), f(ni), f(z), f(b)
        self.qeb, self.qti, self.qni, self.qz, self.qb = f(qeb), f(qti), f(qni), f(qz), f(qb)
        self.qref = float(qref)
        for g, q in ((self.eb, self.qeb), (self.ti, self.qti), (self.ni, self.qni), (self.z, self.qz), (self.b, self.qb)):
            if g.shape != q.shape or g.ndim != 1 or g.size < 1:
                raise ValueError("every CX grid needs a matching 1-D rate array")


class ConstantBeamCXPEC(ConstantRate):
    """Constant effective CX emission coefficient in W m^3 (core/tests/test_beamcxline.py:34-46)."""

    def __init__(self, donor_metastable, value):
        super().__init__(value)
        self.donor_metastable = int(donor_metastable)


class SingleRayAttenuator:
    """singleray.pyx:36-105: attenuation along the beam axis, Gaussian across it."""

    def __init__(self, step=0.01, clamp_to_zero=False, clamp_sigma=5.0, beam=None, plasma=None, atomic_data=None):
        if step <= 0.0:
            raise ValueError("The step size must be greater than zero.")
        if clamp_sigma <= 0.0:
            raise ValueError("The value of clamp_sigma must be greater than zero.")
        self.step, self.clamp_to_zero, self.clamp_sigma = float(step), bool(clamp_to_zero), float(clamp_sigma)
        self.beam, self.plasma, self.atomic_data = beam, plasma, atomic_data


class BeamModel:
    def __init__(self, beam=None, plasma=None, atomic_data=None):
        self.beam, self.plasma, self.atomic_data = beam, plasma, atomic_data


class BeamCXLine(BeamModel):
    """charge_exchange.pyx:36-115: line of (element, charge) emitted after charge exchange between the beam atoms and the
    receiver (element, charge + 1)."""
    kind = _abi.MODEL_BEAM_CX_LINE

    def __init__(self, line, beam=None, plasma=None, atomic_data=None, lineshape=None, lineshape_args=None, lineshape_kwargs=None):
        super().__init__(beam, plasma, atomic_data)
        if not isinstance(line, Line):
            raise TypeError("line must be a Line")
        self.line = line
        self.lineshape_class = lineshape or GaussianLine
        if not (isinstance(self.lineshape_class, type) and issubclass(self.lineshape_class, LineShapeModel)):
            raise TypeError("The attribute lineshape must be a subclass of LineShapeModel.")
        self.lineshape_args = list(lineshape_args) if lineshape_args else []
        self.lineshape_kwargs = dict(lineshape_kwargs) if lineshape_kwargs else {}

    def populate(self, beam, plasma, atomic_data):
        """charge_exchange.pyx:294-352 -> (receiver index, [cx rates by metastable], wavelength, lineshape)."""
        if beam is None:
            raise RuntimeError("The emission model is not connected to a beam object.")
        if plasma is None:
            raise RuntimeError("The emission model is not connected to a plasma object.")
        if atomic_data is None:
            raise RuntimeError("The emission model is not connected to an atomic data source.")
        element, charge = self.line.element, self.line.charge + 1
        try:
            index = plasma.composition.index(element, charge)
        except ValueError:
            raise RuntimeError("The plasma object does not contain the ion species for the specified CX line "
                               "(element={}, ionisation={}).".format(element.symbol, charge))
        rates = list(atomic_data.beam_cx_pec(beam.element, element, charge, self.line.transition))
        ground = [r for r in rates if r.donor_metastable == 1]
        if len(ground) != 1:
            raise RuntimeError("beam_cx_pec must return exactly one rate for the ground-state donor (metastable 1).")
        excited = [r for r in rates if r.donor_metastable != 1]
        if len(excited) > 3:
            raise ValueError("At most three excited donor metastables are supported.")
        # charge_exchange.pyx:341-361: every excited rate is linked with the population coefficients of all plasma species
        self.population = [[atomic_data.beam_population_rate(beam.element, r.donor_metastable, sp.element, sp.charge)
                            for sp in plasma.composition] for r in excited]
        rates = ground + excited
        wavelength = atomic_data.wavelength(element, charge - 1, self.line.transition)
        species = plasma.composition.get(element, charge)
        shape = self.lineshape_class(self.line, wavelength, species, plasma, atomic_data, *self.lineshape_args, **self.lineshape_kwargs)
        return index, rates, wavelength, shape

    def __repr__(self):
        return "<BeamCXLine: element={}, charge={}, transition={}>".format(self.line.element.name, self.line.charge, self.line.transition)


class BeamEmissionLine(BeamModel):
    """beam_emission.pyx:36-98: Balmer-series emission of the beam atoms excited by the plasma, with the motional-Stark-effect
    multiplet line shape (mse.pyx).  The intensity ratios are constants or, as in the reference (mse.pyx:103-121), callables:
    ``sigma_to_pi(ne, beam_energy)`` and ``sigma1_to_sigma0(ne)``, ``pi2_to_pi3(ne)``, ``pi4_to_pi3(ne)``.  Callables are tabulated by
    the flattener on ``ratio_knots`` points uniform in log10(ne) over ``ratio_density_range`` (linear interpolation on the device)."""
    kind = _abi.MODEL_BEAM_EMISSION_LINE

    def __init__(self, line, beam=None, plasma=None, atomic_data=None, sigma_to_pi=0.56, sigma1_to_sigma0=0.7060001671878492,
                 pi2_to_pi3=0.3140003593919741, pi4_to_pi3=0.7279994935840365):
        super().__init__(beam, plasma, atomic_data)
        if not isinstance(line, Line):
            raise TypeError("line must be a Line")
        self.line = line
        self.ratio_functions = (sigma_to_pi, sigma1_to_sigma0, pi2_to_pi3, pi4_to_pi3)
        self.ratio_density_range, self.ratio_knots = (1e16, 1e22), 385
        self.ratios = tuple(0.0 if callable(v) else float(v) for v in self.ratio_functions)

    def ratio_table(self, beam_energy):
        """None when every ratio is a constant, else (log10 ne of the first knot, knot spacing, table[4][knots])."""
        if not any(callable(v) for v in self.ratio_functions):
            return None
        lo, hi = np.log10(self.ratio_density_range[0]), np.log10(self.ratio_density_range[1])
        lne = np.linspace(lo, hi, self.ratio_knots)
        tab = np.empty((4, self.ratio_knots))
        for k, v in enumerate(self.ratio_functions):
            if callable(v):
                tab[k] = [v(n, beam_energy) if k == 0 else v(n) for n in 10.0 ** lne]
            else:
                tab[k] = float(v)
        return float(lo), float(lne[1] - lne[0]), np.ascontiguousarray(tab)

    def populate(self, beam, plasma, atomic_data):
        """beam_emission.pyx:178-216 -> (wavelength, [(species index, rate)])."""
        if beam is None:
            raise RuntimeError("The emission model is not connected to a beam object.")
        if plasma is None:
            raise RuntimeError("The emission model is not connected to a plasma object.")
        if atomic_data is None:
            raise RuntimeError("The emission model is not connected to an atomic data source.")
        if beam.element is not self.line.element:
            raise TypeError("The specified line element '{}' is incompatible with the attached neutral "
                            "beam element '{}'.".format(self.line.element.symbol, beam.element.symbol))
        if self.line.charge != 0:
            raise TypeError("The transition specified does not belong to a neutral atom.")
        wavelength = atomic_data.wavelength(beam.element, 0, self.line.transition)
        rates = [(i, atomic_data.beam_emission_pec(beam.element, sp.element, sp.charge, self.line.transition))
                 for i, sp in enumerate(plasma.composition)]
        return wavelength, rates

    def __repr__(self):
        return "<BeamEmissionLine: element={}, charge={}, transition={}>".format(self.line.element.name, self.line.charge, self.line.transition)


class Beam:
    """beam/node.pyx:100-212, Raysect-free: ``transform`` is the beam -> world 4x4 matrix (z along the beam axis, origin at
    the source).  Defaults as in the reference: energy 0 eV/amu, power 0 W, temperature 0 eV, sigma 0.1 m, no divergence,
    length 1 m, NumericalIntegrator(step=0.001)."""

    _NOTIFYING = ("transform", "energy", "power", "temperature", "element", "sigma", "divergence_x", "divergence_y", "length",
                  "plasma", "atomic_data", "attenuator", "integrator", "models")

    def __init__(self, transform=None, name="Beam"):
        self.__dict__["notifier"] = Notifier()
        self.name = name
        self.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=np.float64)
        self.energy, self.power, self.temperature = 0.0, 0.0, 0.0
        self.element = None
        self.sigma, self.divergence_x, self.divergence_y, self.length = 0.1, 0.0, 0.0, 1.0
        self.plasma, self.atomic_data, self.attenuator = None, None, None
        self.models = []
        self.integrator = NumericalIntegrator(step=0.001)

    def __setattr__(self, name, value):
        if name == "energy" and value < 0:
            raise ValueError("Beam energy cannot be less than zero.")
        if name == "power" and value < 0:
            raise ValueError("Beam power cannot be less than zero.")
        if name == "sigma" and value <= 0:
            raise ValueError("Beam sigma (width) cannot be less than or equal to zero.")
        if name in ("divergence_x", "divergence_y") and value < 0:
            raise ValueError("Beam divergence cannot be less than zero.")
        if name == "length" and value <= 0:
            raise ValueError("Beam length cannot be less than or equal to zero.")
        object.__setattr__(self, name, value)
        if name in self._NOTIFYING:
            self.notifier.notify()

    @property
    def geometry(self):
        """Bounding primitive in the beam frame, chosen exactly as Beam._generate_geometry does (beam/node.pyx:505-554): a
        cylinder of radius clamp_sigma * sigma without divergence — and also when a cone would save less than 10 % of the
        volume — else the cone from clamp_sigma * sigma at the source to clamp_sigma * sqrt(sigma^2 + (length tan(div))^2)."""
        if self.attenuator is None:
            raise ValueError("The beam must have an attenuator model to provide density values.")
        ns = self.attenuator.clamp_sigma
        if self.divergence_x == 0 and self.divergence_y == 0:
            return HollowCylinder(0.0, ns * self.sigma, 0.0, self.length)
        drdz = np.tan(np.deg2rad(max(self.divergence_x, self.divergence_y)))
        radius_start = ns * self.sigma
        radius_end = ns * np.sqrt(self.sigma ** 2 + self.length ** 2 * drdz ** 2)
        distance_apex = radius_start * self.length / (radius_end - radius_start)
        cone_height = self.length + distance_apex
        cylinder_volume = self.length * np.pi * radius_end ** 2
        cone_volume = np.pi * (cone_height * radius_end ** 2 - distance_apex * radius_start ** 2) / 3
        if cone_volume / cylinder_volume > 0.9:
            return HollowCylinder(0.0, ns * self.sigma, 0.0, self.length)
        return TruncatedCone(radius_start, radius_end, self.length)


def _fill_beam_rate(r, rate, keep):
    if rate is None:
        r.n_e = r.n_n = r.n_t = 0
        r.constant = 0.0                       # NullBeamStoppingRate
    elif isinstance(rate, ConstantRate):
        r.n_e = r.n_n = r.n_t = 0
        r.constant = rate.value
    elif isinstance(rate, BeamStoppingTable):
        r.n_e, r.n_n, r.n_t = rate.e.size, rate.n.size, rate.t.size
        r.extrapolate = 1 if rate.extrapolate else 0
        keep.append(rate)
        dp = lambda a: a.ctypes.data_as(_abi.c_double_p)
        r.e, r.n, r.t, r.sen, r.st, r.sref = dp(rate.e), dp(rate.n), dp(rate.t), dp(rate.sen), dp(rate.st), rate.sref
    else:
        raise TypeError("Unsupported beam stopping rate object %r" % (rate,))


def _fill_cx_rate(r, rate, keep):
    if isinstance(rate, ConstantRate):
        r.n_eb = 0
        r.constant = rate.value
    elif isinstance(rate, BeamCXTable):
        keep.append(rate)
        dp = lambda a: a.ctypes.data_as(_abi.c_double_p)
        r.n_eb, r.n_ti, r.n_ni, r.n_z, r.n_b = rate.eb.size, rate.ti.size, rate.ni.size, rate.z.size, rate.b.size
        r.extrapolate = 1 if rate.extrapolate else 0
        r.eb, r.ti, r.ni, r.z, r.b = dp(rate.eb), dp(rate.ti), dp(rate.ni), dp(rate.z), dp(rate.b)
        r.qeb, r.qti, r.qni, r.qz, r.qb = dp(rate.qeb), dp(rate.qti), dp(rate.qni), dp(rate.qz), dp(rate.qb)
        r.qref = rate.qref
    else:
        raise TypeError("Unsupported beam CX rate object %r" % (rate,))


def flatten_beam_scene(beam, min_wavelength, max_wavelength, bins):
    """Flatten ``beam`` (with its models), the plasma it crosses and their atomic data for a Spectrum(min, max, bins):
    what BeamMaterial.__init__ (beam/material.pyx:31-47), SingleRayAttenuator._populate_stopping_data_cache
    (singleray.pyx:319-340) and every beam model's _populate_cache would resolve."""
    from .flatten import flatten_scene
    if beam.plasma is None:
        raise ValueError("The beam must have a reference to a plasma object to be used with an emission model.")
    if beam.atomic_data is None:
        raise ValueError("The beam must have an atomic data source to be used with an emission model.")
    if beam.attenuator is None:
        raise ValueError("The beam must have an attenuator model to provide density values.")
    if beam.element is None:
        raise ValueError("The beam must have an element.")
    plasma = beam.plasma
    # the plasma part of the scene (fields, species, transform) without the plasma's own models
    saved_models, saved_integrator = list(plasma.models), plasma.integrator
    try:
        plasma._models._models = []
        object.__setattr__(plasma, "integrator", beam.integrator)
        fs = flatten_scene(plasma, min_wavelength, max_wavelength, bins)
    finally:
        plasma._models._models = saved_models
        object.__setattr__(plasma, "integrator", saved_integrator)
    d, keep = fs.desc, fs.keep

    bd = _abi.BeamDesc()
    from .flatten import affine_inverse
    b2w = np.asarray(beam.transform, dtype=np.float64)
    w2p = np.eye(4) if plasma.transform is None else affine_inverse(plasma.transform)
    b2p = w2p @ b2w                                    # beam.to(plasma)
    w2b = affine_inverse(b2w)                          # the integrator marches in the beam primitive's local space
    for i in range(3):
        for j in range(4):
            bd.beam_to_plasma[4 * i + j] = b2p[i, j]
            d.world_to_plasma[4 * i + j] = w2b[i, j]
    bd.energy, bd.power, bd.temperature = beam.energy, beam.power, beam.temperature
    bd.atomic_weight = beam.element.atomic_weight
    bd.sigma, bd.divergence_x, bd.divergence_y, bd.length = beam.sigma, beam.divergence_x, beam.divergence_y, beam.length
    att = beam.attenuator
    bd.attenuator_step, bd.clamp_sigma, bd.clamp_to_zero = att.step, att.clamp_sigma, int(att.clamp_to_zero)
    species = list(plasma.composition)
    idx = np.arange(len(species), dtype=np.int32)
    rates = (_abi.BeamRate * max(1, len(species)))()
    for k, sp in enumerate(species):
        _fill_beam_rate(rates[k], beam.atomic_data.beam_stopping_rate(beam.element, sp.element, sp.charge), keep)
    bd.n_stopping = len(species)
    bd.stopping_species = idx.ctypes.data_as(_abi.c_int32_p)
    bd.stopping_rates = C.cast(rates, C.POINTER(_abi.BeamRate))
    keep.extend([idx, rates, bd])
    d.beam = C.pointer(bd)

    models = list(beam.models)
    mo_arr = (_abi.ModelDesc * max(1, len(models)))()
    for i, mdl in enumerate(models):
        mo = mo_arr[i]
        if isinstance(mdl, BeamEmissionLine):
            wavelength, bes = mdl.populate(beam, plasma, mdl.atomic_data or beam.atomic_data)
            mo.kind = mdl.kind
            mo.species = -1
            mo.wavelength = wavelength
            mo.atomic_weight = beam.element.atomic_weight       # the multiplet is broadened by the BEAM temperature (mse.pyx:100-102)
            mo.pec.n_ne = mo.pec.n_te = 0
            mo.pec.constant = 0.0
            mo.shape.kind = _abi.SHAPE_GAUSSIAN
            mo.shape.polarisation = _abi.POL_NO
            ext = _abi.ModelExt()
            bidx = np.ascontiguousarray([k for k, _ in bes], dtype=np.int32)
            barr = (_abi.BeamRate * max(1, len(bes)))()
            for k, (_, r) in enumerate(bes):
                _fill_beam_rate(barr[k], r, keep)
            ext.n_bes = len(bes)
            ext.bes_species = bidx.ctypes.data_as(_abi.c_int32_p)
            ext.bes_rates = C.cast(barr, C.POINTER(_abi.BeamRate))
            for k in range(4):
                ext.mse_ratios[k] = mdl.ratios[k]
            table = mdl.ratio_table(beam.energy)
            if table is not None:
                ext.mse_lne0, ext.mse_dlne, tab = table
                ext.n_mse = tab.shape[1]
                ext.mse_ratio_tab = tab.ctypes.data_as(_abi.c_double_p)
                keep.append(tab)
            keep.extend([bidx, barr, ext])
            mo.ext = C.pointer(ext)
            continue
        if not isinstance(mdl, BeamCXLine):
            raise TypeError("Unsupported BeamModel for the B200 path: %r" % (mdl,))
        index, cx_rates, wavelength, shape = mdl.populate(beam, plasma, mdl.atomic_data or beam.atomic_data)
        mo.kind = mdl.kind
        mo.species = index
        mo.wavelength = wavelength
        mo.atomic_weight = mdl.line.element.atomic_weight
        mo.pec.n_ne = mo.pec.n_te = 0
        mo.pec.constant = 0.0
        shape._fill(mo.shape, keep)
        ext = _abi.ModelExt()
        arr = (_abi.CXRate * max(1, len(cx_rates)))()
        for k, r in enumerate(cx_rates):
            _fill_cx_rate(arr[k], r, keep)
        ext.n_cx = len(cx_rates)
        ext.cx = C.cast(arr, C.POINTER(_abi.CXRate))
        if len(cx_rates) > 1:
            n_sp = len(list(plasma.composition))
            parr = (_abi.BeamRate * ((len(cx_rates) - 1) * n_sp))()
            for k, row in enumerate(mdl.population):
                for s_i, r in enumerate(row):
                    _fill_beam_rate(parr[k * n_sp + s_i], r, keep)
            ext.cx_population = C.cast(parr, C.POINTER(_abi.BeamRate))
            keep.append(parr)
        keep.extend([shape, arr, ext])
        mo.ext = C.pointer(ext)
    keep.append(mo_arr)
    d.n_models = len(models)
    d.models = C.cast(mo_arr, C.POINTER(_abi.ModelDesc))
    return fs


def beam_ray_segments(beam, origins, directions):
    """Chords of world-space rays through the beam's bounding primitive (what Raysect's tracer would hand to
    BeamMaterial's integrator)."""
    from .geometry import ray_segments
    return ray_segments(beam.geometry, origins, directions, beam.transform)
