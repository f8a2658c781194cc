"""Emission models and line shapes: host-side descriptors with the reference's names and constructor arguments.

cherab/core/model/plasma/impact_excitation.pyx:28-136, recombination.pyx:28-139, bremsstrahlung.pyx:93-245;
cherab/core/model/lineshape/{gaussian,multiplet,zeeman,stark}.pyx.  The arithmetic lives in the CUDA library; these
classes only carry what ``_populate_cache`` would resolve (target species, rate table, wavelength, shape parameters).
"""
import numpy as np

from . import _abi
from .atomic import Line


# ------------------------------------------------------------------------------------------------------------------
# line shapes
# ------------------------------------------------------------------------------------------------------------------
class LineShapeModel:
    """cherab/core/model/lineshape/base.pyx:25-45."""
    kind = None

    def __init__(self, line, wavelength, target_species, plasma, atomic_data):
        self.line, self.wavelength, self.target_species = line, wavelength, target_species
        self.plasma, self.atomic_data = plasma, atomic_data

    def _fill(self, sh, keep):
        sh.kind = self.kind
        sh.polarisation = _abi.POL_NO


class GaussianLine(LineShapeModel):
    """gaussian.pyx:93-139."""
    kind = _abi.SHAPE_GAUSSIAN


class MultipletLineShape(LineShapeModel):
    """multiplet.pyx:41-117."""
    kind = _abi.SHAPE_MULTIPLET

    def __init__(self, line, wavelength, target_species, plasma, atomic_data, multiplet):
        super().__init__(line, wavelength, target_species, plasma, atomic_data)
        multiplet = np.array(multiplet, dtype=np.float64)
        if not (len(multiplet.shape) == 2 and multiplet.shape[0] == 2):
            raise ValueError("The multiplet specification must be an array of shape (Nx2).")
        if not multiplet[1, :].sum() == 1.0:
            raise ValueError("The multiplet line ratios should sum to one.")
        self.multiplet = np.ascontiguousarray(multiplet)

    def _fill(self, sh, keep):
        super()._fill(sh, keep)
        sh.n_components = self.multiplet.shape[1]
        keep.append(self.multiplet)
        sh.multiplet = self.multiplet.ctypes.data_as(_abi.c_double_p)


def _polarisation(value):
    v = value.lower()
    if v == "pi":
        return _abi.POL_PI
    if v == "sigma":
        return _abi.POL_SIGMA
    if v == "no":
        return _abi.POL_NO
    raise ValueError('Select between "pi", "sigma" or "no", {} is unsupported.'.format(value))


class ZeemanLineShapeModel(LineShapeModel):
    """zeeman.pyx:42-87."""

    def __init__(self, line, wavelength, target_species, plasma, atomic_data, polarisation="no"):
        super().__init__(line, wavelength, target_species, plasma, atomic_data)
        self.polarisation = polarisation
        self._pol = _polarisation(polarisation)

    def _fill(self, sh, keep):
        sh.kind = self.kind
        sh.polarisation = self._pol


class ZeemanTriplet(ZeemanLineShapeModel):
    """zeeman.pyx:90-162."""
    kind = _abi.SHAPE_ZEEMAN_TRIPLET


class ParametrisedZeemanTriplet(ZeemanLineShapeModel):
    """zeeman.pyx:165-268."""
    kind = _abi.SHAPE_PARAM_ZEEMAN

    def __init__(self, line, wavelength, target_species, plasma, atomic_data, line_parameters=None, polarisation="no"):
        super().__init__(line, wavelength, target_species, plasma, atomic_data, polarisation)
        alpha, beta, gamma = line_parameters or atomic_data.zeeman_triplet_parameters(line)
        if alpha <= 0:
            raise ValueError("Parameter alpha must be positive.")
        if beta < 0:
            raise ValueError("Parameter beta must be non-negative.")
        self.parameters = (float(alpha), float(beta), float(gamma))

    def _fill(self, sh, keep):
        super()._fill(sh, keep)
        sh.param[0], sh.param[1], sh.param[2] = self.parameters


class ZeemanStructure:
    """cherab/core/atomic/zeeman.pyx:28-142: lists of (wavelength(B), ratio(B)) for pi / sigma+ / sigma- components.
    Each function may be a constant, a callable of B, or a (b, values) table; the flattener tabulates callables on
    ``b_grid`` (linear interpolation on device)."""

    def __init__(self, pi_components, sigma_plus_components, sigma_minus_components, b_grid=None):
        for name, comps in (("pi_components", pi_components), ("sigma_plus_components", sigma_plus_components),
                            ("sigma_minus_components", sigma_minus_components)):
            for c in comps:
                if len(c) != 2:
                    raise ValueError('Argument "%s" must be a list of 2-tuples.' % name)
        self.groups = (list(pi_components), list(sigma_plus_components), list(sigma_minus_components))
        self.b_grid = np.linspace(0.0, 20.0, 2001) if b_grid is None else np.ascontiguousarray(b_grid, dtype=np.float64)

    def tabulate(self):
        def tab(f):
            if callable(f):
                return np.array([f(b) for b in self.b_grid], dtype=np.float64)
            return np.full(self.b_grid.size, float(f))
        wl = [tab(c[0]) for g in self.groups for c in g]
        ra = [tab(c[1]) for g in self.groups for c in g]
        return np.ascontiguousarray(wl), np.ascontiguousarray(ra)


class ZeemanMultiplet(ZeemanLineShapeModel):
    """zeeman.pyx:271-365."""
    kind = _abi.SHAPE_ZEEMAN_MULTIPLET

    def __init__(self, line, wavelength, target_species, plasma, atomic_data, zeeman_structure=None, polarisation="no"):
        super().__init__(line, wavelength, target_species, plasma, atomic_data, polarisation)
        self.zeeman_structure = zeeman_structure or atomic_data.zeeman_structure(line)

    def _fill(self, sh, keep):
        super()._fill(sh, keep)
        zs = self.zeeman_structure
        wl, ra = zs.tabulate()
        keep.extend([wl, ra, zs.b_grid])
        sh.n_b = zs.b_grid.size
        sh.n_pi, sh.n_sigma_plus, sh.n_sigma_minus = (len(g) for g in zs.groups)
        sh.b_grid = zs.b_grid.ctypes.data_as(_abi.c_double_p)
        sh.zeeman_wavelength = wl.ctypes.data_as(_abi.c_double_p)
        sh.zeeman_ratio = ra.ctypes.data_as(_abi.c_double_p)


class StarkBroadenedLine(ZeemanLineShapeModel):
    """stark.pyx:150-348."""
    kind = _abi.SHAPE_STARK

    def __init__(self, line, wavelength, target_species, plasma, atomic_data, stark_model_coefficients=None,
                 integrator=None, polarisation="no"):
        super().__init__(line, wavelength, target_species, plasma, atomic_data, polarisation)
        try:
            cij, aij, bij = stark_model_coefficients or atomic_data.stark_model_coefficients(line)
        except IndexError:
            raise ValueError("Stark broadening coefficients for {} is not currently available.".format(line))
        if cij <= 0:
            raise ValueError("Coefficient c_ij must be positive.")
        if aij <= 0:
            raise ValueError("Coefficient a_ij must be positive.")
        if bij <= 0:
            raise ValueError("Coefficient b_ij must be positive.")
        self.coefficients = (float(cij), float(aij), float(bij))

    def _fill(self, sh, keep):
        super()._fill(sh, keep)
        sh.param[0], sh.param[1], sh.param[2] = self.coefficients


# ------------------------------------------------------------------------------------------------------------------
# plasma models
# ------------------------------------------------------------------------------------------------------------------
class PlasmaModel:
    """cherab/core/plasma/model.pyx:19-113."""

    def __init__(self, plasma=None, atomic_data=None):
        self.plasma, self.atomic_data = plasma, atomic_data


class _LineModel(PlasmaModel):
    kind = None

    def __init__(self, line, plasma=None, atomic_data=None, lineshape=None, lineshape_args=None, lineshape_kwargs=None):
        super().__init__(plasma, atomic_data)
        if not isinstance(line, Line):
            raise TypeError("line must be a Line")
        self.line = line
        self.lineshape_class = lineshape or GaussianLine
        if not (isinstance(self.lineshape_class, type) and issubclass(self.lineshape_class, LineShapeModel)):
            raise TypeError("The attribute lineshape must be a subclass of LineShapeModel.")
        self.lineshape_args = list(lineshape_args) if lineshape_args else []
        self.lineshape_kwargs = dict(lineshape_kwargs) if lineshape_kwargs else {}

    def _target(self, plasma):
        raise NotImplementedError

    def _rate(self, atomic_data):
        raise NotImplementedError

    def populate(self, plasma, atomic_data):
        """What _populate_cache resolves (impact_excitation.pyx:102-128): (species index, rate, wavelength, lineshape)."""
        if plasma is None:
            raise RuntimeError("The emission model is not connected to a plasma object.")
        if atomic_data is None:
            raise RuntimeError("The emission model is not connected to an atomic data source.")
        element, charge = self._target(plasma)
        try:
            index = plasma.composition.index(element, charge)
        except ValueError:
            raise RuntimeError("The plasma object does not contain the ion species for the specified line "
                               "(element={}, ionisation={}).".format(element.symbol, charge))
        rate = self._rate(atomic_data)
        wavelength = atomic_data.wavelength(self.line.element, self.line.charge, self.line.transition)
        species = plasma.composition.get(element, charge)
        shape = self.lineshape_class(self.line, wavelength, species, plasma, atomic_data,
                                     *self.lineshape_args, **self.lineshape_kwargs)
        return index, rate, wavelength, shape


class ExcitationLine(_LineModel):
    """impact_excitation.pyx:28-136."""
    kind = _abi.MODEL_EXCITATION_LINE

    def _target(self, plasma):
        return self.line.element, self.line.charge

    def _rate(self, atomic_data):
        return atomic_data.impact_excitation_pec(self.line.element, self.line.charge, self.line.transition)

    def __repr__(self):
        return "<ExcitationLine: element={}, charge={}, transition={}>".format(self.line.element.name, self.line.charge, self.line.transition)


class RecombinationLine(_LineModel):
    """recombination.pyx:28-139: the target species is the recombining ion, charge + 1 (:113-121)."""
    kind = _abi.MODEL_RECOMBINATION_LINE

    def _target(self, plasma):
        return self.line.element, self.line.charge + 1

    def _rate(self, atomic_data):
        return atomic_data.recombination_pec(self.line.element, self.line.charge, self.line.transition)

    def __repr__(self):
        return "<RecombinationLine: element={}, charge={}, transition={}>".format(self.line.element.name, self.line.charge, self.line.transition)


class ThermalCXLine(_LineModel):
    """thermal_cx.pyx:28-163: line emission of the receiver (element, charge + 1) after thermal charge exchange with every
    other species of the composition that is not fully ionised."""
    kind = _abi.MODEL_THERMAL_CX_LINE

    def _target(self, plasma):
        return self.line.element, self.line.charge + 1

    def _rate(self, atomic_data):
        return None

    def donors(self, plasma, atomic_data):
        """[(species index, rate)] — thermal_cx.pyx:140-148."""
        receiver = plasma.composition.get(self.line.element, self.line.charge + 1)
        out = []
        for i, sp in enumerate(plasma.composition):
            if sp is not receiver and sp.charge < sp.element.atomic_number:
                out.append((i, atomic_data.thermal_cx_pec(sp.element, sp.charge, self.line.element, self.line.charge + 1, self.line.transition)))
        return out

    def __repr__(self):
        return "<ThermalCXLine: element={}, charge={}, transition={}>".format(self.line.element.name, self.line.charge, self.line.transition)


class TotalRadiatedPower(PlasmaModel):
    """total_radiated_power.pyx:30-175: line + recombination/continuum + charge-exchange radiated power of one charge state,
    spread evenly over the observed spectral range."""
    kind = _abi.MODEL_TOTAL_RADIATED_POWER

    def __init__(self, element, charge, plasma=None, atomic_data=None):
        if not 0 <= charge < element.atomic_number:
            raise ValueError("TotalRadiatedPower cannot be calculated for charge state (element={}, ionisation={})."
                             "".format(element.symbol, charge))
        super().__init__(plasma, atomic_data)
        self.element, self.charge = element, charge

    def populate(self, plasma, atomic_data):
        """total_radiated_power.pyx:120-163 -> (line_rad index, recom index, [hydrogen indices], plt, prb, prc)."""
        from .atomic import hydrogen, deuterium, tritium
        if plasma is None:
            raise RuntimeError("The emission model is not connected to a plasma object.")
        if atomic_data is None:
            raise RuntimeError("The emission model is not connected to an atomic data source.")
        plt = atomic_data.line_radiated_power_rate(self.element, self.charge)
        try:
            i_line = plasma.composition.index(self.element, self.charge)
        except ValueError:
            raise RuntimeError("The plasma object does not contain the required ion species for calculating"
                               "total line radiaton, (element={}, ionisation={}).".format(self.element.symbol, self.charge))
        prb = atomic_data.continuum_radiated_power_rate(self.element, self.charge + 1)
        try:
            i_recom = plasma.composition.index(self.element, self.charge + 1)
        except ValueError:
            raise RuntimeError("The plasma object does not contain the required ion species for calculating"
                               "recombination/continuum emission, (element={}, ionisation={}).".format(self.element.symbol, self.charge + 1))
        prc = atomic_data.cx_radiated_power_rate(self.element, self.charge + 1)
        hyd = []
        for iso in (hydrogen, deuterium, tritium):
            try:
                hyd.append(plasma.composition.index(iso, 0))
            except ValueError:
                pass
        return i_line, i_recom, hyd, plt, prb, prc

    def __repr__(self):
        return "<TotalRadiatedPower: element={}, charge={}>".format(self.element.name, self.charge)


class Bremsstrahlung(PlasmaModel):
    """bremsstrahlung.pyx:93-245.  ``gaunt_factor`` may be a (u, gamma2, table) triple; default: atomic_data's."""
    kind = _abi.MODEL_BREMSSTRAHLUNG

    def __init__(self, plasma=None, atomic_data=None, gaunt_factor=None, integrator=None):
        super().__init__(plasma, atomic_data)
        self.gaunt_factor = gaunt_factor

    def __repr__(self):
        return "<PlasmaModel - Bremsstrahlung>"
