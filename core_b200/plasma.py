"""Plasma scene objects: the host-side mirror of cherab.core.plasma / distribution / species for the emission path.

Same names and argument meaning as the reference (cherab/core/plasma/node.pyx:201-554, distribution.pyx:190-303,
species.pyx:26-80), but the 3-D functions are *flattenable field descriptors* instead of Raysect Function3D
objects: the CUDA kernels evaluate a fixed set of field kinds from device-resident tables (no Python in the loop,
no CPU fallback).  An unsupported field raises TypeError at flatten time.
"""
import numpy as np

from .notify import Notifier

from . import _abi


# ------------------------------------------------------------------------------------------------------------------
# scalar / vector fields
# ------------------------------------------------------------------------------------------------------------------
class ScalarField:
    kind = None

    def _fill(self, f, keep):
        raise NotImplementedError


class Constant3D(ScalarField):
    """raysect Constant3D; what build_constant_slab_plasma uses (tools/plasmas/slab.pyx:198-260)."""
    kind = _abi.FIELD_CONSTANT

    def __init__(self, value):
        self.value = float(value)

    def _fill(self, f, keep):
        f.kind = self.kind
        f.c[0] = self.value


class GaussianVolume(ScalarField):
    """offset + peak * exp(-|p - centre|^2 / (2 sigma^2)) — tools/plasmas/gaussian_volume.pyx:24-59;
    ``offset`` covers the ``1 + GaussianVolume(79, sigma)`` arithmetic of demos/balmer_series.py:59."""
    kind = _abi.FIELD_GAUSSIAN_VOLUME

    def __init__(self, peak, sigma, offset=0.0, centre=(0.0, 0.0, 0.0)):
        self.peak, self.sigma, self.offset, self.centre = float(peak), float(sigma), float(offset), tuple(centre)

    def _fill(self, f, keep):
        f.kind = self.kind
        f.c[0], f.c[1], f.c[2] = self.offset, self.peak, self.sigma
        f.c[3], f.c[4], f.c[5] = self.centre


class SlabIonFunction(ScalarField):
    """IonFunction pedestal along +x (tools/plasmas/slab.pyx:63-110)."""
    kind = _abi.FIELD_SLAB_ION

    def __init__(self, t_core, t_lcfs, p=2, q=2, pedestal_top=1):
        self.args = (float(t_core), float(t_lcfs), float(p), float(q), float(pedestal_top))

    def _fill(self, f, keep):
        f.kind = self.kind
        for i, v in enumerate(self.args):
            f.c[i] = v


class SlabNeutralFunction(ScalarField):
    """NeutralFunction decay along +x (tools/plasmas/slab.pyx:20-60)."""
    kind = _abi.FIELD_SLAB_NEUTRAL

    def __init__(self, peak, sigma, pedestal_top=1):
        self.peak, self.sigma = float(peak), float(sigma)

    def _fill(self, f, keep):
        f.kind = self.kind
        f.c[0], f.c[1] = self.peak, self.sigma


def _dptr(a, keep):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    keep.append(a)
    return a.ctypes.data_as(_abi.c_double_p)


class AxisymBlend(ScalarField):
    """AxisymmetricMapper(Blend2D(Discrete2DMesh(edge), equilibrium.map2d(cubic1d(core)), mask)) — the Generomak
    full-profile recipe (cherab/generomak/plasma/plasma.py:580-638, 132-163).  ``edge`` holds one value per mesh
    triangle, ``core`` one value per core psi_n knot; the mesh/equilibrium/mask live in the plasma's AxisymContext."""
    kind = _abi.FIELD_AXISYM_BLEND

    def __init__(self, edge, core):
        self.edge = None if edge is None else np.ascontiguousarray(edge, dtype=np.float64)
        self.core = None if core is None else np.ascontiguousarray(core, dtype=np.float64)

    def _fill(self, f, keep):
        f.kind = self.kind
        f.edge = _dptr(self.edge, keep)
        f.core = _dptr(self.core, keep)


class VectorField:
    pass


class ConstantVector3D(VectorField):
    def __init__(self, x, y, z):
        self.v = (float(x), float(y), float(z))

    def _fill(self, f, keep):
        f.kind = _abi.FIELD_CONSTANT
        f.c[0], f.c[1], f.c[2] = self.v


class AxisymBlendVector(VectorField):
    """VectorAxisymmetricMapper(BlendVector2D(ConstantVector2D(edge), equilibrium.map_vector2d(vtor, vpol, vnorm), mask))
    — plasma.py:612-636, efit.pyx:280-344."""

    def __init__(self, edge_vector, core_vtor, core_vpol, core_vnorm):
        self.edge = tuple(float(x) for x in edge_vector)
        self.vtor, self.vpol, self.vnorm = core_vtor, core_vpol, core_vnorm

    def _fill(self, f, keep):
        f.kind = _abi.FIELD_AXISYM_BLEND
        f.c[0], f.c[1], f.c[2] = self.edge
        f.core_vtor = _dptr(self.vtor, keep)
        f.core_vpol = _dptr(self.vpol, keep)
        f.core_vnorm = _dptr(self.vnorm, keep)


def _as_scalar_field(x):
    if isinstance(x, ScalarField):
        return x
    if isinstance(x, (int, float, np.floating)):
        return Constant3D(x)
    raise TypeError("Unsupported Function3D for the B200 path: %r (supported: Constant3D, GaussianVolume, "
                    "SlabIonFunction, SlabNeutralFunction, AxisymBlend)" % (x,))


def _as_vector_field(x):
    if isinstance(x, VectorField):
        return x
    if isinstance(x, (tuple, list, np.ndarray)) and len(x) == 3:
        return ConstantVector3D(*x)
    raise TypeError("Unsupported VectorFunction3D for the B200 path: %r" % (x,))


# ------------------------------------------------------------------------------------------------------------------
# equilibrium + mesh context shared by AxisymBlend fields
# ------------------------------------------------------------------------------------------------------------------
class EFITEquilibrium:
    """Inputs of cherab.tools.equilibrium.EFITEquilibrium (efit.pyx:92-142) that the flux mapping needs."""

    def __init__(self, r, z, psi_grid, psi_axis, psi_lcfs, f_profile, b_vacuum_radius, b_vacuum_magnitude, lcfs_polygon):
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.psi = np.ascontiguousarray(psi_grid, dtype=np.float64)
        if self.psi.shape != (self.r.size, self.z.size):
            raise ValueError("psi_grid must have shape (len(r), len(z))")
        self.psi_axis, self.psi_lcfs = float(psi_axis), float(psi_lcfs)
        f_profile = np.asarray(f_profile, dtype=np.float64)
        self.f_psin = np.ascontiguousarray(f_profile[0])
        self.f_value = np.ascontiguousarray(f_profile[1])
        self.b_vacuum_radius, self.b_vacuum_magnitude = float(b_vacuum_radius), float(b_vacuum_magnitude)
        poly = np.asarray(lcfs_polygon, dtype=np.float64)
        if poly.shape[0] == 2 and poly.shape[1] != 2:
            poly = poly.T  # efit.pyx:167-169 transposes 2xN to Nx2
        self.lcfs_polygon = np.ascontiguousarray(poly)
        self.r_range = (self.r.min(), self.r.max())
        self.z_range = (self.z.min(), self.z.max())


class AxisymContext:
    """Equilibrium + edge triangular mesh + core psi_n grid + blend mask (plasma.py:96-129, 233-272, 610)."""

    def __init__(self, equilibrium, vertices, triangles, core_psin, mask_x=(0, 0.94, 1.0, 1.1), mask_y=(1, 1, 0, 0)):
        self.equilibrium = equilibrium
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.core_psin = np.ascontiguousarray(core_psin, dtype=np.float64)
        self.mask_x = np.ascontiguousarray(mask_x, dtype=np.float64)
        self.mask_y = np.ascontiguousarray(mask_y, dtype=np.float64)

    def _fill(self, a, keep):
        e = self.equilibrium
        a.eq.nr, a.eq.nz = e.r.size, e.z.size
        a.eq.r, a.eq.z, a.eq.psi = _dptr(e.r, keep), _dptr(e.z, keep), _dptr(e.psi, keep)
        a.eq.psi_axis, a.eq.psi_lcfs = e.psi_axis, e.psi_lcfs
        a.eq.n_f, a.eq.n_lcfs = e.f_psin.size, e.lcfs_polygon.shape[0]
        a.eq.f_psin, a.eq.f_value = _dptr(e.f_psin, keep), _dptr(e.f_value, keep)
        a.eq.lcfs_polygon = _dptr(e.lcfs_polygon, keep)
        a.eq.b_vacuum_radius, a.eq.b_vacuum_magnitude = e.b_vacuum_radius, e.b_vacuum_magnitude
        a.n_vertices, a.n_triangles = self.vertices.shape[0], self.triangles.shape[0]
        a.vertices = _dptr(self.vertices, keep)
        keep.append(self.triangles)
        a.triangles = self.triangles.ctypes.data_as(_abi.c_int32_p)
        a.n_core, a.n_mask = self.core_psin.size, self.mask_x.size
        a.core_psin = _dptr(self.core_psin, keep)
        a.mask_x, a.mask_y = _dptr(self.mask_x, keep), _dptr(self.mask_y, keep)


class EFITMagneticField:
    """VectorAxisymmetricMapper(equilibrium.b_field) — plasma.py:699, efit.pyx:413-461."""


# ------------------------------------------------------------------------------------------------------------------
# distributions, species, plasma
# ------------------------------------------------------------------------------------------------------------------
class Maxwellian:
    """cherab/core/distribution.pyx:190-303: (density, temperature, velocity, atomic_mass)."""

    def __init__(self, density, temperature, velocity, atomic_mass):
        self.density = _as_scalar_field(density)
        self.temperature = _as_scalar_field(temperature)
        self.velocity = _as_vector_field(velocity)
        self.atomic_mass = float(atomic_mass)


class Species:
    """cherab/core/species.pyx:26-80."""

    def __init__(self, element, charge, distribution):
        if charge > element.atomic_number:
            raise ValueError("Charge state cannot be larger than the atomic number.")
        if charge < 0:
            raise ValueError("Charge state cannot be less than zero.")
        self.element, self.charge, self.distribution = element, int(charge), distribution


class NumericalIntegrator:
    """raysect NumericalIntegrator(step, min_samples=5); Plasma default step 0.001 (plasma/node.pyx:318)."""

    def __init__(self, step=0.001, min_samples=5):
        if step <= 0:
            raise ValueError("Numerical integration step size can not be less than or equal to zero")
        if min_samples < 2:
            raise ValueError("At least two samples are required to perform the numerical integration.")
        self.step, self.min_samples = float(step), int(min_samples)


class Composition:
    """cherab/core/plasma/node.pyx:33-163 (set/add/get/clear by (element, charge)); every change notifies (:75,:96,:141)."""

    def __init__(self):
        self._species = []
        self.notifier = Notifier()

    def __iter__(self):
        return iter(self._species)

    def __len__(self):
        return len(self._species)

    def set(self, species):
        species = list(species)
        for s in species:
            if not isinstance(s, Species):
                raise TypeError("The composition must consist of a sequence of Species objects.")
        self._species = []
        for s in species:
            self._species = [o for o in self._species if not (o.element is s.element and o.charge == s.charge)]
            self._species.append(s)
        self.notifier.notify()

    def add(self, species):
        if not isinstance(species, Species):
            raise TypeError("Only Species objects can be added to a composition.")
        self._species = [s for s in self._species if not (s.element is species.element and s.charge == species.charge)]
        self._species.append(species)
        self.notifier.notify()

    def get(self, element, charge):
        for s in self._species:
            if s.element is element and s.charge == charge:
                return s
        raise ValueError("Could not find a species with the specified element and charge state.")

    def index(self, element, charge):
        for i, s in enumerate(self._species):
            if s.element is element and s.charge == charge:
                return i
        raise ValueError("Could not find a species with the specified element and charge state.")

    def clear(self):
        self._species = []
        self.notifier.notify()


class ModelManager:
    """cherab/core/plasma/node.pyx:166-198: the list of emission models attached to a plasma; every change notifies."""

    def __init__(self):
        self._models = []
        self.notifier = Notifier()

    def __iter__(self):
        return iter(self._models)

    def __len__(self):
        return len(self._models)

    def __getitem__(self, i):
        return self._models[i]

    def set(self, models):
        models = list(models)
        for m in models:
            if not hasattr(m, "kind"):
                raise TypeError("The model list must consist of only PlasmaModel objects.")
        self._models = models
        self.notifier.notify()

    def add(self, model):
        if not hasattr(model, "kind"):
            raise TypeError("The model list must consist of only PlasmaModel objects.")
        self._models.append(model)
        self.notifier.notify()

    def clear(self):
        self._models = []
        self.notifier.notify()


class Plasma:
    """cherab/core/plasma/node.pyx:201-554, Raysect-free: geometry is a primitive from core_b200.geometry,
    geometry_transform a 4x4 (or 3x4) world<-plasma affine matrix."""

    # attributes whose change invalidates anything cached from this plasma (node.pyx:323-331, 464-509, 545-554)
    _NOTIFYING = ("b_field", "electron_distribution", "atomic_data", "geometry", "geometry_transform", "transform", "integrator", "axisym")

    def __init__(self, name="Plasma"):
        self.__dict__["notifier"] = Notifier()
        self.name = name
        self._composition = Composition()
        self._composition.notifier.add(self._modified)
        self._models = ModelManager()
        self._models.notifier.add(self._modified)
        self.b_field = ConstantVector3D(0, 0, 0)
        self.electron_distribution = None
        self.atomic_data = None
        self.geometry = None
        self.geometry_transform = None  # geometry-local -> plasma space (node.pyx:535-540)
        self.transform = None           # plasma space -> world (the Node transform); None = identity
        self.integrator = NumericalIntegrator(step=0.001)
        self.axisym = None

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in self._NOTIFYING:
            self._modified()

    def _modified(self):
        """Plasma._modified (node.pyx:545-554): tell every registered observer that cached data is stale."""
        self.notifier.notify()

    @property
    def models(self):
        return self._models

    @models.setter
    def models(self, values):
        self._models.set(values)

    def geometry_to_world(self):
        """4x4 matrix taking the geometry primitive's local frame to world space."""
        m = np.eye(4)
        if self.transform is not None:
            m = m @ np.asarray(self.transform, dtype=np.float64)
        if self.geometry_transform is not None:
            m = m @ np.asarray(self.geometry_transform, dtype=np.float64)
        return m

    @property
    def composition(self):
        return self._composition

    @composition.setter
    def composition(self, values):
        self._composition.set(values)
