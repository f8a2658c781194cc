"""Pin the oracle's emission models + integrators against the reference's known-answer tests:
cherab/core/tests/test_line_emission.py:99-200 (slab, constant PEC), test_bremsstrahlung.py:41-95,
cherab/core/math/tests/test_integrators.py:68-76."""
import ctypes as C

import numpy as np
import pytest
from scipy import constants as const
from scipy.special import erf, roots_legendre

import core_b200 as cb
from core_b200 import _abi
from core_b200.slab import build_constant_slab_plasma
from oracle import oracle
from helpers import ATOMIC_MASS, ELEMENTARY_CHARGE, SPEED_OF_LIGHT


class MockAtomicData(cb.AtomicData):
    """core/tests/test_line_emission.py:32-97."""

    def impact_excitation_pec(self, ion, charge, transition):
        return cb.ConstantRate(1.4e-39)

    def recombination_pec(self, ion, charge, transition):
        return cb.ConstantRate(8.e-40)

    def wavelength(self, ion, charge, transition):
        return 529.27


def _slab():
    plasma = build_constant_slab_plasma(length=1.2, width=1, height=1, electron_density=1e19, electron_temperature=1000.,
                                        plasma_species=[(cb.carbon, 5, 2.e18, 800., (0, 0, 0)), (cb.carbon, 6, 3.e18, 900., (0, 0, 0))],
                                        b_field=(0, 10., 0))
    plasma.atomic_data = MockAtomicData()
    return plasma


def _trace(plasma, lo, hi, bins):
    fs = cb.flatten_scene(plasma, lo, hi, bins)
    rays = cb.ray_segments(plasma.geometry, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    assert rays.n_segments == 1 and abs(rays.seg_t1[0] - rays.seg_t0[0] - 1.2) < 1e-12
    return oracle.emission_render(fs, rays)


def _gauss_ref(radiance, wavelength, temperature, weight, lo, hi, bins):
    sigma = np.sqrt(temperature * ELEMENTARY_CHARGE / (weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    wl, delta = np.linspace(lo, hi, bins + 1, retstep=True)
    erfs = erf((wl - wavelength) / (np.sqrt(2.) * sigma))
    return 0.5 * radiance * (erfs[1:] - erfs[:-1]) / delta


def test_excitation_line_default_lineshape():
    plasma = _slab()
    line = cb.Line(cb.carbon, 5, (8, 7))
    plasma.models = [cb.ExcitationLine(line)]
    got, _ = _trace(plasma, 529.27 - 1.5, 529.27 + 1.5, 512)
    radiance = 0.25 / np.pi * 1.4e-39 * 2.e18 * 1e19 * 1.2
    ref = _gauss_ref(radiance, 529.27, 800., cb.carbon.atomic_weight, 529.27 - 1.5, 529.27 + 1.5, 512)
    assert np.max(np.abs(got[0] - ref)) < 1e-8
    assert np.max(np.abs(got[0] - ref)) < 1e-12 * ref.max()


def test_recombination_line_uses_charge_plus_one():
    plasma = _slab()
    line = cb.Line(cb.carbon, 5, (8, 7))
    plasma.models = [cb.RecombinationLine(line)]
    got, _ = _trace(plasma, 529.27 - 1.5, 529.27 + 1.5, 512)
    radiance = 0.25 / np.pi * 8.e-40 * 3.e18 * 1e19 * 1.2   # density AND temperature of C6+ (recombination.pyx:113-128)
    ref = _gauss_ref(radiance, 529.27, 900., cb.carbon.atomic_weight, 529.27 - 1.5, 529.27 + 1.5, 512)
    assert np.max(np.abs(got[0] - ref)) < 1e-12 * ref.max()


def test_models_add_not_overwrite():
    plasma = _slab()
    line = cb.Line(cb.carbon, 5, (8, 7))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    both, _ = _trace(plasma, 528, 531, 128)
    plasma.models = [cb.ExcitationLine(line)]
    a, _ = _trace(plasma, 528, 531, 128)
    plasma.models = [cb.RecombinationLine(line)]
    b, _ = _trace(plasma, 528, 531, 128)
    assert np.allclose(both, a + b, rtol=1e-13, atol=0)


def test_missing_species_raises_runtime_error():
    plasma = _slab()
    plasma.models = [cb.ExcitationLine(cb.Line(cb.carbon, 3, (8, 7)))]
    try:
        cb.flatten_scene(plasma, 528, 531, 16)
    except RuntimeError as e:
        assert "does not contain the ion species" in str(e)
    else:
        raise AssertionError("expected RuntimeError (impact_excitation.pyx:110-118)")


def _gauss_legendre_scipy(f, a, b, rtol=1e-5, min_order=1, max_order=50):
    """integrators1d.pyx:189-224 restated with scipy.special.roots_legendre."""
    old, new = np.inf, 0.0
    c, d = 0.5 * (a + b), 0.5 * (b - a)
    for order in range(min_order, max_order + 1):
        x, w = roots_legendre(order)
        new = d * sum(wi * f(c + d * xi) for xi, wi in zip(x, w))
        err = abs(new - old)
        old = new
        if err < rtol * abs(new):
            break
    return new


def test_gaussian_quadrature_erf():
    # core/math/tests/test_integrators.py:68-76
    cb_t = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    fn = cb_t(lambda x, ctx: 2 / np.sqrt(np.pi) * np.exp(-x * x))
    l = oracle.lib()
    l.cb2o_gauss_legendre.restype = C.c_double
    l.cb2o_gauss_legendre.argtypes = [cb_t, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    val = l.cb2o_gauss_legendre(fn, None, -0.5, 3.0, 1e-8, 1, 50)
    assert abs(val - (erf(3.0) - erf(-0.5))) < 1e-8
    ref = _gauss_legendre_scipy(lambda x: 2 / np.sqrt(np.pi) * np.exp(-x * x), -0.5, 3.0, 1e-8)
    assert abs(val - ref) < 1e-14


def _gaunt_struct():
    u, g2, gff = (np.ascontiguousarray(a, dtype=np.float64) for a in cb.AtomicData().free_free_gaunt_factor())
    g = _abi.Gaunt(u.size, g2.size, u.ctypes.data_as(_abi.c_double_p), g2.ctypes.data_as(_abi.c_double_p),
                   gff.ctypes.data_as(_abi.c_double_p))
    return g, (u, g2, gff)


def test_gaunt_factor_limits_and_table():
    g, (u, g2, gff) = _gaunt_struct()
    l = oracle.lib()
    # knots are reproduced exactly by a Hermite interpolant
    for i in (2, 4, 6):
        for j in (10, 40, 80):
            te = 1.0 * 13.605693122994 / g2[j]                 # z = 1 -> gamma2 = Ry/te
            wl = 1239.8419738620933 / (te * u[i])
            assert abs(l.cb2o_gaunt_factor(C.byref(g), 1.0, te, wl) - gff[i, j]) < 1e-9
    assert l.cb2o_gaunt_factor(C.byref(g), 0.0, 10.0, 500.0) == 0.0                    # gaunt.pyx:123
    assert l.cb2o_gaunt_factor(C.byref(g), 1.0, 1e-4, 1e3) == 1.0                      # classical limit :130
    te, wl = 1e12, 500.0                                                              # Born limit :134
    uu = 1239.8419738620933 / (te * wl)
    assert abs(l.cb2o_gaunt_factor(C.byref(g), 1.0, te, wl) - np.sqrt(3) / np.pi * (np.log(4 / uu) - 0.5772156649015329)) < 1e-12


def test_bremsstrahlung_slab():
    # core/tests/test_bremsstrahlung.py:41-95 (D+ 1e19, N7+ 1e18, Te 2 keV, 400-800 nm, 128 bins, 1 m slab)
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=2000.,
                                        plasma_species=[(cb.deuterium, 1, 1.e19, 2000., (0, 0, 0)), (cb.nitrogen, 7, 1.e18, 2000., (0, 0, 0))])
    plasma.atomic_data = cb.AtomicData()
    plasma.models = [cb.Bremsstrahlung()]
    fs = cb.flatten_scene(plasma, 400., 800., 128)
    rays = cb.ray_segments(plasma.geometry, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    got, stats = oracle.emission_render(fs, rays)
    assert stats["brems_bin_evals"] == 128 * 1001

    brems_const = (const.e ** 2 * 0.25 / np.pi / const.epsilon_0) ** 3
    brems_const *= 32 * np.pi ** 2 / (3 * np.sqrt(3) * const.m_e ** 2 * const.c ** 3)
    brems_const *= np.sqrt(2 * const.m_e / (np.pi * const.e))
    brems_const *= const.c * 1e9 * 0.25 / np.pi
    exp_factor = const.h * const.c * 1.e9 / const.e
    ne = te = None
    ne, te = 1e19, 2000.
    g, _keep = _gaunt_struct()
    l = oracle.lib()

    def brems_func(wvl):
        s = 0
        for z, ni in ((1, 1e19), (7, 1e18)):
            s += ni * l.cb2o_gaunt_factor(C.byref(g), float(z), te, wvl) * z * z
        return brems_const * s * ne / (np.sqrt(te) * wvl * wvl) * np.exp(-exp_factor / (te * wvl))

    wl, delta = np.linspace(400., 800., 129, retstep=True)
    ref = np.array([_gauss_legendre_scipy(brems_func, wl[i], wl[i + 1]) / delta for i in range(128)])
    assert np.max(np.abs(got[0] - ref)) < 1e-10
    # scipy.constants (CODATA 2022) vs the reference's hard-coded CODATA 2018 values differ by 4e-9 relative
    assert np.max(np.abs(got[0] / ref - 1)) < 2e-8


# ---- ThermalCXLine and TotalRadiatedPower: the reference's own known-answer tests ----
class _PowerAtomicData(cb.AtomicData):
    """MockAtomicData of core/tests/test_total_radiated_power.py:72-87 and test_line_emission.py:58-84."""

    def wavelength(self, ion, charge, transition):
        return 529.27

    def thermal_cx_pec(self, donor_ion, donor_charge, receiver_ion, receiver_charge, transition):
        return cb.ConstantRate(1.2e-46)

    def line_radiated_power_rate(self, ion, charge):
        return cb.ConstantRate(1.e-32)

    def continuum_radiated_power_rate(self, ion, charge):
        return cb.ConstantRate(1.e-33)

    def cx_radiated_power_rate(self, ion, charge):
        return cb.ConstantRate(1.e-31)


def thermal_cx_scene(atomic=None):
    # core/tests/test_line_emission.py:241-290
    plasma = build_constant_slab_plasma(length=1.2, width=1, height=1, electron_density=1e19, electron_temperature=1000.,
                                        plasma_species=[(cb.carbon, 6, 1.67e18, 800., (0, 0, 0)), (cb.deuterium, 0, 1.e19, 100., (0, 0, 0))],
                                        b_field=(0, 10., 0))
    plasma.atomic_data = atomic if atomic is not None else _PowerAtomicData()
    line = cb.Line(cb.carbon, 5, (8, 7))
    plasma.models = [cb.ThermalCXLine(line)]
    flat = cb.flatten_scene(plasma, 529.27 - 1.5, 529.27 + 1.5, 512)
    rays = cb.ray_segments(plasma.geometry, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    radiance = 0.25 / np.pi * 1.2e-46 * 1.67e18 * 1e19 * 1.2
    sigma = np.sqrt(800. * const.e / (cb.carbon.atomic_weight * const.physical_constants["atomic mass constant"][0])) * 529.27 / const.c
    ref = oracle.add_gaussian_line(radiance, 529.27, sigma, 529.27 - 1.5, 529.27 + 1.5, 512)
    return flat, rays, ref


def total_radiated_power_scene():
    # core/tests/test_total_radiated_power.py:90-140
    species = [(cb.deuterium, 0, 1.e18, 500., (0, 0, 0)), (cb.hydrogen, 0, 1.e18, 500., (0, 0, 0)),
               (cb.nitrogen, 6, 5.e18, 1100., (0, 0, 0)), (cb.nitrogen, 7, 1.e19, 1100., (0, 0, 0))]
    plasma = build_constant_slab_plasma(length=1.2, width=1.2, height=1.2, electron_density=1e19, electron_temperature=1000., plasma_species=species)
    plasma.atomic_data = _PowerAtomicData()
    plasma.models = [cb.TotalRadiatedPower(cb.nitrogen, 6)]
    flat = cb.flatten_scene(plasma, 500., 550., 2)
    rays = cb.ray_segments(plasma.geometry, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    total = 0.25 / np.pi * 1.2 * (1.e-32 * 1e19 * 5e18 + 1.e-33 * 1e19 * 1e19 + 1.e-31 * (1e18 + 1e18) * 1e19)
    return flat, rays, total


def test_thermal_cx_line_slab():
    flat, rays, ref = thermal_cx_scene()
    got, _ = oracle.emission_render(flat, rays)
    assert np.max(np.abs(got[0] - ref)) < 1e-8            # the reference's delta


def test_total_radiated_power_slab():
    flat, rays, total = total_radiated_power_scene()
    got, _ = oracle.emission_render(flat, rays)
    delta = 50.0 / 2
    assert abs(got[0].sum() * delta / total - 1.0) < 1e-8  # Spectrum.total() == sum(samples) * delta_wavelength


def test_total_radiated_power_rejects_bare_nucleus():
    with pytest.raises(ValueError):
        cb.TotalRadiatedPower(cb.nitrogen, 7)


# ---- tabulated thermal-CX rates: Interpolator3DArray 'cubic' restated + ThermalCXPEC (openadas/rates/pec.pyx:153-194) ----
def test_tricubic_reproduces_quadratics_and_reduces_to_the_bicubic():
    rng = np.random.default_rng(0)
    x, y, z = np.sort(rng.uniform(0, 3, 7)), np.sort(rng.uniform(-1, 2, 6)), np.sort(rng.uniform(0, 1, 8))
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")

    def fn(a, b, c):
        return 1 + 2 * a - b + 0.5 * c + a * b - 2 * b * c + a * c + a * a + 0.3 * b * b - c * c + 0.7 * a * b * c
    f = fn(X, Y, Z)
    for _ in range(100):        # interior cells: the 3-point knot derivatives are exact for quadratics there
        p = rng.uniform(x[1], x[-2]), rng.uniform(y[1], y[-2]), rng.uniform(z[1], z[-2])
        assert abs(oracle.interp3d_cubic(x, y, z, f, *p) - fn(*p)) < 1e-12
    for i, j, k in ((0, 0, 0), (6, 5, 7), (3, 2, 4)):       # knots are reproduced
        assert abs(oracle.interp3d_cubic(x, y, z, f, x[i], y[j], z[k]) - f[i, j, k]) < 1e-13
    f2 = np.sin(X[:, :, 0]) + Y[:, :, 0] ** 2
    f3 = np.repeat(f2[:, :, None], z.size, axis=2)
    for _ in range(50):         # no dependence on the third axis -> the 2-D interpolator
        p = rng.uniform(x[0], x[-1]), rng.uniform(y[0], y[-1]), rng.uniform(z[0], z[-1])
        assert abs(oracle.interp3d_cubic(x, y, z, f3, *p) - oracle.interp2d_cubic(x, y, f2, p[0], p[1])) < 1e-13
    assert oracle.interp3d_cubic(x, y, z, f, x[0] - 5, y[2], z[3]) == oracle.interp3d_cubic(x, y, z, f, x[0], y[2], z[3])   # 'nearest'


class _PowerLawCX(_PowerAtomicData):
    """thermal CX rate = A ne^a te^b td^c photon m^3/s: linear in log space, so the tricubic is exact everywhere."""
    A, a, b, c = 3.0e-15, 0.1, -0.3, 0.45

    def thermal_cx_pec(self, donor_ion, donor_charge, receiver_ion, receiver_charge, transition):
        ne, te, td = np.logspace(17, 21, 9), np.logspace(0, 4, 11), np.logspace(-1, 3.5, 8)
        rate = self.A * ne[:, None, None] ** self.a * te[None, :, None] ** self.b * td[None, None, :] ** self.c
        return cb.RateTable3D(ne, te, td, rate, extrapolate=False)


def test_thermal_cx_line_slab_with_a_tabulated_rate():
    flat, rays, _ = thermal_cx_scene(atomic=_PowerLawCX())
    got, st = oracle.emission_render(flat, rays)
    q = _PowerLawCX.A * 1e19 ** _PowerLawCX.a * 1000. ** _PowerLawCX.b * 100. ** _PowerLawCX.c       # ne, te, T(D0) of the slab
    q *= const.h * const.c / (529.27e-9)                                                          # PhotonToJ, conversion.py:44-52
    radiance = 0.25 / np.pi * q * 1.67e18 * 1e19 * 1.2
    sigma = np.sqrt(800. * const.e / (cb.carbon.atomic_weight * const.physical_constants["atomic mass constant"][0])) * 529.27 / const.c
    ref = oracle.add_gaussian_line(radiance, 529.27, sigma, 529.27 - 1.5, 529.27 + 1.5, 512)
    assert st["out_of_domain"] == 0
    assert np.max(np.abs(got[0] - ref)) <= 1e-9 * ref.max()
