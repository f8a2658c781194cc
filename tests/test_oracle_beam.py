"""Pin the oracle's beam path against the reference's own known-answer tests: cherab/core/tests/test_beam.py:64-158
(SingleRayAttenuator density with a constant stopping rate, Beam.direction) and test_beamcxline.py:85-137 (BeamCXLine
radiance against a midpoint sum of Beam.density)."""
import numpy as np
import pytest
from scipy import constants as const

import core_b200 as cb
from core_b200.slab import build_constant_slab_plasma
from oracle import oracle

from helpers import ATOMIC_MASS as AMU, ELEMENTARY_CHARGE as QE
# the reference takes e and m_u for EvAmuToMS / EvToJ from scipy.constants (cherab/core/utility/conversion.py:24-31), whose
# CODATA release depends on the installed scipy (2018 vs 2022 differ by 1.4e-9 in m_u); the oracle and the CUDA library pin the
# CODATA-2018 values of cherab/core/utility/constants.pyx:22-37, so the closed forms below use those


class BeamMockData(cb.AtomicData):
    def __init__(self, stopping, cx=3.4e-34):
        self.stopping, self.cx = stopping, cx

    def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
        return cb.ConstantRate(self.stopping)

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        return [cb.ConstantBeamCXPEC(1, self.cx)]

    def wavelength(self, ion, charge, transition):
        return 656.104


def beam_scene(stopping, density=1e19, temperature=1e3, models=(), **beam_kw):
    atomic = BeamMockData(stopping)
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=density, electron_temperature=temperature,
                                        plasma_species=[(cb.deuterium, 1, density, temperature, (0, 0, 0))], b_field=(0, 10.0, 0))
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.translate(0.5, 0, 0))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 50000, 1e6, 10, cb.deuterium
    for k, v in beam_kw.items():
        setattr(beam, k, v)
    beam.models = list(models)
    return plasma, beam


def test_beam_density_and_direction():
    # test_beam.py:64-158
    plasma, beam = beam_scene(1e-13, sigma=0.2, divergence_x=1.0, divergence_y=2.0, length=10.0)
    flat = cb.flatten_beam_scene(beam, 655.1, 657.1, 16)
    z0, x0, y0 = 0.8, 0.5, 0.5
    dens, dirs = oracle.beam_sample(flat, [[0, 0, z0], [x0, y0, z0], [0, 0, -1]])
    speed = np.sqrt(2 * 50000 * QE / AMU)
    attenuation = np.exp(-z0 * 1e19 * 1e-13 / speed)
    rate = 1e6 / (50000 * cb.deuterium.atomic_weight * QE)
    tx, ty = np.tan(np.deg2rad(1.0)), np.tan(np.deg2rad(2.0))
    sx, sy = np.sqrt(0.2 ** 2 + (z0 * tx) ** 2), np.sqrt(0.2 ** 2 + (z0 * ty) ** 2)
    on = rate / speed / (2 * np.pi * sx * sy) * attenuation
    off = on * np.exp(-0.5 * ((x0 / sx) ** 2 + (y0 / sy) ** 2))
    assert abs(dens[0] / on - 1) < 1e-12 and abs(dens[1] / off - 1) < 1e-12 and dens[2] == 0
    ex, ey = x0 * (z0 * tx) ** 2 / (0.2 ** 2 + (z0 * tx) ** 2), y0 * (z0 * ty) ** 2 / (0.2 ** 2 + (z0 * ty) ** 2)
    ref = np.array([ex, ey, z0]) / np.linalg.norm([ex, ey, z0])
    assert np.array_equal(dirs[0], [0, 0, 1]) and np.array_equal(dirs[2], [0, 0, 1])
    assert np.max(np.abs(dirs[1] - ref)) < 1e-12


def cx_case(shape=None):
    # test_beamcxline.py:85-137: D-alpha, ray along -x through the beam origin plane, 512 bins on [655.1, 657.1]
    line = cb.Line(cb.deuterium, 0, (3, 2))
    plasma, beam = beam_scene(0.0, temperature=200.0, models=[cb.BeamCXLine(line, lineshape=shape)])
    flat = cb.flatten_beam_scene(beam, 655.1, 657.1, 512)
    rays = cb.beam_ray_segments(beam, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    return plasma, beam, flat, rays


def test_beam_cx_line_default_lineshape():
    plasma, beam, flat, rays = cx_case()
    got, stats = oracle.emission_render(flat, rays)
    dx = beam.integrator.step
    xs = dx * (np.arange(-int(0.5 / dx), int(0.5 / dx)) + 0.5)
    dens, _ = oracle.beam_sample(flat, np.stack([xs, np.zeros_like(xs), np.zeros_like(xs)], axis=1))
    radiance = 0.25 * 3.4e-34 * 1e19 * dens.sum() * dx / np.pi
    sigma = np.sqrt(200.0 * QE / (cb.deuterium.atomic_weight * AMU)) * 656.104 / const.c
    ref = oracle.add_gaussian_line(radiance, 656.104, sigma, 655.1, 657.1, 512)
    assert stats["samples"] == 1001
    assert np.max(np.abs(got[0] - ref)) < 1e-8              # the reference's delta


def test_beam_scene_validation():
    plasma, beam = beam_scene(0.0)
    beam.models = [cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    with pytest.raises(RuntimeError):
        cb.flatten_beam_scene(beam, 500, 550, 8)             # no C6+ in the plasma
    with pytest.raises(ValueError):
        beam.sigma = 0.0
    beam.attenuator = None
    with pytest.raises(ValueError):
        cb.flatten_beam_scene(beam, 500, 550, 8)


# ---- BeamEmissionLine + MSE multiplet ----
class BesMockData(BeamMockData):
    def beam_emission_pec(self, beam_ion, plasma_ion, charge, transition):
        return cb.ConstantRate(2.0e-35)


def mse_case(ratio_functions=False):
    """Inputs of core/tests/test_lineshapes.py:391-472 (D-alpha 656.104 nm, beam 60 keV/amu at 10 eV, B = (0, 5, 0) T,
    ne 1e19, line of sight (-1, 1, 0)/sqrt 2, 512 bins on +-3 nm) replayed through the whole path: a parallel beam along +z
    through the slab, so every sample has the same multiplet shape and the spectrum is (integrated radiance) x (unit shape)."""
    atomic = BesMockData(0.0)
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=20.,
                                        plasma_species=[(cb.deuterium, 1, 1.e18, 5., (2.e4, 0, 0)), (cb.nitrogen, 1, 1.e17, 10., (1.e4, 5.e4, 0))],
                                        b_field=(0, 5., 0))
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.translate(0.5, 0.0, -0.5))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 1e6, 10, cb.deuterium
    kw = {}
    if ratio_functions:
        # the reference also takes functions of (ne, beam energy) / (ne) for the intensity ratios (mse.pyx:103-121); these are
        # linear in log10(ne) (exact on the flattener's table) and give the default constants at ne = 1e19, E = 60 keV/amu
        kw = dict(sigma_to_pi=lambda ne, e: 0.56 + 0.05 * (np.log10(ne) - 19.0) + 1e-6 * (e - 60000.0),
                  sigma1_to_sigma0=lambda ne: 0.7060001671878492 - 0.03 * (np.log10(ne) - 19.0),
                  pi2_to_pi3=lambda ne: 0.3140003593919741 * (1.0 + 0.1 * (np.log10(ne) - 19.0)), pi4_to_pi3=0.7279994935840365)
    beam.models = [cb.BeamEmissionLine(cb.Line(cb.deuterium, 0, (3, 2)), **kw)]
    flat = cb.flatten_beam_scene(beam, 656.104 - 3, 656.104 + 3, 512)
    d = np.array([-1.0, 1.0, 0.0]) / np.sqrt(2)
    rays = cb.beam_ray_segments(beam, [np.array([0.5, 0.0, 0.0]) - 2.0 * d], [d])
    return flat, rays, d


def mse_unit_shape(direction):
    from scipy.special import erf
    wavelength, s2p, s1s0, p23, p43 = 656.104, 0.56, 0.7060001671878492, 0.3140003593919741, 0.7279994935840365
    bv = np.array([0, 0, 1.0]) * np.sqrt(2 * 60000 * QE / AMU)
    stark = abs(2.77e-8 * np.linalg.norm(np.cross(bv, [0, 5.0, 0])))
    central = wavelength * (1 + bv.dot(direction) / const.c)
    sigma = np.sqrt(10 * QE / (cb.deuterium.atomic_weight * AMU)) * wavelength / const.c
    wl, delta = np.linspace(wavelength - 3, wavelength + 3, 513, retstep=True)
    g = lambda c: 0.5 * np.diff(erf((wl - c) / (np.sqrt(2.) * sigma))) / delta
    dd = 1 / (1 + s2p)
    isig, ipi = s2p * dd, 0.5 * dd
    is0 = 1 / (s1s0 + 1)
    is1 = 0.5 * s1s0 * is0
    ip3 = 1 / (1 + p23 + p43)
    out = isig * is0 * g(central) + isig * is1 * (g(central + stark) + g(central - stark))
    for k, a in ((2, p23 * ip3), (3, ip3), (4, p43 * ip3)):
        out += ipi * a * (g(central + k * stark) + g(central - k * stark))
    return out, delta


def test_beam_emission_multiplet():
    flat, rays, d = mse_case()
    got, stats = oracle.emission_render(flat, rays)
    shape, delta = mse_unit_shape(d)
    total = got[0].sum() * delta                     # integrated radiance (the multiplet integrates to isig + 2 ipi (..) = 1 x radiance)
    assert total > 0 and stats["samples"] > 100
    assert np.max(np.abs(got[0] / total - shape / (shape.sum() * delta))) < 1e-10
    # beam emission rate: radiance = 1/(4 pi) n_beam sum_s (n_s Z_s) q_s, constant q = 2e-35 W m^3 for both singly charged species
    dens, _ = oracle.beam_sample(flat, [[0, 0, 0.5]])
    sigma_b = 0.1
    expected = 0.25 / np.pi * (1e18 + 1e17) * 2.0e-35 * dens[0] * np.sqrt(2 * np.pi) * sigma_b   # chord through the axis of the round Gaussian beam
    assert abs(total / expected - 1) < 2e-3           # trapezium sum of a Gaussian over a chord clipped at 5 sigma


# ---- excited donor metastables: _composite_cx_rate / _beam_population (charge_exchange.pyx:204-292) ----
class MetastableMockData(BeamMockData):
    """Ground-state rate q1, one excited metastable with rate q2 and a constant relative population k."""
    q2, k = 9.0e-34, 0.25

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        return [cb.ConstantBeamCXPEC(2, self.q2), cb.ConstantBeamCXPEC(1, self.cx)]      # any order: the ground state is found by its label

    def beam_population_rate(self, beam_ion, metastable, plasma_ion, charge):
        assert metastable == 2
        return cb.ConstantRate(self.k)


def metastable_case():
    line = cb.Line(cb.deuterium, 0, (3, 2))
    plasma, beam = beam_scene(0.0, temperature=200.0, models=[cb.BeamCXLine(line)])
    atomic = MetastableMockData(0.0)
    plasma.atomic_data = beam.atomic_data = atomic
    flat = cb.flatten_beam_scene(beam, 655.1, 657.1, 512)
    rays = cb.beam_ray_segments(beam, [[1.5, 0, 0]], [[-1.0, 0, 0]])
    return beam, flat, rays


def test_beam_cx_line_population_weighted_metastables():
    beam, flat, rays = metastable_case()
    got, stats = oracle.emission_render(flat, rays)
    _, flat1, _ = cx_case()[1:]
    ground, _ = oracle.emission_render(flat1, rays)                       # the same scene with the ground state alone (q1)
    q1, q2, k = 3.4e-34, MetastableMockData.q2, MetastableMockData.k
    composite = (q1 + k * q2) / (1 + k)                                   # charge_exchange.pyx:178
    assert np.max(np.abs(got[0] - ground[0] * composite / q1)) <= 1e-12 * got.max()


def test_beam_emission_multiplet_with_ratio_functions():
    # function-valued intensity ratios that equal the defaults at the slab's density reproduce the constant-ratio spectrum
    flat_c, rays, _ = mse_case()
    flat_f, _, _ = mse_case(ratio_functions=True)
    assert flat_f.desc.models[0].ext.contents.n_mse > 1 and flat_c.desc.models[0].ext.contents.n_mse == 0
    const_ratios, _ = oracle.emission_render(flat_c, rays)
    functions, _ = oracle.emission_render(flat_f, rays)
    assert np.max(np.abs(functions - const_ratios)) <= 1e-12 * const_ratios.max()
