"""GPU parity tests of the beam path: Beam.density / direction, SingleRayAttenuator and BeamCXLine through
the CUDA library vs the oracle and vs the reference's closed forms (cherab/core/tests/test_beam.py, test_beamcxline.py)."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200 import _abi, generomak
from core_b200.engine import EmissionScene
from oracle import oracle
from test_oracle_beam import beam_scene, cx_case

pytestmark = pytest.mark.gpu


def parity(got, ref, rtol=1e-4, floor=1e-9):
    tol = rtol * np.abs(ref) + floor * np.abs(ref).max(axis=-1, keepdims=True)
    return float(np.max(np.abs(got - ref) / (tol + 1e-300)))


def beam_sample(scene, pts):
    import ctypes as C
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((pts.shape[0], 4))
    _abi.check(scene._lib, scene._lib.cb2_beam_sample(scene._h, pts.ctypes.data_as(_abi.c_double_p), pts.shape[0], out.ctypes.data_as(_abi.c_double_p)))
    return out[:, 0], out[:, 1:]


def test_beam_density_and_direction():
    plasma, beam = beam_scene(1e-13, sigma=0.2, divergence_x=1.0, divergence_y=2.0, length=10.0)
    flat = cb.flatten_beam_scene(beam, 655.1, 657.1, 16)
    rng = np.random.default_rng(5)
    pts = np.concatenate([[[0, 0, 0.8], [0.5, 0.5, 0.8], [0, 0, -1.0], [0, 0, 10.5]],
                          np.stack([rng.uniform(-1, 1, 200), rng.uniform(-1, 1, 200), rng.uniform(0, 10, 200)], axis=1)])
    scene = EmissionScene(flat)
    dens, dirs = beam_sample(scene, pts)
    scene.close()
    rd, rdir = oracle.beam_sample(flat, pts)
    assert dens[2] == 0 and dens[3] == 0 and np.array_equal(dens == 0, rd == 0)
    nz = rd > 0
    assert np.max(np.abs(dens[nz] / rd[nz] - 1)) < 2e-5          # fp32 evaluation
    assert np.max(np.abs(dirs - rdir)) < 1e-6


@pytest.mark.parametrize("shape", [None, cb.ZeemanTriplet])
def test_beam_cx_line_slab(shape):
    # test_beamcxline.py:85-173 through the CUDA path
    plasma, beam, flat, rays = cx_case(shape)
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"] == 1001
    assert parity(got, ref) <= 1.0


def test_beam_cx_generomak_tabulated_rates():
    # config C5 shape at small size: diverging, attenuated beam through the Generomak plasma, ADF12/ADF21-shaped synthetic
    # tables, C5+ n = 8 -> 7 CX line observed by a fan of sight lines crossing the beam
    plasma = generomak.get_plasma()
    atomic = cb.SyntheticADAS()
    atomic.wavelength = lambda ion, charge, transition: 529.05
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))       # z axis points into the plasma
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.05, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    flat = cb.flatten_beam_scene(beam, 526.0, 532.0, 256)
    # sight lines from above towards points on the beam axis
    axis_pts = (np.asarray(beam.transform) @ np.stack([np.zeros(12), np.zeros(12), np.linspace(0.6, 2.4, 12), np.ones(12)]))[:3].T
    origin = np.tile([[1.8, 0.2, 1.6]], (12, 1))
    rays = cb.beam_ray_segments(beam, origin, axis_pts - origin)
    assert rays.n_segments == 12
    scene = EmissionScene(flat)
    got, st = scene.render(rays, out_of_domain="count")
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"] and ref.max() > 0
    assert abs(st["out_of_domain"] - rst["out_of_domain"]) <= 4     # a few edge samples leave the tabulated ranges: clamped and counted
    assert parity(got, ref) <= 1.0


def test_beam_emission_multiplet_slab():
    # core/tests/test_lineshapes.py:391-472 inputs through the CUDA path: the MSE multiplet against the reference's closed form
    from test_oracle_beam import mse_case, mse_unit_shape
    flat, rays, d = mse_case()
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"]
    assert parity(got, ref) <= 1.0
    shape, delta = mse_unit_shape(d)
    total = got[0].sum() * delta
    assert parity((got[0] / total)[None, :], (shape / (shape.sum() * delta))[None, :]) <= 1.0


def test_beam_cx_and_emission_generomak_tabulated_rates():
    # both beam models on the Generomak plasma with ADF12 / ADF21 / ADF22-shaped synthetic tables (BASELINE config C5 shape)
    plasma = generomak.get_plasma()
    atomic = cb.SyntheticADAS()
    balmer = atomic.wavelength
    atomic.wavelength = lambda ion, charge, transition: 529.05 if ion is cb.carbon else balmer(ion, charge, transition)
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.05, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamEmissionLine(cb.Line(cb.deuterium, 0, (3, 2))), cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    axis_pts = (np.asarray(beam.transform) @ np.stack([np.zeros(16), np.zeros(16), np.linspace(0.6, 2.4, 16), np.ones(16)]))[:3].T
    origin = np.tile([[1.8, 0.2, 1.6]], (16, 1))
    rays = cb.beam_ray_segments(beam, origin, axis_pts - origin)
    for lo, hi, bins in ((526.0, 532.0, 256), (650.0, 662.0, 1024)):       # the CX line window and the Balmer-alpha / MSE window
        flat = cb.flatten_beam_scene(beam, lo, hi, bins)
        scene = EmissionScene(flat)
        got, st = scene.render(rays)
        scene.close()
        ref, rst = oracle.emission_render(flat, rays)
        assert st["samples"] == rst["samples"] and ref.max() > 0
        assert parity(got, ref) <= 1.0, (lo, hi)


def test_beam_cx_metastables_slab_and_generomak():
    # excited donor metastables weighted by their beam populations (charge_exchange.pyx:204-292)
    from test_oracle_beam import metastable_case
    beam, flat, rays = metastable_case()
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"]
    assert parity(got, ref) <= 1.0

    class MetaADAS(cb.SyntheticADAS):
        """ADF12-shaped tables for the donor ground state and n = 2, ADF22-shaped BeamPopulationRate tables per species."""

        def wavelength(self, ion, charge, transition):
            return 529.05

        def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
            g = cb.SyntheticADAS.beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition)[0]
            e = cb.BeamCXTable(2, g.eb, g.ti, g.ni, g.z, g.b, 30.0 * g.qeb * (g.eb / 4e4) ** -0.4, g.qti * 1.1, g.qni, g.qz * 0.9, g.qb, g.qref,
                               extrapolate=self.permit_extrapolation)
            return [g, e]

        def beam_population_rate(self, beam_ion, metastable, plasma_ion, charge):
            if charge == 0:
                return None
            e, n, t = np.logspace(3.5, 5.5, 25), np.logspace(17.0, 21.5, 26), np.logspace(0.0, 4.5, 16)
            sref = 4.0e-3
            sen = sref * (1 + 0.03 * charge) * (e[:, None] / 4e4) ** 0.2 * (n[None, :] / 1e19) ** 0.1
            st = sref * (1 + 0.04 * np.log10(t / 1e3))
            return cb.BeamStoppingTable(e, n, t, sen, st, sref, extrapolate=self.permit_extrapolation)

    plasma = generomak.get_plasma()
    atomic = MetaADAS()
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.05, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    flat = cb.flatten_beam_scene(beam, 526.0, 532.0, 256)
    axis_pts = (np.asarray(beam.transform) @ np.stack([np.zeros(12), np.zeros(12), np.linspace(0.6, 2.4, 12), np.ones(12)]))[:3].T
    origin = np.tile([[1.8, 0.2, 1.6]], (12, 1))
    rays = cb.beam_ray_segments(beam, origin, axis_pts - origin)
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    ground_only = cb.SyntheticADAS()
    ground_only.wavelength = atomic.wavelength
    beam.atomic_data = plasma.atomic_data = ground_only
    ref1, _ = oracle.emission_render(cb.flatten_beam_scene(beam, 526.0, 532.0, 256), rays)
    assert st["samples"] == rst["samples"] and ref.max() > 0
    assert np.abs(ref - ref1).max() > 1e-3 * ref.max()                    # the excited state does change the answer
    assert parity(got, ref) <= 1.0


def test_beam_emission_ratio_functions():
    # MSE intensity ratios as functions of the electron density: constant slab (equals the constant-ratio case) and the Generomak
    # plasma, where ne varies by decades along the beam
    from test_oracle_beam import mse_case
    flat, rays, _ = mse_case(ratio_functions=True)
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"] and parity(got, ref) <= 1.0
    plasma = generomak.get_plasma()
    atomic = cb.SyntheticADAS(permit_extrapolation=True)
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.05, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamEmissionLine(cb.Line(cb.deuterium, 0, (3, 2)),
                                       sigma_to_pi=lambda ne, e: 0.4 + 0.3 / (1.0 + (ne / 3e19) ** 0.7) + 1e-6 * e,
                                       sigma1_to_sigma0=lambda ne: 0.7 + 0.1 * np.tanh(np.log10(ne) - 19.0), pi2_to_pi3=0.31,
                                       pi4_to_pi3=lambda ne: 0.73 * (ne / 1e19) ** 0.05)]
    flat = cb.flatten_beam_scene(beam, 650.0, 662.0, 1024)
    axis_pts = (np.asarray(beam.transform) @ np.stack([np.zeros(10), np.zeros(10), np.linspace(0.6, 2.4, 10), np.ones(10)]))[:3].T
    origin = np.tile([[1.8, 0.2, 1.6]], (10, 1))
    rays = cb.beam_ray_segments(beam, origin, axis_pts - origin)
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"] and ref.max() > 0
    assert parity(got, ref) <= 1.0


class NarrowBeamADAS(cb.SyntheticADAS):
    """Beam tables that stop short of what the Generomak beam scene asks for — interaction energy (60 keV/amu against a table
    ending at 40), target density, temperature — so that every lookup family is taken outside its table."""

    def _cut(self, r):
        e, n, t = slice(0, 18), slice(8, 22), slice(3, 12)
        return cb.BeamStoppingTable(r.e[e], r.n[n], r.t[t], r.sen[e, n], r.st[t], r.sref, extrapolate=r.extrapolate)

    def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
        r = cb.SyntheticADAS.beam_stopping_rate(self, beam_ion, plasma_ion, charge)
        return None if r is None else self._cut(r)

    def beam_emission_pec(self, beam_ion, plasma_ion, charge, transition):
        r = cb.SyntheticADAS.beam_emission_pec(self, beam_ion, plasma_ion, charge, transition)
        return None if r is None else self._cut(r)

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        g = cb.SyntheticADAS.beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition)[0]
        s = slice(0, 12)                                                       # E up to ~30 keV/amu, Ti and Zeff grids cut too
        return [cb.BeamCXTable(1, g.eb[s], g.ti[:8], g.ni, g.z[2:], g.b, g.qeb[s], g.qti[:8], g.qni, g.qz[2:], g.qb, g.qref,
                               extrapolate=g.extrapolate)]


def _narrow_beam_scene(permit):
    plasma = generomak.get_plasma()
    atomic = NarrowBeamADAS(permit_extrapolation=permit)
    balmer = atomic.wavelength
    atomic.wavelength = lambda ion, charge, transition: 529.05 if ion is cb.carbon else balmer(ion, charge, transition)
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.05, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamEmissionLine(cb.Line(cb.deuterium, 0, (3, 2))), cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    axis_pts = (np.asarray(beam.transform) @ np.stack([np.zeros(12), np.zeros(12), np.linspace(0.6, 2.4, 12), np.ones(12)]))[:3].T
    origin = np.tile([[1.8, 0.2, 1.6]], (12, 1))
    return beam, cb.beam_ray_segments(beam, origin, axis_pts - origin)


def test_beam_rates_extrapolate_like_the_reference():
    # extrapolate=True: 'linear' (2-D) / 'quadratic' (1-D) for the ADF21 / ADF22 tables, 'quadratic' in log10 E and 'nearest' for the
    # factors of the ADF12 table (openadas/rates/beam.pyx:73-84, cx.pyx:96-102) — host attenuation, device kernels and oracle agree
    beam, rays = _narrow_beam_scene(True)
    z = np.linspace(0.05, 2.9, 40)
    for lo, hi, bins in ((526.0, 532.0, 256), (650.0, 662.0, 512)):
        flat = cb.flatten_beam_scene(beam, lo, hi, bins)
        scene = EmissionScene(flat)
        got, st = scene.render(rays)
        dens, _ = beam_sample(scene, np.stack([0 * z, 0 * z, z], axis=1))
        scene.close()
        ref, rst = oracle.emission_render(flat, rays)
        rd, _ = oracle.beam_sample(flat, np.stack([0 * z, 0 * z, z], axis=1))
        assert st["out_of_domain"] == 0 and rst["out_of_domain"] == 0
        assert ref.max() > 0 and parity(got, ref) <= 1.0, (lo, hi)
        assert np.max(np.abs(dens / rd - 1)) < 3e-5
    # the extrapolation matters: the clamped ('nearest') result of the same tables differs visibly
    beam_c, _ = _narrow_beam_scene(False)
    flat_c = cb.flatten_beam_scene(beam_c, 650.0, 662.0, 512)
    clamped, cst = oracle.emission_render(flat_c, rays)
    assert cst["out_of_domain"] > 0
    assert np.max(np.abs(clamped - ref)) > 1e-3 * ref.max()


def test_beam_rates_without_extrapolation_raise():
    beam, rays = _narrow_beam_scene(False)
    flat = cb.flatten_beam_scene(beam, 650.0, 662.0, 256)
    scene = EmissionScene(flat)
    with pytest.raises(ValueError):
        scene.render(rays)
    got, st = scene.render(rays, out_of_domain="count")
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["out_of_domain"] > 0 and abs(st["out_of_domain"] - rst["out_of_domain"]) <= 2e-3 * rst["out_of_domain"] + 4
    assert parity(got, ref) <= 1.0
