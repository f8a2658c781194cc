"""Pin the oracle's line shapes against the reference's own known-answer tests.

Right-hand sides are the numpy/scipy closed forms of cherab/core/tests/test_lineshapes.py:62-389 with the same inputs
and tolerances (abs 1e-10 for the Gaussian family, rel 1e-8 for Stark); the left-hand side is the oracle run through
the whole path (flattened slab scene, unit radiance, 1 m chord)."""
import numpy as np
import pytest
from scipy.integrate import quad
from scipy.special import erf, hyp2f1

import core_b200 as cb
from oracle import oracle
from helpers import (ATOMIC_MASS, BOHR_MAGNETON, ELEMENTARY_CHARGE, HC_EV_NM, SPEED_OF_LIGHT, UnitRadianceAtomicData,
                     lineshape_slab_plasma, slab_ray)


def render(plasma, line, wavelength, lo, hi, bins, direction, shape=None, args=None, kwargs=None, **flat_kw):
    target = plasma.composition.get(line.element, line.charge)
    plasma.atomic_data = UnitRadianceAtomicData(1e19, target.distribution.density.value, wavelength)
    plasma.models = [cb.ExcitationLine(line, lineshape=shape, lineshape_args=args, lineshape_kwargs=kwargs)]
    fs = cb.flatten_scene(plasma, lo, hi, bins, **flat_kw)
    out, stats = oracle.emission_render(fs, slab_ray(direction))
    return out[0], stats


def test_stark_norm_constant():
    # stark.pyx:62 STARK_NORM_COEFFICIENT hard-coded in oracle/cb2_oracle.c
    assert abs(4 * 50 * hyp2f1(0.4, 1, 1.4, -(2 * 50) ** 2.5) - 2.641279471021934) < 1e-13


def test_gaussian_line():
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.deuterium, 0, (3, 2))
    wavelength, bins = 656.104, 256
    lo, hi = wavelength - 0.5, wavelength + 0.5
    direction = np.array([-1.0, 0, 0])
    got, stats = render(plasma, line, wavelength, lo, hi, bins, direction)
    temperature, velocity = 5.0, np.array([2e4, 0, 0])
    shifted = wavelength * (1 + velocity.dot(direction) / SPEED_OF_LIGHT)
    sigma = np.sqrt(temperature * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    wl, delta = np.linspace(lo, hi, bins + 1, retstep=True)
    erfs = erf((wl - shifted) / (np.sqrt(2.) * sigma))
    ref = 0.5 * (erfs[1:] - erfs[:-1]) / delta
    assert np.max(np.abs(got - ref)) < 1e-10
    assert stats["gaussian_bin_evals"] > 0 and stats["samples"] == 1001


def test_multiplet_line_shape():
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.nitrogen, 1, ("2s2 2p1 4f1 3G13.0", "2s2 2p1 3d1 3F10.0"))
    multiplet = [[403.509, 404.132, 404.354, 404.479, 405.692], [0.205, 0.562, 0.175, 0.029, 0.029]]
    wavelength, bins = 404.21, 512
    lo, hi = min(multiplet[0]) - 0.5, max(multiplet[0]) + 0.5
    direction = np.array([-1.0, 0, 0])
    got, _ = render(plasma, line, wavelength, lo, hi, bins, direction, cb.MultipletLineShape, [multiplet])
    temperature, velocity = 10.0, np.array([1e4, 5e4, 0])
    sigma = np.sqrt(temperature * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    doppler = 1 + velocity.dot(direction) / SPEED_OF_LIGHT
    wl, delta = np.linspace(lo, hi, bins + 1, retstep=True)
    ref = 0
    for w, ratio in zip(*multiplet):
        erfs = erf((wl - w * doppler) / (np.sqrt(2.) * sigma))
        ref = ref + 0.5 * ratio * (erfs[1:] - erfs[:-1]) / delta
    assert np.max(np.abs(got - ref)) < 1e-10


def _triplet_reference(wavelength, lo, hi, bins, direction, sigma, wl_plus, wl_minus):
    velocity, b_field = np.array([2e4, 0, 0]), np.array([0, 5.0, 0])
    doppler = 1 + velocity.dot(direction) / SPEED_OF_LIGHT
    b_magn = np.linalg.norm(b_field)
    cos_sqr = (b_field.dot(direction) / b_magn) ** 2
    sin_sqr = 1. - cos_sqr
    temp = 1. / (np.sqrt(2.) * sigma)
    wl, delta = np.linspace(lo, hi, bins + 1, retstep=True)
    erfs = erf((wl - wavelength * doppler) * temp)
    g_pi = 0.5 * (erfs[1:] - erfs[:-1]) / delta
    erfs = erf((wl - wl_plus * doppler) * temp)
    g_sigma = 0.5 * (erfs[1:] - erfs[:-1]) / delta
    erfs = erf((wl - wl_minus * doppler) * temp)
    g_sigma = g_sigma + 0.5 * (erfs[1:] - erfs[:-1]) / delta
    tri = {"pi": 0.5 * sin_sqr * g_pi, "sigma": (0.25 * sin_sqr + 0.5 * cos_sqr) * g_sigma}
    tri["no"] = tri["pi"] + tri["sigma"]
    return tri


@pytest.mark.parametrize("pol", ["no", "pi", "sigma"])
def test_zeeman_triplet(pol):
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.deuterium, 0, (3, 2))
    wavelength, bins = 656.104, 256
    lo, hi = wavelength - 0.5, wavelength + 0.5
    direction = np.array([-1.0, 1.0, 0]) / np.sqrt(2)
    got, _ = render(plasma, line, wavelength, lo, hi, bins, direction, cb.ZeemanTriplet, kwargs={"polarisation": pol})
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    pe = HC_EV_NM / wavelength
    ref = _triplet_reference(wavelength, lo, hi, bins, direction, sigma,
                             HC_EV_NM / (pe - BOHR_MAGNETON * 5.0), HC_EV_NM / (pe + BOHR_MAGNETON * 5.0))[pol]
    assert np.max(np.abs(got - ref)) < 1e-10


@pytest.mark.parametrize("pol", ["no", "pi", "sigma"])
def test_parametrised_zeeman_triplet(pol):
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.deuterium, 0, (3, 2))
    wavelength, bins = 656.104, 256
    lo, hi = wavelength - 0.5, wavelength + 0.5
    direction = np.array([-1.0, 1.0, 0]) / np.sqrt(2)
    alpha, beta, gamma = cb.AtomicData().zeeman_triplet_parameters(line)
    got, _ = render(plasma, line, wavelength, lo, hi, bins, direction, cb.ParametrisedZeemanTriplet, kwargs={"polarisation": pol})
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    sigma *= np.sqrt(1. + beta * beta * 5.0 ** (2. * gamma))
    ref = _triplet_reference(wavelength, lo, hi, bins, direction, sigma,
                             wavelength + 0.5 * alpha * 5.0, wavelength - 0.5 * alpha * 5.0)[pol]
    assert np.max(np.abs(got - ref)) < 1e-10


@pytest.mark.parametrize("pol", ["no", "pi", "sigma"])
def test_zeeman_multiplet(pol):
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.deuterium, 0, (3, 2))
    wavelength, bins = 656.104, 256
    lo, hi = wavelength - 0.5, wavelength + 0.5
    direction = np.array([-1.0, 1.0, 0]) / np.sqrt(2)
    pe = HC_EV_NM / wavelength
    zs = cb.ZeemanStructure([(wavelength, 1.0)],
                            [(lambda b: HC_EV_NM / (pe - BOHR_MAGNETON * b), 0.5)],
                            [(lambda b: HC_EV_NM / (pe + BOHR_MAGNETON * b), 0.5)])
    got, _ = render(plasma, line, wavelength, lo, hi, bins, direction, cb.ZeemanMultiplet, [zs], {"polarisation": pol})
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    ref = _triplet_reference(wavelength, lo, hi, bins, direction, sigma,
                             HC_EV_NM / (pe + BOHR_MAGNETON * 5.0), HC_EV_NM / (pe - BOHR_MAGNETON * 5.0))[pol]
    assert np.max(np.abs(got - ref)) < 1e-10


def test_stark_broadened_line():
    plasma = lineshape_slab_plasma()
    line = cb.Line(cb.deuterium, 0, (6, 2))
    wavelength, bins, rtol = 656.104, 512, 1e-8
    lo, hi = wavelength - 0.2, wavelength + 0.2
    direction = np.array([-1.0, 1.0, 0]) / np.sqrt(2)
    got, stats = render(plasma, line, wavelength, lo, hi, bins, direction, cb.StarkBroadenedLine, quad_rtol=rtol)
    assert stats["lorentzian_bin_evals"] > 0
    # a 1 m chord at step 1 mm integrates a constant emissivity: divide nothing, compare per bin
    velocity, b_field = np.array([2e4, 0, 0]), np.array([0, 5.0, 0])
    doppler = 1 + velocity.dot(direction) / SPEED_OF_LIGHT
    b_magn = 5.0
    pe = HC_EV_NM / wavelength
    wl_plus, wl_minus = HC_EV_NM / (pe - BOHR_MAGNETON * b_magn), HC_EV_NM / (pe + BOHR_MAGNETON * b_magn)
    cos_sqr = (b_field.dot(direction) / b_magn) ** 2
    sin_sqr = 1. - cos_sqr
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (line.element.atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    fwhm_gauss = 2 * np.sqrt(2 * np.log(2)) * sigma
    cij, aij, bij = cb.AtomicData().stark_model_coefficients(line)
    fwhm_lorentz = cij * 1e19 ** aij / (20. ** bij)
    if fwhm_gauss <= fwhm_lorentz:
        c = [1., 0, 0.57575, 0.37902, -0.42519, -0.31525, 0.31718]
        fwhm_full = fwhm_lorentz * np.poly1d(c[::-1])(fwhm_gauss / fwhm_lorentz)
    else:
        c = [1., 0.15882, 1.04388, -1.38281, 0.46251, 0.82325, -0.58026]
        fwhm_full = fwhm_gauss * np.poly1d(c[::-1])(fwhm_lorentz / fwhm_gauss)
    wl, delta = np.linspace(lo, hi, bins + 1, retstep=True)
    temp = 2 * np.sqrt(np.log(2)) / fwhm_full
    erfs = erf((wl - wavelength * doppler) * temp)
    gaussian = 0.25 * sin_sqr * (erfs[1:] - erfs[:-1]) / delta
    for w in (wl_plus, wl_minus):
        erfs = erf((wl - w * doppler) * temp)
        gaussian += 0.5 * (0.25 * sin_sqr + 0.5 * cos_sqr) * (erfs[1:] - erfs[:-1]) / delta
    norm = (0.5 * fwhm_full) ** 1.5 / (4 * 50 * hyp2f1(0.4, 1, 1.4, -(2 * 50) ** 2.5))

    def shape(x, x0):
        return norm / (np.abs(x - x0 * doppler) ** 2.5 + (0.5 * fwhm_full) ** 2.5)

    wpoly = [5.14820e-04, 1.38821e+00, -9.60424e-02, -3.83995e-02, -7.40042e-03, -5.47626e-04]
    lw = np.exp(np.poly1d(wpoly[::-1])(np.log(fwhm_lorentz / fwhm_full)))
    for i in range(bins):
        lb = 0.5 * sin_sqr * quad(shape, wl[i], wl[i + 1], args=(wavelength,), epsrel=rtol)[0]
        lb += (0.25 * sin_sqr + 0.5 * cos_sqr) * quad(shape, wl[i], wl[i + 1], args=(wl_plus,), epsrel=rtol)[0]
        lb += (0.25 * sin_sqr + 0.5 * cos_sqr) * quad(shape, wl[i], wl[i + 1], args=(wl_minus,), epsrel=rtol)[0]
        ref = lb / delta * lw + gaussian[i] * (1. - lw)
        assert abs(got[i] / ref - 1.) < rtol * 3, (i, got[i], ref)


def test_gaussian_edge_cases():
    # gaussian.pyx:57-68: sigma <= 0 and lines entirely outside the window add nothing
    z = oracle.add_gaussian_line(1.0, 500.0, 0.0, 499, 501, 16)
    assert not z.any()
    z = oracle.add_gaussian_line(1.0, 600.0, 0.1, 499, 501, 16)
    assert not z.any()
    z = oracle.add_gaussian_line(2.5, 500.0, 0.05, 499, 501, 64)
    assert abs(z.sum() * (2.0 / 64) - 2.5) < 1e-12  # the whole line is inside the window: bins integrate to radiance
