"""Generate tests/golden/*.npz.

Part A ("closed_forms.npz") needs numpy/scipy only: the right-hand sides of the reference's own known-answer tests
(cherab/core/tests/test_lineshapes.py:62-93 Gaussian, :95-131 multiplet, :133-191 Zeeman triplet,
test_line_emission.py:99-134 slab line, test_bremsstrahlung.py:41-95 slab continuum) evaluated with the inputs those
tests state.  Part B ("oracle_frames.npz") stores small frames rendered by the CPU oracle (oracle/cb2_oracle.c) for
the benchmark scenes, so that both the oracle and the CUDA path are regression-pinned at BASELINE's scene shapes.

    python tests/golden/make_golden.py            # rewrites both files
"""
import os
import sys

import numpy as np
from scipy import constants as const
from scipy.integrate import quad
from scipy.special import erf

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ATOMIC_MASS = 1.66053906660e-27
ELEMENTARY_CHARGE = 1.602176634e-19
SPEED_OF_LIGHT = 299792458.0
BOHR_MAGNETON = 5.78838180123e-5
HC_EV_NM = 1239.8419738620933


def gaussian_bins(wl, centre, sigma):
    e = erf((wl - centre) / (np.sqrt(2.0) * sigma))
    return 0.5 * (e[1:] - e[:-1]) / (wl[1] - wl[0])


def closed_forms():
    out = {}
    # test_lineshapes.py:62-93 — D-alpha, T = 5 eV, v = (2e4, 0, 0), line of sight (-1, 0, 0), 256 bins on +-0.5 nm
    wavelength, weight = 656.104, 2.0141017778          # deuterium.atomic_weight (cherab/core/atomic/elements.pyx)
    direction = np.array([-1.0, 0.0, 0.0])
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    wl = np.linspace(wavelength - 0.5, wavelength + 0.5, 257)
    out["gaussian"] = gaussian_bins(wl, wavelength * (1 + np.array([2e4, 0, 0]).dot(direction) / SPEED_OF_LIGHT), sigma)
    # test_lineshapes.py:95-131 — nitrogen II multiplet, T = 10 eV, v = (1e4, 5e4, 0)
    multiplet = [[403.509, 404.132, 404.354, 404.479, 405.692], [0.205, 0.562, 0.175, 0.029, 0.029]]
    w0, weight_n = 404.21, 14.006855                          # nitrogen.atomic_weight
    sigma = np.sqrt(10.0 * ELEMENTARY_CHARGE / (weight_n * ATOMIC_MASS)) * w0 / SPEED_OF_LIGHT
    doppler = 1 + np.array([1e4, 5e4, 0]).dot(direction) / SPEED_OF_LIGHT
    wl = np.linspace(min(multiplet[0]) - 0.5, max(multiplet[0]) + 0.5, 513)
    out["multiplet"] = sum(r * gaussian_bins(wl, w * doppler, sigma) for w, r in zip(*multiplet))
    # test_lineshapes.py:133-191 — Zeeman triplet, B = (0, 5, 0) T, line of sight (-1, 1, 0)/sqrt(2), no polarisation filter
    d2 = np.array([-1.0, 1.0, 0.0]) / np.sqrt(2.0)
    sigma = np.sqrt(5.0 * ELEMENTARY_CHARGE / (weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT
    doppler = 1 + np.array([2e4, 0, 0]).dot(d2) / SPEED_OF_LIGHT
    b = np.array([0, 5.0, 0])
    cos_sqr = (b.dot(d2) / 5.0) ** 2
    sin_sqr = 1.0 - cos_sqr
    photon_energy = HC_EV_NM / wavelength
    wl_plus, wl_minus = HC_EV_NM / (photon_energy - BOHR_MAGNETON * 5.0), HC_EV_NM / (photon_energy + BOHR_MAGNETON * 5.0)
    wl = np.linspace(wavelength - 0.5, wavelength + 0.5, 257)
    out["zeeman_triplet"] = 0.5 * sin_sqr * gaussian_bins(wl, wavelength * doppler, sigma) + \
        (0.25 * sin_sqr + 0.5 * cos_sqr) * (gaussian_bins(wl, wl_plus * doppler, sigma) + gaussian_bins(wl, wl_minus * doppler, sigma))
    # test_line_emission.py:99-134 — slab: ne 1e19, C5+ 2e18, PEC 1.4e-39 W m^3, chord 1.2 m: integrated radiance
    out["slab_line_radiance"] = np.array([1.4e-39 * 2e18 * 1e19 * 1.2 / (4 * np.pi)])
    # test_bremsstrahlung.py:41-95 needs the Gaunt-factor table: stored separately by the oracle part (B)
    return out


def oracle_frames():
    import core_b200 as cb
    from core_b200 import generomak
    from helpers import generomak_camera_rays
    from oracle import oracle
    out = {}
    # BASELINE config C1 at 4x4 px: Generomak H-alpha, excitation + recombination, 512 bins
    plasma = generomak.get_plasma()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.005)
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 512)
    out["c1_frame"], st = oracle.emission_render(flat, generomak_camera_rays(plasma, (4, 4)))
    out["c1_samples"] = np.array([st["samples"]])
    # BASELINE config C3 model mix at 2x2 px: 8 Balmer lines + Bremsstrahlung, 2048 bins
    plasma = generomak.get_plasma()
    lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4, 5, 6)]
    plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines] + [cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=0.02)
    flat = cb.flatten_scene(plasma, 390.0, 700.0, 2048)
    out["c3_frame"], st = oracle.emission_render(flat, generomak_camera_rays(plasma, (2, 2)))
    out["c3_samples"] = np.array([st["samples"]])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "closed_forms.npz"), **closed_forms())
    np.savez_compressed(os.path.join(HERE, "oracle_frames.npz"), **oracle_frames())
    for f in ("closed_forms.npz", "oracle_frames.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
