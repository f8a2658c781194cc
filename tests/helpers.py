"""Shared test scaffolding: the analytic scenes the reference's own tests use."""
import numpy as np

import core_b200 as cb
from core_b200.slab import build_constant_slab_plasma

ATOMIC_MASS = 1.66053906660e-27
ELEMENTARY_CHARGE = 1.602176634e-19
SPEED_OF_LIGHT = 299792458.0
BOHR_MAGNETON = 5.78838180123e-5
HC_EV_NM = 1239.8419738620933


class UnitRadianceAtomicData(cb.AtomicData):
    """Rates chosen so that RECIP_4_PI * PEC * ne * ni == 1: a 1 m slab then returns add_line(radiance=1) itself,
    which is how the reference's test_lineshapes.py right-hand sides are replayed through the full path."""

    def __init__(self, ne, ni, wavelength):
        self._pec = 4.0 * np.pi / (ne * ni)
        self._wl = wavelength

    def wavelength(self, ion, charge, transition):
        return self._wl

    def impact_excitation_pec(self, ion, charge, transition):
        return cb.ConstantRate(self._pec)

    def recombination_pec(self, ion, charge, transition):
        return cb.ConstantRate(self._pec)


def lineshape_slab_plasma():
    """core/tests/test_lineshapes.py:47-57."""
    species = [(cb.deuterium, 0, 1.e18, 5., (2.e4, 0, 0)), (cb.nitrogen, 1, 1.e17, 10., (1.e4, 5.e4, 0))]
    return build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=20.,
                                      plasma_species=species, b_field=(0, 5., 0))


def slab_ray(direction, length=1.0, centre=(0.5, 0.0, 0.0)):
    """One ray crossing the slab box through ``centre`` with a chord of exactly ``length`` (segments given directly)."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    o = np.asarray(centre) - d * (0.5 * length + 1.0)
    return cb.RayBatch(o[None, :], d[None, :], [0, 1], [1.0], [1.0 + length])


def generomak_camera_rays(plasma, pixels, sub=(0.5, 0.5), pixel_index=None):
    """The C1 pose of SURVEY 8(d): pinhole at (2.3, 0, 1.25) looking at (1.0, 0.8, -0.5), 45 degree FoV."""
    cam = cb.PinholeCamera(pixels, fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
    o, d = cam.rays(sub[0], sub[1], pixel_index)
    return cb.ray_segments(plasma.geometry, o, d, plasma.geometry_to_world())
