"""Slab test plasmas — cherab/tools/plasmas/slab.pyx:112-260 with the same arguments."""
from .atomic import hydrogen
from .geometry import Box
from .plasma import Maxwellian, Plasma, SlabIonFunction, SlabNeutralFunction, Species

ATOMIC_MASS = 1.66053906660e-27
ELECTRON_MASS = 9.1093837015e-31


def build_constant_slab_plasma(length=5, width=1, height=1, electron_density=1e19, electron_temperature=2.5e3,
                               plasma_species=None, b_field=(0, 0, 0)):
    """slab.pyx:198-260: constant conditions in Box((0,-w/2,-h/2),(length,w/2,h/2));
    plasma_species = [(element, charge, density, temperature, velocity), ...]."""
    if plasma_species is None:
        plasma_species = [(hydrogen, 1, electron_density, electron_temperature, (0, 0, 0))]
    plasma = Plasma()
    plasma.geometry = Box((0, -width / 2, -height / 2), (length, width / 2, height / 2))
    plasma.electron_distribution = Maxwellian(electron_density, electron_temperature, (0, 0, 0), ELECTRON_MASS)
    plasma.b_field = tuple(b_field)
    plasma.composition = [Species(el, ch, Maxwellian(n, t, tuple(v), el.atomic_weight * ATOMIC_MASS))
                          for el, ch, n, t, v in plasma_species]
    return plasma


def build_slab_plasma(length=5, width=1, height=1, peak_density=1e19, peak_temperature=2500, pedestal_top=1,
                      neutral_temperature=0.5, impurities=None):
    """slab.pyx:112-195: pedestal profiles along +x."""
    plasma = Plasma()
    plasma.geometry = Box((0, -width / 2, -height / 2), (length, width / 2, height / 2))
    zero = (0, 0, 0)
    species = [Species(hydrogen, 0, Maxwellian(SlabNeutralFunction(peak_density, 0.1, pedestal_top=pedestal_top),
                                               neutral_temperature, zero, hydrogen.atomic_weight * ATOMIC_MASS)),
               Species(hydrogen, 1, Maxwellian(SlabIonFunction(peak_density, 0, pedestal_top=pedestal_top),
                                               SlabIonFunction(peak_temperature, 0, pedestal_top=pedestal_top), zero,
                                               hydrogen.atomic_weight * ATOMIC_MASS))]
    for impurity, ionisation, concentration in (impurities or []):
        species.append(Species(impurity, ionisation,
                               Maxwellian(SlabIonFunction(peak_density * concentration, 0, pedestal_top=pedestal_top),
                                          SlabIonFunction(peak_temperature, 0, pedestal_top=pedestal_top), zero,
                                          impurity.atomic_weight * ATOMIC_MASS)))
    plasma.electron_distribution = Maxwellian(SlabIonFunction(peak_density, 0, pedestal_top=pedestal_top),
                                              SlabIonFunction(peak_temperature, 0, pedestal_top=pedestal_top), zero, ELECTRON_MASS)
    plasma.b_field = zero
    plasma.composition = species
    return plasma
