"""The Generomak example tokamak as a flattenable scene (the benchmark scene of BASELINE configs C1/C3).

Follows cherab/generomak/plasma/plasma.py:96-129 (edge interpolators), :233-272 (core interpolators), :580-638
(get_full_profiles: blend of edge mesh and flux-mapped core profiles), :132-163 (Maxwellians), :641-701 (get_plasma),
and cherab/generomak/equilibrium/equilibrium.py:8-40.  The data comes from core_b200/data/generomak.npz, generated
from the reference's JSON files by tools/make_data_tables.py.
"""
import os

import numpy as np

from .atomic import SyntheticADAS, carbon, hydrogen
from .geometry import HollowCylinder, translate
from .plasma import (AxisymBlend, AxisymBlendVector, AxisymContext, EFITEquilibrium, EFITMagneticField, Maxwellian,
                     Plasma, Species)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "generomak.npz")

ATOMIC_MASS = 1.66053906660e-27
ELECTRON_MASS = 9.1093837015e-31


def load_tables():
    return dict(np.load(_DATA))


def load_equilibrium(tables=None):
    """equilibrium.py:8-40."""
    t = tables or load_tables()
    return EFITEquilibrium(t["eq_r"], t["eq_z"], t["eq_psi_grid"], float(t["eq_psi_axis"]), float(t["eq_psi_lcfs"]),
                           t["eq_f_profile"], float(t["eq_b_vacuum_radius"]), float(t["eq_b_vacuum_magnitude"]),
                           t["eq_lcfs_polygon"])


def get_plasma(atomic_data=None, tables=None):
    """Full (core + edge) Generomak plasma — plasma.py:641-701 with get_2d_distributions(get_full_profiles())."""
    t = tables or load_tables()
    eq = load_equilibrium(t)
    plasma = Plasma(name="Generomak plasma")
    plasma.axisym = AxisymContext(eq, t["mesh_vertices"], t["mesh_triangles"], t["core_psi_norm"])
    edge_v = (0.0, 1e-10, 0.0)  # plasma.py:110,120: avoid zero-length vectors for blending

    def dist(key, mass):
        n = AxisymBlend(t["edge_%s_density" % key], t["core_%s_density" % key])
        temp = AxisymBlend(t["edge_%s_temperature" % key], t["core_%s_temperature" % key])
        v = AxisymBlendVector(edge_v, t["core_%s_vtor" % key], t["core_%s_vpol" % key], t["core_%s_vnorm" % key])
        return Maxwellian(n, temp, v, mass)

    plasma.electron_distribution = dist("electron", ELECTRON_MASS)
    composition = []
    for element, name in ((hydrogen, "hydrogen"), (carbon, "carbon")):
        for charge in range(element.atomic_number + 1):
            composition.append(Species(element, charge, dist("%s%d" % (name, charge), element.atomic_weight * ATOMIC_MASS)))
    plasma.composition = composition
    plasma.b_field = EFITMagneticField()
    plasma.atomic_data = atomic_data or SyntheticADAS(permit_extrapolation=True)
    # plasma.py:673-681: Subtract(Cylinder(r_max, h), Cylinder(r_min, h + 2 mm)), base translated to z_min
    r_range, z_range = eq.r_range, eq.z_range
    plasma.geometry = HollowCylinder(r_range[0], r_range[1], 0.0, z_range[1] - z_range[0])
    plasma.geometry_transform = translate(0, 0, z_range[0])
    return plasma
