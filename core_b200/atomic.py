"""Atomic-data side of the hot path: elements, lines and ADAS-format rate providers.

Host-side mirror of the parts of ``cherab.core.atomic`` / ``cherab.openadas`` the emission path touches
(cherab/core/atomic/elements.pyx:243-402, line.pyx:20-80, interface.pyx:24-207; cherab/openadas/rates/pec.pyx:27-140).
Rates stay in their on-disk ADAS/OpenADAS shape ({'ne','te','rate'} in m^-3, eV, photon m^3/s) until the
scene flattener hands them to the CUDA library, which builds the log10-log10 bicubic tables.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class Element:
    """cherab/core/atomic/elements.pyx Element/Isotope: name, symbol, atomic_number, atomic_weight."""

    def __init__(self, name, symbol, atomic_number, atomic_weight, element=None):
        self.name, self.symbol = name, symbol
        self.atomic_number, self.atomic_weight = int(atomic_number), float(atomic_weight)
        self.element = element or self  # isotopes point at their element (elements.pyx:364)

    def __repr__(self):
        return "<Element: %s>" % self.name


hydrogen = Element("hydrogen", "H", 1, (1.00784 + 1.00811) / 2)
helium = Element("helium", "He", 2, 4.002602)
beryllium = Element("beryllium", "Be", 4, 9.0121831)
carbon = Element("carbon", "C", 6, (12.0096 + 12.0116) / 2)
nitrogen = Element("nitrogen", "N", 7, (14.00643 + 14.00728) / 2)
oxygen = Element("oxygen", "O", 8, (15.99903 + 15.99977) / 2)
neon = Element("neon", "Ne", 10, 20.1797)
argon = Element("argon", "Ar", 18, (39.792 + 39.963) / 2)
deuterium = Element("deuterium", "D", 1, 2.0141017778, hydrogen)
tritium = Element("tritium", "T", 1, 3.0160492777, hydrogen)


class Line:
    """cherab/core/atomic/line.pyx:20-80."""

    def __init__(self, element, charge, transition):
        if charge > element.atomic_number - 1:
            raise ValueError("Charge state cannot be larger than one less than the atomic number.")
        if charge < 0:
            raise ValueError("Charge state cannot be less than zero.")
        self.element, self.charge, self.transition = element, int(charge), tuple(transition)

    def __repr__(self):
        return "<Line: %s, %d, %s>" % (self.element.name, self.charge, self.transition)

    def __hash__(self):
        return hash((self.element.name, self.charge, self.transition))

    def __eq__(self, other):
        return isinstance(other, Line) and (self.element.name, self.charge, self.transition) == \
            (other.element.name, other.charge, other.transition)


class RateTable:
    """An ADF15-shaped PEC block: the ``data`` dict ImpactExcitationPEC/RecombinationPEC take (pec.pyx:48-68)."""

    def __init__(self, ne, te, rate, extrapolate=False):
        self.ne = np.ascontiguousarray(ne, dtype=np.float64)
        self.te = np.ascontiguousarray(te, dtype=np.float64)
        self.rate = np.ascontiguousarray(rate, dtype=np.float64)
        if self.rate.shape != (self.ne.size, self.te.size):
            raise ValueError("rate must have shape (len(ne), len(te))")
        self.extrapolate = bool(extrapolate)


class RateTable3D:
    """A thermal-CX PEC block: the ``data`` dict ThermalCXPEC takes, {'ne','te','td','rate'} (openadas/rates/pec.pyx:153-184)."""

    def __init__(self, ne, te, td, rate, extrapolate=False):
        self.ne = np.ascontiguousarray(ne, dtype=np.float64)
        self.te = np.ascontiguousarray(te, dtype=np.float64)
        self.td = np.ascontiguousarray(td, dtype=np.float64)
        self.rate = np.ascontiguousarray(rate, dtype=np.float64)
        if self.rate.shape != (self.ne.size, self.te.size, self.td.size):
            raise ValueError("rate must have shape (len(ne), len(te), len(td))")
        if min(self.rate.shape) < 2:
            raise ValueError("thermal CX rate tables need at least 2 knots per axis")
        self.extrapolate = bool(extrapolate)


class ConstantRate:
    """Constant-valued rate in W m^3, as the mock AtomicData of core/tests/test_line_emission.py:32-88 returns."""

    def __init__(self, value):
        self.value = float(value)


class AtomicData:
    """The subset of cherab/core/atomic/interface.pyx:24-207 the emission path calls."""

    def wavelength(self, ion, charge, transition):
        raise NotImplementedError("The wavelength() virtual method is not implemented for this atomic data source.")

    def impact_excitation_pec(self, ion, charge, transition):
        raise NotImplementedError("The impact_excitation() virtual method is not implemented for this atomic data source.")

    def recombination_pec(self, ion, charge, transition):
        raise NotImplementedError("The recombination() virtual method is not implemented for this atomic data source.")

    def thermal_cx_pec(self, donor_ion, donor_charge, receiver_ion, receiver_charge, transition):
        raise NotImplementedError("The thermal_cx_pec() virtual method is not implemented for this atomic data source.")

    def line_radiated_power_rate(self, ion, charge):
        raise NotImplementedError("The line_radiated_power_rate() virtual method is not implemented for this atomic data source.")

    def continuum_radiated_power_rate(self, ion, charge):
        raise NotImplementedError("The continuum_radiated_power_rate() virtual method is not implemented for this atomic data source.")

    def cx_radiated_power_rate(self, ion, charge):
        raise NotImplementedError("The cx_radiated_power_rate() virtual method is not implemented for this atomic data source.")

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        """List of effective CX emission coefficients, one per donor metastable (interface.pyx:86-95)."""
        raise NotImplementedError("The cxs_rates() virtual method is not implemented for this atomic data source.")

    def beam_population_rate(self, beam_ion, metastable, plasma_ion, charge):
        """Population of a beam metastable relative to the ground state (interface.pyx:97-107); a BeamStoppingTable-shaped
        dimensionless table, a ConstantRate or None (NullBeamPopulationRate)."""
        raise NotImplementedError("The beam_population() virtual method is not implemented for this atomic data source.")

    def beam_emission_pec(self, beam_ion, plasma_ion, charge, transition):
        raise NotImplementedError("The beam_emission() virtual method is not implemented for this atomic data source.")

    def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
        raise NotImplementedError("The beam_stopping() virtual method is not implemented for this atomic data source.")

    def free_free_gaunt_factor(self):
        """MaxwellianFreeFreeGauntFactor table (cherab/core/atomic/gaunt.pyx:143-158): (u, gamma2, gaunt_factor)."""
        t = np.load(os.path.join(_DATA, "atomic_tables.npz"))
        return t["gaunt_u"], t["gaunt_gamma2"], t["gaunt_factor"]

    def stark_model_coefficients(self, line):
        """(c_ij, a_ij, b_ij) — cherab/core/atomic/interface.pyx:176-196."""
        t = np.load(os.path.join(_DATA, "atomic_tables.npz"))
        data = json.loads(str(t["stark_json"]))
        sym = line.element.symbol.lower()
        if sym not in data:
            raise ValueError("Stark broadening coefficients for {} is not currently available.".format(line))
        key = "%s -> %s" % (line.transition[0], line.transition[1])
        try:
            return tuple(data[sym][str(line.charge)][key])
        except KeyError:
            raise ValueError("Stark broadening coefficients for {} is not currently available.".format(line))

    def zeeman_triplet_parameters(self, line):
        """(alpha, beta, gamma) — cherab/core/atomic/interface.pyx:154-174."""
        t = np.load(os.path.join(_DATA, "atomic_tables.npz"))
        data = json.loads(str(t["zeeman_parametrised_json"]))
        sym = line.element.symbol.lower()
        key = "%s -> %s" % (line.transition[0], line.transition[1])
        try:
            return tuple(data[sym][str(line.charge)][key])
        except KeyError:
            raise ValueError("Data for {} is not available.".format(line))


# Balmer wavelengths the OpenADAS repository ships for hydrogen (cherab/openadas/repository/create.py:224-241)
_H_BALMER = {(3, 2): 656.279, (4, 2): 486.135, (5, 2): 434.047, (6, 2): 410.173, (7, 2): 397.008}
# deuterium: reduced-mass scaled from H0, with the observed NIST values where the repository overrides them (create.py:296-311)
_D_BALMER = {(3, 2): 656.101, (4, 2): 486.000, (5, 2): 433.928, (6, 2): 410.062, (7, 2): 396.899}


class SyntheticADAS(AtomicData):
    """Deterministic synthetic rates with exactly the OpenADAS JSON shapes: no repository is present
    in this environment (it is downloaded by cherab.openadas.repository.populate()).

    excitation   : 1e-14 (te/10)^-0.5 exp(-E_n/te) (1 + 0.1 log10(ne/1e19))  photon m^3/s
    recombination: 3e-19 (te/10)^-0.7                                         photon m^3/s
    on ne = logspace(13.7, 21.3, 24) m^-3, te = logspace(-0.7, 4, 29) eV, scaled per upper level n.
    """

    def __init__(self, permit_extrapolation=True):
        self.permit_extrapolation = permit_extrapolation
        self.ne = np.logspace(13.7, 21.3, 24)
        self.te = np.logspace(-0.7, 4.0, 29)

    def wavelength(self, ion, charge, transition):
        table = _D_BALMER if ion.name == "deuterium" else _H_BALMER
        try:
            return table[(int(transition[0]), int(transition[1]))]
        except (KeyError, ValueError, TypeError):
            raise RuntimeError("Requested wavelength is not available: %s %d %s" % (ion.name, charge, transition))

    def _scale(self, transition):
        n = int(transition[0])
        return (3.0 / n) ** 3, 13.605693122994 * (1.0 - 1.0 / n ** 2)

    def impact_excitation_pec(self, ion, charge, transition):
        s, e_n = self._scale(transition)
        ne, te = self.ne[:, None], self.te[None, :]
        rate = s * 1e-14 * (te / 10.0) ** -0.5 * np.exp(-e_n / te) * (1.0 + 0.1 * np.log10(ne / 1e19))
        return RateTable(self.ne, self.te, rate, self.permit_extrapolation)

    def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
        """Synthetic ADF21-shaped stopping coefficient: sen[e 25 x n 26], st[t 16], smooth and positive."""
        from .beam import BeamStoppingTable
        if charge == 0:
            return None                       # no beam stopping data for neutrals (NullBeamStoppingRate)
        e, n, t = np.logspace(3.5, 5.5, 25), np.logspace(17.0, 21.5, 26), np.logspace(0.0, 4.5, 16)
        sref = 1.0e-13
        sen = sref * (1 + 0.05 * charge) * (e[:, None] / 4e4) ** -0.35 * (1 + 0.08 * np.log10(n[None, :] / 1e19))
        st = sref * (1 + 0.05 * np.log10(t / 1e3))
        return BeamStoppingTable(e, n, t, sen, st, sref, extrapolate=self.permit_extrapolation)

    def beam_emission_pec(self, beam_ion, plasma_ion, charge, transition):
        """Synthetic ADF22-shaped beam emission coefficient (photon m^3 s^-1): same grids as the stopping coefficient."""
        from .beam import BeamStoppingTable
        if charge == 0:
            return None
        e, n, t = np.logspace(3.5, 5.5, 25), np.logspace(17.0, 21.5, 26), np.logspace(0.0, 4.5, 16)
        sref = 3.0e-15
        sen = sref * (1 + 0.03 * charge) * (e[:, None] / 4e4) ** 0.2 * (1 - 0.06 * np.log10(n[None, :] / 1e19))
        st = sref * (1 + 0.04 * np.log10(t / 1e3))
        return BeamStoppingTable(e, n, t, sen, st, sref, extrapolate=self.permit_extrapolation)

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        """Synthetic ADF12-shaped effective CX emission coefficient: eb[24], ti[12], ni[24], zeff[12], b[12]."""
        from .beam import BeamCXTable
        eb, ti, ni = np.logspace(3.3, 5.6, 24), np.logspace(0.5, 4.5, 12), np.logspace(17.0, 21.5, 24)
        z, b = np.linspace(1.0, 6.0, 12), np.linspace(0.0, 12.0, 12)
        qref = 2.0e-15
        qeb = qref * np.exp(-0.5 * (np.log(eb / 4e4) / 1.2) ** 2) + 1e-18
        qti = qref * (1 + 0.1 * np.log10(ti / 1e3))
        qni = qref * (1 - 0.05 * np.log10(ni / 1e19))
        qz = qref * (1 + 0.03 * (z - 2.0))
        qb = qref * (1 + 0.004 * b)
        return [BeamCXTable(1, eb, ti, ni, z, b, qeb, qti, qni, qz, qb, qref, extrapolate=self.permit_extrapolation)]

    def thermal_cx_pec(self, donor_ion, donor_charge, receiver_ion, receiver_charge, transition):
        """Synthetic ThermalCXPEC-shaped table on (ne[24], te[29], td[12]): smooth, positive, falling with the donor charge."""
        s, _ = self._scale(transition)
        td = np.logspace(-0.7, 4.0, 12)
        ne, te, tdd = self.ne[:, None, None], self.te[None, :, None], td[None, None, :]
        rate = (s * 2e-15 / (1.0 + donor_charge) * (tdd / 10.0) ** 0.3 / (1.0 + tdd / 3e3) * (te / 10.0) ** -0.1
                * (1.0 + 0.05 * np.log10(ne / 1e19)))
        return RateTable3D(self.ne, self.te, td, rate, self.permit_extrapolation)

    def recombination_pec(self, ion, charge, transition):
        s, _ = self._scale(transition)
        ne, te = self.ne[:, None], self.te[None, :]
        rate = s * 3e-19 * (te / 10.0) ** -0.7 * np.ones_like(ne)
        return RateTable(self.ne, self.te, rate, self.permit_extrapolation)
