"""OpenADAS atomic-data provider reading Cherab's on-disk JSON repository.

Restates the read side of cherab/openadas/openadas.py:24-100,318-395 and cherab/openadas/repository/{pec,wavelength,
radiated_power,utility}.py: the repository that ``cherab.openadas.repository.populate()`` downloads to
``~/.cherab/openadas`` is a tree of JSON files

    wavelength/<symbol>/<charge>.json              {"<upper> -> <lower>": nm}
    pec/excitation/<symbol>/<charge>.json          {"<upper> -> <lower>": {"ne": [...], "te": [...], "rate": [[...]]}}
    pec/recombination/<symbol>/<charge>.json       same shape (photon m^3 s^-1 on ne [m^-3] x te [eV])
    radiated_power/{line,continuum,cx}/<symbol>.json   {"<charge>": {"ne", "te", "rate" [W m^3]}}
    pec/thermal_cx/<donor>/<donor charge>/<receiver>/<receiver charge>.json   {"<upper> -> <lower>": {"ne", "te", "td", "rate"}}
    beam/stopping/<beam>/<target>/<charge>.json         {"e", "n", "t", "sen", "st", "sref", ...}
    beam/population/<beam>/<metastable>/<target>/<charge>.json   same shape, dimensionless
    beam/emission/<beam>/<target>/<charge>.json         {"<upper> -> <lower>": {"e", "n", "t", "sen", "st", "sref"}}
    beam/cx/<donor>/<receiver>/<charge>.json            {"<upper> -> <lower>": {"<metastable>": {"eb", "ti", "ni", "z", "b", "qeb", ..., "qref"}}}

This class only turns those files into the RateTable objects the scene flattener understands; the log-log cubic
interpolation itself happens on the device (cb2_scene_create).  Rates are looked up under the ELEMENT of an isotope (ADAS
has no isotope rates, openadas.py:339-341), wavelengths under the isotope with an optional element fallback (:66-72).
"""
import json
import os

import numpy as np

from .atomic import AtomicData, RateTable, RateTable3D

DEFAULT_REPOSITORY_PATH = os.path.expanduser("~/.cherab/openadas/repository")


def encode_transition(transition):
    """cherab/openadas/repository/utility.py:28-40."""
    upper, lower = transition
    return "{} -> {}".format(str(upper).lower(), str(lower).lower())


class OpenADAS(AtomicData):
    """openadas.py:24-58: ``permit_extrapolation`` selects nearest-neighbour extrapolation of the rate tables,
    ``missing_rates_return_null`` turns a missing rate into None (the model then contributes nothing),
    ``wavelength_element_fallback`` lets an isotope use its element's wavelength."""

    def __init__(self, data_path=None, permit_extrapolation=False, missing_rates_return_null=False, wavelength_element_fallback=False):
        self.data_path = data_path or DEFAULT_REPOSITORY_PATH
        self.permit_extrapolation = permit_extrapolation
        self.missing_rates_return_null = missing_rates_return_null
        self.wavelength_element_fallback = wavelength_element_fallback

    # ---- repository access (repository/*.py) ----
    def _load(self, relative_path, key, what):
        path = os.path.join(self.data_path, relative_path)
        try:
            with open(path, "r") as f:
                return json.load(f)[key]
        except (FileNotFoundError, KeyError):
            raise RuntimeError("Requested %s is not available." % what)

    def _table(self, d):
        return RateTable(np.array(d["ne"], np.float64), np.array(d["te"], np.float64), np.array(d["rate"], np.float64), self.permit_extrapolation)

    def _get_wavelength(self, ion, charge, transition):
        return float(self._load("wavelength/{}/{}.json".format(ion.symbol.lower(), charge), encode_transition(transition),
                                "wavelength (element={}, charge={}, transition={})".format(ion.symbol, charge, transition)))

    # ---- AtomicData interface ----
    def wavelength(self, ion, charge, transition):
        if ion.element is not ion and self.wavelength_element_fallback:
            try:
                return self._get_wavelength(ion, charge, transition)
            except RuntimeError:
                return self._get_wavelength(ion.element, charge, transition)
        return self._get_wavelength(ion, charge, transition)

    def _pec(self, cls, ion, charge, transition):
        element = ion.element                     # no isotope rates in ADAS (openadas.py:339-341)
        try:
            d = self._load("pec/{}/{}/{}.json".format(cls, element.symbol.lower(), charge), encode_transition(transition),
                           "PEC rate (class={}, element={}, charge={}, transition={})".format(cls, element.symbol, charge, transition))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        return self._table(d)

    def impact_excitation_pec(self, ion, charge, transition):
        return self._pec("excitation", ion, charge, transition)

    def recombination_pec(self, ion, charge, transition):
        return self._pec("recombination", ion, charge, transition)

    def _power(self, cls, ion, charge):
        element = ion.element
        try:
            d = self._load("radiated_power/{}/{}.json".format(cls, element.symbol.lower()), str(charge),
                           "radiated power rate (element={}, charge={})".format(element.symbol, charge))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        return self._table(d)

    def line_radiated_power_rate(self, ion, charge):
        return self._power("line", ion, charge)

    def continuum_radiated_power_rate(self, ion, charge):
        return self._power("continuum", ion, charge)

    def cx_radiated_power_rate(self, ion, charge):
        return self._power("cx", ion, charge)

    # ---- thermal charge exchange (repository/pec.py:364-403, openadas.py:389-430) ----
    def thermal_cx_pec(self, donor_ion, donor_charge, receiver_ion, receiver_charge, transition):
        donor, receiver = donor_ion.element, receiver_ion.element
        try:
            d = self._load("pec/thermal_cx/{}/{}/{}/{}.json".format(donor.symbol.lower(), donor_charge, receiver.symbol.lower(), receiver_charge),
                           encode_transition(transition),
                           "thermal charge-exchange PEC (donor={}, donor charge={}, receiver={}, receiver charge={})".format(
                               donor.symbol, donor_charge, receiver.symbol, receiver_charge))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        f = lambda k: np.array(d[k], np.float64)
        return RateTable3D(f("ne"), f("te"), f("td"), f("rate"), self.permit_extrapolation)

    # ---- beam rates (repository/beam/{stopping,population,emission,cx}.py, openadas.py:164-330) ----
    def _load_file(self, relative_path, what):
        path = os.path.join(self.data_path, relative_path)
        try:
            with open(path, "r") as f:
                return json.load(f)
        except FileNotFoundError:
            raise RuntimeError("Requested %s is not available." % what)

    def _beam_table(self, d):
        from .beam import BeamStoppingTable
        f = lambda k: np.array(d[k], np.float64)
        return BeamStoppingTable(f("e"), f("n"), f("t"), f("sen"), f("st"), float(d["sref"]), extrapolate=self.permit_extrapolation)

    def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
        beam, target = beam_ion.element, plasma_ion.element
        try:
            d = self._load_file("beam/stopping/{}/{}/{}.json".format(beam.symbol.lower(), target.symbol.lower(), charge),
                                "beam stopping rate (beam species={}, target ion={}, target charge={})".format(beam.symbol, target.symbol, charge))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        return self._beam_table(d)

    def beam_population_rate(self, beam_ion, metastable, plasma_ion, charge):
        beam, target = beam_ion.element, plasma_ion.element
        try:
            d = self._load_file("beam/population/{}/{}/{}/{}.json".format(beam.symbol.lower(), metastable, target.symbol.lower(), charge),
                                "beam population rate (beam species={}, metastable={}, target ion={}, target charge={})".format(
                                    beam.symbol, metastable, target.symbol, charge))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        return self._beam_table(d)

    def beam_emission_pec(self, beam_ion, plasma_ion, charge, transition):
        beam, target = beam_ion.element, plasma_ion.element
        try:
            d = self._load("beam/emission/{}/{}/{}.json".format(beam.symbol.lower(), target.symbol.lower(), charge), encode_transition(transition),
                           "beam emission rate (beam species={}, target ion={}, target charge={}, transition={})".format(
                               beam.symbol, target.symbol, charge, transition))
        except RuntimeError:
            if self.missing_rates_return_null:
                return None
            raise
        return self._beam_table(d)

    def beam_cx_pec(self, donor_ion, receiver_ion, receiver_charge, transition):
        """One BeamCXTable per donor metastable, in the file's order (repository/beam/cx.py:210-262)."""
        from .beam import BeamCXTable, ConstantBeamCXPEC
        donor, receiver = donor_ion.element, receiver_ion.element
        try:
            rates = self._load("beam/cx/{}/{}/{}.json".format(donor.symbol.lower(), receiver.symbol.lower(), receiver_charge),
                               encode_transition(transition),
                               "beam CX effective emission rates (donor={}, receiver={}, charge={}, transition={})".format(
                                   donor.symbol, receiver.symbol, receiver_charge, transition))
        except RuntimeError:
            if self.missing_rates_return_null:
                return [ConstantBeamCXPEC(1, 0.0)]                # [NullBeamCXPEC()], openadas.py:197-199
            raise
        out = []
        for metastable, d in rates.items():
            f = lambda k: np.array(d[k], np.float64)
            out.append(BeamCXTable(int(metastable), f("eb"), f("ti"), f("ni"), f("z"), f("b"), f("qeb"), f("qti"), f("qni"), f("qz"), f("qb"),
                                   float(d["qref"]), extrapolate=self.permit_extrapolation))
        return out
