"""OpenADAS atomic-data provider reading Cherab's on-disk JSON repository.

Restates the read side of cherab/openadas/openadas.py:24-100,318-395 and cherab/openadas/repository/{pec,wavelength,
radiated_power,utility}.py: the repository that ``cherab.openadas.repository.populate()`` downloads to
``~/.cherab/openadas`` is a tree of JSON files

    wavelength/<symbol>/<charge>.json              {"<upper> -> <lower>": nm}
    pec/excitation/<symbol>/<charge>.json          {"<upper> -> <lower>": {"ne": [...], "te": [...], "rate": [[...]]}}
    pec/recombination/<symbol>/<charge>.json       same shape (photon m^3 s^-1 on ne [m^-3] x te [eV])
    radiated_power/{line,continuum,cx}/<symbol>.json   {"<charge>": {"ne", "te", "rate" [W m^3]}}

This class only turns those files into the RateTable objects the scene flattener understands; the log-log cubic
interpolation itself happens on the device (cb2_scene_create).  Rates are looked up under the ELEMENT of an isotope (ADAS
has no isotope rates, openadas.py:339-341), wavelengths under the isotope with an optional element fallback (:66-72).
"""
import json
import os

import numpy as np

from .atomic import AtomicData, RateTable

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
