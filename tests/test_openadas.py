"""OpenADAS repository reader: a repository written in the reference's on-disk JSON layout
(cherab/openadas/repository/pec.py:344-361, wavelength.py:108-128, radiated_power.py:254-290) must feed the path exactly
like the in-memory tables it was written from."""
import json
import os

import numpy as np
import pytest

import core_b200 as cb
from core_b200 import generomak
from core_b200.openadas import encode_transition
from oracle import oracle
from helpers import generomak_camera_rays


def write_repository(root):
    syn = cb.SyntheticADAS()
    for cls, fn in (("excitation", syn.impact_excitation_pec), ("recombination", syn.recombination_pec)):
        content = {}
        for n in (3, 4):
            t = fn(cb.hydrogen, 0, (n, 2))
            content[encode_transition((n, 2))] = {"ne": t.ne.tolist(), "te": t.te.tolist(), "rate": t.rate.tolist()}
        path = os.path.join(root, "pec", cls, "h", "0.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(content, open(path, "w"))
    path = os.path.join(root, "wavelength", "h", "0.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump({encode_transition((n, 2)): syn.wavelength(cb.hydrogen, 0, (n, 2)) for n in (3, 4)}, open(path, "w"))
    ne, te = np.logspace(17, 21, 9), np.logspace(0, 4, 11)
    for cls, scale in (("line", 1e-32), ("continuum", 1e-33), ("cx", 1e-31)):
        path = os.path.join(root, "radiated_power", cls, "c.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        rate = scale * (1 + 0.2 * np.log10(ne[:, None] / 1e19)) * (te[None, :] / 100.0) ** 0.3
        json.dump({"2": {"ne": ne.tolist(), "te": te.tolist(), "rate": rate.tolist()},
                   "3": {"ne": ne.tolist(), "te": te.tolist(), "rate": (2 * rate).tolist()}}, open(path, "w"))


def scenes(tmp_path):
    write_repository(str(tmp_path))
    out = []
    for atomic in (cb.OpenADAS(data_path=str(tmp_path), permit_extrapolation=True), cb.SyntheticADAS()):
        plasma = generomak.get_plasma()
        plasma.atomic_data = atomic
        lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4)]
        plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines]
        plasma.integrator = cb.NumericalIntegrator(step=0.01)
        out.append((plasma, cb.flatten_scene(plasma, 480.0, 660.0, 600)))
    return out


def test_repository_feeds_the_same_scene_as_in_memory_tables(tmp_path):
    (plasma, flat_repo), (_, flat_mem) = scenes(tmp_path)
    rays = generomak_camera_rays(plasma, (2, 2))
    a, _ = oracle.emission_render(flat_repo, rays)
    b, _ = oracle.emission_render(flat_mem, rays)
    assert a.max() > 0 and np.array_equal(a, b)


def test_missing_rates_and_fallbacks(tmp_path):
    write_repository(str(tmp_path))
    strict = cb.OpenADAS(data_path=str(tmp_path))
    with pytest.raises(RuntimeError):
        strict.impact_excitation_pec(cb.hydrogen, 0, (7, 2))
    with pytest.raises(RuntimeError):
        strict.wavelength(cb.deuterium, 0, (3, 2))              # no isotope file and no fallback
    lenient = cb.OpenADAS(data_path=str(tmp_path), missing_rates_return_null=True, wavelength_element_fallback=True)
    assert lenient.impact_excitation_pec(cb.hydrogen, 0, (7, 2)) is None
    assert lenient.wavelength(cb.deuterium, 0, (3, 2)) == strict.wavelength(cb.hydrogen, 0, (3, 2))
    # isotope rates come from the element (openadas.py:339-341)
    assert np.array_equal(lenient.impact_excitation_pec(cb.deuterium, 0, (3, 2)).rate, strict.impact_excitation_pec(cb.hydrogen, 0, (3, 2)).rate)
    assert strict.line_radiated_power_rate(cb.carbon, 2).rate.shape == (9, 11)
    assert lenient.cx_radiated_power_rate(cb.carbon, 5) is None


@pytest.mark.gpu
def test_repository_scene_on_the_device(tmp_path):
    from core_b200.engine import EmissionScene
    write_repository(str(tmp_path))
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.OpenADAS(data_path=str(tmp_path), permit_extrapolation=True)
    plasma.models = [cb.ExcitationLine(cb.Line(cb.hydrogen, 0, (3, 2))), cb.RecombinationLine(cb.Line(cb.hydrogen, 0, (4, 2))),
                     cb.TotalRadiatedPower(cb.carbon, 2)]
    plasma.integrator = cb.NumericalIntegrator(step=0.01)
    flat = cb.flatten_scene(plasma, 480.0, 660.0, 600)
    rays = generomak_camera_rays(plasma, (3, 3))
    scene = EmissionScene(flat)
    got, st = scene.render(rays)
    scene.close()
    ref, rst = oracle.emission_render(flat, rays)
    assert st["samples"] == rst["samples"]
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(got - ref) <= tol)
