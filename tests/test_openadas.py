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


def test_beam_and_thermal_cx_repository_files(tmp_path):
    """The beam / thermal-CX part of the repository layout (cherab/openadas/repository/beam/*.py, pec.py:364-403) feeds the
    beam models and ThermalCXLine; a whole beam scene flattens from files on disk."""
    import json
    from core_b200.beam import BeamCXTable, BeamStoppingTable
    synth = cb.SyntheticADAS()

    def dump(rel, obj):
        path = tmp_path / rel
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_text(json.dumps(obj))

    def beam_dict(t):
        return {"e": t.e.tolist(), "n": t.n.tolist(), "t": t.t.tolist(), "sen": t.sen.tolist(), "st": t.st.tolist(), "sref": t.sref, "eref": 4e4, "nref": 1e19, "tref": 1e3}

    st = synth.beam_stopping_rate(cb.deuterium, cb.deuterium, 1)
    dump("beam/stopping/h/h/1.json", beam_dict(st))                                  # isotope -> element (openadas.py:227-232)
    dump("beam/stopping/h/c/6.json", beam_dict(synth.beam_stopping_rate(cb.deuterium, cb.carbon, 6)))
    dump("beam/population/h/2/h/1.json", beam_dict(st))
    dump("beam/population/h/2/c/6.json", beam_dict(st))
    em = synth.beam_emission_pec(cb.deuterium, cb.deuterium, 1, (3, 2))
    dump("beam/emission/h/h/1.json", {"3 -> 2": beam_dict(em)})
    cx = synth.beam_cx_pec(cb.deuterium, cb.carbon, 6, (8, 7))[0]
    cxd = {k: getattr(cx, k).tolist() for k in ("eb", "ti", "ni", "z", "b", "qeb", "qti", "qni", "qz", "qb")}
    cxd["qref"] = cx.qref
    dump("beam/cx/h/c/6.json", {"8 -> 7": {"1": cxd, "2": cxd}})
    tc = synth.thermal_cx_pec(cb.deuterium, 0, cb.carbon, 6, (8, 7))
    dump("pec/thermal_cx/h/0/c/6.json", {"8 -> 7": {"ne": tc.ne.tolist(), "te": tc.te.tolist(), "td": tc.td.tolist(), "rate": tc.rate.tolist()}})
    dump("wavelength/c/5.json", {"8 -> 7": 529.05})

    adas = cb.OpenADAS(data_path=str(tmp_path), permit_extrapolation=True)
    got = adas.beam_stopping_rate(cb.deuterium, cb.deuterium, 1)
    assert isinstance(got, BeamStoppingTable) and np.array_equal(got.sen, st.sen) and got.sref == st.sref
    assert np.array_equal(adas.beam_population_rate(cb.deuterium, 2, cb.deuterium, 1).st, st.st)
    assert np.array_equal(adas.beam_emission_pec(cb.deuterium, cb.deuterium, 1, (3, 2)).sen, em.sen)
    rates = adas.beam_cx_pec(cb.deuterium, cb.carbon, 6, (8, 7))
    assert [r.donor_metastable for r in rates] == [1, 2] and all(isinstance(r, BeamCXTable) for r in rates)
    assert np.array_equal(rates[0].qeb, cx.qeb)
    t3 = adas.thermal_cx_pec(cb.deuterium, 0, cb.carbon, 6, (8, 7))
    assert isinstance(t3, cb.RateTable3D) and t3.rate.shape == tc.rate.shape and t3.extrapolate
    with pytest.raises(RuntimeError):
        adas.beam_stopping_rate(cb.deuterium, cb.neon, 10)
    null = cb.OpenADAS(data_path=str(tmp_path), missing_rates_return_null=True)
    assert null.beam_stopping_rate(cb.deuterium, cb.neon, 10) is None and null.thermal_cx_pec(cb.deuterium, 0, cb.neon, 10, (6, 5)) is None
    assert null.beam_cx_pec(cb.deuterium, cb.neon, 10, (6, 5))[0].value == 0.0

    # a beam scene straight from the files: slab plasma of D+ and C6+, beam CX line of C5+ 8 -> 7
    from core_b200.slab import build_constant_slab_plasma
    plasma = build_constant_slab_plasma(length=1, width=1, height=1, electron_density=1e19, electron_temperature=1e3,
                                        plasma_species=[(cb.deuterium, 1, 1e19, 1e3, (0, 0, 0)), (cb.carbon, 6, 1e17, 1e3, (0, 0, 0))], b_field=(0, 2.0, 0))
    plasma.atomic_data = adas
    beam = cb.Beam(transform=cb.translate(0.5, 0, 0))
    beam.atomic_data, beam.plasma = adas, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 50000, 1e6, 10, cb.deuterium
    beam.models = [cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    flat = cb.flatten_beam_scene(beam, 526.0, 532.0, 64)
    assert flat.desc.models[0].ext.contents.n_cx == 2
    from oracle import oracle
    ref, _ = oracle.emission_render(flat, cb.beam_ray_segments(beam, [[1.5, 0, 0.5]], [[-1.0, 0, 0]]))
    assert ref.max() > 0 and np.all(np.isfinite(ref))
