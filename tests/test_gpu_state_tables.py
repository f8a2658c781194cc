"""The table-driven state kernel (state tables per mesh triangle and on a psi_n grid) against the generic
kernel and the oracle: the Generomak scenes it takes (plasma lines with Gaussian / multiplet shapes, Bremsstrahlung moments,
constant and flux-mapped velocities), and the scenes it must leave to the generic kernel."""
import numpy as np
import pytest

import core_b200 as cb
from core_b200 import generomak
from core_b200.engine import EmissionScene
from oracle import oracle
from helpers import generomak_camera_rays

pytestmark = pytest.mark.gpu
RTOL, FLOOR = 1e-4, 1e-9      # |gpu - ref| <= 1e-4 |ref| + 1e-9 max_bin |ref[ray]|


def worst_ratio(got, ref):
    tol = RTOL * np.abs(ref) + FLOOR * np.abs(ref).max(axis=1, keepdims=True)
    return float(np.max(np.abs(got - ref) / (tol + 1e-300)))


def render_both_ways(flat, rays, monkeypatch, expect_tables=True):
    monkeypatch.setenv("CB2_STATE_MEMO", "0")
    generic = EmissionScene(flat)
    assert generic.info()["state_table_intervals"] == 0
    a, sa = generic.render(rays)
    generic.close()
    monkeypatch.setenv("CB2_STATE_MEMO", "1")
    tabled = EmissionScene(flat)
    info = tabled.info()
    assert (info["state_table_intervals"] > 0) == expect_tables, info
    if expect_tables:
        assert info["state_table_error_1e9"] <= 4000      # accepted mid-interval error of the psi_n grid: <= 4e-6
    b, sb = tabled.render(rays)
    tabled.close()
    assert sa["samples"] == sb["samples"]
    return a, b, sa, sb


def test_c1_halpha_tables_vs_generic_vs_oracle(monkeypatch):
    plasma = generomak.get_plasma()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 512)
    rays = generomak_camera_rays(plasma, (24, 24))
    a, b, sa, sb = render_both_ways(flat, rays, monkeypatch)
    ref, rst = oracle.emission_render(flat, rays)
    assert sb["samples"] == rst["samples"] and sb["out_of_domain"] == 0
    assert worst_ratio(b, ref) <= 1.0 and worst_ratio(a, ref) <= 1.0
    assert worst_ratio(b, a) <= 0.2                        # the two kernels agree far inside the tolerance


def test_c3_mix_with_moments(monkeypatch):
    plasma = generomak.get_plasma()
    lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4, 5, 6)]
    plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines] + [cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=0.004)
    flat = cb.flatten_scene(plasma, 390.0, 700.0, 2048)
    rays = generomak_camera_rays(plasma, (5, 5))
    a, b, sa, sb = render_both_ways(flat, rays, monkeypatch)
    assert sa["brems_bin_evals"] == sb["brems_bin_evals"]
    ref, _ = oracle.emission_render(flat, rays)
    assert worst_ratio(b, ref) <= 1.0
    assert worst_ratio(b, a) <= 0.2


def test_multiplet_and_constant_velocity(monkeypatch):
    plasma = generomak.get_plasma()
    h0 = plasma.composition.get(cb.hydrogen, 0)
    d = h0.distribution
    plasma.composition.add(cb.Species(cb.hydrogen, 0, cb.Maxwellian(d.density, d.temperature, cb.ConstantVector3D(1.5e4, -2.0e4, 5.0e3), d.atomic_mass)))
    multiplet = [[655.9, 656.1, 656.28, 656.5], [0.2, 0.5, 0.25, 0.05]]
    l3, l4 = cb.Line(cb.hydrogen, 0, (3, 2)), cb.Line(cb.hydrogen, 0, (4, 2))
    plasma.models = [cb.ExcitationLine(l3, lineshape=cb.MultipletLineShape, lineshape_args=[multiplet]), cb.RecombinationLine(l3),
                     cb.ExcitationLine(l4)]
    flat = cb.flatten_scene(plasma, 480.0, 660.0, 1024)
    rays = generomak_camera_rays(plasma, (12, 12))
    a, b, _, _ = render_both_ways(flat, rays, monkeypatch)
    ref, _ = oracle.emission_render(flat, rays)
    assert worst_ratio(b, ref) <= 1.0
    assert worst_ratio(b, a) <= 0.2


def test_ineligible_scenes_keep_the_generic_kernel(monkeypatch):
    # a Zeeman shape needs the field direction against the ray: not a function of the plasma state alone
    plasma = generomak.get_plasma()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line, lineshape=cb.ZeemanTriplet)]
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 256)
    rays = generomak_camera_rays(plasma, (4, 4))
    a, b, _, _ = render_both_ways(flat, rays, monkeypatch, expect_tables=False)
    assert np.array_equal(a, b) or worst_ratio(b, a) <= 1e-3


def test_accumulate_and_f32_frames(monkeypatch):
    monkeypatch.setenv("CB2_STATE_MEMO", "1")
    plasma = generomak.get_plasma()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line), cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=0.003)
    flat = cb.flatten_scene(plasma, 640.0, 670.0, 300)
    rays = generomak_camera_rays(plasma, (9, 9))
    scene = EmissionScene(flat)
    assert scene.info()["state_table_intervals"] > 0
    a, _ = scene.render(rays)
    again, _ = scene.render(rays)
    assert np.array_equal(a, again)                        # fixed summation order: bit-reproducible
    f32, _ = scene.render(rays, dtype=np.float32)
    acc = a.copy()
    scene.render(rays, out=acc, scale=0.5, accumulate=True)
    scene.close()
    assert np.allclose(f32, a, rtol=3e-7, atol=0)
    assert np.allclose(acc, 1.5 * a, rtol=1e-12, atol=0)


@pytest.mark.parametrize("contract", ["tcgen05", "ffma"])
def test_contraction_kernels_against_oracle(monkeypatch, contract):
    """Both contraction kernels of the moment formulation — the tcgen05 3xTF32 kernel (default) and the FFMA tile kernel
    (CB2_CONTRACT=ffma) — on ragged shapes: 144 rays (one 256-row tile, partly empty), 130 / 257 / 2048 bins, float32 and
    float64 frames, accumulate."""
    if contract == "ffma":
        monkeypatch.setenv("CB2_CONTRACT", "ffma")
    else:
        monkeypatch.delenv("CB2_CONTRACT", raising=False)
    plasma = generomak.get_plasma()
    plasma.models = [cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=0.02)
    rays = generomak_camera_rays(plasma, (12, 12))
    for lo, hi, bins in ((500.0, 600.0, 130), (500.0, 600.0, 257), (390.0, 700.0, 2048)):
        flat = cb.flatten_scene(plasma, lo, hi, bins)
        scene = EmissionScene(flat)
        info = scene.info()
        assert info["brems_mode"] == "moments"
        assert info["contraction_on_tensor_cores"] == (1 if contract == "tcgen05" else 0)
        a64, _ = scene.render(rays)
        a32, _ = scene.render(rays, dtype=np.float32)
        twice = a64.copy()
        scene.render(rays, out=twice, scale=0.5, accumulate=True)
        scene.close()
        ref, _ = oracle.emission_render(flat, rays)
        assert worst_ratio(a64, ref) <= 1.0, (contract, bins)
        assert worst_ratio(a32.astype(np.float64), 0.5 * ref + 0.5 * ref) <= 2.0, (contract, bins)
        assert worst_ratio(twice, 1.5 * ref) <= 1.0, (contract, bins)


def test_fused_kernel_matches_the_two_kernel_path(monkeypatch):
    # CB2_FUSED=1: state and binning of the table-driven scene in one kernel (no line records through HBM), blend-zone samples
    # through the fix-up pass and a bin pass over the flagged groups; C3's model mix with the moment contraction
    plasma = generomak.get_plasma()
    lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4)]
    plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines] + [cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=0.003)
    flat = cb.flatten_scene(plasma, 390.0, 700.0, 2048)
    rays = generomak_camera_rays(plasma, (10, 10))
    monkeypatch.setenv("CB2_FUSED", "0")
    two = EmissionScene(flat)
    a, sa = two.render(rays)
    two.close()
    monkeypatch.setenv("CB2_FUSED", "1")
    one = EmissionScene(flat)
    b, sb = one.render(rays)
    acc, _ = one.render(rays, out=b.copy(), accumulate=True)
    one.close()
    for k in ("samples", "gaussian_bin_evals", "brems_bin_evals", "out_of_domain"):
        assert sa[k] == sb[k], k
    assert worst_ratio(b, a) <= 0.05                     # same arithmetic per sample; only the summation order of a ray's groups differs
    assert worst_ratio(acc, 2 * b) <= 0.05
    ref, _ = oracle.emission_render(flat, rays)
    assert worst_ratio(b, ref) <= 1.0
