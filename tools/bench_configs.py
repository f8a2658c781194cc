"""Device-timed throughput of the BASELINE configurations bench.py does not run (C1, C2, C5; C4 is tools/bench_rt.py, the headline
C3 is bench.py).  One JSON line per configuration: samples/s, the per-kernel split of cb2_scene_profile, and the oracle on a
bounded sample of the same rays as the CPU figure.

    python tools/bench_configs.py [c1] [c2] [c5] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import core_b200 as cb  # noqa: E402
from core_b200 import generomak  # noqa: E402
from core_b200.engine import DeviceRays, EmissionScene  # noqa: E402


def c1():
    """Generomak, H-alpha excitation + recombination, Gaussian line shape, 512 bins on [651.279, 661.279] nm, 128 x 128 pinhole."""
    plasma = generomak.get_plasma()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    flat = cb.flatten_scene(plasma, 651.279, 661.279, 512)
    cam = cb.PinholeCamera((128, 128), fov=45, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    rays = cb.ray_segments(plasma.geometry, *cam.rays(), plasma.geometry_to_world())
    return "C1 Generomak H-alpha 128x128, 512 bins, step 1 mm", flat, rays


def c2():
    """demos/balmer_series.py shape: 5 Balmer lines x {excitation, recombination}, Stark + Doppler + Zeeman, 1024 bins, 64 sight lines."""
    from test_gpu_emission import balmer_series_scene
    plasma, flat = balmer_series_scene()
    xs = np.linspace(-5, 5, 64)
    o = np.stack([xs, np.zeros(64), np.full(64, -5.0)], axis=1)
    rays = cb.ray_segments(plasma.geometry, o, -o / np.linalg.norm(o, axis=1, keepdims=True))
    return "C2 Balmer series (Stark-broadened), 64 sight lines, 1024 bins, step 0.05 m", flat, rays


def c5():
    """demos/beam.py shape on the Generomak plasma: BeamCXLine(C5+ 8->7) + BeamEmissionLine(D 3->2) with ADF12/21/22-shaped tables,
    256 fibres x 64 cone samples crossing the beam axis, 1024 bins on the Balmer-alpha / MSE window."""
    plasma = generomak.get_plasma()
    atomic = cb.SyntheticADAS()
    balmer = atomic.wavelength
    atomic.wavelength = lambda ion, charge, transition: 529.05 if ion is cb.carbon else balmer(ion, charge, transition)
    plasma.atomic_data = atomic
    beam = cb.Beam(transform=cb.look_at((3.2, -0.4, 0.0), (1.0, 0.3, 0.05)))
    beam.atomic_data, beam.plasma = atomic, plasma
    beam.attenuator = cb.SingleRayAttenuator(clamp_to_zero=True)
    beam.energy, beam.power, beam.temperature, beam.element = 60000, 3e6, 10, cb.deuterium
    beam.sigma, beam.divergence_x, beam.divergence_y, beam.length = 0.025, 0.5, 0.5, 3.0
    beam.integrator = cb.NumericalIntegrator(step=0.0025, min_samples=10)
    beam.models = [cb.BeamEmissionLine(cb.Line(cb.deuterium, 0, (3, 2))), cb.BeamCXLine(cb.Line(cb.carbon, 5, (8, 7)))]
    flat = cb.flatten_beam_scene(beam, 650.0, 662.0, 1024)
    group = cb.FibreOpticGroup()
    b2w = np.asarray(beam.transform)
    zs, ys = np.meshgrid(np.linspace(0.6, 2.4, 16), np.linspace(-0.03, 0.03, 16), indexing="ij")
    for z, y in zip(zs.ravel(), ys.ravel()):
        target = (b2w @ np.array([0.0, y, z, 1.0]))[:3]
        group.add_observer(cb.FibreOptic(transform=cb.look_at((1.8, 0.2, 1.6), tuple(target)), acceptance_angle=0.5, radius=0.001, pixel_samples=64))
    o, d, _, _ = group.gather_rays()
    rays = cb.beam_ray_segments(beam, o, d)
    return "C5 beam CX + beam emission (MSE), 256 fibres x 64 cone samples, 1024 bins, step 2.5 mm", flat, rays


def run(name, builder, reps, cpu_rays):
    from oracle import oracle
    label, flat, rays = builder()
    scene = EmissionScene(flat)
    dev = torch.device("cuda", 0)
    dr = DeviceRays(rays, device=dev)
    out = torch.zeros((rays.n_rays, scene.bins), dtype=torch.float32, device=dev)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(3):
        scene.render_device(dr, out, stats=stats)
    torch.cuda.synchronize()
    stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        scene.render_device(dr, out, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = stats.cpu().numpy() // reps
    scene.profile(True)
    scene.render_device(dr, out)
    torch.cuda.synchronize()
    kernels = {k: {"ms": v[0], "launches": v[1]} for k, v in scene.profile(False).items()}
    # CPU figure: the oracle on a bounded, seeded sample of the same rays, all host threads
    pick = np.sort(np.random.default_rng(3).choice(rays.n_rays, size=min(cpu_rays, rays.n_rays), replace=False))
    sub = rays.subset(pick)
    t0 = time.perf_counter()
    ref, rst = oracle.emission_render(flat, sub)
    dt = time.perf_counter() - t0
    got = out[torch.as_tensor(pick, device=dev)].double().cpu().numpy()
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True) + 6e-8 * np.abs(ref)
    line = {"config": label, "rays": int(rays.n_rays), "samples": int(st[0]), "ms_per_frame": ms, "msamples_per_s": st[0] / ms * 1e-3,
            "gaussian_bin_evals": int(st[1]), "lorentzian_bin_evals": int(st[2]), "kernels": kernels, "plan": scene.info(),
            "cpu_oracle": {"rays": int(pick.size), "samples": int(rst["samples"]), "seconds": dt, "msamples_per_s": rst["samples"] / dt * 1e-6,
                           "threads": os.cpu_count()},
            "parity_on_sample": bool(np.all(np.abs(got - ref) <= tol))}
    scene.close()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c5"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-rays", type=int, default=8)
    a = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    for name in a.configs:
        run(name, {"c1": c1, "c2": c2, "c5": c5}[name], a.reps, a.cpu_rays)


if __name__ == "__main__":
    main()
