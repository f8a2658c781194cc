"""The reference arm through the REAL Cherab + Raysect API, for a box where both import (CPU baseline (ii)).

Neither package is installed in the build container or on the GPU boxes of this round (raysect 0.8.1 is not vendored and
there is no network), so this script is **untested there**: `python tools/run_reference.py --probe` prints whether the reference
can run, and `bench.py --impl reference` keeps timing the oracle port when it cannot.

What it does when they import: builds the benchmark scene with the reference's own objects — `cherab.generomak.plasma.get_plasma()`,
`ExcitationLine` / `RecombinationLine` / `Bremsstrahlung` from `cherab.core.model`, an `AtomicData` subclass that wraps the same
synthetic ADF15-shaped tables (core_b200.SyntheticADAS) in the reference's `ImpactExcitationPEC` / `RecombinationPEC` — and traces
the listed pixels' rays with `Ray(origin, direction, min_wavelength, max_wavelength, bins).trace(world)` exactly as
`cherab/core/tests/test_line_emission.py:116-118` does, one process per core.  Output: one JSON line (samples/s) and, with
`--save`, an `.npz` of the spectra for a parity check against the CUDA path.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def available():
    try:
        import raysect.optical  # noqa: F401
        import cherab.core  # noqa: F401
        import cherab.generomak.plasma  # noqa: F401
        return None
    except Exception as exc:                      # ImportError, or a binary built against another numpy
        return "%s: %s" % (type(exc).__name__, exc)


def _world(bins, lo, hi):
    """World with the Generomak plasma and the benchmark's models, all reference objects."""
    from raysect.optical import World
    from cherab.core.atomic import AtomicData, Line, hydrogen
    from cherab.core.model import Bremsstrahlung, ExcitationLine, RecombinationLine
    from cherab.generomak.plasma import get_plasma
    from cherab.openadas.rates.pec import ImpactExcitationPEC, RecombinationPEC
    import core_b200 as cb

    synth = cb.SyntheticADAS(permit_extrapolation=True)

    class SyntheticAtomicData(AtomicData):
        """The same synthetic tables the CUDA path and the oracle see, as the reference's own rate objects."""

        def wavelength(self, ion, charge, transition):
            return synth.wavelength(getattr(cb, ion.name.lower(), cb.hydrogen), charge, transition)

        def _data(self, table):
            # ADF15 units of the reference's reader: the rate objects take photon m^3/s on ne [m^-3] x te [eV] (pec.pyx:48-68)
            return {"ne": table.ne, "te": table.te, "rate": table.rate}

        def impact_excitation_pec(self, ion, charge, transition):
            t = synth.impact_excitation_pec(cb.hydrogen, charge, transition)
            return ImpactExcitationPEC(self.wavelength(ion, charge, transition), self._data(t), extrapolate=True)

        def recombination_pec(self, ion, charge, transition):
            t = synth.recombination_pec(cb.hydrogen, charge, transition)
            return RecombinationPEC(self.wavelength(ion, charge, transition), self._data(t), extrapolate=True)

    world = World()
    plasma = get_plasma(atomic_data=SyntheticAtomicData(), parent=world)
    lines = [Line(hydrogen, 0, (n, 2)) for n in (3, 4, 5, 6)]
    plasma.models = [ExcitationLine(l) for l in lines] + [RecombinationLine(l) for l in lines] + [Bremsstrahlung()]
    return world, plasma


def _trace(args):
    """Worker: spectra of a slice of the ray list."""
    origins, directions, lo, hi, bins = args
    from raysect.core import Point3D, Vector3D
    from raysect.optical import Ray
    world, _ = _world(bins, lo, hi)
    out = np.zeros((len(origins), bins))
    for i, (o, d) in enumerate(zip(origins, directions)):
        ray = Ray(origin=Point3D(*o), direction=Vector3D(*d), min_wavelength=lo, max_wavelength=hi, bins=bins, extinction_prob=0.0)
        out[i] = ray.trace(world).samples
    return out


def run(pixels, bins, pixel_index, sample_id, lo=390.0, hi=700.0, processes=None):
    """Spectra [len(pixel_index), bins] of the listed pixels' rays and the wall time, through the reference."""
    import multiprocessing as mp
    import core_b200 as cb
    import bench
    sx, sy = cb.stratified_offsets(4)[sample_id % 16]
    cam = cb.PinholeCamera((pixels, pixels), fov=45, transform=cb.look_at(bench.CAMERA_POS, bench.CAMERA_TARGET))
    o, d = cam.rays(sx, sy, pixel_index)
    processes = processes or os.cpu_count()
    chunks = [(o[k::processes], d[k::processes], lo, hi, bins) for k in range(processes) if len(o[k::processes])]
    t0 = time.perf_counter()
    with mp.Pool(len(chunks)) as pool:
        parts = pool.map(_trace, chunks)
    dt = time.perf_counter() - t0
    out = np.zeros((len(o), bins))
    for k, part in enumerate(parts):
        out[k::processes] = part
    return out, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--pixels", type=int, default=1024)
    ap.add_argument("--bins", type=int, default=2048)
    ap.add_argument("--rays", type=int, default=8)
    ap.add_argument("--save", default=None)
    a = ap.parse_args()
    why = available()
    if why is not None or a.probe:
        print(json.dumps({"impl": "reference", "available": why is None, "unavailable": why}))
        return
    pick = np.sort(np.random.default_rng(1234).choice(a.pixels * a.pixels, size=a.rays, replace=False))
    spectra, dt = run(a.pixels, a.bins, pick, 0)
    # samples = intervals + 1 per chord, counted the way the integrator does (the reference does not report it)
    import bench
    import core_b200 as cb
    plasma, _ = bench.build_scene(a.bins)
    rays = bench.make_rays(plasma, a.pixels, pick, 0)
    length = rays.seg_t1 - rays.seg_t0
    samples = int((np.maximum(4, np.ceil(length / 1e-3)) + 1).sum())
    print(json.dumps({"impl": "reference", "kind": "reference", "rays": int(a.rays), "samples": samples, "seconds": dt,
                      "msamples_per_s": samples / dt * 1e-6, "cores": os.cpu_count()}))
    if a.save:
        np.savez_compressed(a.save, pixel_index=pick, spectra=spectra)


if __name__ == "__main__":
    main()
