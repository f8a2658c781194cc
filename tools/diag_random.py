import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_gpu_random as m
from core_b200.engine import EmissionScene
from oracle import oracle
for seed in range(40):
    rng = np.random.default_rng(1000 + seed)
    flat, rays, kind, flat_wide = m._scene(rng)
    sc = EmissionScene(flat); got, st = sc.render(rays); sc.close()
    ref, rst = oracle.emission_render(flat, rays)
    g = flat.desc.grid
    tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
    if True:
        wide, _ = oracle.emission_render(flat_wide, rays)
        gw = flat_wide.desc.grid
        total = wide.sum(axis=1, keepdims=True) * (gw.max_wavelength - gw.min_wavelength) / gw.bins
        tol = tol + 4e-16 * total / ((g.max_wavelength - g.min_wavelength) / g.bins)
    err = np.abs(got - ref)
    mo = flat.desc.models[0]
    print("seed", seed, "kind", kind, "bins", g.bins, "window", g.min_wavelength, g.max_wavelength, "line", mo.wavelength, "step", flat.desc.step, "min_samples", flat.desc.min_samples)
    for r in range(rays.n_rays):
        bad = np.nonzero(err[r] > tol[r])[0]
        if bad.size:
            b = bad[np.argmax((err[r] / (tol[r] + 1e-300))[bad])]
            print("  ray", r, "nbad", bad.size, "worst bin", b, "got", got[r, b], "ref", ref[r, b], "raymax", np.abs(ref[r]).max(), "err/tol", err[r, b] / tol[r, b], "bad bins", bad[:6], bad[-3:])
