"""Where bin_kernel's per-component set-up goes (CPU prototype, oracle state on a sample of the C3 rays): for every (sample, line) the
Gaussian width in bins decides the evaluation form and how many 32 / 16 / 8-bin windows a component pass walks.  Prints the mix of
(live sample, component) pairs by width class and the evaluated bins per pair — the numbers behind DESIGN.md section 8 item 1.

    python tools/proto_bin_pass_mix.py [n_rays]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                        # noqa: E402
from oracle import oracle                           # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 96
oracle.build()
plasma, flat = bench.build_scene(2048)
pick = np.sort(np.random.default_rng(3).choice(1024 * 1024, size=n_rays, replace=False))
rays = bench.make_rays(plasma, 1024, pick, 0)
d = flat.desc
pts = []
for r in range(rays.n_rays):
    for s in range(rays.seg_offset[r], rays.seg_offset[r + 1]):
        L = rays.seg_t1[s] - rays.seg_t0[s]
        iv = max(d.min_samples - 1, int(np.ceil(L / d.step)))
        t = rays.seg_t0[s] + L * np.arange(iv + 1) / iv
        pts.append(rays.origin[r] + t[:, None] * rays.direction[r])
pts = np.concatenate(pts)
state = oracle.sample_state(flat, pts)              # [n, 2 + 5 species + 3]: ne, te, then (n, T, vx, vy, vz) per species
ne, te = state[:, 0], state[:, 1]
live = (ne > 0) & (te > 0)
delta = (d.grid.max_wavelength - d.grid.min_wavelength) / d.grid.bins
print("%d rays, %d samples, %.1f %% live" % (n_rays, pts.shape[0], 100 * live.mean()))
rows = []
for m in range(d.n_models):
    M = d.models[m]
    if M.kind == 2:                                 # Bremsstrahlung
        continue
    sp = M.species
    n_s, t_s = state[:, 2 + 5 * sp], state[:, 3 + 5 * sp]
    on = live & (n_s > 0) & (t_s > 0)
    sigma = np.sqrt(t_s[on] * 1.602176634e-19 / (d.species[sp].atomic_weight * 1.66053906660e-27)) * M.wavelength / 299792458.0 / delta
    rows.append((M.wavelength, M.kind, on.mean(), sigma))
allsig = np.concatenate([r[3] for r in rows])
print("(live sample, line) pairs: %d; width in bins: median %.2f, mean %.2f" % (allsig.size, np.median(allsig), allsig.mean()))
edges = [0, 0.98, 2.0, 4.0, 8.0, 1e9]
names = ["< 0.98 (erfc differences, 10 sigma)", "0.98 - 2 (series: one 32-bin window or a tail)", "2 - 4 (one window + a tail)",
         "4 - 8 (2 - 4 windows)", ">= 8 (4+ windows)"]
for a, b, nm in zip(edges[:-1], edges[1:], names):
    sel = (allsig >= a) & (allsig < b)
    span = np.where(allsig[sel] < 0.98, 20.0, 14.0) * allsig[sel] + 1
    print("  sigma %-52s %5.1f %% of the pairs, %5.1f %% of the evaluated bins, %5.1f bins per pair" % (
        nm, 100 * sel.mean(), 100 * span.sum() / (np.where(allsig < 0.98, 20.0, 14.0) * allsig + 1).sum(), span.mean() if sel.any() else 0))
for wl, kind, frac, sigma in rows:
    print("  line %.1f nm %-13s live in %4.1f %% of the samples, sigma median %.2f bins (10th / 90th percentile %.2f / %.2f)" % (
        wl, "excitation" if kind == 0 else "recombination", 100 * frac, np.median(sigma), np.percentile(sigma, 10), np.percentile(sigma, 90)))
