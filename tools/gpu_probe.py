"""Ad-hoc GPU probe: timing of scene variants + parity diagnostics dumped to gpurun_out/ (scratch)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import core_b200 as cb
from core_b200 import generomak
from core_b200.engine import EmissionScene, DeviceRays
from oracle import oracle
from helpers import generomak_camera_rays

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
what = sys.argv[1:] or ["time", "diag"]

def scene_variant(kind, bins, lo, hi, step=1e-3):
    plasma = generomak.get_plasma()
    lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4, 5, 6)]
    if kind == "c1":
        plasma.models = [cb.ExcitationLine(lines[0]), cb.RecombinationLine(lines[0])]
    elif kind == "lines8":
        plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines]
    elif kind == "brems":
        plasma.models = [cb.Bremsstrahlung()]
    else:
        plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines] + [cb.Bremsstrahlung()]
    plasma.integrator = cb.NumericalIntegrator(step=step)
    return plasma, cb.flatten_scene(plasma, lo, hi, bins)

def timeit(kind, bins, lo, hi, pixels=128, reps=3):
    plasma, flat = scene_variant(kind, bins, lo, hi)
    rays = generomak_camera_rays(plasma, (pixels, pixels))
    sc = EmissionScene(flat)
    dr = DeviceRays(rays)
    out = torch.zeros((rays.n_rays, bins), dtype=torch.float32, device="cuda:0")
    stats = torch.zeros(6, dtype=torch.int64, device="cuda:0")
    sc.render_device(dr, out, stats=stats); torch.cuda.synchronize(); stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sc.render_device(dr, out, stats=stats)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = stats.cpu().numpy() // reps
    print("%-8s bins=%4d rays=%6d samples=%.3e gauss/sample=%.1f brems/sample=%.1f  %.2f ms  %.1f Msamples/s" % (
        kind, bins, rays.n_rays, st[0], st[1] / st[0], st[3] / st[0], ms, st[0] / ms * 1e-3), flush=True)
    sc.close()

if "time" in what:
    timeit("c1", 512, 651.279, 661.279)
    timeit("lines8", 2048, 390., 700.)
    timeit("brems", 2048, 390., 700.)
    timeit("c3", 2048, 390., 700.)
    timeit("brems", 512, 651.279, 661.279)

if "diag" in what:
    plasma, flat = scene_variant("c1", 512, 651.279, 661.279)
    rng = np.random.default_rng(7)
    n = 20000
    r = rng.uniform(0.74, 2.40, n); phi = rng.uniform(-np.pi, np.pi, n); z = rng.uniform(-1.79, 1.54, n)
    pts = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    sc = EmissionScene(flat)
    got = sc.sample_state(pts); ref = oracle.sample_state(flat, pts)
    rays = generomak_camera_rays(plasma, (24, 24))
    g2, st = sc.render(rays); r2, rst = oracle.emission_render(flat, rays)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag.npz"), pts=pts, got=got, ref=ref, g2=g2, r2=r2,
                        seg_offset=rays.seg_offset, t0=rays.seg_t0, t1=rays.seg_t1, origin=rays.origin, direction=rays.direction)
    print("diag saved", st, rst)

def one_launch(kind, bins, lo, hi, pixels):
    plasma, flat = scene_variant(kind, bins, lo, hi)
    rays = generomak_camera_rays(plasma, (pixels, pixels))
    sc = EmissionScene(flat)
    dr = DeviceRays(rays)
    out = torch.zeros((rays.n_rays, bins), dtype=torch.float32, device="cuda:0")
    for _ in range(2):
        sc.render_device(dr, out)
        torch.cuda.synchronize()
    sc.close()

if "prof_c1" in what:
    one_launch("c1", 512, 651.279, 661.279, 96)
if "prof_c3" in what:
    one_launch("c3", 2048, 390., 700., 48)
if "prof_brems" in what:
    one_launch("brems", 2048, 390., 700., 48)
if "prof_lines8" in what:
    one_launch("lines8", 2048, 390., 700., 48)
if "tune" in what:
    for nw, bpl in ((4, 4), (2, 8), (1, 16)):
        os.environ["CB2_NW"], os.environ["CB2_BPL"] = str(nw), str(bpl)
        print("NW=%d BPL=%d" % (nw, bpl)); timeit("c1", 512, 651.279, 661.279)
    for nw, bpl in ((8, 8), (4, 16), (2, 16)):
        os.environ["CB2_NW"], os.environ["CB2_BPL"] = str(nw), str(bpl)
        print("NW=%d BPL=%d" % (nw, bpl)); timeit("lines8", 2048, 390., 700.); timeit("brems", 2048, 390., 700.); timeit("c3", 2048, 390., 700.)
