"""Diagnostic: C1 full frame, state tables on / off against the oracle; lists the rays outside the tolerance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import core_b200 as cb
from core_b200 import generomak
from core_b200.engine import EmissionScene
from oracle import oracle
from helpers import generomak_camera_rays

plasma = generomak.get_plasma()
line = cb.Line(cb.hydrogen, 0, (3, 2))
plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
flat = cb.flatten_scene(plasma, 651.279, 661.279, 512)
rays = generomak_camera_rays(plasma, (128, 128))
ref, rst = oracle.emission_render(flat, rays)
tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
res = {}
for memo in ("0", "1"):
    os.environ["CB2_STATE_MEMO"] = memo
    sc = EmissionScene(flat)
    got, st = sc.render(rays)
    sc.close()
    ratio = np.abs(got - ref) / (tol + 1e-300)
    per_ray = ratio.max(axis=1)
    bad = np.nonzero(per_ray > 1.0)[0]
    print("memo=%s worst %.3g, rays outside tolerance: %d %s" % (memo, per_ray.max(), bad.size, bad[:20]))
    for r in bad[:6]:
        b = int(ratio[r].argmax())
        print("   ray %d bin %d got %.6e ref %.6e rel %.3g  rowmax %.3e  sum got/ref %.8f" % (r, b, got[r, b], ref[r, b], got[r, b] / ref[r, b] - 1, ref[r].max(), got[r].sum() / ref[r].sum()))
    res[memo] = got
d = np.abs(res["1"] - res["0"]) / (tol + 1e-300)
print("memo 1 vs 0: worst %.3g at ray %d" % (d.max(), d.max(axis=1).argmax()))
bad = np.nonzero((np.abs(res["1"] - ref) / (tol + 1e-300)).max(axis=1) > 1.0)[0]
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_c1.npz"), bad=bad, got1=res["1"][bad], got0=res["0"][bad], ref=ref[bad],
                    origin=rays.origin[bad], direction=rays.direction[bad], seg_offset=rays.seg_offset, t0=rays.seg_t0, t1=rays.seg_t1)
