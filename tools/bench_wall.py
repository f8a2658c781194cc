"""First-wall occlusion measurement: cb2_wall_clip_device on the rays of the C3 camera (1024 x 1024 pinhole at the C1 pose) against the
Generomak wall (2.88e6 triangles behind the BVH), device-timed; the brute-force oracle on a bounded sample beside it.

    python tools/bench_wall.py [pixels] [reps]      -> one JSON line
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                        # noqa: E402
import core_b200 as cb                              # noqa: E402
from core_b200 import generomak                     # noqa: E402

px = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
plasma = generomak.get_plasma()
cam = cb.PinholeCamera((px, px), fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
pin = cb.DevicePinhole(cam, plasma.geometry, to_world=plasma.geometry_to_world())
t0 = time.perf_counter()
wall = cb.FirstWall()
create_s = time.perf_counter() - t0
rays = pin.rays()
hit = torch.empty(pin.n_rays, dtype=torch.float64, device=pin.device)
wall.clip_device(rays, hit_out=hit)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    wall.clip_device(rays, hit_out=hit)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t = hit.cpu().numpy()
from oracle import oracle                           # noqa: E402  (the CPU figure beside it)
sel = np.sort(np.random.default_rng(1).choice(pin.n_rays, size=64, replace=False))
host = rays.to_host()
t0 = time.perf_counter()
ref = oracle.wall_hit(wall.triangles, host.origin[sel], host.direction[sel])
cpu_s = time.perf_counter() - t0
ok = bool(np.array_equal(np.isfinite(ref), np.isfinite(t[sel])) and np.allclose(t[sel][np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-12))
print(json.dumps({"config": "first-wall clip, %dx%d pinhole rays, Generomak wall" % (px, px), "rays": pin.n_rays, "triangles": wall.n_triangles,
                  "device_ms": ms, "Mrays_per_s": pin.n_rays / ms * 1e-3, "hit_fraction": float(np.isfinite(t).mean()),
                  "bvh_build_s": create_s, "oracle_64_rays_s": cpu_s, "oracle_Mrays_per_s": 64 / cpu_s * 1e-6, "oracle_agrees": ok}))
