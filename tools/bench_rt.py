"""Ray-transfer (geometry-matrix) measurement — BASELINE config C4:
RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=400, n_height=800, radius_inner=0.73) at z = -1.8, default step
0.1*min(dr, dz), 512x512 pinhole rays of the C1 pose, CSR output with the identity voxel map (320 000 sources).

    python tools/bench_rt.py [pixels] [reps]      -> one JSON line (midpoint steps/s device-timed, nnz, bytes)
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
from core_b200.engine import DeviceRays, RayTransferScene   # noqa: E402
from core_b200.raytransfer import RayTransferCylinder       # noqa: E402

px = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=400, n_height=800, radius_inner=0.73, transform=cb.translate(0, 0, -1.8))
cam = cb.PinholeCamera((px, px), fov=45, transform=cb.look_at((2.3, 0, 1.25), (1.0, 0.8, -0.5)))
o, d = cam.rays()
rays = cb.ray_segments(rtc.primitive, o, d, rtc.transform)
scene = RayTransferScene(rtc)
dr = DeviceRays(rays)
# steps per ray: n = max(min_samples, int(L / step))  (emitters.pyx:112-116)
length = rays.seg_t1 - rays.seg_t0
steps = int(np.maximum(2, (length[length > 0] / rtc.step).astype(np.int64)).sum())
cap = int(4000 * rays.n_rays)
ro, cols, lens = scene.render_csr_device(dr, cap)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ro, cols, lens = scene.render_csr_device(dr, cap)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t0 = time.perf_counter()
r2, c2, l2, st = scene.render_csr(rays, capacity=cap)
host_ms = (time.perf_counter() - t0) * 1e3
nnz = int(cols.numel())
print(json.dumps({"config": "C4 ray transfer %dx%d rays, cylinder 400x1x800, step %.3e m, CSR" % (px, px, rtc.step),
                  "rays": rays.n_rays, "steps": steps, "rt_steps_counter": st["rt_steps"], "nnz": nnz,
                  "device_ms": ms, "Gsteps_per_s": steps / ms * 1e-6, "host_buffer_ms": host_ms,
                  "csr_bytes": nnz * 12 + (rays.n_rays + 1) * 8,
                  "algorithmic": {"flop_fp64_per_step": 22, "fp64_tflops": steps * 22 / ms * 1e-9}}))
