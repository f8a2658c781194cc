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
t0 = time.perf_counter()
r3, c3, l3, _ = scene.render_csr(rays, capacity=cap)
host_ms_warm = (time.perf_counter() - t0) * 1e3
from core_b200.engine import measure_peaks           # noqa: E402
peaks = measure_peaks()
peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
F_STEP = 22.0                                        # float64 flop per midpoint step + 1 sqrt
fp64 = steps * F_STEP / ms * 1e-9
csr_bytes = nnz * 12 + (rays.n_rays + 1) * 8
hbm_peak = float(peaks_file.get("hbm_gbs", 6500.0))
print(json.dumps({"config": "C4 ray transfer %dx%d rays, cylinder 400x1x800, step %.3e m, CSR" % (px, px, rtc.step),
                  "rays": rays.n_rays, "steps": steps, "rt_steps_counter": st["rt_steps"], "nnz": nnz,
                  "device_ms": ms, "Gsteps_per_s": steps / ms * 1e-6,
                  "host_buffer_ms": host_ms, "host_buffer_ms_second_call": host_ms_warm,
                  "host_buffer_note": "cb2_rt_render_csr into fresh pageable numpy arrays: rays H2D + kernels + CSR D2H through the "
                                      "pinned chunk ring (cb2_d2h) + first-touch faults of the destination",
                  "pcie_floor_ms": csr_bytes / 55e9 * 1e3, "csr_bytes": csr_bytes,
                  "roofline": {"bound": "fp64", "achieved": fp64, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
                               "frac": fp64 / peaks["fp64_tflops"], "flop_fp64_per_step": F_STEP,
                               "peak_source": "measured live (cb2_measure_peak_fp64: DFMA issue-rate microbenchmark, same process)",
                               "hbm": {"achieved_gbs": csr_bytes / ms * 1e-6, "peak_gbs": hbm_peak, "frac": csr_bytes / ms * 1e-6 / hbm_peak}}}))
