"""SART at the C4 scale: the ray-transfer CSR of the 512 x 512 pinhole frame over the (400, 1, 800) Generomak grid is built on
the device and inverted in place for a fixed number of iterations.  Prints one JSON line with the HBM roofline of the
iteration kernels (algorithmic bytes = CSR pass + CSC pass of the stored matrix per iteration, cb2_sart_info[2]).

    python tools/bench_sart.py [--pixels 512] [--iterations 50] [--frames 1] [--f32]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import core_b200 as cb  # noqa: E402
from core_b200.engine import DeviceRays, RayTransferScene  # noqa: E402
from core_b200.raytransfer import RayTransferCylinder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pixels", type=int, default=512)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="also time the reference's Cython SART on this many detectors")
    a = ap.parse_args()
    import torch
    rtc = RayTransferCylinder(radius_outer=2.41, height=3.35, n_radius=400, n_height=800, radius_inner=0.73, transform=cb.translate(0, 0, -1.8))
    prim = cb.HollowCylinder(0.73, 2.41, -1.8, 1.55)
    cam = cb.PinholeCamera((a.pixels, a.pixels), fov=45.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    batch = cb.ray_segments(prim, *cam.rays())
    scene = RayTransferScene(rtc)
    dev = DeviceRays(batch)
    cap = int(batch.n_rays) * 1300
    row_offset, columns, lengths = scene.render_csr_device(dev, capacity=cap)
    torch.cuda.synchronize()
    t0 = time.time()
    solver = cb.SartSolver.from_device_csr(row_offset, columns, lengths, rtc.bins, value_dtype=np.float32 if a.f32 else np.float64)
    t_setup = time.time() - t0
    info = solver.info()
    # measurements: projections of a smooth emissivity through the same matrix
    rr, zz = np.meshgrid(np.linspace(-1, 1, 400), np.linspace(-1, 1, 800), indexing="ij")
    truth = np.exp(-(rr ** 2 + zz ** 2) / 0.3).reshape(-1)
    x = torch.from_numpy(truth).cuda()
    crow = row_offset
    csr = torch.sparse_csr_tensor(crow, columns.to(torch.int64), lengths, size=(batch.n_rays, rtc.bins))
    m = (csr @ x).cpu().numpy()
    frames = np.stack([m * (1.0 + 0.1 * f) for f in range(a.frames)])
    solver(frames, max_iterations=3, conv_tol=0.0)                       # warm-up
    sols, convs = solver(frames, max_iterations=a.iterations, conv_tol=0.0)
    info = solver.info()
    ms = info["solve_ms"]
    it = info["iterations_launched"]
    groups = (a.frames + 3) // 4
    gbs = info["bytes_per_iteration"] * it * groups / (ms * 1e-3) / 1e9
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {"metric": "sart_iterations_per_second", "value": it * a.frames / (ms * 1e-3), "unit": "frame-iterations/s",
           "config": {"workload": "C4 geometry matrix %dx%d rays x %d sources" % (a.pixels, a.pixels, rtc.bins), "nnz": int(info["nnz"]),
                      "frames": a.frames, "iterations": int(it), "value_dtype": "f32" if a.f32 else "f64"},
           "ms_per_iteration": ms / it / groups, "setup_s": t_setup, "final_convergence": convs[0][-1],
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None}}
    if a.cpu_sample:
        from oracle import sart
        ref = sart.ref_module()
        n = a.cpu_sample
        ro = row_offset[:n + 1].cpu().numpy(); co = columns[:ro[-1]].cpu().numpy(); le = lengths[:ro[-1]].cpu().numpy()
        used = np.unique(co)
        remap = -np.ones(rtc.bins, dtype=np.int64); remap[used] = np.arange(used.size)
        dense = np.zeros((n, used.size))
        for r in range(n):
            dense[r, remap[co[ro[r]:ro[r + 1]]]] = le[ro[r]:ro[r + 1]]
        fn = ref.invert_sart if ref is not None else sart.invert_sart
        t0 = time.time()
        _, c = fn(dense, m[:n].copy(), max_iterations=3, conv_tol=0.0)
        dt = (time.time() - t0) / len(c)
        out["cpu_baseline"] = {"kind": "reference" if ref is not None else "port", "cores": 1, "sample": "%d detectors x %d sources dense" % (n, used.size),
                               "s_per_iteration": dt, "matrix_entries_per_s": n * used.size / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
