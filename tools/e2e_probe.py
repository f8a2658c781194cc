"""Where the end-to-end time goes: device-resident render vs host-buffer render (with / without the per-batch D2H overlap)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from core_b200.engine import DeviceRays, EmissionScene
px = int(sys.argv[1]) if len(sys.argv) > 1 else 512
plasma, flat = bench.build_scene(2048)
scene = EmissionScene(flat)
pix = bench.rank_pixels(px, 0, 1)
rays = bench.make_rays(plasma, px, pix, 0)
dr = DeviceRays(rays)
frame = torch.zeros((pix.size, 2048), dtype=torch.float32, device="cuda:0")
host = torch.empty((pix.size, 2048), dtype=torch.float32, pin_memory=True)
hnp = host.numpy()
for _ in range(2):
    scene.render_device(dr, frame); torch.cuda.synchronize()
t0 = time.perf_counter(); scene.render_device(dr, frame); torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
scene.render(rays, out=hnp)
t0 = time.perf_counter(); scene.render(rays, out=hnp); t_host = time.perf_counter() - t0
t0 = time.perf_counter(); host.copy_(frame); torch.cuda.synchronize(); t_copy = time.perf_counter() - t0
t0 = time.perf_counter(); d2 = DeviceRays(rays); torch.cuda.synchronize(); t_h2d = time.perf_counter() - t0
print("px %d rays %d: device render %.1f ms, host-buffer render %.1f ms, plain D2H of the frame %.1f ms (%.1f GB/s), rays H2D %.1f ms, overlap=%s"
      % (px, pix.size, t_dev * 1e3, t_host * 1e3, t_copy * 1e3, frame.numel() * 4 / t_copy * 1e-9, t_h2d * 1e3, os.environ.get("CB2_D2H_OVERLAP", "1")))
