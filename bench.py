#!/usr/bin/env python
"""bench.py — the headline measurement: Generomak camera spectral ray samples/s.

Workload (BASELINE.json configs[2]): Generomak full blended plasma, 1024x1024 pinhole camera at
(2.3, 0, 1.25) looking at (1.0, 0.8, -0.5), 2048 bins on [390, 700] nm, models = H alpha..delta x {Excitation,
Recombination} (GaussianLine) + Bremsstrahlung, synthetic ADF15-shaped rates, step 1 mm.  One "step" = one pass of the
whole camera with one of the 16 stratified sub-pixel sample positions (16 steps = the 16 samples/pixel frame); the
pass accumulates into the device-resident frame.  With N > 1 (torchrun) the 16x16-pixel image tiles are dealt
round-robin to the ranks (fixed total work: strong scaling); NCCL is used only to gather the frame (e2e leg).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pixels P] [--bins B]

One JSON line on stdout (rank 0).  `value` is device-timed (CUDA events, inputs resident in HBM); `e2e` is the same
metric through the host-buffer C-ABI call (ray segments H2D + kernel + frame D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Generomak camera spectral ray samples/s (M)"
UNIT = "Msamples/s"
CAMERA_POS, CAMERA_TARGET = (2.3, 0.0, 1.25), (1.0, 0.8, -0.5)
TILE = 16

# Algorithmic work per unit: minimal formulation, counted from the kernels' own counters
F_GAUSS_BIN = 14.0      # flop per Gaussian bin evaluation (5-term series form: 7 FMA-class) + 1 SFU            [bin_kernel]
F_FIXED_BASE = 20.0 + 3 * 40.0 + 30.0   # transform/R, three bicubics, masks/blend                [state_kernel]
F_MOMENT = 80.0         # flop per live sample for the Bremsstrahlung moment update (weights + 4 nodes x charges) [state_kernel]
REC_BYTES = 384.0       # bytes per live (32-sample group, line component) record row, written once and read once


def build_scene(bins, lo=390.0, hi=700.0):
    import core_b200 as cb
    from core_b200 import generomak
    plasma = generomak.get_plasma()
    lines = [cb.Line(cb.hydrogen, 0, (n, 2)) for n in (3, 4, 5, 6)]
    plasma.models = [cb.ExcitationLine(l) for l in lines] + [cb.RecombinationLine(l) for l in lines] + [cb.Bremsstrahlung()]
    flat = cb.flatten_scene(plasma, lo, hi, bins)
    return plasma, flat


def fixed_flops_per_sample(flat):
    d = flat.desc
    n_line = sum(1 for i in range(d.n_models) if d.models[i].kind != 2)
    species = {d.models[i].species for i in range(d.n_models) if d.models[i].kind != 2}
    n_charged = sum(1 for i in range(d.n_species) if d.species[i].charge > 0) if n_line < d.n_models else 0
    p = 2 + 4 * len(species) + n_charged
    return F_FIXED_BASE + 10.0 * p + 60.0 * n_line


def rank_pixels(pixels, rank, world):
    from core_b200.sharding import tile_pixels
    return tile_pixels(pixels, rank, world)


def make_rays(plasma, pixels, pixel_index, sample_id):
    import core_b200 as cb
    sx, sy = cb.stratified_offsets(4)[sample_id % 16]
    cam = cb.PinholeCamera((pixels, pixels), fov=45, transform=cb.look_at(CAMERA_POS, CAMERA_TARGET))
    o, d = cam.rays(sx, sy, pixel_index)
    return cb.ray_segments(plasma.geometry, o, d, plasma.geometry_to_world())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(flat, plasma, pixels, n_rays, threads, seed=1234):
    """Oracle (CPU restatement, fp64) timed on a bounded sample of the same workload."""
    from oracle import oracle
    rng = np.random.default_rng(seed)
    pix = np.sort(rng.choice(pixels * pixels, size=n_rays, replace=False))
    rays = make_rays(plasma, pixels, pix, 0)
    t0 = time.perf_counter()
    _, st = oracle.emission_render(flat, rays, n_threads=threads)
    dt = time.perf_counter() - t0
    return st["samples"] / dt * 1e-6, st["samples"], dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the real Cython/Raysect path cannot be built here —
    raysect is absent) on the box's host cores, same config/metric, each step a bounded sample."""
    if rank != 0:
        return
    import __graft_entry__ as g
    from oracle import oracle
    oracle.build()
    plasma, flat = build_scene(args.bins)
    threads = os.cpu_count() or 1
    n_rays = args.ref_rays
    vals = []
    # the real Cherab + Raysect API when both import on this box (tools/run_reference.py; untested where they are absent),
    # else the oracle port
    kind = "port"
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_reference
        if run_reference.available() is None:
            kind = "reference"
    except Exception:
        kind = "port"
    for k in range(args.warmup + args.steps):
        if kind == "reference":
            rng = np.random.default_rng(1234 + k)
            pick = np.sort(rng.choice(args.pixels * args.pixels, size=n_rays, replace=False))
            rays = make_rays(plasma, args.pixels, pick, 0)
            length = rays.seg_t1 - rays.seg_t0
            samples = int((np.maximum(flat.desc.min_samples - 1, np.ceil(length / flat.desc.step)) + 1).sum())
            _, dt = run_reference.run(args.pixels, args.bins, pick, 0)
        else:
            v, samples, dt = cpu_baseline(flat, plasma, args.pixels, n_rays, threads, seed=1234 + k)
        if k >= args.warmup:
            vals.append((samples, dt))
    tot_s = sum(s for s, _ in vals)
    tot_t = sum(t for _, t in vals)
    value = tot_s / tot_t * 1e-6
    sample = "%d rays of the %dx%d frame per step (random pixels, seeded), %d bins, all models" % (n_rays, args.pixels, args.pixels, args.bins)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_t / max(len(vals), 1) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(args, world):
    return {"workload": "Generomak full-frame %dx%d pinhole camera, H alpha..delta x {excitation, recombination} + Bremsstrahlung, "
                        "%d bins on [390,700] nm, 1 of 16 stratified samples/pixel per step, step 1 mm" % (args.pixels, args.pixels, args.bins),
            "pixels": [args.pixels, args.pixels], "bins": args.bins, "models": 9, "tiles": "16x16 px round-robin over %d rank(s)" % world,
            "l2": "the %0.1f GB fp32 frame written per step exceeds L2; scene tables (few MB) are L2-resident by design"
                  % (args.pixels * args.pixels * args.bins * 4 / 1e9)}


def shared_host_frame(rows_total, bins, rank, dist, torch):
    """The whole [rows_total, bins] fp32 frame in shared host memory (/dev/shm), mapped and page-locked by every rank: each rank's
    strided D2H copies put its tiles on their pixels, so the frame a consumer maps is in image order.  Returns (numpy frame or
    None, description); None when /dev/shm cannot hold the frame or the mapping cannot be page-locked on some rank."""
    from core_b200.sharding import open_shared_frame, unlink_shared_frame
    name = "cb2_frame_%s" % os.environ.get("MASTER_PORT", "0")
    ok = torch.zeros(1, dtype=torch.int32, device="cuda")
    frame = None
    if rank == 0:
        try:
            open_shared_frame(name, rows_total, bins, create=True)
        except OSError:
            pass
    dist.barrier()
    try:
        frame, _mm = open_shared_frame(name, rows_total, bins, create=False)
        rc = torch.cuda.cudart().cudaHostRegister(frame.ctypes.data, frame.nbytes, 0)
        if int(getattr(rc, "value", rc)) != 0:
            frame = None
    except Exception:
        frame = None
    ok[0] = 1 if frame is not None else 0
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        unlink_shared_frame(name)                         # the mappings keep the memory alive; nothing is left behind
    if int(ok.item()) == 1:
        return frame, ("image-ordered frame [nx*ny, bins] (pixel ix*ny+iy) in shared host memory (/dev/shm): every rank writes its tiles to "
                       "their pixels over its own PCIe link (strided D2H inside the timed region)")
    return None, ""


_RESULT_FD = None


def emit(line):
    """The result line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pixels", type=int, default=1024)
    ap.add_argument("--bins", type=int, default=2048)
    ap.add_argument("--ref-rays", type=int, default=8)
    ap.add_argument("--cpu-rays", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--frame", default="host", choices=["host", "gather"],
                    help="N > 1 end-to-end path: every rank reads its shard back into the frame in shared host memory (default), "
                         "or the frame is gathered on rank 0's GPU with NCCL and read back over rank 0's PCIe link")
    args = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner at the first
    # communicator): point file descriptor 1 at stderr for the whole run and keep the real stdout for the result line.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from core_b200.engine import DeviceRays, EmissionScene, measure_peaks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the frame gather overlaps the next chunk's kernels: keep NCCL's (spin-waiting) send/recv kernels on a few SMs
        os.environ.setdefault("NCCL_MAX_CTAS", "4")
        os.environ.setdefault("NCCL_MIN_CTAS", "1")
        dist.init_process_group("nccl", device_id=dev)

    plasma, flat = build_scene(args.bins)
    scene = EmissionScene(flat, device=local_rank)
    pix = rank_pixels(args.pixels, rank, world)
    n_pass = args.warmup + args.steps
    host_rays = [make_rays(plasma, args.pixels, pix, k).pin() for k in range(min(n_pass, 16))]     # inputs of the e2e leg: pinned host memory
    dev_rays = [DeviceRays(r, device=dev) for r in host_rays]
    frame = torch.zeros((pix.size, args.bins), dtype=torch.float32, device=dev)
    stats = torch.zeros(6, dtype=torch.int64, device=dev)

    peaks = measure_peaks(local_rank) if rank == 0 else None
    plan = scene.info()

    def step(k):
        scene.render_device(dev_rays[k % len(dev_rays)], frame, scale=1.0 / 16.0, accumulate=(k % 16) != 0, stats=stats)

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    stats.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    st = stats.clone()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
    st = st.cpu().numpy()
    samples, gauss, brems, ood = int(st[0]), int(st[1]), int(st[3]), int(st[5])
    value = samples / (ms * 1e-3) * 1e-6

    # ---- per-kernel split: one extra pass bracketed by CUDA events inside the library (cb2_scene_profile) ----
    kernel_ms = None
    if rank == 0 and plan["two_kernel_line_path"]:
        scratch_stats = torch.zeros(6, dtype=torch.int64, device=dev)
        scene.profile(True)
        scene.render_device(dev_rays[args.warmup % len(dev_rays)], frame, scale=1.0 / 16.0, accumulate=True, stats=scratch_stats)
        torch.cuda.synchronize()
        kernel_ms = scene.profile(False)
    launches_per_step = sum(v[1] for v in kernel_ms.values()) if kernel_ms else 1

    # ---- e2e: host buffers through the C-ABI call (H2D rays + kernel + [NCCL gather] + D2H frame) ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.steps
        # The frame the consumer gets is in IMAGE order at every N: row p = ix * ny + iy of a [nx * ny, bins] float32 array.  The rays are
        # listed tile by tile (16 x 16 pixels), so the host-buffer call takes the destination row of every ray (cb2_emission_render_rows)
        # and copies each tile to its pixels with strided D2H copies that overlap the next batch's kernels.
        rows = pix
        if world == 1:
            host_frame = torch.empty((args.pixels * args.pixels, args.bins), dtype=torch.float32, pin_memory=True).numpy()
            scene.render(host_rays[0], out=host_frame, rows=rows)      # warm the staging buffers once
        n_chunks = 8
        gather = world > 1 and args.frame == "gather"
        frame_kind = ("image-ordered frame [nx*ny, bins] (pixel ix*ny+iy) in rank 0's pinned host memory, tiles copied to their pixels with "
                      "strided D2H inside the timed region") if world == 1 else "NCCL gather to rank 0, then D2H (rank-major tile order)"
        if world > 1 and not gather:
            # The path shards by pixel tiles and has no exchange step: every rank renders its tiles through the same host-buffer call
            # as at N = 1 and writes them over ITS OWN PCIe link to their pixels of one frame in shared host memory (/dev/shm file
            # mapped and page-locked by every rank).  No device collective.
            host_frame, frame_kind = shared_host_frame(args.pixels * args.pixels, args.bins, rank, dist, torch)
            if host_frame is None:
                host_frame = torch.empty((pix.size, args.bins), dtype=torch.float32, pin_memory=True).numpy()
                rows = np.arange(pix.size)
                frame_kind = "rank-local pinned host buffers in the rank's tile order (no /dev/shm room for the frame)"
            scene.render(host_rays[0], out=host_frame, rows=rows)      # warm the staging buffers once
            dist.barrier()
        if gather:
            # NCCL only gathers the frame.  The rank's rays are cut into chunks: while chunk c+1 renders, chunk c is gathered
            # into one device buffer on rank 0 (NCCL stream) and read back into pinned host memory (copy stream), so the
            # single PCIe link of rank 0 works in the shadow of the compute.
            counts = [rank_pixels(args.pixels, r, world).size for r in range(world)]
            cmax = -(-max(counts) // n_chunks)                 # rows per chunk (last chunk of a rank may be shorter: padded)
            bounds = [min(c * cmax, pix.size) for c in range(n_chunks + 1)]
            chunk_rays = [[hr.subset(np.arange(bounds[c], bounds[c + 1])) for c in range(n_chunks)] for hr in host_rays]
            send = [torch.zeros((cmax, args.bins), dtype=torch.float32, device=dev) for _ in range(n_chunks)]
            full = host_full = gathered = None
            if rank == 0:
                full = [torch.empty((world * cmax, args.bins), dtype=torch.float32, device=dev) for _ in range(n_chunks)]
                gathered = [[f[r * cmax:(r + 1) * cmax] for r in range(world)] for f in full]
                host_full = torch.empty((n_chunks, world * cmax, args.bins), dtype=torch.float32, pin_memory=True)
            # neither the compute nor the copy may run on the legacy default stream: it synchronises implicitly with every
            # other blocking stream, which would serialise render, gather and read-back
            copy_stream = torch.cuda.Stream(device=dev)
            compute_stream = torch.cuda.Stream(device=dev)
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_samples = 0
        for k in range(e2e_steps):
            if not gather:
                r = host_rays[(args.warmup + k) % len(host_rays)]
                _, s = scene.render(r, out=host_frame, rows=rows)
                e2e_samples += s["samples"]
            else:
                works = []
                with torch.cuda.stream(compute_stream):
                    for c in range(n_chunks):
                        rc_ = chunk_rays[(args.warmup + k) % len(host_rays)][c]
                        if rc_.n_rays:
                            dr = DeviceRays(rc_, device=dev)
                            scene.render_device(dr, send[c][:rc_.n_rays], scale=1.0, accumulate=False, stats=stats)
                        w = dist.gather(send[c], gathered[c] if rank == 0 else None, dst=0, async_op=True)
                        if rank == 0:
                            with torch.cuda.stream(copy_stream):
                                w.wait()                          # the copy stream (not the compute stream) waits for the gather
                                host_full[c].copy_(full[c], non_blocking=True)
                        works.append(w)
                    for w in works:
                        w.wait()
                torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            e2e_samples = samples * e2e_steps // max(args.steps, 1)
            h2d_all = torch.tensor([sum(a.nbytes for a in (host_rays[0].origin, host_rays[0].direction, host_rays[0].seg_offset,
                                                              host_rays[0].seg_t0, host_rays[0].seg_t1))], dtype=torch.float64, device=dev)
            dist.all_reduce(h2d_all)
        h2d = sum(host_rays[(args.warmup + k) % len(host_rays)].origin.nbytes * 2 + host_rays[(args.warmup + k) % len(host_rays)].seg_offset.nbytes
                  + host_rays[(args.warmup + k) % len(host_rays)].seg_t0.nbytes * 2 for k in range(e2e_steps)) // max(e2e_steps, 1)
        if world > 1:
            h2d = int(h2d_all.item())                     # all ranks' ray shards
        d2h_rows = n_chunks * world * cmax if gather else args.pixels * args.pixels
        e2e = {"value": e2e_samples / dt * 1e-6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h_rows * args.bins * 4 + 48 * world), "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3, "frame": frame_kind}

    if rank == 0:
        fixed = fixed_flops_per_sample(flat)
        n_line = sum(1 for i in range(flat.desc.n_models) if flat.desc.models[i].kind != 2)
        step_s = ms * 1e-3 / args.steps
        per_step = lambda x: x / args.steps / world                      # counters are summed over steps and ranks
        live = brems / max(args.bins, 1)                                   # samples with ne, te > 0 (Bremsstrahlung counter / bins)
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        hbm_src = "of measured" if "hbm_gbs" in peaks_file else "of fallback"
        kernels = {}
        if kernel_ms:
            tot = sum(v[0] for v in kernel_ms.values()) or 1.0
            k_pad, n_pad = plan["moment_row"], -(-args.bins // 128) * 128
            work = {
                # (flop, sfu ops) per step on this GPU, algorithmic
                "state_kernel": (per_step(fixed * samples + F_MOMENT * live), per_step((4.0 + 2.0 * n_line) * live)),
                "bin_kernel": (per_step(F_GAUSS_BIN * gauss), per_step(gauss)),
                "contract_kernel": (2.0 * pix.size * k_pad * n_pad, 0.0),
                "helpers": (0.0, 0.0)}
            for name, (kms, n) in kernel_ms.items():
                flop, sfu = work[name]
                t = kms * 1e-3
                e = {"ms_per_step": kms, "share": kms / tot, "launches_per_step": n}
                if t > 0 and flop > 0:
                    e["fp32"] = {"achieved": flop / t * 1e-12, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s", "frac": flop / t * 1e-12 / peaks["fp32_tflops"]}
                if t > 0 and sfu > 0:
                    e["sfu"] = {"achieved": sfu / t * 1e-12, "peak": peaks["sfu_tops"], "unit": "Tops/s", "frac": sfu / t * 1e-12 / peaks["sfu_tops"]}
                if name == "contract_kernel" and plan.get("contraction_on_tensor_cores") and t > 0:
                    # error-compensated 3xTF32: three tensor-core GEMMs per algorithmic float32 GEMM (+ the hi/lo split of the moments).
                    # Peak = the dense TF32 rate, half the measured bf16 one (B200_PROFILING.md: 1.1 vs 2.25 PFLOP/s nominal).
                    tf32_peak = 0.5 * float(peaks_file.get("bf16_tflops", 2250.0))
                    e.pop("fp32", None)
                    e["tensor"] = {"achieved": 3.0 * flop / t * 1e-12, "peak": tf32_peak, "unit": "TFLOP/s", "frac": 3.0 * flop / t * 1e-12 / tf32_peak,
                                   "fp32_equivalent_tflops": flop / t * 1e-12,
                                   "note": "hand-written tcgen05 kind::tf32 kernel (contract_tc_kernel: 3 MMAs per k-block on hi/lo operand tiles, frame += fused); peak = measured bf16 / 2"}
                kernels[name] = e
            dom = max((k for k in kernels if k != "helpers"), key=lambda k: kernels[k]["ms_per_step"])
        else:
            dom = "emission_kernel"
            kernels[dom] = {"ms_per_step": step_s * 1e3, "share": 1.0, "launches_per_step": 1,
                            "fp32": {"achieved": per_step(fixed * samples + F_GAUSS_BIN * gauss + 10.0 * brems) / step_s * 1e-12,
                                     "peak": peaks["fp32_tflops"], "unit": "TFLOP/s"},
                            "sfu": {"achieved": per_step(gauss + brems) / step_s * 1e-12, "peak": peaks["sfu_tops"], "unit": "Tops/s"}}
            for kk in ("fp32", "sfu"):
                kernels[dom][kk]["frac"] = kernels[dom][kk]["achieved"] / kernels[dom][kk]["peak"]
        d = kernels[dom]
        cands = [("fp32", d.get("fp32")), ("sfu", d.get("sfu"))]
        bound, best = max(((b_, r_) for b_, r_ in cands if r_), key=lambda x: x[1]["frac"])
        # whole-step view: all algorithmic flops of the step over the step time
        step_flop = per_step(fixed * samples + F_MOMENT * live + F_GAUSS_BIN * gauss) + (2.0 * pix.size * plan["moment_row"] * (-(-args.bins // 128) * 128) if kernel_ms else per_step(10.0 * brems))
        frame_bytes = pix.size * args.bins * 4.0
        # DRAM bytes per launch of the dominant kernel from one `ncu --set full` capture at the default configuration
        # (profiles/r2q_bin_kernel_c3_ncu_summary.txt: 3.88 GB read = the line records the state kernels hand over, 0.132 GB written =
        # the batch's spectra); unknown for any other size
        traffic = None
        if dom == "bin_kernel" and args.pixels == 1024 and args.bins == 2048 and world == 1:
            # (captured on a 16 384-ray launch; a launch of batch_rays rays moves proportionally more)
            traffic = {"bytes_per_launch": 4.014e9 * plan["batch_rays"] / 16384.0, "algorithmic_bytes_per_launch": plan["batch_rays"] * args.bins * 4.0,
                       "source": "profiles/r2q_bin_kernel_c3_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum)",
                       "note": "the line records handed from the state kernels to bin_kernel are re-read from HBM (0.5 TB/s, 8 % of the HBM peak): "
                               "not the bound of an issue-bound kernel; the opt-in fused kernel (CB2_FUSED=1) moves 0.30 GB per launch "
                               "(profiles/r2b_fused_kernel_ncu_summary.txt) at 10 % more time"}
        roofline = {"bound": bound, "achieved": best["achieved"], "peak": best["peak"], "unit": best["unit"], "frac": best["frac"],
                    "traffic": traffic, "kernel": dom, "launch_ms": d["ms_per_step"] / max(d["launches_per_step"], 1),
                    "peak_source": "measured live (cb2_measure_peaks: FFMA / MUFU.EX2 issue-rate microbenchmarks, same process)",
                    "kernels": kernels,
                    "step": {"ms": step_s * 1e3, "fp32_tflops": step_flop / step_s * 1e-12, "fp32_frac": step_flop / step_s * 1e-12 / peaks["fp32_tflops"],
                             "flop_per_sample": step_flop / max(per_step(samples), 1)},
                    "algorithmic": {"samples": samples, "live_samples": int(live), "gaussian_bin_evals": gauss,
                                    "flop_per_gauss_bin": F_GAUSS_BIN, "fixed_flop_per_sample": fixed, "moment_flop_per_live_sample": F_MOMENT,
                                    "contraction_flop_per_ray": 2.0 * plan["moment_row"] * (-(-args.bins // 128) * 128)},
                    "hbm": {"achieved": frame_bytes * 3.0 / step_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                            "note": "frame written by bin_kernel and read-modify-written by contract_kernel; not a bound; " + hbm_src}}
        threads = os.cpu_count() or 1
        cpu_v, cpu_s, cpu_t = cpu_baseline(flat, plasma, args.pixels, args.cpu_rays, threads)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args, world), "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline,
                "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": "%d rays (random pixels, seeded) of the same frame, %d samples in %.1f s" % (args.cpu_rays, cpu_s, cpu_t)},
                "out_of_domain_samples": ood, "plan": plan}
        if e2e:
            line["e2e"] = e2e
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
