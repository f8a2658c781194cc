"""GPU experiment driver for the state kernels: runs the C3 scene on a sub-frame once per environment variant (each in its
own process, so the library re-reads its tuning variables), prints the per-kernel device times and compares the frames.

  python tools/probe_state.py [pixels] [VAR=val,VAR=val ...]        parent: one child per variant (first = reference frame)
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def child(pixels, tag, bins):
    import torch
    import bench
    from core_b200.engine import DeviceRays, EmissionScene
    plasma, flat = bench.build_scene(bins) if bins == 2048 else bench.build_scene(bins, 651.279, 661.279)
    pix = np.arange(pixels * pixels)
    rays = bench.make_rays(plasma, pixels, pix, 0)
    t0 = time.perf_counter()
    sc = EmissionScene(flat)
    t_create = time.perf_counter() - t0
    dr = DeviceRays(rays)
    out = torch.zeros((rays.n_rays, bins), dtype=torch.float32, device="cuda:0")
    stats = torch.zeros(6, dtype=torch.int64, device="cuda:0")
    for _ in range(2):
        sc.render_device(dr, out, stats=stats)
    torch.cuda.synchronize()
    stats.zero_()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sc.render_device(dr, out, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = (stats.cpu().numpy() // reps).tolist()
    sc.profile(True)
    sc.render_device(dr, out)
    torch.cuda.synchronize()
    prof = sc.profile(False)
    lib = sc._lib
    info = {k: int(lib.cb2_scene_info(sc._h, k)) for k in (9, 10, 11)}
    frame = out.cpu().numpy()
    np.save(os.path.join(OUT, "probe_%s.npy" % tag), frame[:: max(1, rays.n_rays // 1024)])
    res = {"tag": tag, "ms": ms, "samples": st[0], "msamples_s": st[0] / ms * 1e-3, "ood": st[5], "create_s": t_create,
           "prof": {k: v[0] for k, v in prof.items()}, "memo_rows": info[9], "memo_err_1e9": info[10], "fixup_ms": info[11] * 1e-3}
    if os.environ.get("PROBE_ORACLE"):
        from oracle import oracle
        rng = np.random.default_rng(5)
        sel = np.sort(rng.choice(rays.n_rays, size=int(os.environ["PROBE_ORACLE"]), replace=False))
        sub = bench.make_rays(plasma, pixels, sel, 0)
        ref, _ = oracle.emission_render(flat, sub)
        got = frame[sel].astype(np.float64)
        tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
        res["oracle_worst"] = float(np.max(np.abs(got - ref) / (tol + 1e-300)))
    print("PROBE " + json.dumps(res), flush=True)
    sc.close()


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]), sys.argv[3], int(sys.argv[4]))
    pixels = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    bins = int(os.environ.get("PROBE_BINS", "2048"))
    variants = sys.argv[2:] or ["CB2_STATE_MEMO=0", "CB2_STATE_MEMO=1"]
    ref = None
    for i, v in enumerate(variants):
        env = dict(os.environ)
        for kv in v.split(","):
            if "=" in kv:
                k, val = kv.split("=", 1)
                env[k] = val
        tag = "v%d" % i
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(pixels), tag, str(bins)], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
        for l in r.stderr.splitlines():
            if l.startswith("state table"):
                print("   " + l)
        if not line:
            print("variant %s FAILED\n%s\n%s" % (v, r.stdout[-2000:], r.stderr[-3000:]), flush=True)
            continue
        res = json.loads(line[0][6:])
        frame = np.load(os.path.join(OUT, "probe_%s.npy" % tag)).astype(np.float64)
        os.remove(os.path.join(OUT, "probe_%s.npy" % tag))
        cmp = ""
        if ref is None:
            ref = frame
        else:
            tol = 1e-4 * np.abs(ref) + 1e-9 * np.abs(ref).max(axis=1, keepdims=True)
            cmp = " worst|d|/tol vs v0 = %.3g" % float(np.max(np.abs(frame - ref) / (tol + 1e-300)))
        print("%-40s %8.2f ms %7.1f Msamples/s  state %.2f bin %.2f contract %.2f helpers %.2f  rows=%d err=%.2g ood=%d create=%.2fs fixup=%.2f%s%s" % (
            v, res["ms"], res["msamples_s"], res["prof"]["state_kernel"], res["prof"]["bin_kernel"], res["prof"]["contract_kernel"],
            res["prof"]["helpers"], res["memo_rows"], res["memo_err_1e9"] * 1e-9, res["ood"], res["create_s"], res["fixup_ms"], cmp,
            "  oracle worst=%.3g" % res["oracle_worst"] if "oracle_worst" in res else ""), flush=True)


if __name__ == "__main__":
    main()
