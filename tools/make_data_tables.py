#!/usr/bin/env python
"""Generate the data fixtures the hot path needs from the reference checkout's DATA files.

Run in the build container only (needs /root/reference); the outputs are committed so that tests, smoke()
and bench.py never read /root/reference at run time.

  core_b200/data/generomak.npz      Generomak equilibrium + edge mesh + edge/core profiles
                                    (cherab/generomak/equilibrium/data/generomak_equilibrium.json,
                                     cherab/generomak/plasma/data/{edge,core}/*.json — SURVEY Appendix D)
  core_b200/data/generomak_first_wall.npz  the first-wall component meshes (cherab/generomak/machine/data/first_wall/*.obj: vertices
                                    float64 [n, 3] and triangles int32 [m, 3] per component, faces fan-triangulated)
  core_b200/data/atomic_tables.npz  free-free Gaunt factor table (cherab/core/atomic/data/maxwellian_free_free_gaunt_factor.json)
                                    and the Stark model coefficients (cherab/core/atomic/data/lineshape/stark/{h,d,t}.json)
"""
import json
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "core_b200", "data")


def jload(*parts):
    with open(os.path.join(REF, *parts)) as f:
        return json.load(f)


def generomak():
    eq = jload("cherab/generomak/equilibrium/data/generomak_equilibrium.json")
    out = {}
    for k in ("r", "z", "psi_grid", "f_profile", "q_profile", "lcfs_polygon", "limiter_polygon",
              "magnetic_axis", "x_points", "strike_points"):
        out["eq_" + k] = np.asarray(eq[k], dtype=np.float64)
    for k in ("psi_axis", "psi_lcfs", "b_vacuum_radius", "b_vacuum_magnitude", "time"):
        out["eq_" + k] = np.float64(eq[k])
    pdir = "cherab/generomak/plasma/data"
    mesh = jload(pdir, "edge/mesh.json")
    out["mesh_vertices"] = np.asarray(mesh["vertex_coords"], dtype=np.float64)
    out["mesh_triangles"] = np.asarray(mesh["triangles"], dtype=np.int32)
    out["core_psi_norm"] = np.asarray(jload(pdir, "core/psi_norm.json")["psi_norm"], dtype=np.float64)
    species = [("electrons", "electron")] + [("hydrogen%d" % i, "hydrogen%d" % i) for i in range(2)] + \
              [("carbon%d" % i, "carbon%d" % i) for i in range(7)]
    for fname, key in species:
        e = jload(pdir, "edge/%s.json" % fname)
        c = jload(pdir, "core/%s.json" % fname)
        out["edge_%s_density" % key] = np.asarray(e["density"], dtype=np.float64)
        out["edge_%s_temperature" % key] = np.asarray(e["temperature"], dtype=np.float64)
        for q in ("density", "temperature", "vtor", "vpol", "vnorm"):
            out["core_%s_%s" % (key, q)] = np.asarray(c[q], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "generomak.npz"), **out)


def atomic_tables():
    g = jload("cherab/core/atomic/data/maxwellian_free_free_gaunt_factor.json")
    out = {"gaunt_u": np.asarray(g["u"], dtype=np.float64),
           "gaunt_gamma2": np.asarray(g["gamma2"], dtype=np.float64),
           "gaunt_factor": np.asarray(g["gaunt_factor"], dtype=np.float64)}
    stark = {}
    for iso in ("h", "d", "t"):
        stark[iso] = jload("cherab/core/atomic/data/lineshape/stark/%s.json" % iso)
    out["stark_json"] = np.asarray(json.dumps(stark))
    zee = {}
    zdir = os.path.join(REF, "cherab/core/atomic/data/lineshape/zeeman/parametrised")
    for fn in sorted(os.listdir(zdir)):
        zee[fn[:-5]] = jload("cherab/core/atomic/data/lineshape/zeeman/parametrised", fn)
    out["zeeman_parametrised_json"] = np.asarray(json.dumps(zee))
    np.savez_compressed(os.path.join(OUT, "atomic_tables.npz"), **out)


def read_obj(path):
    """Wavefront OBJ -> (vertices [n, 3] float64, triangles [m, 3] int32); polygons are fan-triangulated, indices may be
    negative (relative) and may carry /vt/vn suffixes."""
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(v) + i)
                for k in range(1, len(idx) - 1):
                    f.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int32)


def first_wall():
    wdir = os.path.join(REF, "cherab/generomak/machine/data/first_wall")
    out = {}
    for fn in sorted(os.listdir(wdir)):
        if fn.endswith(".obj"):
            v, f = read_obj(os.path.join(wdir, fn))
            out[fn[:-4] + "_vertices"] = v
            out[fn[:-4] + "_triangles"] = f
    np.savez_compressed(os.path.join(OUT, "generomak_first_wall.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    generomak()
    atomic_tables()
    first_wall()
    for fn in ("generomak.npz", "atomic_tables.npz", "generomak_first_wall.npz"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)), "bytes")
