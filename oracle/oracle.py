"""ctypes wrapper of the CPU oracle (oracle/libcb2_oracle.so).  TEST INFRASTRUCTURE — only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from core_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcb2_oracle.so")
ORACLE_SYMBOLS = ["cb2o_abi_version", "cb2o_last_error", "cb2o_emission_render", "cb2o_sample_state", "cb2o_state_width", "cb2o_beam_sample",
                  "cb2o_rt_render_dense", "cb2o_add_gaussian_line", "cb2o_add_lorentzian_line", "cb2o_interp1d_cubic",
                  "cb2o_interp2d_cubic", "cb2o_gauss_legendre", "cb2o_gaunt_factor", "cb2o_pec_evaluate", "cb2o_interp3d_cubic",
                  "cb2o_thermal_cx_pec_evaluate", "cb2o_wall_hit"]
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cb2_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "cherab_b200.h")
    # content hash, not file times: times do not survive the copy to the GPU box
    import hashlib
    want = hashlib.sha256(open(src, "rb").read() + open(hdr, "rb").read() + open(os.path.join(_HERE, "cb2_oracle.h"), "rb").read()).hexdigest()
    stamp = LIB_PATH + ".sha"
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    if force or not os.path.exists(LIB_PATH) or have != want:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcb2_oracle.so"], stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(want)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        dp = _abi.c_double_p
        l.cb2o_last_error.restype = C.c_char_p
        l.cb2o_emission_render.argtypes = [C.POINTER(_abi.SceneDesc), C.POINTER(_abi.Rays), dp, C.c_double, C.c_int, C.c_int, C.POINTER(_abi.Stats)]
        l.cb2o_sample_state.argtypes = [C.POINTER(_abi.SceneDesc), dp, C.c_int64, dp]
        l.cb2o_state_width.argtypes = [C.POINTER(_abi.SceneDesc)]
        l.cb2o_beam_sample.argtypes = [C.POINTER(_abi.SceneDesc), dp, C.c_int64, dp]
        l.cb2o_rt_render_dense.argtypes = [C.POINTER(_abi.RTDesc), C.POINTER(_abi.Rays), dp, C.c_int, C.c_int, C.POINTER(_abi.Stats)]
        l.cb2o_add_gaussian_line.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(_abi.SpectralGrid), dp]
        l.cb2o_add_lorentzian_line.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(_abi.SpectralGrid), dp, C.c_double, C.c_int, C.c_int]
        l.cb2o_interp1d_cubic.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int]
        l.cb2o_interp1d_cubic.restype = C.c_double
        l.cb2o_interp2d_cubic.argtypes = [dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        l.cb2o_interp2d_cubic.restype = C.c_double
        l.cb2o_gaunt_factor.argtypes = [C.POINTER(_abi.Gaunt), C.c_double, C.c_double, C.c_double]
        l.cb2o_gaunt_factor.restype = C.c_double
        l.cb2o_pec_evaluate.argtypes = [C.POINTER(_abi.Rate2D), C.c_double, C.c_double, C.c_double]
        l.cb2o_pec_evaluate.restype = C.c_double
        l.cb2o_interp3d_cubic.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        l.cb2o_interp3d_cubic.restype = C.c_double
        l.cb2o_thermal_cx_pec_evaluate.argtypes = [C.POINTER(_abi.Rate3D), C.c_double, C.c_double, C.c_double, C.c_double]
        l.cb2o_thermal_cx_pec_evaluate.restype = C.c_double
        for s in ORACLE_SYMBOLS:
            getattr(l, s)
        _lib = l
    return _lib


def _dp(a):
    return a.ctypes.data_as(_abi.c_double_p)


def emission_render(flat, rays, scale=1.0, n_threads=0, out=None):
    """fp64 oracle render of ``rays`` (RayBatch) for a FlatScene -> (spectra[n_rays, bins], stats dict)."""
    l = lib()
    bins = flat.desc.grid.bins
    accumulate = out is not None
    if out is None:
        out = np.zeros((rays.n_rays, bins), dtype=np.float64)
    st = _abi.Stats()
    rs = rays.as_struct()
    _abi.check(l, l.cb2o_emission_render(C.byref(flat.desc), C.byref(rs), _dp(out), scale, int(accumulate), n_threads, C.byref(st)),
               "cb2o_last_error")
    return out, st.as_dict()


def sample_state(flat, points):
    l = lib()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    w = l.cb2o_state_width(C.byref(flat.desc))
    out = np.zeros((pts.shape[0], w), dtype=np.float64)
    _abi.check(l, l.cb2o_sample_state(C.byref(flat.desc), _dp(pts), pts.shape[0], _dp(out)), "cb2o_last_error")
    return out


def wall_hit(triangles, origin, direction):
    """Distance to the first hit of every ray with the triangle soup [n, 3, 3] (+inf: miss), brute force in float64."""
    l = lib()
    tri = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
    o = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
    t = np.empty(o.shape[0])
    l.cb2o_wall_hit.argtypes = [_abi.c_double_p, C.c_int64, _abi.c_double_p, _abi.c_double_p, C.c_int64, _abi.c_double_p]
    _abi.check(l, l.cb2o_wall_hit(_dp(tri), tri.shape[0], _dp(o), _dp(d), o.shape[0], _dp(t)), "cb2o_last_error")
    return t


def beam_sample(flat, beam_points):
    """Beam.density / Beam.direction at points in beam coordinates -> (density[n], direction[n, 3])."""
    l = lib()
    pts = np.ascontiguousarray(beam_points, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((pts.shape[0], 4), dtype=np.float64)
    _abi.check(l, l.cb2o_beam_sample(C.byref(flat.desc), _dp(pts), pts.shape[0], _dp(out)), "cb2o_last_error")
    return out[:, 0], out[:, 1:]


def rt_render_dense(rt_desc, rays, n_threads=0):
    l = lib()
    out = np.zeros((rays.n_rays, rt_desc.bins), dtype=np.float64)
    st = _abi.Stats()
    rs = rays.as_struct()
    _abi.check(l, l.cb2o_rt_render_dense(C.byref(rt_desc), C.byref(rs), _dp(out), 0, n_threads, C.byref(st)), "cb2o_last_error")
    return out, st.as_dict()


def add_gaussian_line(radiance, wavelength, sigma, min_wavelength, max_wavelength, bins, samples=None):
    g = _abi.SpectralGrid(min_wavelength, max_wavelength, bins, 0)
    if samples is None:
        samples = np.zeros(bins)
    lib().cb2o_add_gaussian_line(radiance, wavelength, sigma, C.byref(g), _dp(samples))
    return samples


def add_lorentzian_line(radiance, wavelength, lambda_1_2, min_wavelength, max_wavelength, bins, samples=None,
                        rtol=1e-5, min_order=1, max_order=50):
    g = _abi.SpectralGrid(min_wavelength, max_wavelength, bins, 0)
    if samples is None:
        samples = np.zeros(bins)
    lib().cb2o_add_lorentzian_line(radiance, wavelength, lambda_1_2, C.byref(g), _dp(samples), rtol, min_order, max_order)
    return samples


def interp1d_cubic(x, f, px):
    x, f = np.ascontiguousarray(x, float), np.ascontiguousarray(f, float)
    return lib().cb2o_interp1d_cubic(_dp(x), _dp(f), x.size, px, 1)


def interp2d_cubic(x, y, f, px, py):
    x, y, f = (np.ascontiguousarray(a, float) for a in (x, y, f))
    return lib().cb2o_interp2d_cubic(_dp(x), _dp(y), _dp(f), x.size, y.size, px, py, 1)


def interp3d_cubic(x, y, z, f, px, py, pz):
    x, y, z, f = (np.ascontiguousarray(a, float) for a in (x, y, z, f))
    return lib().cb2o_interp3d_cubic(_dp(x), _dp(y), _dp(z), _dp(f), x.size, y.size, z.size, px, py, pz)
