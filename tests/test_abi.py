"""The C-ABI library loads without a GPU and exports exactly what include/cherab_b200.h declares (no compute calls)."""
import os
import re

import core_b200
from core_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(prefix, header=("include", "cherab_b200.h")):
    text = open(os.path.join(ROOT, *header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, text)))


def test_product_library_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    lib = _abi.load_library()
    names = _declared("cb2_")
    assert names == sorted(_abi.PRODUCT_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cb2_abi_version() == _abi.ABI_VERSION


def test_oracle_library_exports_header_symbols():
    from oracle import oracle
    lib = oracle.lib()
    names = _declared("cb2o_", ("oracle", "cb2_oracle.h"))
    assert names == sorted(oracle.ORACLE_SYMBOLS)
    assert _declared("cb2o_") == []          # the product header declares nothing of the oracle
    for n in names:
        assert hasattr(lib, n), n


def test_struct_sizes_match_header():
    # compile-time sizes from a tiny C program built against the header
    import subprocess, tempfile, ctypes
    structs = {"cb2_spectral_grid": _abi.SpectralGrid, "cb2_scalar_field": _abi.ScalarField, "cb2_vector_field": _abi.VectorField,
               "cb2_equilibrium": _abi.Equilibrium, "cb2_axisym": _abi.Axisym, "cb2_species": _abi.SpeciesDesc,
               "cb2_rate2d": _abi.Rate2D, "cb2_rate3d": _abi.Rate3D, "cb2_model_ext": _abi.ModelExt, "cb2_beam_rate": _abi.BeamRate, "cb2_cx_rate": _abi.CXRate, "cb2_beam_desc": _abi.BeamDesc, "cb2_gaunt": _abi.Gaunt, "cb2_lineshape": _abi.LineShape, "cb2_model": _abi.ModelDesc,
               "cb2_scene_desc": _abi.SceneDesc, "cb2_rays": _abi.Rays, "cb2_stats": _abi.Stats, "cb2_rt_desc": _abi.RTDesc, "cb2_sart_desc": _abi.SartDesc, "cb2_primitive": _abi.PrimitiveDesc, "cb2_pinhole": _abi.PinholeDesc}
    src = '#include <stdio.h>\n#include "cherab_b200.h"\nint main(){' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    sizes = dict(l.split() for l in out if l)
    for n, t in structs.items():
        assert int(sizes[n]) == ctypes.sizeof(t), n


def test_no_cpu_fallback_without_gpu():
    """On a box without a usable CUDA device the product must raise, never compute."""
    import pytest
    lib = _abi.load_library()
    if lib.cb2_device_count() > 0:
        pytest.skip("GPU present")
    import core_b200 as cb
    from core_b200.engine import EmissionScene
    from core_b200.slab import build_constant_slab_plasma
    plasma = build_constant_slab_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    plasma.models = [cb.RecombinationLine(cb.Line(cb.hydrogen, 0, (3, 2)))]
    flat = cb.flatten_scene(plasma, 650, 660, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EmissionScene(flat)


def test_product_does_not_import_oracle():
    pkg = os.path.dirname(core_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text and "cb2o_" not in text, fn
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        if fn.endswith((".cu", ".h")):
            assert "cb2o_" not in open(os.path.join(pkg, "csrc", fn)).read(), fn


def test_rows_plan_covers_every_ray_once():
    # cb2_rows_plan: the strided-copy plan of cb2_emission_render_rows (host logic, no device): a rank's tile-ordered rays become one
    # 2-D copy per 16 x 16 tile; any other row list decomposes into runs that cover every ray exactly once
    import ctypes as C
    import numpy as np
    from core_b200 import _abi
    from core_b200.sharding import tile_pixels
    lib = _abi.load_library()

    def plan(rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.zeros(4 * max(rows.size, 1), dtype=np.int64)
        k = lib.cb2_rows_plan(rows.ctypes.data_as(_abi.c_int64_p), rows.size, out.ctypes.data_as(_abi.c_int64_p), rows.size)
        ops = out[:4 * k].reshape(k, 4)
        # replay: copy k moves rays [first + r len, first + (r + 1) len) to rows dest[first] + r pitch + [0, len)
        seen = np.full(rows.size, -1, dtype=np.int64)
        for first, length, repeats, pitch in ops:
            for r in range(repeats):
                idx = first + r * length + np.arange(length)
                assert np.all(seen[idx] == -1)
                seen[idx] = rows[first] + r * pitch + np.arange(length)
        assert np.array_equal(seen, rows)
        return ops

    ops = plan(tile_pixels((64, 48), 1, 3))                       # 4 x 3 tiles dealt to 3 ranks: rank 1 holds the middle tile of
    assert ops.tolist() == [[0, 16, 64, 48]]                      # every tile row — one pitch throughout, a single 2-D copy
    ops = plan(tile_pixels((64, 48), 1, 4))                       # rank 1 of 4: tiles (0, 1), (1, 2), (3, 0)
    assert len(ops) == 3 and np.all(ops[:, 1] == 16) and np.all(ops[:, 2] == 16) and np.all(ops[:, 3] == 48)
    ops = plan(tile_pixels((40, 24), 0, 1))                       # ragged tiles at the frame's edge (40 = 2.5 tiles, 24 = 1.5 tiles)
    assert ops[:, 1].max() == 16 and ops[:, 1].min() == 8
    assert len(plan(np.arange(1000))) == 1                        # one contiguous run
    rng = np.random.default_rng(11)
    plan(rng.permutation(5000)[:700])                             # single rows
    plan(np.concatenate([np.arange(100, 137), np.arange(10, 12), [5], np.arange(400, 464), np.arange(300, 364), np.arange(200, 264), [1599]]))
    assert lib.cb2_rows_plan(None, 0, None, 0) == 0
