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
