"""The Cython side of the boundary (integration/): shim.pyx — B200Integrator(VolumeIntegrator) and the batch call — cythonized and
compiled against include/cherab_b200.h and the Raysect stand-in, imported, and (on the GPU) called the way Raysect calls a
VolumeIntegrator: start_point = far end, end_point = near end, `spectrum` is added to."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import build_shim
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "core_b200", "csrc", "libcherab_b200.so")):
        g.build()
    out = build_shim.build(str(tmp_path_factory.mktemp("shim")))
    sys.path.insert(0, out)
    for m in [k for k in sys.modules if k == "raysect" or k.startswith("raysect.")]:
        del sys.modules[m]
    from cherab_b200_shim import shim as mod
    yield mod
    sys.path.remove(out)
    for m in [k for k in sys.modules if k == "raysect" or k.startswith("raysect.") or k.startswith("cherab_b200_shim")]:
        del sys.modules[m]


def test_shim_compiles_imports_and_binds_the_library(shim):
    from core_b200 import _abi
    from raysect.optical.material.emitter.inhomogeneous import VolumeIntegrator
    assert shim.abi_version() == _abi.ABI_VERSION
    assert issubclass(shim.B200Integrator, VolumeIntegrator)


@pytest.mark.gpu
def test_b200_integrator_is_called_like_a_volume_integrator(shim):
    import core_b200 as cb
    from core_b200 import generomak
    from core_b200.engine import EmissionScene
    from raysect.optical import AffineMatrix3D, Point3D, Primitive, Ray, Spectrum, World
    from raysect.optical.material.emitter.inhomogeneous import InhomogeneousVolumeEmitter
    plasma = generomak.get_plasma()
    plasma.atomic_data = cb.SyntheticADAS()
    line = cb.Line(cb.hydrogen, 0, (3, 2))
    plasma.models = [cb.ExcitationLine(line), cb.RecombinationLine(line)]
    plasma.integrator = cb.NumericalIntegrator(step=0.005)
    flat = cb.flatten_scene(plasma, 655.0, 657.5, 128)
    cam = cb.PinholeCamera((3, 3), fov=30.0, transform=cb.look_at((2.3, 0.0, 1.25), (1.0, 0.8, -0.5)))
    rays = cb.ray_segments(plasma.geometry, *cam.rays(), plasma.geometry_to_world())
    scene = EmissionScene(flat)
    ref, st = scene.render(rays)
    scene.close()
    integ = shim.B200Integrator(flat, C.addressof(flat.desc), 0)
    material = InhomogeneousVolumeEmitter(integ)
    eye = AffineMatrix3D()
    for r in range(rays.n_rays):
        spectrum = Spectrum(655.0, 657.5, 128)
        spectrum.samples[:] = 1.0                                   # integrate() adds to what the spectrum holds
        o, d = rays.origin[r], rays.direction[r]
        for s in range(rays.seg_offset[r], rays.seg_offset[r + 1]):
            near, far = o + rays.seg_t0[s] * d, o + rays.seg_t1[s] * d
            out = material.integrator.integrate(spectrum, World(), Ray(Point3D(*o)), Primitive(), material, Point3D(*far), Point3D(*near), eye, eye)
            assert out is spectrum
        got = np.asarray(spectrum.samples) - 1.0
        assert np.all(np.abs(got - ref[r]) <= 1e-6 * np.abs(ref[r]) + 1e-9 * ref[r].max() + 1e-12)
    # the batch call (seam S4) and the exception types
    out = np.zeros((rays.n_rays, 128))
    n = integ.render_segments(rays.origin, rays.direction, rays.seg_offset, rays.seg_t0, rays.seg_t1, out)
    assert n == st["samples"] and np.array_equal(out, ref)
    with pytest.raises(ValueError):
        integ.render_segments(rays.origin, rays.direction[:-1], rays.seg_offset, rays.seg_t0, rays.seg_t1, out)
    with pytest.raises(ValueError):
        bad = rays.seg_offset.copy()
        bad[1] = -3
        integ.render_segments(rays.origin, rays.direction, bad, rays.seg_t0, rays.seg_t1, out)
