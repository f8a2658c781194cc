"""Cythonize and compile integration/cherab_b200_shim against the Raysect stand-in (integration/stubs) and include/cherab_b200.h.

    python integration/build_shim.py [build_dir]      -> build_dir holds the importable packages `raysect` (stub) and `cherab_b200_shim`

With a real Raysect installed, drop `stubs` from the include path: shim.pyx compiles unchanged.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(out_dir):
    import numpy
    os.makedirs(out_dir, exist_ok=True)
    for pkg in ("stubs/raysect", "cherab_b200_shim"):
        dst = os.path.join(out_dir, os.path.basename(pkg))
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(HERE, pkg), dst)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-I" + os.path.join(ROOT, "include")]
    lib_dir = os.path.join(ROOT, "core_b200", "csrc")
    mods = [("raysect/optical/_stub.pyx", "raysect/optical/_stub", []),
            ("raysect/optical/material/emitter/inhomogeneous.pyx", "raysect/optical/material/emitter/inhomogeneous", []),
            ("cherab_b200_shim/shim.pyx", "cherab_b200_shim/shim", ["-L" + lib_dir, "-lcherab_b200", "-Wl,-rpath," + lib_dir])]
    for src, stem, link in mods:
        subprocess.run([sys.executable, "-m", "cython", "-3", "-I", out_dir, os.path.join(out_dir, src)], check=True, cwd=out_dir)
        c = os.path.join(out_dir, src[:-4] + ".c")
        subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", *inc, c, "-o",
                        os.path.join(out_dir, stem + ext), *link], check=True)
    return out_dir


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_build")))
