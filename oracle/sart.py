"""CPU restatement of the reference's SART inversion — TEST INFRASTRUCTURE, never on the product path.

Follows cherab/tools/inversions/sart.pyx:26-155 (invert_sart) and :161-302 (invert_constrained_sart) in numpy float64:
the per-cell loop over detectors (sart.pyx:118-142) is the column sum written as one matrix-vector product.  Pinned on
outputs of the reference's own Cython module run in the build container on the reference's fixtures
(tests/golden/make_sart_golden.py -> tests/golden/sart_golden.npz) and on the reference's acceptance test
(cherab/tools/tests/test_sart_opencl.py:57-80: |solution - true_emissivity| <= 1e-2).

``build_ref()`` compiles the UNMODIFIED reference module from where it lies under /root/reference into oracle/_ref/
(cython -> gcc, the module only needs numpy); ``ref_module()`` imports it when present.
"""
import os
import subprocess
import sys
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_REF_SRC = "/root/reference/cherab/tools/inversions/sart.pyx"


def invert_sart(geometry_matrix, measurement_vector, initial_guess=None, max_iterations=250, relaxation=1.0, conv_tol=1.0e-4,
                laplacian_matrix=None, beta_laplace=0.01):
    g = np.asarray(geometry_matrix, dtype=np.float64)
    m = np.asarray(measurement_vector, dtype=np.float64)
    n_sources = g.shape[1]
    if initial_guess is None:                                   # sart.pyx:88-93
        x = np.zeros(n_sources) + np.exp(-1)
    elif isinstance(initial_guess, (float, int)):
        x = np.zeros(n_sources) + initial_guess
    else:
        x = np.array(initial_guess, dtype=np.float64)
    density = g.sum(axis=0)                                     # A_(+,j)  sart.pyx:107-108
    length = g.sum(axis=1)                                      # A_(i,+)  sart.pyx:111-113
    with np.errstate(divide="ignore"):
        inv_length = np.where(length == 0, 0.0, 1.0 / length)   # rays of zero length are skipped, sart.pyx:127-128
    seen = density > 0.0
    gain = np.where(seen, relaxation / np.where(seen, density, 1.0), 0.0)
    y_hat = g @ x
    m_sq = m @ m
    convergence = []
    for k in range(max_iterations):
        penalty = beta_laplace * (np.asarray(laplacian_matrix, dtype=np.float64) @ x) if laplacian_matrix is not None else 0.0   # sart.pyx:255
        x = x + gain * (g.T @ (inv_length * (m - y_hat))) - penalty     # sart.pyx:118-134 / :258-281
        x = np.where(x < 0, 0.0, x)                                     # sart.pyx:137-138
        y_hat = g @ x
        convergence.append((m_sq - y_hat @ y_hat) / m_sq)               # sart.pyx:144-148
        if k > 0 and abs(convergence[k] - convergence[k - 1]) < conv_tol:
            break
    return x, convergence


def invert_constrained_sart(geometry_matrix, laplacian_matrix, measurement_vector, initial_guess=None, max_iterations=250,
                            relaxation=1.0, beta_laplace=0.01, conv_tol=1.0e-4):
    return invert_sart(geometry_matrix, measurement_vector, initial_guess, max_iterations, relaxation, conv_tol,
                       laplacian_matrix=laplacian_matrix, beta_laplace=beta_laplace)


def _ref_path():
    return os.path.join(_REF_DIR, "cherab_ref_sart" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ref():
    """oracle/_ref/cherab_ref_sart*.so from the reference's sart.pyx, if /root/reference is present.  Returns the path or None."""
    out = _ref_path()
    if os.path.exists(out):
        return out
    if not os.path.exists(_REF_SRC):
        return None
    import numpy
    os.makedirs(_REF_DIR, exist_ok=True)
    c_file = os.path.join(_REF_DIR, "cherab_ref_sart.c")
    # the C file is generated straight from the reference source into the git-ignored oracle/_ref/; module name = file name
    pyx = os.path.join(_REF_DIR, "cherab_ref_sart.pyx")
    if os.path.lexists(pyx):
        os.remove(pyx)
    os.symlink(_REF_SRC, pyx)
    try:
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", c_file], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
                               c_file, "-o", out])
    finally:
        os.remove(pyx)
        if os.path.exists(c_file):
            os.remove(c_file)
    return out


def ref_module():
    """The compiled reference module (invert_sart, invert_constrained_sart) or None when it was never built."""
    path = _ref_path()
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("cherab_ref_sart", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
