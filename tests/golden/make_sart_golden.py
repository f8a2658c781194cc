"""Generates tests/golden/sart_golden.npz by running the REFERENCE's own SART (cherab/tools/inversions/sart.pyx, compiled
unmodified into oracle/_ref by oracle.sart.build_ref) on the reference's fixtures cherab/tools/tests/data/*.npy
(the inputs of cherab/tools/tests/test_sart_opencl.py:43-52).  Run in the build container only: needs /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sart  # noqa: E402

DATA = "/root/reference/cherab/tools/tests/data"


def laplacian_11x8():
    """5-point isotropic Laplacian (C x_l - sum of neighbours, sart.pyx:178-180) on the fixture's 11 x 8 emissivity grid."""
    n0, n1 = 11, 8
    lap = np.zeros((n0 * n1, n0 * n1))
    for a in range(n0):
        for b in range(n1):
            i = a * n1 + b
            for da, db in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                if 0 <= a + da < n0 and 0 <= b + db < n1:
                    lap[i, i] += 1
                    lap[i, (a + da) * n1 + b + db] = -1
    return lap


def main():
    sart.build_ref()
    ref = sart.ref_module()
    gm = np.load(os.path.join(DATA, "geometry_matrix.npy"))
    gm = gm.reshape(gm.shape[0] * gm.shape[1], gm.shape[2])
    receiver = np.load(os.path.join(DATA, "receiver.npy")).flatten()
    truth = np.load(os.path.join(DATA, "true_emissivity.npy")).flatten()
    g64, m64 = gm.astype(np.float64), receiver.astype(np.float64)
    out = {"geometry_matrix": gm, "receiver": receiver, "true_emissivity": truth, "laplacian": laplacian_11x8()}
    cases = {
        "plain": dict(),
        "relaxed": dict(relaxation=0.6, initial_guess=0.25, conv_tol=1e-6),
        "capped": dict(max_iterations=5),
    }
    for name, kw in cases.items():
        s, c = ref.invert_sart(g64, m64, **kw)
        out["sol_" + name], out["conv_" + name] = np.array(s), np.array(c)
    s, c = ref.invert_constrained_sart(g64, np.identity(g64.shape[1]), m64, beta_laplace=0.001)      # test_sart_opencl.py:72-80
    out["sol_identity"], out["conv_identity"] = np.array(s), np.array(c)
    s, c = ref.invert_constrained_sart(g64, out["laplacian"], m64, beta_laplace=0.01, conv_tol=1e-5)
    out["sol_laplace"], out["conv_laplace"] = np.array(s), np.array(c)
    rng = np.random.default_rng(7)
    guess = rng.uniform(0.0, 1.0, g64.shape[1])
    out["guess_array"] = guess
    s, c = ref.invert_sart(g64, m64 * 3.0, initial_guess=guess.copy(), conv_tol=1e-5)
    out["sol_array_guess"], out["conv_array_guess"] = np.array(s), np.array(c)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sart_golden.npz"), **out)
    for k in out:
        print(k, out[k].shape)


if __name__ == "__main__":
    main()
