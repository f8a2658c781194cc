"""The SART oracle (oracle/sart.py) against outputs of the reference's own Cython module on the reference's fixtures
(tests/golden/sart_golden.npz, made by tests/golden/make_sart_golden.py) and the reference's acceptance criterion."""
import os

import numpy as np
import pytest

from oracle import sart

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sart_golden.npz"))
G = GOLD["geometry_matrix"].astype(np.float64)
M = GOLD["receiver"].astype(np.float64)

CASES = {
    "plain": (lambda f: f(G, M), None),
    "relaxed": (lambda f: f(G, M, relaxation=0.6, initial_guess=0.25, conv_tol=1e-6), None),
    "capped": (lambda f: f(G, M, max_iterations=5), None),
    "array_guess": (lambda f: f(G, M * 3.0, initial_guess=GOLD["guess_array"].copy(), conv_tol=1e-5), None),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_run(name):
    sol, conv = CASES[name][0](sart.invert_sart)
    assert len(conv) == len(GOLD["conv_" + name])
    np.testing.assert_allclose(sol, GOLD["sol_" + name], rtol=1e-11, atol=1e-13 * GOLD["sol_" + name].max())
    np.testing.assert_allclose(conv, GOLD["conv_" + name], rtol=1e-9, atol=1e-13)


def test_oracle_constrained_matches_reference_run():
    sol, conv = sart.invert_constrained_sart(G, np.identity(G.shape[1]), M, beta_laplace=0.001)
    assert len(conv) == len(GOLD["conv_identity"])
    np.testing.assert_allclose(sol, GOLD["sol_identity"], rtol=1e-11, atol=1e-13 * GOLD["sol_identity"].max())
    sol, conv = sart.invert_constrained_sart(G, GOLD["laplacian"], M, beta_laplace=0.01, conv_tol=1e-5)
    assert len(conv) == len(GOLD["conv_laplace"])
    np.testing.assert_allclose(sol, GOLD["sol_laplace"], rtol=1e-11, atol=1e-13 * GOLD["sol_laplace"].max())


def test_reference_acceptance_criterion():
    # cherab/tools/tests/test_sart_opencl.py:57-80
    sol, _ = sart.invert_sart(G, M)
    assert np.allclose(sol, GOLD["true_emissivity"], atol=1e-2)
    sol, _ = sart.invert_constrained_sart(G, np.identity(G.shape[1]), M, beta_laplace=0.001)
    assert np.allclose(sol / sol.max(), GOLD["true_emissivity"], atol=1e-2)


def test_zero_length_rays_and_unseen_cells_are_skipped():
    # sart.pyx:127-128 (ray length 0 -> skipped) and :136-137 (cell seen by no ray keeps its value)
    rng = np.random.default_rng(3)
    g = rng.uniform(0, 1, (40, 12)) * (rng.uniform(0, 1, (40, 12)) < 0.4)
    g[5] = 0.0
    g[:, 7] = 0.0
    x_true = rng.uniform(0.5, 2, 12)
    m = g @ x_true
    m[5] = 3.0                          # a measurement on a ray that crosses nothing must not matter
    sol, conv = sart.invert_sart(g, m, initial_guess=0.3, max_iterations=40, conv_tol=0)
    assert sol[7] == 0.3 and len(conv) == 40 and np.all(np.isfinite(sol))


@pytest.mark.skipif(sart.ref_module() is None and not os.path.exists("/root/reference"), reason="reference module not built on this box")
def test_oracle_matches_compiled_reference_on_random_input():
    sart.build_ref()
    ref = sart.ref_module()
    rng = np.random.default_rng(11)
    g = rng.uniform(0, 1, (200, 60)) * (rng.uniform(0, 1, (200, 60)) < 0.2)
    m = g @ rng.uniform(0, 3, 60)
    lap = np.diag(np.full(60, 2.0)) - np.diag(np.ones(59), 1) - np.diag(np.ones(59), -1)
    s0, c0 = ref.invert_constrained_sart(g, lap, m, beta_laplace=0.02, relaxation=0.8)
    s1, c1 = sart.invert_constrained_sart(g, lap, m, beta_laplace=0.02, relaxation=0.8)
    assert len(c0) == len(c1)
    np.testing.assert_allclose(s1, np.array(s0), rtol=1e-10, atol=1e-13)
