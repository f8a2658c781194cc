"""SART on the CUDA path (cb2_sart_* through the C ABI) against the reference's own outputs (tests/golden/sart_golden.npz),
the oracle on seeded inputs, and size-independent properties."""
import os

import numpy as np
import pytest

import core_b200 as cb
from oracle import sart

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sart_golden.npz"))
G = GOLD["geometry_matrix"].astype(np.float64)
M = GOLD["receiver"].astype(np.float64)
RTOL = 1e-9          # float64 on both sides; only the order of the sums differs


def close(got, ref, rtol=RTOL):
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max())


def test_reference_fixture_plain_relaxed_capped():
    with cb.SartSolver(G) as solver:
        for name, kw in (("plain", {}), ("relaxed", dict(relaxation=0.6, initial_guess=0.25, conv_tol=1e-6)), ("capped", dict(max_iterations=5))):
            sol, conv = solver(M, **kw)
            assert len(conv) == len(GOLD["conv_" + name]), name
            close(sol, GOLD["sol_" + name])
            close(np.array(conv), GOLD["conv_" + name])
        sol, conv = solver(M * 3.0, initial_guess=GOLD["guess_array"], conv_tol=1e-5)
        assert len(conv) == len(GOLD["conv_array_guess"])
        close(sol, GOLD["sol_array_guess"])
    # the reference's acceptance test, cherab/tools/tests/test_sart_opencl.py:57-62
    sol, _ = cb.invert_sart(G, M)
    assert np.allclose(sol, GOLD["true_emissivity"], atol=1e-2)


def test_reference_fixture_constrained():
    sol, conv = cb.invert_constrained_sart(G, np.identity(G.shape[1]), M, beta_laplace=0.001)
    assert len(conv) == len(GOLD["conv_identity"])
    close(sol, GOLD["sol_identity"])
    assert np.allclose(sol / sol.max(), GOLD["true_emissivity"], atol=1e-2)      # test_sart_opencl.py:72-80
    with cb.SartSolver(G, laplacian_matrix=np.identity(G.shape[1])) as solver:
        solver.update_laplacian_matrix(GOLD["laplacian"])
        sol, conv = solver(M, beta_laplace=0.01, conv_tol=1e-5)
        assert len(conv) == len(GOLD["conv_laplace"])
        close(sol, GOLD["sol_laplace"])
        assert solver.info()["laplacian_nnz"] == np.count_nonzero(GOLD["laplacian"])


def test_float32_fixture_and_float32_storage():
    # the fixture as shipped (float32, what SartOpencl takes) gives the same answer as its float64 copy
    sol, conv = cb.invert_sart(GOLD["geometry_matrix"], GOLD["receiver"])
    assert len(conv) == len(GOLD["conv_plain"])
    close(sol, GOLD["sol_plain"])
    with cb.SartSolver(G, value_dtype=np.float32) as solver:
        sol, _ = solver(M)
        close(sol, GOLD["sol_plain"])
        assert solver.info()["bytes_per_iteration"] < 2 * np.count_nonzero(G) * 8.5 + 8 * sum(G.shape) + 64


def _random_problem(seed, n_det, n_src, fill):
    rng = np.random.default_rng(seed)
    g = rng.uniform(0, 1, (n_det, n_src)) * (rng.uniform(0, 1, (n_det, n_src)) < fill)
    g[n_det // 3] = 0.0                                      # a ray that crosses nothing
    g[:, n_src // 2] = 0.0                                   # a cell no ray sees
    m = g @ rng.uniform(0, 3, n_src)
    m[n_det // 3] = 1.5
    return g, m


def test_csr_input_frames_and_edge_cases_against_oracle():
    g, m = _random_problem(5, 700, 300, 0.08)
    frames = np.stack([m, 0.5 * m, m[::-1].copy(), 2.0 * m, m + 0.01, 0.1 * m])       # 6 frames: one full group of 4 and a ragged one
    row_offset = np.concatenate([[0], np.cumsum((g != 0).sum(axis=1))])
    columns = np.nonzero(g)[1].astype(np.int32)
    values = g[g != 0]
    lap = np.diag(np.full(300, 2.0)) - np.diag(np.ones(299), 1) - np.diag(np.ones(299), -1)
    lap_csr = (np.concatenate([[0], np.cumsum((lap != 0).sum(axis=1))]), np.nonzero(lap)[1], lap[lap != 0])
    with cb.SartSolver.from_csr(row_offset, columns, values, 300, laplacian_matrix=lap_csr) as solver:
        assert solver.info()["nnz"] == values.size
        sols, convs = solver(frames, beta_laplace=0.02, relaxation=0.8, initial_guess=0.3, conv_tol=1e-5)
        for f in range(frames.shape[0]):
            ref, rconv = sart.invert_constrained_sart(g, lap, frames[f], beta_laplace=0.02, relaxation=0.8, initial_guess=0.3, conv_tol=1e-5)
            assert len(convs[f]) == len(rconv), f
            close(sols[f], ref)
            close(np.array(convs[f]), np.array(rconv), rtol=1e-7)
        # a frame inverted alone is bit-identical to the same frame inverted in a group (fixed summation order)
        one, _ = solver(frames[2], beta_laplace=0.02, relaxation=0.8, initial_guess=0.3, conv_tol=1e-5)
        assert np.array_equal(one, sols[2])
    with cb.SartSolver(g) as solver:
        sol, conv = solver(m, initial_guess=0.3, max_iterations=40, conv_tol=0)
        assert sol[150] == 0.3 and len(conv) == 40               # unseen cell untouched, no early stop with conv_tol = 0
        ref, _ = sart.invert_sart(g, m, initial_guess=0.3, max_iterations=40, conv_tol=0)
        close(sol, ref)


def test_bad_arguments_raise():
    with pytest.raises(ValueError):
        cb.SartSolver(np.zeros(5))
    with pytest.raises(ValueError):
        cb.SartSolver(G, laplacian_matrix=np.identity(3))
    with cb.SartSolver(G) as solver:
        with pytest.raises(ValueError):
            solver(M[:-1])
        # no iteration: the reference's loop does not run and the seed comes back with an empty convergence list (sart.pyx:103-152)
        x0, conv0 = solver(M, max_iterations=0, initial_guess=0.25)
        assert conv0 == [] and np.array_equal(x0, np.full(G.shape[1], 0.25))


def test_geometry_matrix_from_the_ray_transfer_kernel_stays_on_the_device():
    """C4-shaped chain at a small size: ray-transfer CSR built on the device -> SART without a host round trip; a known
    emissivity is recovered from its own projections (round-trip property)."""
    import torch
    from core_b200.engine import DeviceRays, RayTransferScene
    from core_b200.raytransfer import RayTransferCylinder
    rtc = RayTransferCylinder(radius_outer=2.0, height=2.0, n_radius=12, n_height=12, radius_inner=1.0, transform=cb.translate(0, 0, -1.0))
    prim = cb.HollowCylinder(1.0, 2.0, -1.0, 1.0)
    rays = []
    for pos in ((3.5, 0.0, 0.0), (0.0, 3.5, 0.8), (-3.5, 0.2, -0.7), (2.5, 2.5, 1.5), (0.5, -3.5, 0.3)):
        cam = cb.PinholeCamera((24, 24), fov=50.0, transform=cb.look_at(pos, (0.0, 0.0, 0.0)))
        o, d = cam.rays()
        rays.append((o, d))
    o = np.concatenate([r[0] for r in rays]); d = np.concatenate([r[1] for r in rays])
    batch = cb.ray_segments(prim, o, d)
    scene = RayTransferScene(rtc)
    dev = DeviceRays(batch)
    row_offset, columns, lengths = scene.render_csr_device(dev, capacity=batch.n_rays * 64)
    h_ro, h_co, h_le, _ = scene.render_csr(batch)
    scene.close()
    dense = np.zeros((batch.n_rays, rtc.bins))
    for r in range(batch.n_rays):
        dense[r, h_co[h_ro[r]:h_ro[r + 1]]] = h_le[h_ro[r]:h_ro[r + 1]]
    rr, zz = np.meshgrid(np.arange(12), np.arange(12), indexing="ij")
    truth = (1.0 + np.exp(-((rr - 6.0) ** 2 + (zz - 5.0) ** 2) / 8.0)).reshape(-1)
    m = dense @ truth
    with cb.SartSolver.from_device_csr(row_offset, columns, lengths, rtc.bins) as solver:
        assert solver.info()["nnz"] == h_co.size
        sol, conv = solver(m, max_iterations=2000, conv_tol=1e-12)
    ref, rconv = sart.invert_sart(dense, m, max_iterations=2000, conv_tol=1e-12)
    assert len(conv) == len(rconv)
    close(sol, ref, rtol=1e-8)
    seen = dense.sum(axis=0) > 0
    assert np.abs(dense @ sol - m).max() <= 2e-3 * m.max() and seen.sum() > 100
