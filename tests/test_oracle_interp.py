"""Properties of the restated Raysect cubic interpolators (SURVEY Appendix B.4 — parity unpinned: the reference holds
no golden vectors for them, so these tests pin the restated formula against independent closed forms)."""
import numpy as np

from oracle import oracle


def catmull_rom(f0, f1, f2, f3, t):
    return 0.5 * (2 * f1 + (-f0 + f2) * t + (2 * f0 - 5 * f1 + 4 * f2 - f3) * t ** 2 + (-f0 + 3 * f1 - 3 * f2 + f3) * t ** 3)


def test_1d_uniform_interior_is_catmull_rom():
    rng = np.random.default_rng(1)
    x = np.linspace(0, 1, 11)
    f = rng.normal(size=11)
    for i in range(1, 9):
        for t in (0.0, 0.13, 0.5, 0.99):
            got = oracle.interp1d_cubic(x, f, x[i] + t * 0.1)
            assert abs(got - catmull_rom(f[i - 1], f[i], f[i + 1], f[i + 2], t)) < 1e-12


def test_1d_nonuniform_reproduces_quadratics_and_knots():
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(0, 5, 20))
    q = lambda v: 1.5 - 2 * v + 0.7 * v * v
    f = q(x)
    for px in rng.uniform(x[1], x[-2], 50):   # interior cells: 3-point derivative is exact for quadratics
        assert abs(oracle.interp1d_cubic(x, f, px) - q(px)) < 1e-11
    g = rng.normal(size=20)
    for i in range(20):
        assert abs(oracle.interp1d_cubic(x, g, x[i]) - g[i]) < 1e-13
    # 'nearest' extrapolation clamps
    assert oracle.interp1d_cubic(x, g, x[0] - 1) == g[0] and oracle.interp1d_cubic(x, g, x[-1] + 1) == g[-1]


def test_1d_c1_continuity():
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 1, 12))
    f = rng.normal(size=12)
    for i in range(1, 11):
        e = 1e-5 * min(x[i] - x[i - 1], x[i + 1] - x[i])
        p = lambda v: oracle.interp1d_cubic(x, f, v)
        dl = (3 * p(x[i]) - 4 * p(x[i] - e) + p(x[i] - 2 * e)) / (2 * e)   # 2nd-order one-sided differences
        dr = (-3 * p(x[i]) + 4 * p(x[i] + e) - p(x[i] + 2 * e)) / (2 * e)
        assert abs(dl - dr) < 1e-5 * max(1.0, abs(dl))


def test_2d_separable_and_bilinear_terms():
    rng = np.random.default_rng(4)
    x = np.sort(rng.uniform(0, 3, 9))
    y = np.sort(rng.uniform(-1, 2, 8))
    # f = a + b x + c y + d x y + e x^2 + g y^2 is reproduced in interior cells (2nd-order derivative stencils,
    # four-corner cross derivative exact for xy)
    fn = lambda X, Y: 0.3 + 1.2 * X - 0.7 * Y + 0.9 * X * Y + 0.5 * X * X - 0.4 * Y * Y
    F = fn(x[:, None], y[None, :])
    for _ in range(60):
        px, py = rng.uniform(x[1], x[-2]), rng.uniform(y[1], y[-2])
        assert abs(oracle.interp2d_cubic(x, y, F, px, py) - fn(px, py)) < 1e-11
    G = rng.normal(size=(9, 8))
    for i in range(9):
        for j in range(8):
            assert abs(oracle.interp2d_cubic(x, y, G, x[i], y[j]) - G[i, j]) < 1e-13


def test_2d_uniform_matches_tensor_catmull_rom():
    rng = np.random.default_rng(5)
    x = np.arange(8) * 0.03
    y = np.arange(7) * 0.05
    F = rng.normal(size=(8, 7))
    i, j, t, u = 3, 2, 0.37, 0.81
    rows = [catmull_rom(F[i - 1 + a, j - 1], F[i - 1 + a, j], F[i - 1 + a, j + 1], F[i - 1 + a, j + 2], u) for a in range(4)]
    ref = catmull_rom(rows[0], rows[1], rows[2], rows[3], t)
    assert abs(oracle.interp2d_cubic(x, y, F, x[i] + t * 0.03, y[j] + u * 0.05) - ref) < 1e-12
