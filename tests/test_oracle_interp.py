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


def test_beam_rate_extrapolation_rules():
    """'quadratic' (1-D) and 'linear' (2-D) extrapolation as the oracle restates them for the beam tables (beam.pyx:73-84): reached
    through the beam-density probe of a one-species slab whose stopping table is a power law in every argument — log-log linear
    data, which both rules continue exactly — so the attenuation outside the table equals the analytic power law."""
    import core_b200 as cb
    from oracle import oracle
    from test_oracle_beam import beam_scene

    class PowerLawADAS(cb.AtomicData):
        def __init__(self, extrapolate):
            self.extrapolate = extrapolate

        def beam_stopping_rate(self, beam_ion, plasma_ion, charge):
            e, n, t = np.logspace(4.8, 5.5, 6), np.logspace(19.5, 21.0, 7), np.logspace(3.5, 4.5, 5)     # none contains the slab's state
            sref = 1e-13
            sen = sref * (e[:, None] / 1e5) ** -0.4 * (n[None, :] / 1e20) ** 0.1
            st = sref * (t / 1e4) ** 0.05
            return cb.BeamStoppingTable(e, n, t, sen, st, sref, extrapolate=self.extrapolate)

    def density_on_axis(extrapolate):
        plasma, beam = beam_scene(1e-13, sigma=0.2, divergence_x=0.0, divergence_y=0.0, length=10.0)
        beam.atomic_data = PowerLawADAS(extrapolate)
        flat = cb.flatten_beam_scene(beam, 655.1, 657.1, 16)
        z = np.linspace(0.5, 9.5, 10)
        d, _ = oracle.beam_sample(flat, np.stack([0 * z, 0 * z, z], axis=1))
        return plasma, beam, z, d

    plasma, beam, z, d = density_on_axis(True)
    sp = [s for s in plasma.composition if s.charge > 0][0]
    ni, ti = sp.distribution.density.value, sp.distribution.temperature.value
    energy = beam.energy                                    # the slab species are at rest in the reference's test scene
    rate = 1e-13 * (energy / 1e5) ** -0.4 * ((sp.charge ** 2 * ni / sp.charge) / 1e20) ** 0.1 * (ti / 1e4) ** 0.05
    speed = np.sqrt(2 * energy * 1.602176634e-19 / 1.66053906660e-27)
    ratio = d[1:] / d[:-1]
    expect = np.exp(-ni * sp.charge * rate * np.diff(z) / speed)
    assert np.allclose(ratio, expect, rtol=1e-9), (ratio, expect)
    _, _, _, dc = density_on_axis(False)                    # clamped tables attenuate differently
    assert not np.allclose(dc[1:] / dc[:-1], expect, rtol=1e-4)
