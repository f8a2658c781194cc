"""Prototype behind the Bremsstrahlung moment formulation: accuracy of 4-point Lagrange interpolation of
Phi_z(T; bin) = g_ff(Z, T, lambda) lambda^-2 exp(-(hc/lambda - x_ref)/T) on temperature nodes uniform in
s = ln(tau) + tau/tau_c, tau = 1/T, against direct evaluation.  The worst case sits at the second-derivative kinks of the
reference's bicubic Gaunt table (u = 1 crossings at 2-3 eV in the visible).  Uses the oracle's Gaunt factor (test
infrastructure) — this is a design tool, not product code.

    python tools/proto_brems_moments.py        # prints nodes and worst relative error for several (tau_c, ds)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core_b200 as cb                      # noqa: E402
from core_b200 import _abi                  # noqa: E402
from oracle import oracle                   # noqa: E402

HC = 1239.8419738620933
lib = oracle.lib()
u, g2, gff = (np.ascontiguousarray(a, dtype=np.float64) for a in cb.AtomicData().free_free_gaunt_factor())
G = _abi.Gaunt(u.size, g2.size, u.ctypes.data_as(_abi.c_double_p), g2.ctypes.data_as(_abi.c_double_p), gff.ctypes.data_as(_abi.c_double_p))
lam = np.linspace(390, 700, 2049)
lamc = (0.5 * (lam[1:] + lam[:-1]))[::64]
x_ref = 0.5 * (HC / 390.0 + HC / 700.0)


def phi(z, te):
    return np.array([lib.cb2o_gaunt_factor(C.byref(G), float(z), float(te), float(w)) * np.exp(-(HC / w - x_ref) / te) / w ** 2 for w in lamc])


def nodes(tau_c, ds, te_min=0.14, te_max=3100.0):
    smap = lambda tau: np.log(tau) + tau / tau_c
    s0, s1 = smap(1 / te_max), smap(1 / te_min)
    s = s0 - ds + ds * np.arange(int(np.ceil((s1 - s0) / ds)) + 4)
    tau = np.exp(np.minimum(s, np.log(tau_c)))
    for _ in range(60):
        tau = np.maximum(tau - (np.log(tau) + tau / tau_c - s) / (1 / tau + 1 / tau_c), 1e-12)
    return s, tau, smap


def lagrange(t):
    return np.array([-t * (t - 1) * (t - 2) / 6, (t + 1) * (t - 1) * (t - 2) / 2, -(t + 1) * t * (t - 2) / 2, (t + 1) * t * (t - 1) / 6])


if __name__ == "__main__":
    for tau_c, ds in [(3.0, 0.05), (3.0, 0.07), (4.0, 0.06)]:
        s, tau, smap = nodes(tau_c, ds)
        worst = (0.0, 0.0, 0)
        for z in (1, 3, 6):
            tab = np.array([phi(z, 1 / t) for t in tau])
            for te in np.exp(np.random.default_rng(0).uniform(np.log(0.145), np.log(3000.0), 600)):
                f = (float(smap(1 / te)) - s[0]) / ds
                i = int(f)
                err = np.max(np.abs(lagrange(f - i) @ tab[i - 1:i + 3] / phi(z, te) - 1))
                worst = max(worst, (err, te, z))
        print("tau_c %.1f ds %.2f nodes %d worst rel err %.3g at Te = %.3g eV, Z = %d" % (tau_c, ds, tau.size, *worst))
