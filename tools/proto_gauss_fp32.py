"""Numerical prototype (numpy float32 emulation) of the two device paths for Gaussian bin integrals:
series path (small h) and erfc-difference path (large h). Compares with float64 erf differences."""
import numpy as np
from scipy.special import erf, erfcx
from numpy.polynomial import chebyshev as Ch
from numpy.polynomial import Polynomial

f32 = np.float32

def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

def ex2(x):  # MUFU.EX2 model: 2 ulp
    return (np.exp2(x.astype(np.float64))).astype(f32)

def fit_q(k=0.5, deg=9, amax=8.0):
    smin = 1/(1+k*amax)
    s = smin + (1-smin)*(np.cos(np.pi*(np.arange(4000)+0.5)/4000)+1)/2
    a = (1/s-1)/k
    f = 0.5*erfcx(a)/s
    V = Ch.chebvander(2*(s-smin)/(1-smin)-1, deg)
    c, *_ = np.linalg.lstsq(V/f[:, None], np.ones_like(f), rcond=None)
    Q = Polynomial(Ch.cheb2poly(c))(Polynomial([-1-2*smin/(1-smin), 2/(1-smin)]))
    return Q.coef

K = 0.5
QC = fit_q(K, 9)

def t_half_erfc(a, comp=True):
    """0.5*erfc(a), a>=0 float32"""
    s = f32(1)/(fma(f32(K)*np.ones_like(a), a, f32(1)*np.ones_like(a)))
    acc = np.full_like(a, f32(QC[-1]))
    for c in QC[-2::-1]:
        acc = fma(acc, s, np.full_like(a, f32(c)))
    q = acc*s
    hi = a*a
    L2E = f32(1.4426950408889634)
    p = hi*L2E
    e = ex2(-p)
    if comp:
        lo = fma(a, a, -hi)
        plo = fma(hi, np.full_like(a, L2E), -p)
        eps = fma(plo, np.full_like(a, f32(0.6931471805599453)), lo)
        e = fma(-e, eps, e)
    return q*e

def bins_erf(cf, kb, nb, comp=True):
    """bin integrals I_b = 0.5[erf(xu)-erf(xl)], x = (edge - cf)*kb, edges 0..nb in rel-bin units"""
    rel = np.arange(nb+1).astype(f32)
    x = fma(rel, np.full_like(rel, f32(kb)), np.full_like(rel, f32(-cf*kb)))  # (rel - cf)*kb
    t = t_half_erfc(np.abs(x), comp)
    xl, xu, tl, tu = x[:-1], x[1:], t[:-1], t[1:]
    D = np.where(xl >= 0, tl-tu, np.where(xu <= 0, tu-tl, f32(1)-tl-tu))
    return D

def bins_series(cf, kb, nb):
    rel = np.arange(nb).astype(f32)
    kbf = f32(kb)
    x = fma(rel, np.full_like(rel, kbf), np.full_like(rel, f32((0.5-cf)*kb)))
    h2 = f32(0.25)*kbf*kbf
    s0 = f32(1) - h2/f32(3) + h2*h2/f32(10)
    s1 = f32(2)*h2/f32(3) - f32(0.4)*h2*h2
    s2 = f32(4)*h2*h2/f32(30)
    m2 = x*x
    e = ex2(m2*f32(-1.4426950408889634))
    S = fma(fma(np.full_like(m2, s2), m2, np.full_like(m2, s1)), m2, np.full_like(m2, s0))
    A = kbf*f32(0.5641895835477563)
    return A*e*S

def ref(cf, kb, nb):
    e = (np.arange(nb+1) - cf)*kb
    er = erf(e)
    return 0.5*(er[1:]-er[:-1])

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    print("QC", repr(QC.astype(np.float32)))
    for sigma_b in [1.0, 1.5, 3, 6, 12, 12.2, 20, 40, 100]:
        kb = 1/(np.sqrt(2)*sigma_b); h = kb/2
        worst = {}
        for trial in range(20):
            nb = int(min(2048, 22*sigma_b+4)); cf = nb/2 + rng.uniform(0, 1)
            r = ref(cf, kb, nb)
            floor = 1e-9*r.max()
            for name, fn in (("erf", lambda: bins_erf(cf, kb, nb, False)), ("erf+comp", lambda: bins_erf(cf, kb, nb, True)), ("series", lambda: bins_series(cf, kb, nb))):
                v = fn().astype(np.float64)
                err = np.abs(v-r)/(np.abs(r)+floor/1e-4)   # in units where 1e-4 is the tolerance
                worst[name] = max(worst.get(name, 0), err.max())
        print("sigma_b=%6.1f h=%.4f" % (sigma_b, h), {k: "%.1e" % v for k, v in worst.items()})


def bins_series5(cf, kb, nb):
    """5-term series (through h^8 H_8(m)/9!) with the sqrt(log2 e) folding used on the device."""
    rel = np.arange(nb).astype(f32)
    kbf = f32(kb)
    SQ = f32(1.2011224087864498); IL = f32(0.6931471805599453)
    kx = kbf*SQ
    xo = f32(0.5-cf)*kx
    x = fma(rel, np.full_like(rel, kx), np.full_like(rel, xo))
    h2 = f32(0.25)*kbf*kbf
    t1 = h2/f32(6); t2 = h2*h2/f32(120); t3 = h2*h2*h2/f32(5040); t4 = h2*h2*h2*h2/f32(362880)
    A = kbf*f32(0.5641895835477563)
    s0 = A*(f32(1) - f32(2)*t1 + f32(12)*t2 - f32(120)*t3 + f32(1680)*t4)
    s1 = A*(f32(4)*t1 - f32(48)*t2 + f32(720)*t3 - f32(13440)*t4)*IL
    s2 = A*(f32(16)*t2 - f32(480)*t3 + f32(13440)*t4)*IL*IL
    s3 = A*(f32(64)*t3 - f32(3584)*t4)*IL*IL*IL
    s4 = A*(f32(256)*t4)*IL*IL*IL*IL
    m2 = x*x
    e = ex2(-m2)
    S = np.full_like(m2, s4)
    for c in (s3, s2, s1, s0):
        S = fma(S, m2, np.full_like(m2, c))
    return e*S


if __name__ == "__main__":
    print("5-term series")
    rng = np.random.default_rng(1)
    for sigma_b in [1.0, 1.41, 2.0, 2.83, 4.0, 5.66, 8, 20]:
        kb = 1/(np.sqrt(2)*sigma_b)
        worst = 0
        for trial in range(20):
            nb = int(min(2048, 22*sigma_b+4)); cf = nb/2 + rng.uniform(0, 1)
            r = ref(cf, kb, nb); floor = 1e-9*r.max()
            v = bins_series5(cf, kb, nb).astype(np.float64)
            worst = max(worst, (np.abs(v-r)/(np.abs(r)+floor/1e-4)).max())
        print("sigma_b=%5.2f h=%.4f worst %.1e" % (sigma_b, kb/2, worst))
