/*
 * cb2_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A scalar float64 restatement of the reference's algorithm for the hot path, written from the
 * reference's formulas (cherab/core v1.5.0; every function cites the file:line it follows) and,
 * for the Raysect 0.8.1 pieces whose source is not in /root/reference (NumericalIntegrator,
 * Interpolator1DArray/2DArray, Discrete2DMesh, Blend2D), from their published behaviour
 * (SURVEY.md Appendix B).  Parity status: the Cherab-side arithmetic is pinned against the
 * reference's own known-answer tests (tests/test_oracle_*.py); the Raysect-side interpolation
 * semantics are "parity unpinned" (no golden vectors exist in the reference for them).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It deliberately evaluates the function tree the way the reference does
 * (every quantity re-walks mapper -> blend -> mask -> psi_n), so it doubles as the CPU baseline.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC, -ffp-contract=off).
 */
#include "cb2_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants: cherab/core/utility/constants.pyx:22-37 ---- */
#define RECIP_4_PI (1.0 / (4.0 * M_PI))
#define ATOMIC_MASS 1.66053906660e-27
#define ELEMENTARY_CHARGE 1.602176634e-19
#define SPEED_OF_LIGHT 299792458.0
#define PLANCK_CONSTANT 6.62607015e-34
#define HC_EV_NM 1239.8419738620933
#define ELECTRON_REST_MASS 9.1093837015e-31
#define RYDBERG_CONSTANT_EV 13.605693122994
#define VACUUM_PERMITTIVITY 8.8541878128e-12
#define BOHR_MAGNETON 5.78838180123e-5
#define EULER_GAMMA 0.5772156649015329 /* gaunt.pyx:31 */

#define GAUSSIAN_CUTOFF_SIGMA 10.0   /* gaussian.pyx:33 */
#define LORENTZIAN_CUTOFF_GAMMA 50.0 /* stark.pyx:45 */
/* 4 * 50 * hyp2f1(0.4, 1, 1.4, -(2*50)**2.5)  (stark.pyx:62); value pinned against scipy in tests/test_oracle_lineshapes.py */
#define STARK_NORM_COEFFICIENT 2.641279471021934

static __thread char g_err[512];
static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
const char* cb2o_last_error(void) { return g_err; }
int cb2o_abi_version(void) { return CB2_ABI_VERSION; }

/* =================================================================================================
 * Raysect interpolators (restated)
 * ============================================================================================== */

/* index i with x[i] <= v < x[i+1]; the last knot belongs to the last cell. Caller guarantees x[0]<=v<=x[n-1]. */
static int find_cell(const double* x, int n, double v) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

/* derivative estimate at knot i along a strided line f[i*stride]: interior = 2nd-order 3-point formula on an
 * unevenly spaced grid, ends = one-sided first difference (raysect _ArrayDerivative1D) */
static double knot_derivative(const double* x, const double* f, int n, int stride, int i) {
    if (n < 2) return 0.0;
    if (i == 0) return (f[stride] - f[0]) / (x[1] - x[0]);
    if (i == n - 1) return (f[(n - 1) * stride] - f[(n - 2) * stride]) / (x[n - 1] - x[n - 2]);
    double hm = x[i] - x[i - 1], hp = x[i + 1] - x[i];
    double fm = f[(i - 1) * stride], f0 = f[i * stride], fp = f[(i + 1) * stride];
    return (fp * hm * hm - fm * hp * hp + f0 * (hp * hp - hm * hm)) / (hm * hp * (hm + hp));
}

/* cubic Hermite on [0,1]: values f0,f1 and derivatives d0,d1 already scaled by the cell width */
static double hermite(double f0, double f1, double d0, double d1, double t) {
    double a2 = 3.0 * (f1 - f0) - 2.0 * d0 - d1;
    double a3 = 2.0 * (f0 - f1) + d0 + d1;
    return f0 + t * (d0 + t * (a2 + t * a3));
}

/* Interpolator1DArray(x, f, 'cubic', extrap, range): extrapolate!=0 -> 'nearest' (clamp), else clamp too but the
 * caller counts the out-of-domain event (the reference raises ValueError). */
double cb2o_interp1d_cubic(const double* x, const double* f, int n, double px, int extrapolate) {
    (void)extrapolate;
    if (n == 1) return f[0];
    if (px < x[0]) px = x[0];
    if (px > x[n - 1]) px = x[n - 1];
    int i = find_cell(x, n, px);
    double h = x[i + 1] - x[i];
    double t = (px - x[i]) / h;
    double d0 = knot_derivative(x, f, n, 1, i) * h;
    double d1 = knot_derivative(x, f, n, 1, i + 1) * h;
    return hermite(f[i], f[i + 1], d0, d1, t);
}

/* derivative of hermite() with respect to t */
static double hermite_d(double f0, double f1, double d0, double d1, double t) {
    double a2 = 3.0 * (f1 - f0) - 2.0 * d0 - d1;
    double a3 = 2.0 * (f0 - f1) + d0 + d1;
    return d0 + t * (2.0 * a2 + 3.0 * t * a3);
}

/* Interpolator1DArray(x, f, 'cubic', 'quadratic', INFINITY) — the extrapolation Cherab asks for in openadas/rates/beam.pyx:76
 * and cx.pyx:96.  [raysect, restated from memory: parity unpinned]  Outside the knots the value follows a parabola in the
 * normalised coordinate t of the edge interval that matches the spline's value at the edge knot and its first derivative at
 * BOTH knots of that interval: q(t) = a0 t^2 + a1 t + a2 with a0 = (d1 - d0)/2, a1 = d0 and a2 = f0 at the lower end,
 * a2 = f1 - (d0 + d1)/2 at the upper end (d = knot derivatives scaled by the interval width). */
static double interp1d_cubic_quadratic(const double* x, const double* f, int n, double px) {
    if (n == 1) return f[0];
    if (px >= x[0] && px <= x[n - 1]) return cb2o_interp1d_cubic(x, f, n, px, 1);
    int i = px < x[0] ? 0 : n - 2;
    double h = x[i + 1] - x[i];
    double t = (px - x[i]) / h;
    double d0 = knot_derivative(x, f, n, 1, i) * h, d1 = knot_derivative(x, f, n, 1, i + 1) * h;
    double a0 = 0.5 * (d1 - d0), a1 = d0;
    double a2 = px < x[0] ? f[i] : f[i + 1] - 0.5 * d0 - 0.5 * d1;
    return (a0 * t + a1) * t + a2;
}

static double interp1d_linear(const double* x, const double* f, int n, double px) {
    if (n == 1) return f[0];
    if (px < x[0]) px = x[0];
    if (px > x[n - 1]) px = x[n - 1];
    int i = find_cell(x, n, px);
    double t = (px - x[i]) / (x[i + 1] - x[i]);
    return f[i] + t * (f[i + 1] - f[i]);
}

/* cross derivative at knot (i,j): four-corner difference over the neighbouring knots that exist
 * (raysect _ArrayDerivative2D; interior denominator (dx0+dx1)(dy0+dy1)) */
static double knot_cross_derivative(const double* x, const double* y, const double* f, int nx, int ny, int i, int j) {
    int il = i > 0 ? i - 1 : i, ih = i < nx - 1 ? i + 1 : i;
    int jl = j > 0 ? j - 1 : j, jh = j < ny - 1 ? j + 1 : j;
    if (il == ih || jl == jh) return 0.0;
    return (f[ih * ny + jh] - f[ih * ny + jl] - f[il * ny + jh] + f[il * ny + jl]) / ((x[ih] - x[il]) * (y[jh] - y[jl]));
}

/* Interpolator2DArray(x, y, f, 'cubic', ...): bicubic Hermite patch from f, fx, fy, fxy at the 4 cell corners */
double cb2o_interp2d_cubic(const double* x, const double* y, const double* f, int nx, int ny,
                           double px, double py, int extrapolate) {
    (void)extrapolate;
    if (px < x[0]) px = x[0];
    if (px > x[nx - 1]) px = x[nx - 1];
    if (py < y[0]) py = y[0];
    if (py > y[ny - 1]) py = y[ny - 1];
    int i = find_cell(x, nx, px), j = find_cell(y, ny, py);
    double hx = x[i + 1] - x[i], hy = y[j + 1] - y[j];
    double t = (px - x[i]) / hx, u = (py - y[j]) / hy;
    double g[2], gy[2]; /* value and d/dy (scaled) along y at x-knots i, i+1, then Hermite in x */
    double gx[2], gxy[2];
    for (int a = 0; a < 2; a++) {
        int ia = i + a;
        double f0 = f[ia * ny + j], f1 = f[ia * ny + j + 1];
        double dy0 = knot_derivative(y, f + ia * ny, ny, 1, j) * hy;
        double dy1 = knot_derivative(y, f + ia * ny, ny, 1, j + 1) * hy;
        g[a] = hermite(f0, f1, dy0, dy1, u);
        /* x-derivative of the patch along the edge x = x[ia]: Hermite in y of fx with fxy */
        double dx0 = knot_derivative(x, f + j, nx, ny, ia) * hx;
        double dx1 = knot_derivative(x, f + j + 1, nx, ny, ia) * hx;
        double dxy0 = knot_cross_derivative(x, y, f, nx, ny, ia, j) * hx * hy;
        double dxy1 = knot_cross_derivative(x, y, f, nx, ny, ia, j + 1) * hx * hy;
        gx[a] = hermite(dx0, dx1, dxy0, dxy1, u);
        (void)gy; (void)gxy;
    }
    return hermite(g[0], g[1], gx[0], gx[1], t);
}

/* Interpolator2DArray(x, y, f, 'cubic', 'linear', INFINITY, INFINITY) — beam.pyx:75,84.  [raysect, restated from memory: parity
 * unpinned]  Outside the grid: the spline's value, gradient and cross derivative at the nearest point of the grid boundary,
 * f + fx dx + fy dy + fxy dx dy (dx, dy the distances beyond the boundary; one of them is zero next to an edge). */
static double interp2d_cubic_linear(const double* x, const double* y, const double* f, int nx, int ny, double px, double py) {
    double cx = px < x[0] ? x[0] : (px > x[nx - 1] ? x[nx - 1] : px);
    double cy = py < y[0] ? y[0] : (py > y[ny - 1] ? y[ny - 1] : py);
    if (cx == px && cy == py) return cb2o_interp2d_cubic(x, y, f, nx, ny, px, py, 1);
    int i = find_cell(x, nx, cx), j = find_cell(y, ny, cy);
    double hx = x[i + 1] - x[i], hy = y[j + 1] - y[j];
    double t = (cx - x[i]) / hx, u = (cy - y[j]) / hy;
    double g[2], gx[2], dg[2], dgx[2];
    for (int a = 0; a < 2; a++) {
        int ia = i + a;
        double f0 = f[ia * ny + j], f1 = f[ia * ny + j + 1];
        double dy0 = knot_derivative(y, f + ia * ny, ny, 1, j) * hy, dy1 = knot_derivative(y, f + ia * ny, ny, 1, j + 1) * hy;
        double dx0 = knot_derivative(x, f + j, nx, ny, ia) * hx, dx1 = knot_derivative(x, f + j + 1, nx, ny, ia) * hx;
        double dxy0 = knot_cross_derivative(x, y, f, nx, ny, ia, j) * hx * hy, dxy1 = knot_cross_derivative(x, y, f, nx, ny, ia, j + 1) * hx * hy;
        g[a] = hermite(f0, f1, dy0, dy1, u);        dg[a] = hermite_d(f0, f1, dy0, dy1, u);
        gx[a] = hermite(dx0, dx1, dxy0, dxy1, u);   dgx[a] = hermite_d(dx0, dx1, dxy0, dxy1, u);
    }
    double v = hermite(g[0], g[1], gx[0], gx[1], t);
    double fx = hermite_d(g[0], g[1], gx[0], gx[1], t) / hx;
    double fy = hermite(dg[0], dg[1], dgx[0], dgx[1], t) / hy;
    double fxy = hermite_d(dg[0], dg[1], dgx[0], dgx[1], t) / (hx * hy);
    double dx = px - cx, dy = py - cy;
    return v + fx * dx + fy * dy + fxy * dx * dy;
}

/* third cross derivative at knot (i,j,k): eight-corner difference over the neighbouring knots that exist
 * (the 3-D analogue of knot_cross_derivative; raysect _ArrayDerivative3D) */
static double knot_cross3(const double* x, const double* y, const double* z, const double* f, int nx, int ny, int nz, int i, int j, int k) {
    int il = i > 0 ? i - 1 : i, ih = i < nx - 1 ? i + 1 : i;
    int jl = j > 0 ? j - 1 : j, jh = j < ny - 1 ? j + 1 : j;
    int kl = k > 0 ? k - 1 : k, kh = k < nz - 1 ? k + 1 : k;
    if (il == ih || jl == jh || kl == kh) return 0.0;
#define F3(a, b, c) f[((size_t)(a) * ny + (b)) * nz + (c)]
    double v = F3(ih, jh, kh) - F3(ih, jh, kl) - F3(ih, jl, kh) + F3(ih, jl, kl)
             - F3(il, jh, kh) + F3(il, jh, kl) + F3(il, jl, kh) - F3(il, jl, kl);
#undef F3
    return v / ((x[ih] - x[il]) * (y[jh] - y[jl]) * (z[kh] - z[kl]));
}

/* four-corner cross difference of a strided 2-D slice: rows stride sa (n_a knots a), columns stride sb (n_b knots b) */
static double knot_cross_strided(const double* a, const double* b, const double* f, int na, int nb, size_t sa, size_t sb, int i, int j) {
    int il = i > 0 ? i - 1 : i, ih = i < na - 1 ? i + 1 : i;
    int jl = j > 0 ? j - 1 : j, jh = j < nb - 1 ? j + 1 : j;
    if (il == ih || jl == jh) return 0.0;
    return (f[ih * sa + jh * sb] - f[ih * sa + jl * sb] - f[il * sa + jh * sb] + f[il * sa + jl * sb]) / ((a[ih] - a[il]) * (b[jh] - b[jl]));
}

/* Interpolator3DArray(x, y, z, f, 'cubic', ...): tricubic Hermite patch from f and its seven first/cross derivatives at
 * the 8 cell corners, evaluated as nested 1-D Hermite cubics (z, then y, then x).  Out-of-range arguments are clamped
 * ('nearest'); the caller counts the out-of-domain event when extrapolation is not permitted. */
double cb2o_interp3d_cubic(const double* x, const double* y, const double* z, const double* f, int nx, int ny, int nz,
                           double px, double py, double pz) {
    if (px < x[0]) px = x[0];
    if (px > x[nx - 1]) px = x[nx - 1];
    if (py < y[0]) py = y[0];
    if (py > y[ny - 1]) py = y[ny - 1];
    if (pz < z[0]) pz = z[0];
    if (pz > z[nz - 1]) pz = z[nz - 1];
    int i = find_cell(x, nx, px), j = find_cell(y, ny, py), k = find_cell(z, nz, pz);
    double hx = x[i + 1] - x[i], hy = y[j + 1] - y[j], hz = z[k + 1] - z[k];
    double t = (px - x[i]) / hx, u = (py - y[j]) / hy, w = (pz - z[k]) / hz;
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    double gx[2][2]; /* [value | d/dx * hx][x-knot] of the (y,z) patch */
    for (int a = 0; a < 2; a++) {
        int ia = i + a;
        double gy[2][2][2]; /* [x-order][value | d/dy * hy][y-knot] after the z reduction */
        for (int b = 0; b < 2; b++) {
            int jb = j + b;
            double q[2][2][2]; /* [x-order][y-order][z-knot] scaled derivative data, then their z-derivatives */
            double qz[2][2][2];
            for (int c = 0; c < 2; c++) {
                int kc = k + c;
                const double* f0 = f + (size_t)ia * sx + (size_t)jb * sy + kc;
                q[0][0][c] = *f0;
                q[1][0][c] = knot_derivative(x, f + (size_t)jb * sy + kc, nx, (int)sx, ia) * hx;
                q[0][1][c] = knot_derivative(y, f + (size_t)ia * sx + kc, ny, (int)sy, jb) * hy;
                q[1][1][c] = knot_cross_strided(x, y, f + kc, nx, ny, sx, sy, ia, jb) * hx * hy;
                qz[0][0][c] = knot_derivative(z, f + (size_t)ia * sx + (size_t)jb * sy, nz, 1, kc) * hz;
                qz[1][0][c] = knot_cross_strided(x, z, f + (size_t)jb * sy, nx, nz, sx, 1, ia, kc) * hx * hz;
                qz[0][1][c] = knot_cross_strided(y, z, f + (size_t)ia * sx, ny, nz, sy, 1, jb, kc) * hy * hz;
                qz[1][1][c] = knot_cross3(x, y, z, f, nx, ny, nz, ia, jb, kc) * hx * hy * hz;
            }
            for (int ox = 0; ox < 2; ox++)
                for (int oy = 0; oy < 2; oy++)
                    gy[ox][oy][b] = hermite(q[ox][oy][0], q[ox][oy][1], qz[ox][oy][0], qz[ox][oy][1], w);
        }
        for (int ox = 0; ox < 2; ox++) gx[ox][a] = hermite(gy[ox][0][0], gy[ox][0][1], gy[ox][1][0], gy[ox][1][1], u);
    }
    return hermite(gx[0][0], gx[0][1], gx[1][0], gx[1][1], t);
}

/* =================================================================================================
 * GaussianQuadrature (cherab/core/math/integrators/integrators1d.pyx:164-224)
 * ============================================================================================== */
#define GL_MAX_ORDER 64
static double gl_roots[GL_MAX_ORDER + 1][GL_MAX_ORDER];
static double gl_weights[GL_MAX_ORDER + 1][GL_MAX_ORDER];
static int gl_ready = 0;

static void gl_build(void) {
    /* Newton iteration on Legendre polynomials; agrees with scipy.special.roots_legendre to ~1e-15 */
    for (int n = 1; n <= GL_MAX_ORDER; n++) {
        for (int k = 0; k < n; k++) {
            double xk = cos(M_PI * (k + 0.75) / (n + 0.5)), dp = 1.0;
            for (int it = 0; it < 100; it++) {
                double p0 = 1.0, p1 = xk;
                for (int m = 2; m <= n; m++) { double p2 = ((2 * m - 1) * xk * p1 - (m - 1) * p0) / m; p0 = p1; p1 = p2; }
                if (n == 1) { p0 = 1.0; p1 = xk; }
                dp = n * (xk * p1 - p0) / (xk * xk - 1.0);
                double dx = p1 / dp;
                xk -= dx;
                if (fabs(dx) < 1e-16) break;
            }
            { /* recompute derivative at the converged root */
                double p0 = 1.0, p1 = xk;
                for (int m = 2; m <= n; m++) { double p2 = ((2 * m - 1) * xk * p1 - (m - 1) * p0) / m; p0 = p1; p1 = p2; }
                dp = n * (xk * p1 - p0) / (xk * xk - 1.0);
            }
            /* ascending order like scipy */
            gl_roots[n][n - 1 - k] = xk;
            gl_weights[n][n - 1 - k] = 2.0 / ((1.0 - xk * xk) * dp * dp);
        }
        if (n == 1) { gl_roots[1][0] = 0.0; gl_weights[1][0] = 2.0; }
    }
    gl_ready = 1;
}

double cb2o_gauss_legendre(double (*fn)(double, void*), void* ctx, double a, double b,
                           double rtol, int min_order, int max_order) {
    if (!gl_ready) {
#pragma omp critical(cb2o_gl)
        { if (!gl_ready) gl_build(); }
    }
    if (max_order > GL_MAX_ORDER) max_order = GL_MAX_ORDER;
    double oldval = INFINITY, newval = 0.0;
    double c = 0.5 * (a + b), d = 0.5 * (b - a);
    for (int order = min_order; order <= max_order; order++) {
        newval = 0.0;
        for (int i = 0; i < order; i++) newval += gl_weights[order][i] * fn(c + d * gl_roots[order][i], ctx);
        newval *= d;
        double error = fabs(newval - oldval);
        oldval = newval;
        if (error < rtol * fabs(newval)) break;
    }
    return newval;
}

/* =================================================================================================
 * Line shapes
 * ============================================================================================== */

/* add_gaussian_line — cherab/core/model/lineshape/gaussian.pyx:40-90 */
static int64_t add_gaussian_line(double radiance, double wavelength, double sigma,
                                 const cb2_spectral_grid* g, double* samples) {
    if (sigma <= 0) return 0;
    double delta = (g->max_wavelength - g->min_wavelength) / g->bins;
    double lo = wavelength - GAUSSIAN_CUTOFF_SIGMA * sigma;
    if (g->max_wavelength < lo) return 0;
    double hi = wavelength + GAUSSIAN_CUTOFF_SIGMA * sigma;
    if (g->min_wavelength > hi) return 0;
    int start = (int)floor((lo - g->min_wavelength) / delta);
    if (start < 0) start = 0;
    int end = (int)ceil((hi - g->min_wavelength) / delta);
    if (end > g->bins) end = g->bins;
    double temp = 1.0 / (M_SQRT2 * sigma);
    double lower_wavelength = g->min_wavelength + start * delta;
    double lower_integral = erf((lower_wavelength - wavelength) * temp);
    for (int i = start; i < end; i++) {
        double upper_wavelength = g->min_wavelength + delta * (i + 1);
        double upper_integral = erf((upper_wavelength - wavelength) * temp);
        samples[i] += radiance * 0.5 * (upper_integral - lower_integral) / delta;
        lower_integral = upper_integral;
    }
    return end > start ? (int64_t)(end - start) + 1 : 0;
}

int cb2o_add_gaussian_line(double radiance, double wavelength, double sigma,
                           const cb2_spectral_grid* grid, double* samples) {
    add_gaussian_line(radiance, wavelength, sigma, grid, samples);
    return CB2_OK;
}

/* StarkFunction — stark.pyx:52-81 */
typedef struct { double a, x0, norm; } stark_fn;
static double stark_eval(double x, void* p) {
    const stark_fn* s = (const stark_fn*)p;
    return s->norm / (pow(fabs(x - s->x0), 2.5) + s->a);
}

/* add_lorentzian_line — stark.pyx:88-147 */
static int64_t add_lorentzian_line(double radiance, double wavelength, double lambda_1_2,
                                   const cb2_spectral_grid* g, double* samples,
                                   double rtol, int min_order, int max_order) {
    if (lambda_1_2 <= 0) return 0;
    stark_fn s;
    s.x0 = wavelength;
    s.a = pow(0.5 * lambda_1_2, 2.5);
    s.norm = pow(0.5 * lambda_1_2, 1.5) / STARK_NORM_COEFFICIENT;
    double delta = (g->max_wavelength - g->min_wavelength) / g->bins;
    double lo = wavelength - LORENTZIAN_CUTOFF_GAMMA * lambda_1_2;
    if (g->max_wavelength < lo) return 0;
    double hi = wavelength + LORENTZIAN_CUTOFF_GAMMA * lambda_1_2;
    if (g->min_wavelength > hi) return 0;
    int start = (int)floor((lo - g->min_wavelength) / delta);
    if (start < 0) start = 0;
    int end = (int)ceil((hi - g->min_wavelength) / delta);
    if (end > g->bins) end = g->bins;
    double lower_wavelength = g->min_wavelength + start * delta;
    for (int i = start; i < end; i++) {
        double upper_wavelength = g->min_wavelength + delta * (i + 1);
        double bin_integral = cb2o_gauss_legendre(stark_eval, &s, lower_wavelength, upper_wavelength, rtol, min_order, max_order);
        samples[i] += radiance * bin_integral / delta;
        lower_wavelength = upper_wavelength;
    }
    return end > start ? (int64_t)(end - start) + 1 : 0;
}

int cb2o_add_lorentzian_line(double radiance, double wavelength, double lambda_1_2,
                             const cb2_spectral_grid* grid, double* samples,
                             double rtol, int min_order, int max_order) {
    add_lorentzian_line(radiance, wavelength, lambda_1_2, grid, samples, rtol, min_order, max_order);
    return CB2_OK;
}

/* doppler_shift / thermal_broadening — doppler.pyx:29-59 */
static double doppler_shift(double wavelength, const double d[3], const double v[3]) {
    double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double proj = (v[0] * d[0] + v[1] * d[1] + v[2] * d[2]) / len;
    return wavelength * (1.0 + proj / SPEED_OF_LIGHT);
}
static double thermal_broadening(double wavelength, double temperature, double atomic_weight) {
    return sqrt(temperature * ELEMENTARY_CHARGE / (atomic_weight * ATOMIC_MASS)) * wavelength / SPEED_OF_LIGHT;
}

/* =================================================================================================
 * Rates
 * ============================================================================================== */

/* per-call cache of log10 tables so the bicubic sees exactly what pec.pyx:59-68 builds */
typedef struct { int n_ne, n_te; double *lne, *lte, *lrate; } pec_table;

static void pec_table_build(pec_table* t, const cb2_rate2d* p, double wavelength) {
    memset(t, 0, sizeof *t);
    if (p->n_ne <= 0) return;
    t->n_ne = p->n_ne; t->n_te = p->n_te;
    t->lne = (double*)malloc(sizeof(double) * p->n_ne);
    t->lte = (double*)malloc(sizeof(double) * p->n_te);
    t->lrate = (double*)malloc(sizeof(double) * p->n_ne * p->n_te);
    for (int i = 0; i < p->n_ne; i++) t->lne[i] = log10(p->ne[i]);
    for (int j = 0; j < p->n_te; j++) t->lte[j] = log10(p->te[j]);
    double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9; /* PhotonToJ, conversion.py:44-52 */
    for (int k = 0; k < p->n_ne * p->n_te; k++) t->lrate[k] = log10(p->rate[k] / wavelength * conv);
}
static void pec_table_free(pec_table* t) { free(t->lne); free(t->lte); free(t->lrate); }

/* LineRadiationPower / ContinuumPower / CXRadiationPower tables: log10(rate [W m^3]) on (log10 ne, log10 te) —
 * openadas/rates/radiated_power.pyx:48-76 (no photon-to-energy conversion) */
static void power_table_build(pec_table* t, const cb2_rate2d* p) {
    memset(t, 0, sizeof *t);
    if (p->n_ne <= 0) return;
    t->n_ne = p->n_ne; t->n_te = p->n_te;
    t->lne = (double*)malloc(sizeof(double) * p->n_ne);
    t->lte = (double*)malloc(sizeof(double) * p->n_te);
    t->lrate = (double*)malloc(sizeof(double) * p->n_ne * p->n_te);
    for (int i = 0; i < p->n_ne; i++) t->lne[i] = log10(p->ne[i]);
    for (int j = 0; j < p->n_te; j++) t->lte[j] = log10(p->te[j]);
    for (int k = 0; k < p->n_ne * p->n_te; k++) t->lrate[k] = log10(p->rate[k]);
}

/* ThermalCXPEC (openadas/rates/pec.pyx:153-194): log10(PhotonToJ(rate)) on (log10 ne, log10 te, log10 td), tricubic */
typedef struct tcx_table { int n_ne, n_te, n_td; double *lne, *lte, *ltd, *lrate; } tcx_table;

static void tcx_table_build(tcx_table* t, const cb2_rate3d* p, double wavelength) {
    memset(t, 0, sizeof *t);
    if (p->n_ne <= 0) return;
    t->n_ne = p->n_ne; t->n_te = p->n_te; t->n_td = p->n_td;
    size_t n = (size_t)p->n_ne * p->n_te * p->n_td;
    t->lne = (double*)malloc(sizeof(double) * p->n_ne);
    t->lte = (double*)malloc(sizeof(double) * p->n_te);
    t->ltd = (double*)malloc(sizeof(double) * p->n_td);
    t->lrate = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < p->n_ne; i++) t->lne[i] = log10(p->ne[i]);
    for (int j = 0; j < p->n_te; j++) t->lte[j] = log10(p->te[j]);
    for (int k = 0; k < p->n_td; k++) t->ltd[k] = log10(p->td[k]);
    double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9;
    for (size_t k = 0; k < n; k++) t->lrate[k] = log10(p->rate[k] / wavelength * conv);
}
static void tcx_table_free(tcx_table* t) { free(t->lne); free(t->lte); free(t->ltd); free(t->lrate); }

/* ThermalCXPEC.evaluate — pec.pyx:186-194 */
static double tcx_eval(const tcx_table* t, const cb2_rate3d* p, double ne, double te, double td, int64_t* ood) {
    if (p->n_ne <= 0) return p->constant;
    if (ne <= 0 || te <= 0 || td <= 0) return 0.0;
    double x = log10(ne), y = log10(te), z = log10(td);
    if (!p->extrapolate && (x < t->lne[0] || x > t->lne[t->n_ne - 1] || y < t->lte[0] || y > t->lte[t->n_te - 1] ||
                            z < t->ltd[0] || z > t->ltd[t->n_td - 1])) (*ood)++;
    return pow(10.0, cb2o_interp3d_cubic(t->lne, t->lte, t->ltd, t->lrate, t->n_ne, t->n_te, t->n_td, x, y, z));
}

double cb2o_thermal_cx_pec_evaluate(const cb2_rate3d* pec, double wavelength, double ne, double te, double td) {
    tcx_table t; int64_t ood = 0;
    tcx_table_build(&t, pec, wavelength);
    double v = tcx_eval(&t, pec, ne, te, td, &ood);
    tcx_table_free(&t);
    return v;
}

/* ImpactExcitationPEC.evaluate — pec.pyx:70-77 */
static double pec_eval(const pec_table* t, const cb2_rate2d* p, double ne, double te, int64_t* ood) {
    if (p->n_ne <= 0) return p->constant;
    if (ne <= 0 || te <= 0) return 0.0;
    double x = log10(ne), y = log10(te);
    if (!p->extrapolate && (x < t->lne[0] || x > t->lne[t->n_ne - 1] || y < t->lte[0] || y > t->lte[t->n_te - 1])) (*ood)++;
    return pow(10.0, cb2o_interp2d_cubic(t->lne, t->lte, t->lrate, t->n_ne, t->n_te, x, y, 1));
}

double cb2o_pec_evaluate(const cb2_rate2d* pec, double wavelength, double ne, double te) {
    pec_table t; int64_t ood = 0;
    pec_table_build(&t, pec, wavelength);
    double v = pec_eval(&t, pec, ne, te, &ood);
    pec_table_free(&t);
    return v;
}

/* InterpolatedFreeFreeGauntFactor.evaluate — gaunt.pyx:109-140 */
typedef struct { int n_u, n_g; double *lu, *lg; const double* gff; double u_min, u_max, g_min, g_max; } gaunt_table;

static void gaunt_table_build(gaunt_table* t, const cb2_gaunt* g) {
    memset(t, 0, sizeof *t);
    if (g->n_u <= 0) return;
    t->n_u = g->n_u; t->n_g = g->n_gamma2; t->gff = g->gaunt;
    t->lu = (double*)malloc(sizeof(double) * g->n_u);
    t->lg = (double*)malloc(sizeof(double) * g->n_gamma2);
    t->u_min = t->g_min = INFINITY; t->u_max = t->g_max = -INFINITY;
    for (int i = 0; i < g->n_u; i++) { t->lu[i] = log10(g->u[i]); t->u_min = fmin(t->u_min, g->u[i]); t->u_max = fmax(t->u_max, g->u[i]); }
    for (int i = 0; i < g->n_gamma2; i++) { t->lg[i] = log10(g->gamma2[i]); t->g_min = fmin(t->g_min, g->gamma2[i]); t->g_max = fmax(t->g_max, g->gamma2[i]); }
}
static void gaunt_table_free(gaunt_table* t) { free(t->lu); free(t->lg); }

static double gaunt_eval(const gaunt_table* t, double z, double temperature, double wavelength) {
    if (z == 0) return 0.0;
    double gamma2 = z * z * RYDBERG_CONSTANT_EV / temperature;
    double u = HC_EV_NM / (temperature * wavelength); /* PH_TO_EV_FACTOR gaunt.pyx:30 */
    if (u >= t->u_max || gamma2 >= t->g_max) return 1.0;
    if (u < t->u_min || gamma2 < t->g_min) return sqrt(3.0) / M_PI * (log(4.0 / u) - EULER_GAMMA);
    return cb2o_interp2d_cubic(t->lu, t->lg, t->gff, t->n_u, t->n_g, log10(u), log10(gamma2), 0);
}

double cb2o_gaunt_factor(const cb2_gaunt* g, double z, double te, double wavelength) {
    gaunt_table t;
    gaunt_table_build(&t, g);
    double v = gaunt_eval(&t, z, te, wavelength);
    gaunt_table_free(&t);
    return v;
}

/* =================================================================================================
 * Axisymmetric context: EFIT equilibrium + edge mesh (efit.pyx, generomak/plasma/plasma.py)
 * ============================================================================================== */
typedef struct {
    const cb2_axisym* ax;
    double* psin;     /* (psi - psi_axis)/(psi_lcfs - psi_axis), efit.pyx:116 */
    double* dpsi_dr;  /* efit.pyx:184-194 */
    double* dpsi_dz;
    /* uniform bucket grid over the mesh bounding box (stands in for Discrete2DMesh's kd-tree) */
    int gx, gy; double x0, y0, inv_cx, inv_cy;
    int* cell_start; int* cell_tris;
} axisym_ctx;

/* np.gradient(f, edge_order=2) with unit spacing, along a strided line */
static void np_gradient_unit(const double* f, int n, int stride, double* out, int ostride) {
    for (int i = 0; i < n; i++) {
        double v;
        if (n == 1) v = 0.0;
        else if (n == 2) v = f[stride] - f[0];
        else if (i == 0) v = -(3.0 * f[0] - 4.0 * f[stride] + f[2 * stride]) / 2.0;
        else if (i == n - 1) v = (3.0 * f[(n - 1) * stride] - 4.0 * f[(n - 2) * stride] + f[(n - 3) * stride]) / 2.0;
        else v = (f[(i + 1) * stride] - f[(i - 1) * stride]) / 2.0;
        out[i * ostride] = v;
    }
}

static int tri_contains(const double* v, const int32_t* t, double px, double py) {
    double ax = v[2 * t[0]], ay = v[2 * t[0] + 1], bx = v[2 * t[1]], by = v[2 * t[1] + 1], cx = v[2 * t[2]], cy = v[2 * t[2] + 1];
    double d1 = (px - bx) * (ay - by) - (ax - bx) * (py - by);
    double d2 = (px - cx) * (by - cy) - (bx - cx) * (py - cy);
    double d3 = (px - ax) * (cy - ay) - (cx - ax) * (py - ay);
    int neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
    int pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
    return !(neg && pos);
}

static int axisym_ctx_build(axisym_ctx* c, const cb2_axisym* ax) {
    memset(c, 0, sizeof *c);
    c->ax = ax;
    const cb2_equilibrium* e = &ax->eq;
    int nr = e->nr, nz = e->nz;
    c->psin = (double*)malloc(sizeof(double) * nr * nz);
    c->dpsi_dr = (double*)malloc(sizeof(double) * nr * nz);
    c->dpsi_dz = (double*)malloc(sizeof(double) * nr * nz);
    double* dr_di = (double*)malloc(sizeof(double) * nr);
    double* dz_di = (double*)malloc(sizeof(double) * nz);
    for (int k = 0; k < nr * nz; k++) c->psin[k] = (e->psi[k] - e->psi_axis) / (e->psi_lcfs - e->psi_axis);
    np_gradient_unit(e->r, nr, 1, dr_di, 1);
    np_gradient_unit(e->z, nz, 1, dz_di, 1);
    for (int j = 0; j < nz; j++) np_gradient_unit(e->psi + j, nr, nz, c->dpsi_dr + j, nz);
    for (int i = 0; i < nr; i++) np_gradient_unit(e->psi + i * nz, nz, 1, c->dpsi_dz + i * nz, 1);
    for (int i = 0; i < nr; i++)
        for (int j = 0; j < nz; j++) {
            c->dpsi_dr[i * nz + j] *= 1.0 / dr_di[i];
            c->dpsi_dz[i * nz + j] *= 1.0 / dz_di[j];
        }
    free(dr_di); free(dz_di);
    /* bucket grid */
    if (ax->n_triangles > 0) {
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int i = 0; i < ax->n_vertices; i++) {
            xmin = fmin(xmin, ax->vertices[2 * i]); xmax = fmax(xmax, ax->vertices[2 * i]);
            ymin = fmin(ymin, ax->vertices[2 * i + 1]); ymax = fmax(ymax, ax->vertices[2 * i + 1]);
        }
        c->gx = 128; c->gy = 256;
        c->x0 = xmin; c->y0 = ymin;
        c->inv_cx = c->gx / (xmax - xmin) * (1.0 - 1e-12); c->inv_cy = c->gy / (ymax - ymin) * (1.0 - 1e-12);
        int ncell = c->gx * c->gy;
        int* count = (int*)calloc(ncell + 1, sizeof(int));
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {
                c->cell_start = (int*)malloc(sizeof(int) * (ncell + 1));
                int acc = 0;
                for (int k = 0; k < ncell; k++) { c->cell_start[k] = acc; acc += count[k]; count[k] = 0; }
                c->cell_start[ncell] = acc;
                c->cell_tris = (int*)malloc(sizeof(int) * (acc > 0 ? acc : 1));
            }
            for (int t = 0; t < ax->n_triangles; t++) {
                const int32_t* tr = ax->triangles + 3 * t;
                double txmin = INFINITY, txmax = -INFINITY, tymin = INFINITY, tymax = -INFINITY;
                for (int k = 0; k < 3; k++) {
                    double vx = ax->vertices[2 * tr[k]], vy = ax->vertices[2 * tr[k] + 1];
                    txmin = fmin(txmin, vx); txmax = fmax(txmax, vx); tymin = fmin(tymin, vy); tymax = fmax(tymax, vy);
                }
                int i0 = (int)floor((txmin - c->x0) * c->inv_cx) - 1, i1 = (int)floor((txmax - c->x0) * c->inv_cx) + 1;
                int j0 = (int)floor((tymin - c->y0) * c->inv_cy) - 1, j1 = (int)floor((tymax - c->y0) * c->inv_cy) + 1;
                if (i0 < 0) i0 = 0; if (j0 < 0) j0 = 0; if (i1 >= c->gx) i1 = c->gx - 1; if (j1 >= c->gy) j1 = c->gy - 1;
                for (int i = i0; i <= i1; i++)
                    for (int j = j0; j <= j1; j++) {
                        int cell = i * c->gy + j;
                        if (pass == 1) c->cell_tris[c->cell_start[cell] + count[cell]] = t;
                        count[cell]++;
                    }
            }
        }
        free(count);
    }
    return CB2_OK;
}

static void axisym_ctx_free(axisym_ctx* c) {
    free(c->psin); free(c->dpsi_dr); free(c->dpsi_dz); free(c->cell_start); free(c->cell_tris);
}

/* Discrete2DMesh lookup: triangle index or -1 (limit=False -> default value 0 outside). Lowest triangle id wins ties. */
static int mesh_locate(const axisym_ctx* c, double r, double z) {
    const cb2_axisym* ax = c->ax;
    if (ax->n_triangles <= 0) return -1;
    double fx = (r - c->x0) * c->inv_cx, fy = (z - c->y0) * c->inv_cy;
    if (fx < 0 || fy < 0) return -1;
    int i = (int)fx, j = (int)fy;
    if (i >= c->gx || j >= c->gy) return -1;
    int cell = i * c->gy + j, best = -1;
    for (int k = c->cell_start[cell]; k < c->cell_start[cell + 1]; k++) {
        int t = c->cell_tris[k];
        if (tri_contains(ax->vertices, ax->triangles + 3 * t, r, z)) { if (best < 0 || t < best) best = t; }
    }
    return best;
}

/* PolygonMask2D — cherab/core/math/mask.pyx:53-67 (even-odd crossing test of the closed polygon) */
static int polygon_contains(const double* poly, int n, double px, double py) {
    int inside = 0;
    for (int i = 0, j = n - 1; i < n; j = i++) {
        double xi = poly[2 * i], yi = poly[2 * i + 1], xj = poly[2 * j], yj = poly[2 * j + 1];
        if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) inside = !inside;
    }
    return inside;
}

static int in_psi_grid(const cb2_equilibrium* e, double r, double z) {
    return r >= e->r[0] && r <= e->r[e->nr - 1] && z >= e->z[0] && z <= e->z[e->nz - 1];
}

/* psi_normalised = ClampOutput2D(bicubic, min=0) — efit.pyx:116; clamp.pyx:89-90 */
static double psi_normalised(const axisym_ctx* c, double r, double z, int64_t* ood) {
    const cb2_equilibrium* e = &c->ax->eq;
    if (!in_psi_grid(e, r, z)) (*ood)++;
    double v = cb2o_interp2d_cubic(e->r, e->z, c->psin, e->nr, e->nz, r, z, 0);
    return v < 0 ? 0 : v;
}

/* EFITLCFSMask.evaluate — efit.pyx:405-410 */
static int inside_lcfs(const axisym_ctx* c, double r, double z, int64_t* ood) {
    const cb2_equilibrium* e = &c->ax->eq;
    if (!polygon_contains(e->lcfs_polygon, e->n_lcfs, r, z)) return 0;
    return psi_normalised(c, r, z, ood) <= 1.0;
}

/* blend mask: equilibrium.map2d(Interpolator1DArray(mask_x, mask_y, 'linear')) — plasma.py:610, efit.pyx:250-253 */
static double blend_mask(const axisym_ctx* c, double r, double z, int64_t* ood) {
    if (!inside_lcfs(c, r, z, ood)) return 0.0;
    return interp1d_linear(c->ax->mask_x, c->ax->mask_y, c->ax->n_mask, psi_normalised(c, r, z, ood));
}

/* equilibrium.map2d(core 1-D cubic) — efit.pyx:219-253 */
static double core_map2d(const axisym_ctx* c, const double* core, double r, double z, int64_t* ood) {
    if (!core) return 0.0;
    if (!inside_lcfs(c, r, z, ood)) return 0.0;
    return cb2o_interp1d_cubic(c->ax->core_psin, core, c->ax->n_core, psi_normalised(c, r, z, ood), 1);
}

/* MagneticField.evaluate — efit.pyx:437-461; returns (B_R, B_phi, B_Z) */
static void efit_b_field(const axisym_ctx* c, double r, double z, double b[3], int64_t* ood) {
    const cb2_equilibrium* e = &c->ax->eq;
    if (!in_psi_grid(e, r, z)) (*ood)++;
    double br = -cb2o_interp2d_cubic(e->r, e->z, c->dpsi_dz, e->nr, e->nz, r, z, 0) / r;
    double bz = cb2o_interp2d_cubic(e->r, e->z, c->dpsi_dr, e->nr, e->nz, r, z, 0) / r;
    double bt;
    if (inside_lcfs(c, r, z, ood)) {
        double psi_n = psi_normalised(c, r, z, ood);
        bt = cb2o_interp1d_cubic(e->f_psin, e->f_value, e->n_f, psi_n, 0) / r;
    } else {
        bt = e->b_vacuum_magnitude * e->b_vacuum_radius / r;
    }
    b[0] = br; b[1] = bt; b[2] = bz;
}

/* FluxCoordToCartesian.evaluate inside VectorBlend2D(0, v, inside_lcfs) — efit.pyx:280-344,521-546 */
static void core_map_vector2d(const axisym_ctx* c, const cb2_vector_field* f, double r, double z, double v[3], int64_t* ood) {
    v[0] = v[1] = v[2] = 0.0;
    if (!inside_lcfs(c, r, z, ood)) return;
    double b[3];
    efit_b_field(c, r, z, b, ood);
    double psi = psi_normalised(c, r, z, ood);
    const cb2_axisym* ax = c->ax;
    double vtor = f->core_vtor ? cb2o_interp1d_cubic(ax->core_psin, f->core_vtor, ax->n_core, psi, 1) : 0.0;
    double px = 0, pz = 0, nx = 0, nz = 0;
    if (!(b[0] == 0 && b[2] == 0)) {
        double vpol = f->core_vpol ? cb2o_interp1d_cubic(ax->core_psin, f->core_vpol, ax->n_core, psi, 1) : 0.0;
        double vnorm = f->core_vnorm ? cb2o_interp1d_cubic(ax->core_psin, f->core_vnorm, ax->n_core, psi, 1) : 0.0;
        double len = sqrt(b[0] * b[0] + b[2] * b[2]);
        /* Vector3D.set_length(v): normalise then scale by v */
        px = b[0] / len * vpol; pz = b[2] / len * vpol;
        nx = -b[2] / len * vnorm; nz = b[0] / len * vnorm;
    }
    v[0] = px + nx; v[1] = vtor; v[2] = pz + nz;
}

/* =================================================================================================
 * Field evaluation (each call re-walks the tree, like the reference)
 * ============================================================================================== */
typedef struct {
    const cb2_scene_desc* d;
    axisym_ctx ax; int has_ax;
    pec_table* pec;   /* [n_models] */
    pec_table* trp;   /* [n_models][3] TotalRadiatedPower plt, prb, prc */
    struct tcx_table** tcx; /* [n_models] -> [n_donors] ThermalCXPEC log tables */
    gaunt_table gaunt;
    /* SingleRayAttenuator axis table (singleray.pyx:182-224): z_k = linspace(0, length, n_axis), density on the axis */
    int n_axis; double* axis_z; double* axis_density; double tanxdiv, tanydiv;
} scene_ctx;

static double eval_scalar(const scene_ctx* s, const cb2_scalar_field* f, const double p[3], int64_t* ood) {
    switch (f->kind) {
    case CB2_FIELD_CONSTANT: return f->c[0];
    case CB2_FIELD_GAUSSIAN_VOLUME: {
        double dx = p[0] - f->c[3], dy = p[1] - f->c[4], dz = p[2] - f->c[5];
        return f->c[0] + f->c[1] * exp(-(dx * dx + dy * dy + dz * dz) / (2 * f->c[2] * f->c[2]));
    }
    case CB2_FIELD_SLAB_ION: { /* slab.pyx:88-110 */
        double x_norm = p[0] / f->c[4];
        if (x_norm >= 0 && x_norm <= 1) return (f->c[0] - f->c[1]) * pow(1 - pow(1 - x_norm, f->c[2]), f->c[3]) + f->c[1];
        if (x_norm >= 1) return f->c[0];
        return 0.0;
    }
    case CB2_FIELD_SLAB_NEUTRAL: /* slab.pyx:39-60 */
        return p[0] >= 0 ? f->c[0] * exp(-(p[0] * p[0]) / (2 * f->c[1] * f->c[1])) : f->c[0];
    case CB2_FIELD_AXISYM_BLEND: { /* AxisymmetricMapper (mappers.pyx:260-264) of Blend2D(edge, core, mask) */
        double r = sqrt(p[0] * p[0] + p[1] * p[1]), z = p[2];
        double m = blend_mask(&s->ax, r, z, ood);
        double edge = 0.0, core = 0.0;
        if (m < 1.0) { int t = mesh_locate(&s->ax, r, z); edge = (t >= 0 && f->edge) ? f->edge[t] : 0.0; }
        if (m > 0.0) core = core_map2d(&s->ax, f->core, r, z, ood);
        if (m <= 0.0) return edge;
        if (m >= 1.0) return core;
        return (1.0 - m) * edge + m * core;
    }
    }
    return 0.0;
}

static void rotate_axisym(const double v2[3], const double p[3], double out[3]) {
    /* VectorAxisymmetricMapper — mappers.pyx:302-312: rotate (vR, vphi, vZ) by phi = atan2(y,x) about z */
    double phi = atan2(p[1], p[0]) / M_PI * 180.0;
    double cs = cos(phi * M_PI / 180.0), sn = sin(phi * M_PI / 180.0);
    out[0] = cs * v2[0] - sn * v2[1];
    out[1] = sn * v2[0] + cs * v2[1];
    out[2] = v2[2];
}

static void eval_vector(const scene_ctx* s, const cb2_vector_field* f, const double p[3], double out[3], int64_t* ood) {
    if (f->kind == CB2_FIELD_CONSTANT) { out[0] = f->c[0]; out[1] = f->c[1]; out[2] = f->c[2]; return; }
    if (f->kind == CB2_FIELD_AXISYM_BLEND) {
        double r = sqrt(p[0] * p[0] + p[1] * p[1]), z = p[2];
        double m = blend_mask(&s->ax, r, z, ood);
        double v2[3];
        if (m <= 0.0) { v2[0] = f->c[0]; v2[1] = f->c[1]; v2[2] = f->c[2]; }
        else {
            double core[3];
            core_map_vector2d(&s->ax, f, r, z, core, ood);
            if (m >= 1.0) { v2[0] = core[0]; v2[1] = core[1]; v2[2] = core[2]; }
            else for (int k = 0; k < 3; k++) v2[k] = (1.0 - m) * f->c[k] + m * core[k];
        }
        rotate_axisym(v2, p, out);
        return;
    }
    out[0] = out[1] = out[2] = 0.0;
}

static void eval_b_field(const scene_ctx* s, const double p[3], double out[3], int64_t* ood) {
    const cb2_scene_desc* d = s->d;
    if (d->b_field_kind == 0) { out[0] = d->b_field[0]; out[1] = d->b_field[1]; out[2] = d->b_field[2]; return; }
    double r = sqrt(p[0] * p[0] + p[1] * p[1]), b2[3];
    efit_b_field(&s->ax, r, p[2], b2, ood);
    rotate_axisym(b2, p, out);
}

static int beam_ctx_build(scene_ctx* s);

static int scene_ctx_build(scene_ctx* s, const cb2_scene_desc* d) {
    memset(s, 0, sizeof *s);
    if (d->abi_version != CB2_ABI_VERSION) return fail(CB2_ERR_VALUE, "abi_version mismatch");
    s->d = d;
    if (d->axisym) { axisym_ctx_build(&s->ax, d->axisym); s->has_ax = 1; }
    s->pec = (pec_table*)calloc(d->n_models > 0 ? d->n_models : 1, sizeof(pec_table));
    for (int m = 0; m < d->n_models; m++) {
        const cb2_model* mo = &d->models[m];
        if (mo->kind == CB2_MODEL_EXCITATION_LINE || mo->kind == CB2_MODEL_RECOMBINATION_LINE) {
            if (mo->species < 0 || mo->species >= d->n_species)
                return fail(CB2_ERR_RUNTIME, "The plasma object does not contain the ion species for the specified line");
            pec_table_build(&s->pec[m], &mo->pec, mo->wavelength);
        } else if (mo->kind == CB2_MODEL_BREMSSTRAHLUNG) {
            if (d->gaunt.n_u <= 0) return fail(CB2_ERR_RUNTIME, "Bremsstrahlung needs a free-free Gaunt factor table");
        } else if (mo->kind == CB2_MODEL_THERMAL_CX_LINE) {
            if (mo->species < 0 || mo->species >= d->n_species || !mo->ext)
                return fail(CB2_ERR_RUNTIME, "The plasma object does not contain the ion species for the specified CX line");
            if (!s->tcx) s->tcx = (tcx_table**)calloc((size_t)d->n_models, sizeof(tcx_table*));
            s->tcx[m] = (tcx_table*)calloc(mo->ext->n_donors > 0 ? mo->ext->n_donors : 1, sizeof(tcx_table));
            for (int k = 0; k < mo->ext->n_donors; k++) {
                const cb2_rate3d* r3 = &mo->ext->donor_rates[k];
                if (r3->n_ne > 0 && (r3->n_ne < 2 || r3->n_te < 2 || r3->n_td < 2)) return fail(CB2_ERR_VALUE, "thermal CX rate tables need at least 2 knots per axis");
                tcx_table_build(&s->tcx[m][k], r3, mo->wavelength);
            }
        } else if (mo->kind == CB2_MODEL_TOTAL_RADIATED_POWER) {
            if (!mo->ext) return fail(CB2_ERR_RUNTIME, "TotalRadiatedPower needs its resolved species and rates");
            if (!s->trp) s->trp = (pec_table*)calloc(3 * (size_t)d->n_models, sizeof(pec_table));
            if (mo->ext->has_plt) power_table_build(&s->trp[3 * m], &mo->ext->plt);
            if (mo->ext->has_prb) power_table_build(&s->trp[3 * m + 1], &mo->ext->prb);
            if (mo->ext->has_prc) power_table_build(&s->trp[3 * m + 2], &mo->ext->prc);
        } else if (mo->kind == CB2_MODEL_BEAM_CX_LINE) {
            if (!d->beam) return fail(CB2_ERR_RUNTIME, "The emission model is not connected to a beam object.");
            if (mo->species < 0 || mo->species >= d->n_species || !mo->ext || mo->ext->n_cx < 1)
                return fail(CB2_ERR_RUNTIME, "The plasma object does not contain the ion species for the specified CX line");
            if (mo->ext->n_cx > 4) return fail(CB2_ERR_VALUE, "at most four donor metastables are supported");
            if (mo->ext->n_cx > 1 && !mo->ext->cx_population) return fail(CB2_ERR_RUNTIME, "excited donor metastables need their beam population rates");
        } else if (mo->kind == CB2_MODEL_BEAM_EMISSION_LINE) {
            if (!d->beam) return fail(CB2_ERR_RUNTIME, "The emission model is not connected to a beam object.");
            if (!mo->ext) return fail(CB2_ERR_RUNTIME, "BeamEmissionLine needs its resolved rates");
        } else return fail(CB2_ERR_TYPE, "unsupported model kind");
    }
    gaunt_table_build(&s->gaunt, &d->gaunt);
    if (d->beam) {
        for (int m = 0; m < d->n_models; m++)
            if (d->models[m].kind != CB2_MODEL_BEAM_CX_LINE && d->models[m].kind != CB2_MODEL_BEAM_EMISSION_LINE)
                return fail(CB2_ERR_TYPE, "a beam scene renders beam models only");
        if (!(d->beam->energy > 0)) return fail(CB2_ERR_VALUE, "Beam energy must be positive");
        return beam_ctx_build(s);
    }
    return CB2_OK;
}

static void scene_ctx_free(scene_ctx* s) {
    if (s->has_ax) axisym_ctx_free(&s->ax);
    if (s->pec) { for (int m = 0; m < s->d->n_models; m++) pec_table_free(&s->pec[m]); free(s->pec); }
    if (s->trp) { for (int m = 0; m < 3 * s->d->n_models; m++) pec_table_free(&s->trp[m]); free(s->trp); }
    if (s->tcx) {
        for (int m = 0; m < s->d->n_models; m++)
            if (s->tcx[m]) { for (int k = 0; k < s->d->models[m].ext->n_donors; k++) tcx_table_free(&s->tcx[m][k]); free(s->tcx[m]); }
        free(s->tcx);
    }
    gaunt_table_free(&s->gaunt);
    free(s->axis_z); free(s->axis_density);
}

static void xform_point(const double m[12], const double p[3], double o[3]);
static void xform_vector(const double m[12], const double p[3], double o[3]);

/* =================================================================================================
 * Beam: rates, attenuation, density, direction (beam/node.pyx, attenuator/singleray.pyx, openadas/rates/{beam,cx}.pyx)
 * ============================================================================================== */
#define EVAMU_TO_MS2 (2.0 * ELEMENTARY_CHARGE / ATOMIC_MASS)   /* EvAmuToMS.conversion_factor, conversion.py:31 */

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* BeamStoppingRate.evaluate — openadas/rates/beam.pyx:93-103 */
/* wavelength > 0: BeamEmissionPEC, 'sen' is converted from photon m^3/s to W m^3 first (beam.pyx:229-239) */
static double beam_rate_eval_w(const cb2_beam_rate* r, double wavelength, double energy, double density, double temperature, int64_t* ood) {
    if (r->n_e <= 0) return r->constant;
    if (energy <= 0 || density <= 0 || temperature <= 0) return 0.0;
    double le = log10(energy), ln = log10(density), lt = log10(temperature);
    double lo, hi;
    double *xe = (double*)malloc(sizeof(double) * (r->n_e + r->n_n + r->n_t + r->n_e * r->n_n + r->n_t));
    double *xn = xe + r->n_e, *xt = xn + r->n_n, *lsen = xt + r->n_t, *lst = lsen + r->n_e * r->n_n;
    for (int i = 0; i < r->n_e; i++) xe[i] = log10(r->e[i]);
    for (int i = 0; i < r->n_n; i++) xn[i] = log10(r->n[i]);
    for (int i = 0; i < r->n_t; i++) xt[i] = log10(r->t[i]);
    {
        double conv = wavelength > 0 ? PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9 / wavelength : 1.0;   /* PhotonToJ */
        for (int i = 0; i < r->n_e * r->n_n; i++) lsen[i] = log10(r->sen[i] * conv);
    }
    for (int i = 0; i < r->n_t; i++) lst[i] = log10(r->st[i] / r->sref);
    double a, b;
    if (r->extrapolate) {
        /* extrapolate=True: 'linear' for the 2-D part, 'quadratic' for the 1-D parts (beam.pyx:73-84) */
        if (r->n_e == 1 && r->n_n == 1) a = lsen[0];
        else if (r->n_e == 1) a = interp1d_cubic_quadratic(xn, lsen, r->n_n, ln);
        else if (r->n_n == 1) a = interp1d_cubic_quadratic(xe, lsen, r->n_e, le);
        else a = interp2d_cubic_linear(xe, xn, lsen, r->n_e, r->n_n, le, ln);
        b = r->n_t > 1 ? interp1d_cubic_quadratic(xt, lst, r->n_t, lt) : lst[0];
    } else {
        /* 'none': the reference raises ValueError; here the argument is clamped to the edge and the lookup counted */
        lo = xe[0]; hi = xe[r->n_e - 1]; if (le < lo || le > hi) { (*ood)++; le = clampd(le, lo, hi); }
        lo = xn[0]; hi = xn[r->n_n - 1]; if (ln < lo || ln > hi) { (*ood)++; ln = clampd(ln, lo, hi); }
        lo = xt[0]; hi = xt[r->n_t - 1]; if (lt < lo || lt > hi) { (*ood)++; lt = clampd(lt, lo, hi); }
        if (r->n_e == 1 && r->n_n == 1) a = lsen[0];
        else if (r->n_e == 1) a = cb2o_interp1d_cubic(xn, lsen, r->n_n, ln, 1);
        else if (r->n_n == 1) a = cb2o_interp1d_cubic(xe, lsen, r->n_e, le, 1);
        else a = cb2o_interp2d_cubic(xe, xn, lsen, r->n_e, r->n_n, le, ln, 1);
        b = r->n_t > 1 ? cb2o_interp1d_cubic(xt, lst, r->n_t, lt, 1) : lst[0];
    }
    free(xe);
    return pow(10.0, a + b);
}

static double beam_rate_eval(const cb2_beam_rate* r, double energy, double density, double temperature, int64_t* ood) {
    return beam_rate_eval_w(r, 0.0, energy, density, temperature, ood);
}

static double cx_factor(const double* x, const double* q, int n, double scale, double v, int extrapolate, int64_t* ood) {
    if (n == 1) return q[0] * scale;               /* Constant1D */
    /* extrapolate=True: 'nearest' for the four linear-space factors (cx.pyx:97,99-102) */
    if (v < x[0] || v > x[n - 1]) { if (!extrapolate) (*ood)++; v = clampd(v, x[0], x[n - 1]); }
    double* f = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) f[i] = q[i] * scale;
    double r = cb2o_interp1d_cubic(x, f, n, v, 1);
    free(f);
    return r;
}

/* BeamCXPEC.evaluate — openadas/rates/cx.pyx:104-142 */
static double cx_rate_eval(const cb2_cx_rate* r, double wavelength, double energy, double temperature, double density, double zeff,
                           double bmag, int64_t* ood) {
    if (r->n_eb <= 0) return r->constant;
    if (energy <= 0) return 0.0;
    double rate;
    {
        double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9;      /* PhotonToJ */
        double* x = (double*)malloc(sizeof(double) * 2 * r->n_eb);
        double* f = x + r->n_eb;
        for (int i = 0; i < r->n_eb; i++) { x[i] = log10(r->eb[i]); f[i] = log10(r->qeb[i] / wavelength * conv); }
        double le = log10(energy);
        if (r->n_eb == 1) rate = pow(10.0, f[0]);
        else {
            if (r->extrapolate) rate = pow(10.0, interp1d_cubic_quadratic(x, f, r->n_eb, le));    /* 'quadratic', cx.pyx:96,98 */
            else {
                if (le < x[0] || le > x[r->n_eb - 1]) { (*ood)++; le = clampd(le, x[0], x[r->n_eb - 1]); }
                rate = pow(10.0, cb2o_interp1d_cubic(x, f, r->n_eb, le, 1));
            }
        }
        free(x);
    }
    rate *= cx_factor(r->ti, r->qti, r->n_ti, 1.0 / r->qref, temperature, r->extrapolate, ood);
    if (rate <= 0) return 0.0;
    rate *= cx_factor(r->ni, r->qni, r->n_ni, 1.0 / r->qref, density, r->extrapolate, ood);
    if (rate <= 0) return 0.0;
    rate *= cx_factor(r->z, r->qz, r->n_z, 1.0 / r->qref, zeff, r->extrapolate, ood);
    if (rate <= 0) return 0.0;
    rate *= cx_factor(r->b, r->qb, r->n_b, 1.0 / r->qref, bmag, r->extrapolate, ood);
    if (rate <= 0) return 0.0;
    return rate;
}

/* SingleRayAttenuator._beam_stopping — singleray.pyx:266-313 (species with charge 0 are skipped, SURVEY A.7 caveat) */
static double beam_stopping(const scene_ctx* s, const double p[3], const double beam_velocity[3], int64_t* ood) {
    const cb2_scene_desc* d = s->d;
    const cb2_beam_desc* b = d->beam;
    double density_sum = 0;
    for (int k = 0; k < b->n_stopping; k++) {
        const cb2_species* sp = &d->species[b->stopping_species[k]];
        density_sum += (double)sp->charge * sp->charge * eval_scalar(s, &sp->density, p, ood);
    }
    double coeff = 0;
    for (int k = 0; k < b->n_stopping; k++) {
        const cb2_species* sp = &d->species[b->stopping_species[k]];
        if (sp->charge == 0) continue;
        double target_ne = eval_scalar(s, &sp->density, p, ood) * sp->charge;
        double target_ti = eval_scalar(s, &sp->temperature, p, ood);
        double tv[3];
        eval_vector(s, &sp->velocity, p, tv, ood);
        double iv[3] = {beam_velocity[0] - tv[0], beam_velocity[1] - tv[1], beam_velocity[2] - tv[2]};
        double speed = sqrt(iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2]);
        double energy = speed * speed / EVAMU_TO_MS2;
        coeff += target_ne * beam_rate_eval(&b->stopping_rates[k], energy, density_sum / sp->charge, target_ti, ood);
    }
    return coeff;
}

/* SingleRayAttenuator._calc_attenuation / _beam_attenuation — singleray.pyx:182-264 */
static int beam_ctx_build(scene_ctx* s) {
    const cb2_beam_desc* b = s->d->beam;
    int64_t ood = 0;
    int n = 1 + (int)ceil(b->length / b->attenuator_step);
    if (n < 4) n = 4;
    s->n_axis = n;
    s->axis_z = (double*)malloc(sizeof(double) * n);
    s->axis_density = (double*)malloc(sizeof(double) * n);
    double* stop = (double*)malloc(sizeof(double) * n);
    double axis[3] = {0, 0, 1}, dir[3];
    xform_vector(b->beam_to_plasma, axis, dir);
    double dl = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    double speed = sqrt(b->energy * EVAMU_TO_MS2);
    double bv[3] = {dir[0] / dl * speed, dir[1] / dl * speed, dir[2] / dl * speed};
    double n0 = b->power / (b->energy * b->atomic_weight * ELEMENTARY_CHARGE) / speed;
    for (int k = 0; k < n; k++) {
        s->axis_z[k] = (n > 1) ? b->length * k / (n - 1) : 0.0;      /* np.linspace(0, length, nbeam) */
        if (k == n - 1) s->axis_z[k] = b->length;
        double pb[3] = {0, 0, s->axis_z[k]}, pp[3];
        xform_point(b->beam_to_plasma, pb, pp);
        stop[k] = beam_stopping(s, pp, bv, &ood);
    }
    double cum = 0;                                                   /* cumulative_trapezoid(..., initial=0) */
    for (int k = 0; k < n; k++) {
        if (k > 0) cum += 0.5 * (stop[k] + stop[k - 1]) * (s->axis_z[k] - s->axis_z[k - 1]);
        s->axis_density[k] = n0 * exp(-cum / speed);
    }
    free(stop);
    s->tanxdiv = tan(b->divergence_x * M_PI / 180.0);
    s->tanydiv = tan(b->divergence_y * M_PI / 180.0);
    return CB2_OK;
}

/* Beam.density (beam/node.pyx:214-234) -> SingleRayAttenuator.density (singleray.pyx:107-168) */
static double beam_density(const scene_ctx* s, const double pb[3]) {
    const cb2_beam_desc* b = s->d->beam;
    double x = pb[0], y = pb[1], z = pb[2];
    if (z < 0 || z > b->length) return 0.0;
    double s0 = b->sigma * b->sigma;
    double sx = sqrt(s0 + (z * s->tanxdiv) * (z * s->tanxdiv)), sy = sqrt(s0 + (z * s->tanydiv) * (z * s->tanydiv));
    double nr2 = (x / sx) * (x / sx) + (y / sy) * (y / sy);
    if (b->clamp_to_zero && nr2 > b->clamp_sigma * b->clamp_sigma) return 0.0;
    double g = exp(-0.5 * nr2) / (2 * M_PI * sx * sy);
    return interp1d_linear(s->axis_z, s->axis_density, s->n_axis, z) * g;
}

/* Beam.direction — beam/node.pyx:236-279 */
static void beam_direction(const scene_ctx* s, const double pb[3], double out[3]) {
    const cb2_beam_desc* b = s->d->beam;
    double x = pb[0], y = pb[1], z = pb[2];
    if (z <= 0) { out[0] = 0; out[1] = 0; out[2] = 1; return; }
    double ztx = z * z * s->tanxdiv * s->tanxdiv, zty = z * z * s->tanydiv * s->tanydiv, s0 = b->sigma * b->sigma;
    double ex = x * ztx / (s0 + ztx), ey = y * zty / (s0 + zty);
    double l = sqrt(ex * ex + ey * ey + z * z);
    out[0] = ex / l; out[1] = ey / l; out[2] = z / l;
}

/* =================================================================================================
 * Line-shape models (add_line)
 * ============================================================================================== */
typedef struct { int64_t gauss, lorentz, brems, ood; } counters;

static double species_temperature(const scene_ctx* s, int sp, const double p[3], int64_t* ood) {
    return eval_scalar(s, &s->d->species[sp].temperature, p, ood);
}

/* Zeeman pi / sigma component weights shared by zeeman.pyx:139-160 and stark.pyx:323-346 */
static void add_shape(const scene_ctx* s, const cb2_model* mo, double radiance, const double p[3], const double dir[3],
                      double* samples, counters* cn) {
    const cb2_scene_desc* d = s->d;
    const cb2_lineshape* sh = &mo->shape;
    const cb2_spectral_grid* g = &d->grid;
    int sp = mo->species;
    double v[3];

    if (sh->kind == CB2_SHAPE_STARK) { /* StarkBroadenedLine.add_line — stark.pyx:251-348 */
        static const double pg[7] = {1., 0, 0.57575, 0.37902, -0.42519, -0.31525, 0.31718};
        static const double pl[7] = {1., 0.15882, 1.04388, -1.38281, 0.46251, 0.82325, -0.58026};
        static const double pw[6] = {5.14820e-04, 1.38821e+00, -9.60424e-02, -3.83995e-02, -7.40042e-03, -5.47626e-04};
        const double SIGMA2FWHM = 2 * sqrt(2 * log(2.0));
        double ne = eval_scalar(s, &d->electron_density, p, &cn->ood);
        double te = eval_scalar(s, &d->electron_temperature, p, &cn->ood);
        double fwhm_lorentz = (ne > 0 && te > 0) ? sh->param[0] * pow(ne, sh->param[1]) / pow(te, sh->param[2]) : 0;
        double ts = species_temperature(s, sp, p, &cn->ood);
        double fwhm_gauss = ts > 0 ? SIGMA2FWHM * thermal_broadening(mo->wavelength, ts, mo->atomic_weight) : 0;
        if (fwhm_lorentz == 0 && fwhm_gauss == 0) return;
        double fwhm_full, ratio;
        if (fwhm_gauss <= fwhm_lorentz) {
            ratio = fwhm_gauss / fwhm_lorentz; fwhm_full = pg[0];
            for (int i = 1; i < 7; i++) fwhm_full += pg[i] * pow(ratio, i);
            fwhm_full *= fwhm_lorentz;
        } else {
            ratio = fwhm_lorentz / fwhm_gauss; fwhm_full = pl[0];
            for (int i = 1; i < 7; i++) fwhm_full += pl[i] * pow(ratio, i);
            fwhm_full *= fwhm_gauss;
        }
        double sigma = fwhm_full / SIGMA2FWHM;
        double l2t = fwhm_lorentz / fwhm_full, lw;
        if (l2t < 0.01) { lw = 0; fwhm_full = 0; }
        else if (l2t > 0.999) { lw = 1; sigma = 0; }
        else { lw = pw[0]; for (int i = 1; i < 6; i++) lw += pw[i] * pow(log(l2t), i); lw = exp(lw); }
        double gw = 1 - lw;
        eval_vector(s, &d->species[sp].velocity, p, v, &cn->ood);
        double shifted = doppler_shift(mo->wavelength, dir, v);
        double b[3];
        eval_b_field(s, p, b, &cn->ood);
        double bm = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
#define STARK_ADD(R, W)                                                                                              \
    do {                                                                                                             \
        cn->gauss += add_gaussian_line(gw * (R), (W), sigma, g, samples);                                            \
        cn->lorentz += add_lorentzian_line(lw * (R), (W), fwhm_full, g, samples, d->quad_rtol, d->quad_min_order,   \
                                           d->quad_max_order);                                                       \
    } while (0)
        if (bm == 0) {
            if (sh->polarisation != CB2_POL_NO) radiance *= 0.5;
            STARK_ADD(radiance, shifted);
            return;
        }
        double dl = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        double c = (b[0] * dir[0] + b[1] * dir[1] + b[2] * dir[2]) / dl / bm;
        double cos_sqr = c * c, sin_sqr = 1. - cos_sqr;
        if (sh->polarisation != CB2_POL_SIGMA) { double cr = 0.5 * sin_sqr * radiance; STARK_ADD(cr, shifted); }
        if (sh->polarisation != CB2_POL_PI) {
            double cr = (0.25 * sin_sqr + 0.5 * cos_sqr) * radiance;
            double pe = HC_EV_NM / mo->wavelength;
            shifted = doppler_shift(HC_EV_NM / (pe - BOHR_MAGNETON * bm), dir, v);
            STARK_ADD(cr, shifted);
            shifted = doppler_shift(HC_EV_NM / (pe + BOHR_MAGNETON * bm), dir, v);
            STARK_ADD(cr, shifted);
        }
#undef STARK_ADD
        return;
    }

    /* Gaussian family: all start with ts <= 0 -> return (gaussian.pyx:127-129, zeeman.pyx:118-120, multiplet.pyx:98-100) */
    double ts = species_temperature(s, sp, p, &cn->ood);
    if (ts <= 0.0) return;
    eval_vector(s, &d->species[sp].velocity, p, v, &cn->ood);
    double sigma = thermal_broadening(mo->wavelength, ts, mo->atomic_weight);

    if (sh->kind == CB2_SHAPE_GAUSSIAN) {
        cn->gauss += add_gaussian_line(radiance, doppler_shift(mo->wavelength, dir, v), sigma, g, samples);
        return;
    }
    if (sh->kind == CB2_SHAPE_MULTIPLET) { /* multiplet.pyx:93-117 */
        int n = sh->n_components;
        for (int i = 0; i < n; i++)
            cn->gauss += add_gaussian_line(radiance * sh->multiplet[n + i], doppler_shift(sh->multiplet[i], dir, v), sigma, g, samples);
        return;
    }
    /* Zeeman family */
    if (sh->kind == CB2_SHAPE_PARAM_ZEEMAN) sigma *= sqrt(1. + sh->param[1] * sh->param[1] * pow(ts, 2. * sh->param[2])); /* zeeman.pyx:236 */
    double b[3];
    eval_b_field(s, p, b, &cn->ood);
    double bm = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    if (bm == 0) {
        double rr = sh->polarisation == CB2_POL_NO ? radiance : 0.5 * radiance;
        cn->gauss += add_gaussian_line(rr, doppler_shift(mo->wavelength, dir, v), sigma, g, samples);
        return;
    }
    double dl = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    double c = (b[0] * dir[0] + b[1] * dir[1] + b[2] * dir[2]) / dl / bm;
    double cos_sqr = c * c, sin_sqr = 1. - cos_sqr;
    double r_pi = 0.5 * sin_sqr * radiance, r_sigma = (0.25 * sin_sqr + 0.5 * cos_sqr) * radiance;
    if (sh->kind == CB2_SHAPE_ZEEMAN_TRIPLET || sh->kind == CB2_SHAPE_PARAM_ZEEMAN) {
        if (sh->polarisation != CB2_POL_SIGMA)
            cn->gauss += add_gaussian_line(r_pi, doppler_shift(mo->wavelength, dir, v), sigma, g, samples);
        if (sh->polarisation != CB2_POL_PI) {
            double w1, w2;
            if (sh->kind == CB2_SHAPE_ZEEMAN_TRIPLET) { /* zeeman.pyx:152-158 */
                double pe = HC_EV_NM / mo->wavelength;
                w1 = HC_EV_NM / (pe - BOHR_MAGNETON * bm); w2 = HC_EV_NM / (pe + BOHR_MAGNETON * bm);
            } else { /* zeeman.pyx:260-264 */
                w1 = mo->wavelength + 0.5 * sh->param[0] * bm; w2 = mo->wavelength - 0.5 * sh->param[0] * bm;
            }
            cn->gauss += add_gaussian_line(r_sigma, doppler_shift(w1, dir, v), sigma, g, samples);
            cn->gauss += add_gaussian_line(r_sigma, doppler_shift(w2, dir, v), sigma, g, samples);
        }
        return;
    }
    if (sh->kind == CB2_SHAPE_ZEEMAN_MULTIPLET) { /* zeeman.pyx:308-365 + atomic/zeeman.pyx:87-129 */
        int offs[4] = {0, sh->n_pi, sh->n_pi + sh->n_sigma_plus, sh->n_pi + sh->n_sigma_plus + sh->n_sigma_minus};
        for (int grp = 0; grp < 3; grp++) {
            if (grp == 0 && sh->polarisation == CB2_POL_SIGMA) continue;
            if (grp > 0 && sh->polarisation == CB2_POL_PI) continue;
            double cr = grp == 0 ? r_pi : r_sigma;
            int n = offs[grp + 1] - offs[grp];
            double ratio_sum = 0;
            for (int i = 0; i < n; i++) ratio_sum += interp1d_linear(sh->b_grid, sh->zeeman_ratio + (size_t)(offs[grp] + i) * sh->n_b, sh->n_b, bm);
            for (int i = 0; i < n; i++) {
                double wl = interp1d_linear(sh->b_grid, sh->zeeman_wavelength + (size_t)(offs[grp] + i) * sh->n_b, sh->n_b, bm);
                double ra = interp1d_linear(sh->b_grid, sh->zeeman_ratio + (size_t)(offs[grp] + i) * sh->n_b, sh->n_b, bm);
                if (ratio_sum > 0) ra /= ratio_sum;
                cn->gauss += add_gaussian_line(cr * ra, doppler_shift(wl, dir, v), sigma, g, samples);
            }
        }
        return;
    }
}

/* =================================================================================================
 * Emission models
 * ============================================================================================== */
typedef struct { const scene_ctx* s; double ne, te; int n; double dens[64]; double charge[64]; } brems_fn;

/* BremsFunction.evaluate — bremsstrahlung.pyx:70-90 */
static double brems_eval(double wvl, void* ctx) {
    const brems_fn* b = (const brems_fn*)ctx;
    static const double EXP_FACTOR = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9 / ELEMENTARY_CHARGE;
    double BREMS_CONST = pow(ELEMENTARY_CHARGE * ELEMENTARY_CHARGE * RECIP_4_PI / VACUUM_PERMITTIVITY, 3);
    BREMS_CONST *= 32 * M_PI * M_PI / (3 * sqrt(3.0) * ELECTRON_REST_MASS * ELECTRON_REST_MASS * SPEED_OF_LIGHT * SPEED_OF_LIGHT * SPEED_OF_LIGHT);
    BREMS_CONST *= sqrt(2 * ELECTRON_REST_MASS / (M_PI * ELEMENTARY_CHARGE));
    BREMS_CONST *= SPEED_OF_LIGHT * 1e9 * RECIP_4_PI;
    double ni_gff_z2 = 0;
    for (int i = 0; i < b->n; i++) {
        double z = b->charge[i], ni = b->dens[i];
        if (ni > 0) ni_gff_z2 += ni * gaunt_eval(&b->s->gaunt, z, b->te, wvl) * z * z;
    }
    double pre = BREMS_CONST / (sqrt(b->te) * wvl * wvl) * b->ne * ni_gff_z2;
    return pre * exp(-EXP_FACTOR / (b->te * wvl));
}

/* PlasmaMaterial.emission_function — plasma/material.pyx:48-63: all models at one (plasma-space) point */
/* BeamMaterial.emission_function (beam/material.pyx:49-71) + BeamCXLine.emission (charge_exchange.pyx:117-167): pb and dirb
 * are the sample point and the ray direction in the BEAM frame */
static void beam_emission_function(const scene_ctx* s, const double pb[3], const double dirb[3], double* samples, counters* cn) {
    const cb2_scene_desc* d = s->d;
    const cb2_beam_desc* b = d->beam;
    double bdir_b[3], bdir[3], p[3], obs[3];
    beam_direction(s, pb, bdir_b);
    xform_point(b->beam_to_plasma, pb, p);
    xform_vector(b->beam_to_plasma, bdir_b, bdir);
    xform_vector(b->beam_to_plasma, dirb, obs);
    for (int m = 0; m < d->n_models; m++) {
        const cb2_model* mo = &d->models[m];
        double donor = beam_density(s, pb);
        if (donor == 0.0) continue;
        if (mo->kind == CB2_MODEL_BEAM_EMISSION_LINE) {
            /* BeamEmissionLine.emission / _beam_emission_rate (beam_emission.pyx:100-176); neutrals are skipped */
            const cb2_model_ext* x = mo->ext;
            double bl = sqrt(bdir[0] * bdir[0] + bdir[1] * bdir[1] + bdir[2] * bdir[2]);
            double speed = sqrt(b->energy * EVAMU_TO_MS2);
            double bv[3] = {bdir[0] / bl * speed, bdir[1] / bl * speed, bdir[2] / bl * speed};
            double density_sum = 0, rate = 0;
            for (int k = 0; k < x->n_bes; k++) {
                const cb2_species* sp = &d->species[x->bes_species[k]];
                density_sum += (double)sp->charge * sp->charge * eval_scalar(s, &sp->density, p, &cn->ood);
            }
            for (int k = 0; k < x->n_bes; k++) {
                const cb2_species* sp = &d->species[x->bes_species[k]];
                if (sp->charge == 0) continue;
                double target_ne = eval_scalar(s, &sp->density, p, &cn->ood) * sp->charge;
                double target_ti = eval_scalar(s, &sp->temperature, p, &cn->ood);
                double tv[3];
                eval_vector(s, &sp->velocity, p, tv, &cn->ood);
                double iv[3] = {bv[0] - tv[0], bv[1] - tv[1], bv[2] - tv[2]};
                double e_int = (iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2]) / EVAMU_TO_MS2;
                rate += target_ne * beam_rate_eval_w(&x->bes_rates[k], mo->wavelength, e_int, density_sum / sp->charge, target_ti, &cn->ood);
            }
            double radiance = RECIP_4_PI * donor * rate;
            /* BeamEmissionMultiplet.add_line — mse.pyx:62-135 */
            double te = eval_scalar(s, &d->electron_temperature, p, &cn->ood);
            if (te <= 0) continue;
            double ne = eval_scalar(s, &d->electron_density, p, &cn->ood);
            if (ne <= 0) continue;
            double bf[3];
            eval_b_field(s, p, bf, &cn->ood);
            double cx_[3] = {bv[1] * bf[2] - bv[2] * bf[1], bv[2] * bf[0] - bv[0] * bf[2], bv[0] * bf[1] - bv[1] * bf[0]};
            double stark = fabs(2.77e-8 * sqrt(cx_[0] * cx_[0] + cx_[1] * cx_[1] + cx_[2] * cx_[2]));
            double central = doppler_shift(mo->wavelength, obs, bv);
            double sigma = thermal_broadening(mo->wavelength, b->temperature, b->atomic_weight);
            double rt[4] = {x->mse_ratios[0], x->mse_ratios[1], x->mse_ratios[2], x->mse_ratios[3]};
            if (x->n_mse > 1) { /* ratios as functions of ne (mse.pyx:103-121), tabulated on knots uniform in log10(ne) */
                double f = (log10(ne) - x->mse_lne0) / x->mse_dlne;
                f = f < 0 ? 0 : (f > x->n_mse - 1 ? x->n_mse - 1 : f);
                int i0 = (int)f;
                if (i0 > x->n_mse - 2) i0 = x->n_mse - 2;
                double w = f - i0;
                for (int q = 0; q < 4; q++) rt[q] = x->mse_ratio_tab[q * x->n_mse + i0] + w * (x->mse_ratio_tab[q * x->n_mse + i0 + 1] - x->mse_ratio_tab[q * x->n_mse + i0]);
            }
            double s2p = rt[0], s1s0 = rt[1], p23 = rt[2], p43 = rt[3];
            double dd = 1 / (1 + s2p), isig = s2p * dd * radiance, ipi = 0.5 * dd * radiance;
            double is0 = 1 / (s1s0 + 1), is1 = 0.5 * s1s0 * is0;
            double ip3 = 1 / (1 + p23 + p43), ip2 = p23 * ip3, ip4 = p43 * ip3;
            const double amp[9] = {isig * is0, isig * is1, isig * is1, ipi * ip2, ipi * ip2, ipi * ip3, ipi * ip3, ipi * ip4, ipi * ip4};
            const double off[9] = {0, 1, -1, 2, -2, 3, -3, 4, -4};
            for (int k = 0; k < 9; k++) cn->gauss += add_gaussian_line(amp[k], central + off[k] * stark, sigma, &d->grid, samples);
            continue;
        }
        double nr = eval_scalar(s, &d->species[mo->species].density, p, &cn->ood);
        if (nr == 0) continue;
        double tr = eval_scalar(s, &d->species[mo->species].temperature, p, &cn->ood);
        if (tr == 0) continue;
        double vr[3];
        eval_vector(s, &d->species[mo->species].velocity, p, vr, &cn->ood);
        double bl = sqrt(bdir[0] * bdir[0] + bdir[1] * bdir[1] + bdir[2] * bdir[2]);
        double speed = sqrt(b->energy * EVAMU_TO_MS2);
        double iv[3] = {bdir[0] / bl * speed - vr[0], bdir[1] / bl * speed - vr[1], bdir[2] / bl * speed - vr[2]};
        double ispeed = sqrt(iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2]);
        double energy = ispeed * ispeed / EVAMU_TO_MS2;
        /* _composite_cx_rate (charge_exchange.pyx:204-236); Plasma.ion_density / z_effective (node.pyx:396-462) */
        double ion_density = 0, snz = 0, snz2 = 0;
        for (int i = 0; i < d->n_species; i++) {
            double n = eval_scalar(s, &d->species[i].density, p, &cn->ood);
            ion_density += n;
            if (d->species[i].charge > 0) { snz += n * d->species[i].charge; snz2 += n * d->species[i].charge * d->species[i].charge; }
        }
        double zeff = snz > 0 ? snz2 / snz : 0.0;
        double bf[3];
        eval_b_field(s, p, bf, &cn->ood);
        double bmag = sqrt(bf[0] * bf[0] + bf[1] * bf[1] + bf[2] * bf[2]);
        double rate = cx_rate_eval(&mo->ext->cx[0], mo->wavelength, energy, tr, ion_density, zeff, bmag, &cn->ood);
        if (mo->ext->n_cx > 1) {
            /* excited donor metastables weighted by their population relative to the ground state:
             * _composite_cx_rate :215-234 and _beam_population :241-292 (species of charge 0 skipped) */
            double bv[3] = {bdir[0] / bl * speed, bdir[1] / bl * speed, bdir[2] / bl * speed};
            double density_sum = 0, total_population = 1;
            for (int i = 0; i < d->n_species; i++)
                density_sum += (double)d->species[i].charge * d->species[i].charge * eval_scalar(s, &d->species[i].density, p, &cn->ood);
            for (int k = 1; k < mo->ext->n_cx; k++) {
                double pop_coeff = 0, total_ne = 0;
                for (int i = 0; i < d->n_species; i++) {
                    const cb2_species* sp = &d->species[i];
                    if (sp->charge == 0) continue;
                    double target_ne = eval_scalar(s, &sp->density, p, &cn->ood) * sp->charge;
                    double target_ti = eval_scalar(s, &sp->temperature, p, &cn->ood);
                    double tv[3];
                    eval_vector(s, &sp->velocity, p, tv, &cn->ood);
                    double jv[3] = {bv[0] - tv[0], bv[1] - tv[1], bv[2] - tv[2]};
                    double e_int = (jv[0] * jv[0] + jv[1] * jv[1] + jv[2] * jv[2]) / EVAMU_TO_MS2;
                    pop_coeff += target_ne * beam_rate_eval(&mo->ext->cx_population[(size_t)(k - 1) * d->n_species + i], e_int,
                                                            density_sum / sp->charge, target_ti, &cn->ood);
                    total_ne += target_ne;
                }
                double population = total_ne > 0 ? pop_coeff / total_ne : 0.0;
                rate += population * cx_rate_eval(&mo->ext->cx[k], mo->wavelength, energy, tr, ion_density, zeff, bmag, &cn->ood);
                total_population += population;
            }
            rate /= total_population;
        }
        add_shape(s, mo, RECIP_4_PI * donor * nr * rate, p, obs, samples, cn);
    }
}

static void emission_function(const scene_ctx* s, const double p[3], const double dir[3], double* samples, counters* cn) {
    const cb2_scene_desc* d = s->d;
    if (d->beam) { beam_emission_function(s, p, dir, samples, cn); return; }
    for (int m = 0; m < d->n_models; m++) {
        const cb2_model* mo = &d->models[m];
        double ne = eval_scalar(s, &d->electron_density, p, &cn->ood);
        if (ne <= 0) continue;
        double te = eval_scalar(s, &d->electron_temperature, p, &cn->ood);
        if (te <= 0) continue;
        if (mo->kind == CB2_MODEL_BREMSSTRAHLUNG) { /* bremsstrahlung.pyx:169-208 */
            brems_fn b; b.s = s; b.ne = ne; b.te = te; b.n = 0;
            for (int i = 0; i < d->n_species && b.n < 64; i++)
                if (d->species[i].charge > 0) {
                    b.charge[b.n] = d->species[i].charge;
                    b.dens[b.n] = eval_scalar(s, &d->species[i].density, p, &cn->ood);
                    b.n++;
                }
            double delta = (d->grid.max_wavelength - d->grid.min_wavelength) / d->grid.bins;
            double lower = d->grid.min_wavelength;
            for (int i = 0; i < d->grid.bins; i++) {
                double upper = d->grid.min_wavelength + delta * (i + 1);
                samples[i] += cb2o_gauss_legendre(brems_eval, &b, lower, upper, d->quad_rtol, d->quad_min_order, d->quad_max_order) / delta;
                lower = upper;
            }
            cn->brems += d->grid.bins;
            continue;
        }
        if (mo->kind == CB2_MODEL_TOTAL_RADIATED_POWER) { /* total_radiated_power.pyx:70-118 */
            const cb2_model_ext* x = mo->ext;
            double ni = eval_scalar(s, &d->species[x->line_rad_species].density, p, &cn->ood);
            double ni_upper = eval_scalar(s, &d->species[x->recom_species].density, p, &cn->ood);
            double nhyd = 0;
            for (int k = 0; k < x->n_hydrogen; k++) nhyd += eval_scalar(s, &d->species[x->hydrogen_species[k]].density, p, &cn->ood);
            double power = 0;
            if (x->has_plt && ni > 0) power += pec_eval(&s->trp[3 * m], &x->plt, ne, te, &cn->ood) * ne * ni;
            if (x->has_prb && ni_upper > 0) power += pec_eval(&s->trp[3 * m + 1], &x->prb, ne, te, &cn->ood) * ne * ni_upper;
            if (x->has_prc && ni_upper > 0 && nhyd > 0) power += pec_eval(&s->trp[3 * m + 2], &x->prc, ne, te, &cn->ood) * nhyd * ni_upper;
            double radiance = RECIP_4_PI * power / (d->grid.max_wavelength - d->grid.min_wavelength);
            for (int i = 0; i < d->grid.bins; i++) samples[i] += radiance;
            continue;
        }
        if (mo->kind == CB2_MODEL_THERMAL_CX_LINE) { /* thermal_cx.pyx:79-112 */
            const cb2_model_ext* x = mo->ext;
            double nr = eval_scalar(s, &d->species[mo->species].density, p, &cn->ood);
            if (nr <= 0) continue;
            double weighted = 0;
            for (int k = 0; k < x->n_donors; k++) {
                double nd = eval_scalar(s, &d->species[x->donor_species[k]].density, p, &cn->ood);
                double td = eval_scalar(s, &d->species[x->donor_species[k]].temperature, p, &cn->ood);
                weighted += nd * tcx_eval(&s->tcx[m][k], &x->donor_rates[k], ne, te, td, &cn->ood);
            }
            add_shape(s, mo, RECIP_4_PI * weighted * nr, p, dir, samples, cn);
            continue;
        }
        /* ExcitationLine / RecombinationLine — impact_excitation.pyx:86-100, recombination.pyx:86-100 */
        double ni = eval_scalar(s, &d->species[mo->species].density, p, &cn->ood);
        if (ni <= 0) continue;
        double radiance = RECIP_4_PI * pec_eval(&s->pec[m], &mo->pec, ne, te, &cn->ood) * ne * ni;
        add_shape(s, mo, radiance, p, dir, samples, cn);
    }
}

static void xform_point(const double m[12], const double p[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = m[4 * i] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2] + m[4 * i + 3];
}
static void xform_vector(const double m[12], const double p[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = m[4 * i] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2];
}

/* NumericalIntegrator.integrate [raysect 0.8.1] — SURVEY Appendix B.2: trapezium rule over one segment */
static int64_t integrate_segment(const scene_ctx* s, const double o[3], const double dvec[3], double t0, double t1,
                                 double* spectrum, double* cur, double* prev, counters* cn) {
    const cb2_scene_desc* d = s->d;
    int bins = d->grid.bins;
    double start[3], end[3], dir[3];
    /* raysect hands the integrator start_point = far end (where light enters the volume) and end_point = near end,
       so that ray_direction = -(integration direction) */
    for (int k = 0; k < 3; k++) { start[k] = o[k] + t1 * dvec[k]; end[k] = o[k] + t0 * dvec[k]; }
    double sl[3], el[3];
    xform_point(d->world_to_plasma, start, sl);
    xform_point(d->world_to_plasma, end, el);
    double iv[3] = {el[0] - sl[0], el[1] - sl[1], el[2] - sl[2]};
    double length = sqrt(iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2]);
    if (length == 0) return 0;
    for (int k = 0; k < 3; k++) iv[k] /= length;
    xform_vector(d->world_to_plasma, dvec, dir); /* ray direction in plasma space */
    int intervals = (int)ceil(length / d->step);
    if (intervals < d->min_samples - 1) intervals = d->min_samples - 1;
    if (intervals < 1) intervals = 1;
    double h = length / intervals;
    for (int k = 0; k <= intervals; k++) {
        double t = k * h, p[3] = {sl[0] + t * iv[0], sl[1] + t * iv[1], sl[2] + t * iv[2]};
        memset(cur, 0, sizeof(double) * bins);
        emission_function(s, p, dir, cur, cn);
        if (k > 0) for (int i = 0; i < bins; i++) spectrum[i] += 0.5 * h * (cur[i] + prev[i]);
        double* tmp = prev; prev = cur; cur = tmp;
    }
    return (int64_t)intervals + 1;
}

int cb2o_emission_render(const cb2_scene_desc* desc, const cb2_rays* rays, double* out,
                         double scale, int accumulate, int n_threads, cb2_stats* stats) {
    scene_ctx s;
    int rc = scene_ctx_build(&s, desc);
    if (rc != CB2_OK) { scene_ctx_free(&s); return rc; }
    int bins = desc->grid.bins;
    int64_t tot_samples = 0, tot_g = 0, tot_l = 0, tot_b = 0, tot_ood = 0;
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
#pragma omp parallel num_threads(n_threads) reduction(+ : tot_samples, tot_g, tot_l, tot_b, tot_ood)
    {
        double* spectrum = (double*)malloc(sizeof(double) * bins * 3);
        double *cur = spectrum + bins, *prev = spectrum + 2 * bins;
#pragma omp for schedule(dynamic, 1)
        for (int64_t r = 0; r < rays->n_rays; r++) {
            memset(spectrum, 0, sizeof(double) * bins);
            counters cn = {0, 0, 0, 0};
            for (int64_t sg = rays->seg_offset[r]; sg < rays->seg_offset[r + 1]; sg++)
                tot_samples += integrate_segment(&s, rays->origin + 3 * r, rays->direction + 3 * r, rays->seg_t0[sg], rays->seg_t1[sg],
                                                 spectrum, cur, prev, &cn);
            double* o = out + r * (int64_t)bins;
            for (int i = 0; i < bins; i++) o[i] = (accumulate ? o[i] : 0.0) + scale * spectrum[i];
            tot_g += cn.gauss; tot_l += cn.lorentz; tot_b += cn.brems; tot_ood += cn.ood;
        }
        free(spectrum);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->samples = tot_samples; stats->gaussian_bin_evals = tot_g; stats->lorentzian_bin_evals = tot_l;
        stats->brems_bin_evals = tot_b; stats->out_of_domain = tot_ood;
    }
    scene_ctx_free(&s);
    return CB2_OK;
}

/* Beam.density / Beam.direction at points given in beam coordinates: out[n][4] */
int cb2o_beam_sample(const cb2_scene_desc* desc, const double* beam_points, int64_t n, double* out) {
    scene_ctx s;
    int rc = scene_ctx_build(&s, desc);
    if (rc != CB2_OK) { scene_ctx_free(&s); return rc; }
    if (!desc->beam) { scene_ctx_free(&s); return fail(CB2_ERR_VALUE, "the scene has no beam"); }
    for (int64_t i = 0; i < n; i++) {
        out[4 * i] = beam_density(&s, beam_points + 3 * i);
        beam_direction(&s, beam_points + 3 * i, out + 4 * i + 1);
    }
    scene_ctx_free(&s);
    return CB2_OK;
}

int cb2o_state_width(const cb2_scene_desc* d) { return 2 + 5 * d->n_species + 3; }

int cb2o_sample_state(const cb2_scene_desc* desc, const double* points, int64_t n, double* out) {
    scene_ctx s;
    int rc = scene_ctx_build(&s, desc);
    if (rc != CB2_OK) { scene_ctx_free(&s); return rc; }
    int w = cb2o_state_width(desc);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        int64_t ood = 0;
        double p[3];
        xform_point(desc->world_to_plasma, points + 3 * i, p);
        double* o = out + i * w;
        o[0] = eval_scalar(&s, &desc->electron_density, p, &ood);
        o[1] = eval_scalar(&s, &desc->electron_temperature, p, &ood);
        for (int k = 0; k < desc->n_species; k++) {
            o[2 + 5 * k] = eval_scalar(&s, &desc->species[k].density, p, &ood);
            o[3 + 5 * k] = eval_scalar(&s, &desc->species[k].temperature, p, &ood);
            eval_vector(&s, &desc->species[k].velocity, p, o + 4 + 5 * k, &ood);
        }
        eval_b_field(&s, p, o + 2 + 5 * desc->n_species, &ood);
    }
    scene_ctx_free(&s);
    return CB2_OK;
}

/* =================================================================================================
 * Ray transfer — cherab/tools/raytransfer/emitters.pyx:88-224
 * ============================================================================================== */
static int64_t rt_integrate(const cb2_rt_desc* d, const double o[3], const double dvec[3], double t0, double t1, double* samples) {
    double sw[3], ew[3], start[3], end[3];
    for (int k = 0; k < 3; k++) { sw[k] = o[k] + t1 * dvec[k]; ew[k] = o[k] + t0 * dvec[k]; } /* far -> near, as raysect */
    xform_point(d->world_to_local, sw, start);
    xform_point(d->world_to_local, ew, end);
    double dir[3] = {end[0] - start[0], end[1] - start[1], end[2] - start[2]};
    double length = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    int n0 = d->grid_shape[0], n1 = d->grid_shape[1], n2 = d->grid_shape[2];
    if (d->integrator == 1) {
        /* a foreign NumericalIntegrator [raysect, SURVEY Appendix B.2] over emission_function (emitters.pyx:452-473, 557-571):
           unit emissivity in the cell of every sample, trapezium rule */
        if (length == 0) return 0;
        for (int k = 0; k < 3; k++) dir[k] /= length;
        int intervals = (int)ceil(length / d->step);
        if (intervals < d->min_samples - 1) intervals = d->min_samples - 1;
        if (intervals < 1) intervals = 1;
        double h = length / intervals;
        for (int k = 0; k <= intervals; k++) {
            double t = k * h;
            double x = start[0] + t * dir[0], y = start[1] + t * dir[1], z = start[2] + t * dir[2];
            int i0, i1, i2;
            if (d->kind == CB2_RT_CYLINDRICAL) {
                i2 = (int)(z / d->grid_steps[2]);
                double r = sqrt(x * x + y * y);
                i0 = (int)((r - d->rmin) / d->grid_steps[0]);
                if (n1 == 1) i1 = 0;
                else {
                    double phi = (180. / M_PI) * atan2(y, x);
                    phi = fmod(phi + 360., d->period);
                    i1 = (int)(phi / d->grid_steps[1]);
                }
            } else {
                i0 = (int)(x / d->grid_steps[0]); i1 = (int)(y / d->grid_steps[1]); i2 = (int)(z / d->grid_steps[2]);
            }
            if (i0 < 0 || i0 >= n0 || i1 < 0 || i1 >= n1 || i2 < 0 || i2 >= n2) continue;   /* (the reference does not check bounds here) */
            int src = d->voxel_map[((int64_t)i0 * n1 + i1) * n2 + i2];
            if (src < 0) continue;
            samples[src] += (k == 0 || k == intervals) ? 0.5 * h : h;
        }
        return (int64_t)intervals + 1;
    }
    if (length < 0.1 * d->step) return 0;
    for (int k = 0; k < 3; k++) dir[k] /= length;
    int n = (int)(length / d->step);
    if (n < d->min_samples) n = d->min_samples;
    double dt = length / n;
    int i0c = -1, i1c = -1, i2c = -1, isource = -1, isource_current = -1;
    double res = 0;
    for (int it = 0; it < n; it++) {
        double t = (it + 0.5) * dt;
        double x = start[0] + dir[0] * t, y = start[1] + dir[1] * t, z = start[2] + dir[2] * t;
        int i0, i1, i2;
        if (d->kind == CB2_RT_CYLINDRICAL) {
            i2 = (int)(z / d->grid_steps[2]);
            double r = sqrt(x * x + y * y);
            i0 = (int)((r - d->rmin) / d->grid_steps[0]);
            if (n1 == 1) i1 = 0;
            else {
                double phi = (180. / M_PI) * atan2(y, x);
                phi = fmod(phi + 360., d->period);
                i1 = (int)(phi / d->grid_steps[1]);
            }
        } else {
            i0 = (int)(x / d->grid_steps[0]); i1 = (int)(y / d->grid_steps[1]); i2 = (int)(z / d->grid_steps[2]);
        }
        if (i0 != i0c || i1 != i1c || i2 != i2c) {
            i0c = i0; i1c = i1; i2c = i2;
            /* the reference indexes voxel_map unchecked-but-bounds-checked (IndexError); out-of-grid cells are skipped here */
            if (i0 < 0 || i0 >= n0 || i1 < 0 || i1 >= n1 || i2 < 0 || i2 >= n2) isource = -1;
            else isource = d->voxel_map[((int64_t)i0 * n1 + i1) * n2 + i2];
            if (isource != isource_current) {
                if (isource_current > -1) samples[isource_current] += res;
                isource_current = isource;
                res = 0;
            }
        }
        if (isource_current > -1) res += dt;
    }
    if (isource_current > -1) samples[isource_current] += res;
    return n;
}

int cb2o_rt_render_dense(const cb2_rt_desc* desc, const cb2_rays* rays, double* out, int accumulate,
                         int n_threads, cb2_stats* stats) {
    if (desc->abi_version != CB2_ABI_VERSION) return fail(CB2_ERR_VALUE, "abi_version mismatch");
    if (desc->step <= 0) return fail(CB2_ERR_VALUE, "Numerical integration step size can not be less than or equal to zero.");
    if (desc->min_samples < 2) return fail(CB2_ERR_VALUE, "At least two samples are required to perform the numerical integration.");
    int64_t steps = 0;
    (void)n_threads;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads) reduction(+ : steps)
    for (int64_t r = 0; r < rays->n_rays; r++) {
        double* row = out + r * (int64_t)desc->bins;
        if (!accumulate) memset(row, 0, sizeof(double) * desc->bins);
        for (int64_t sg = rays->seg_offset[r]; sg < rays->seg_offset[r + 1]; sg++)
            steps += rt_integrate(desc, rays->origin + 3 * r, rays->direction + 3 * r, rays->seg_t0[sg], rays->seg_t1[sg], row);
    }
    if (stats) { memset(stats, 0, sizeof *stats); stats->rt_steps = steps; }
    return CB2_OK;
}

/* =================================================================================================
 * First-wall occlusion — the role of the wall meshes of cherab/generomak/machine/first_wall.py:120-184 in Raysect's tracer:
 * distance to the first opaque hit.  Brute force: every ray against every triangle, Moeller-Trumbore in float64.
 * [raysect: its Mesh primitive stores float32 vertices and uses a watertight test — hit distances of the real reference carry
 * float32 rounding; parity unpinned]
 * ============================================================================================== */
static double tri_hit(const double* v, const double o[3], const double d[3], double t_min) {
    double e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
    double e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
    double px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
    double det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.0) return INFINITY;
    double inv = 1.0 / det;
    double sx = o[0] - v[0], sy = o[1] - v[1], sz = o[2] - v[2];
    double u = (sx * px + sy * py + sz * pz) * inv;
    if (u < 0.0 || u > 1.0) return INFINITY;
    double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
    double w = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
    if (w < 0.0 || u + w > 1.0) return INFINITY;
    double t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    return t > t_min ? t : INFINITY;
}

int cb2o_wall_hit(const double* vertices, int64_t n_triangles, const double* origin, const double* direction, int64_t n, double* t_hit) {
    if (!vertices || !origin || !direction || !t_hit) return fail(CB2_ERR_VALUE, "null argument");
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; i++) {
        double best = INFINITY;
        for (int64_t k = 0; k < n_triangles; k++) {
            double t = tri_hit(vertices + 9 * k, origin + 3 * i, direction + 3 * i, 1e-9);
            if (t < best) best = t;
        }
        t_hit[i] = best;
    }
    return CB2_OK;
}
