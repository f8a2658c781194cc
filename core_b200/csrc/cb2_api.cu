// cb2_api.cu — C ABI entry points and the host-side scene builder of libcherab_b200.so.
//
// cb2_scene_create turns the flat descriptor (include/cherab_b200.h) into device-resident fp32 tables: this is the
// "interpolators flattened into device-resident tables" step of the north star.  The cubic tables restate Raysect's
// Interpolator1DArray/2DArray 'cubic' (local Hermite, 2nd-order finite-difference knot derivatives — SURVEY
// Appendix B.4); the coefficients are computed here in fp64 and only then rounded to fp32.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "cb2_internal.h"

// ------------------------------------------------------------------------------------------------------------------
// error channel
// ------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

int cb2_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cb2_cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CB2_OK;
    return cb2_fail(CB2_ERR_CUDA, "CUDA error %s in %s (libcherab_b200 has no CPU fallback)", cudaGetErrorString(e), what);
}

extern "C" const char* cb2_last_error(void) { return g_err; }
extern "C" int cb2_abi_version(void) { return CB2_ABI_VERSION; }
extern "C" int cb2_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cb2_cuda_check(e, "cudaGetDeviceCount");
        return -1;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------------------------
// device arena
// ------------------------------------------------------------------------------------------------------------------
struct Arena {
    std::vector<void*> ptrs;
    int rc = CB2_OK;
    template <typename T>
    const T* upload(const std::vector<T>& v) {
        if (rc != CB2_OK) return nullptr;
        void* p = nullptr;
        size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
        rc = cb2_cuda_check(cudaMalloc(&p, bytes), "cudaMalloc(table)");
        if (rc != CB2_OK) return nullptr;
        ptrs.push_back(p);
        if (!v.empty()) rc = cb2_cuda_check(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(table)");
        return (const T*)p;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// cubic table construction (Raysect semantics restated, fp64)
// ------------------------------------------------------------------------------------------------------------------
// first derivative at knot i of samples f[i*stride] on an uneven grid: 3-point 2nd-order in the interior,
// one-sided first difference at the ends
static double d1(const double* x, const double* f, int n, int stride, int i) {
    if (n < 2) return 0.0;
    if (i == 0) return (f[stride] - f[0]) / (x[1] - x[0]);
    if (i == n - 1) return (f[(size_t)(n - 1) * stride] - f[(size_t)(n - 2) * stride]) / (x[n - 1] - x[n - 2]);
    const double a = x[i] - x[i - 1], b = x[i + 1] - x[i];
    const double fm = f[(size_t)(i - 1) * stride], f0 = f[(size_t)i * stride], fp = f[(size_t)(i + 1) * stride];
    return (a * a * fp - b * b * fm + (b * b - a * a) * f0) / (a * b * (a + b));
}

// mixed derivative at knot (i,j): four-corner difference over the neighbours that exist
static double d2(const double* x, const double* y, const double* f, int nx, int ny, int i, int j) {
    const int i0 = std::max(i - 1, 0), i1 = std::min(i + 1, nx - 1), j0 = std::max(j - 1, 0), j1 = std::min(j + 1, ny - 1);
    if (i0 == i1 || j0 == j1) return 0.0;
    return (f[(size_t)i1 * ny + j1] - f[(size_t)i1 * ny + j0] - f[(size_t)i0 * ny + j1] + f[(size_t)i0 * ny + j0]) /
           ((x[i1] - x[i0]) * (y[j1] - y[j0]));
}

static bool is_uniform(const double* x, int n) {
    if (n < 2) return true;
    const double h = (x[n - 1] - x[0]) / (n - 1);
    for (int i = 0; i < n; i++)
        if (fabs(x[i] - (x[0] + i * h)) > 1e-9 * fabs(h)) return false;
    return true;
}

// Hermite basis matrix: p(t) = sum_k a_k t^k with (f0, f1, d0, d1) -> a = H * (f0, f1, d0, d1)
static void hermite_coef(double f0, double f1, double e0, double e1, double a[4]) {
    a[0] = f0;
    a[1] = e0;
    a[2] = 3.0 * (f1 - f0) - 2.0 * e0 - e1;
    a[3] = 2.0 * (f0 - f1) + e0 + e1;
}

static DevTable1D make_knots1d(Arena& A, const double* x, int n) {
    DevTable1D t;
    memset(&t, 0, sizeof t);
    t.n = n;
    std::vector<float> xs(n), iw(std::max(n - 1, 1));
    for (int i = 0; i < n; i++) xs[i] = (float)x[i];
    for (int i = 0; i + 1 < n; i++) iw[i] = (float)(1.0 / (x[i + 1] - x[i]));
    t.xmin = (float)x[0];
    t.xmax = (float)x[n - 1];
    t.uniform = n > 1 && is_uniform(x, n);
    t.x0 = (float)x[0];
    t.inv_dx = n > 1 ? (float)((n - 1) / (x[n - 1] - x[0])) : 0.f;
    t.x = A.upload(xs);
    t.inv_w = A.upload(iw);
    return t;
}

static const float4* make_coef1d(Arena& A, const double* x, const double* f, int n, double scale) {
    std::vector<float4> c(std::max(n - 1, 1));
    if (n == 1) c[0] = make_float4((float)(f[0] * scale), 0, 0, 0);
    for (int i = 0; i + 1 < n; i++) {
        const double h = x[i + 1] - x[i];
        double a[4];
        hermite_coef(f[i], f[i + 1], d1(x, f, n, 1, i) * h, d1(x, f, n, 1, i + 1) * h, a);
        c[i] = make_float4((float)(a[0] * scale), (float)(a[1] * scale), (float)(a[2] * scale), (float)(a[3] * scale));
    }
    return A.upload(c);
}

// fp64 per-cell polynomial coefficients of the bicubic: c[cell*16 + 4*p + q] multiplies t^p u^q (t along x, u along y)
static void build_coef2d(const double* x, const double* y, const double* f, int nx, int ny, std::vector<double>& coef) {
    std::vector<double> fx((size_t)nx * ny), fy((size_t)nx * ny), fxy((size_t)nx * ny);
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++) {
            fx[(size_t)i * ny + j] = d1(x, f + j, nx, ny, i);
            fy[(size_t)i * ny + j] = d1(y, f + (size_t)i * ny, ny, 1, j);
            fxy[(size_t)i * ny + j] = d2(x, y, f, nx, ny, i, j);
        }
    coef.assign((size_t)(nx - 1) * (ny - 1) * 16, 0.0);
    for (int i = 0; i + 1 < nx; i++)
        for (int j = 0; j + 1 < ny; j++) {
            const double hx = x[i + 1] - x[i], hy = y[j + 1] - y[j];
            // For every power of t, a Hermite cubic in u: first reduce along x for the four "u-quantities"
            // (value and u-derivative at u=0 and u=1), then along u.
            double cx[4][4];  // [quantity q][power of t]; q: 0 f(.,j) 1 f(.,j+1) 2 fy*hy(.,j) 3 fy*hy(.,j+1)
            for (int q = 0; q < 4; q++) {
                const int jj = j + (q & 1);
                double v0, v1, e0, e1;
                if (q < 2) {
                    v0 = f[(size_t)i * ny + jj];
                    v1 = f[(size_t)(i + 1) * ny + jj];
                    e0 = fx[(size_t)i * ny + jj] * hx;
                    e1 = fx[(size_t)(i + 1) * ny + jj] * hx;
                } else {
                    v0 = fy[(size_t)i * ny + jj] * hy;
                    v1 = fy[(size_t)(i + 1) * ny + jj] * hy;
                    e0 = fxy[(size_t)i * ny + jj] * hx * hy;
                    e1 = fxy[(size_t)(i + 1) * ny + jj] * hx * hy;
                }
                hermite_coef(v0, v1, e0, e1, cx[q]);
            }
            for (int p = 0; p < 4; p++) {  // power of t
                double a[4];
                hermite_coef(cx[0][p], cx[1][p], cx[2][p], cx[3][p], a);
                for (int q = 0; q < 4; q++) coef[((size_t)i * (ny - 1) + j) * 16 + 4 * p + q] = a[q];
            }
        }
}

static DevTable2D make_table2d(Arena& A, const double* x, const double* y, const double* f, int nx, int ny) {
    DevTable2D t;
    memset(&t, 0, sizeof t);
    t.nx = nx;
    t.ny = ny;
    std::vector<float> xs(nx), ys(ny), iwx(std::max(nx - 1, 1)), iwy(std::max(ny - 1, 1));
    for (int i = 0; i < nx; i++) xs[i] = (float)x[i];
    for (int j = 0; j < ny; j++) ys[j] = (float)y[j];
    for (int i = 0; i + 1 < nx; i++) iwx[i] = (float)(1.0 / (x[i + 1] - x[i]));
    for (int j = 0; j + 1 < ny; j++) iwy[j] = (float)(1.0 / (y[j + 1] - y[j]));
    t.uniform = is_uniform(x, nx) && is_uniform(y, ny);
    t.x0 = (float)x[0];
    t.y0 = (float)y[0];
    t.inv_dx = (float)((nx - 1) / (x[nx - 1] - x[0]));
    t.inv_dy = (float)((ny - 1) / (y[ny - 1] - y[0]));
    t.xmin = (float)x[0];
    t.xmax = (float)x[nx - 1];
    t.ymin = (float)y[0];
    t.ymax = (float)y[ny - 1];
    std::vector<double> c64;
    build_coef2d(x, y, f, nx, ny, c64);
    std::vector<float4> coef((size_t)(nx - 1) * (ny - 1) * 4);
    for (size_t k = 0; k < coef.size(); k++)
        coef[k] = make_float4((float)c64[4 * k], (float)c64[4 * k + 1], (float)c64[4 * k + 2], (float)c64[4 * k + 3]);
    t.x = A.upload(xs);
    t.y = A.upload(ys);
    t.inv_wx = A.upload(iwx);
    t.inv_wy = A.upload(iwy);
    t.coef = A.upload(coef);
    return t;
}

// four-corner cross difference on a strided 2-D slice (rows: knots a, stride sa; columns: knots b, stride sb)
static double d2s(const double* a, const double* b, const double* f, int na, int nb, size_t sa, size_t sb, int i, int j) {
    const int il = i > 0 ? i - 1 : i, ih = i < na - 1 ? i + 1 : i;
    const int jl = j > 0 ? j - 1 : j, jh = j < nb - 1 ? j + 1 : j;
    if (il == ih || jl == jh) return 0.0;
    return (f[ih * sa + jh * sb] - f[ih * sa + jl * sb] - f[il * sa + jh * sb] + f[il * sa + jl * sb]) / ((a[ih] - a[il]) * (b[jh] - b[jl]));
}

// Tricubic Hermite table (raysect Interpolator3DArray 'cubic', restated): f and its seven first / cross derivatives at the
// knots (3-point first derivatives, four- and eight-corner cross differences), reduced per cell to the 64 monomial
// coefficients of t^p u^q w^r in fp64 and stored as 16 float4 (the powers of w innermost).
static DevTable3D make_table3d(Arena& A, const double* x, const double* y, const double* z, const double* f, int nx, int ny, int nz) {
    DevTable3D T;
    memset(&T, 0, sizeof T);
    T.nx = nx; T.ny = ny; T.nz = nz;
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz, n = (size_t)nx * sx;
    // D[ox][oy][oz]: derivative of order (ox, oy, oz) at every knot
    std::vector<double> D[2][2][2];
    for (int a = 0; a < 8; a++) D[a >> 2][(a >> 1) & 1][a & 1].resize(n);
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++) {
                const size_t o = i * sx + j * sy + k;
                D[0][0][0][o] = f[o];
                D[1][0][0][o] = d1(x, f + j * sy + k, nx, (int)sx, i);
                D[0][1][0][o] = d1(y, f + i * sx + k, ny, (int)sy, j);
                D[0][0][1][o] = d1(z, f + i * sx + j * sy, nz, 1, k);
                D[1][1][0][o] = d2s(x, y, f + k, nx, ny, sx, sy, i, j);
                D[1][0][1][o] = d2s(x, z, f + j * sy, nx, nz, sx, 1, i, k);
                D[0][1][1][o] = d2s(y, z, f + i * sx, ny, nz, sy, 1, j, k);
                const int il = i > 0 ? i - 1 : i, ih = i < nx - 1 ? i + 1 : i, jl = j > 0 ? j - 1 : j, jh = j < ny - 1 ? j + 1 : j;
                const int kl = k > 0 ? k - 1 : k, kh = k < nz - 1 ? k + 1 : k;
                double c3 = 0.0;
                if (il != ih && jl != jh && kl != kh) {
                    auto F = [&](int a, int b, int c) { return f[a * sx + b * sy + c]; };
                    c3 = (F(ih, jh, kh) - F(ih, jh, kl) - F(ih, jl, kh) + F(ih, jl, kl) - F(il, jh, kh) + F(il, jh, kl) + F(il, jl, kh) - F(il, jl, kl)) /
                         ((x[ih] - x[il]) * (y[jh] - y[jl]) * (z[kh] - z[kl]));
                }
                D[1][1][1][o] = c3;
            }
    std::vector<float4> coef((size_t)(nx - 1) * (ny - 1) * (nz - 1) * 16);
    for (int i = 0; i + 1 < nx; i++)
        for (int j = 0; j + 1 < ny; j++)
            for (int k = 0; k + 1 < nz; k++) {
                const double h[3] = {x[i + 1] - x[i], y[j + 1] - y[j], z[k + 1] - z[k]};
                // Hermite "quantities" along an axis: 0 value at knot 0, 1 value at knot 1, 2 scaled derivative at 0, 3 at 1
                double cx[4][4][4];   // [qy][qz][power of t]
                for (int qy = 0; qy < 4; qy++)
                    for (int qz = 0; qz < 4; qz++) {
                        const int jj = j + (qy & 1), kk = k + (qz & 1), oy = qy >> 1, oz = qz >> 1;
                        const double sc = (oy ? h[1] : 1.0) * (oz ? h[2] : 1.0);
                        const size_t o0 = i * sx + jj * sy + kk, o1 = o0 + sx;
                        hermite_coef(D[0][oy][oz][o0] * sc, D[0][oy][oz][o1] * sc, D[1][oy][oz][o0] * sc * h[0], D[1][oy][oz][o1] * sc * h[0], cx[qy][qz]);
                    }
                for (int p = 0; p < 4; p++) {
                    double cy[4][4];  // [qz][power of u]
                    for (int qz = 0; qz < 4; qz++) hermite_coef(cx[0][qz][p], cx[1][qz][p], cx[2][qz][p], cx[3][qz][p], cy[qz]);
                    for (int q = 0; q < 4; q++) {
                        double a[4];
                        hermite_coef(cy[0][q], cy[1][q], cy[2][q], cy[3][q], a);
                        coef[(((size_t)i * (ny - 1) + j) * (nz - 1) + k) * 16 + 4 * p + q] = make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
                    }
                }
            }
    auto knots = [&](const double* v, int m, const float*& kv, const float*& iw, float& lo, float& hi) {
        std::vector<float> xs(m), w(std::max(m - 1, 1));
        for (int q = 0; q < m; q++) xs[q] = (float)v[q];
        for (int q = 0; q + 1 < m; q++) w[q] = (float)(1.0 / (v[q + 1] - v[q]));
        kv = A.upload(xs); iw = A.upload(w); lo = (float)v[0]; hi = (float)v[m - 1];
    };
    knots(x, nx, T.x, T.inv_wx, T.xmin, T.xmax);
    knots(y, ny, T.y, T.inv_wy, T.ymin, T.ymax);
    knots(z, nz, T.z, T.inv_wz, T.zmin, T.zmax);
    T.coef = A.upload(coef);
    return T;
}

// np.gradient(f, edge_order=2) along a strided line with unit spacing (efit.pyx:187-189)
static void np_gradient_unit(const double* f, int n, int stride, double* out, int ostride) {
    for (int i = 0; i < n; i++) {
        double v;
        if (n == 1) v = 0;
        else if (n == 2) v = f[stride] - f[0];
        else if (i == 0) v = -(3.0 * f[0] - 4.0 * f[stride] + f[2 * (size_t)stride]) / 2.0;
        else if (i == n - 1) v = (3.0 * f[(size_t)(n - 1) * stride] - 4.0 * f[(size_t)(n - 2) * stride] + f[(size_t)(n - 3) * stride]) / 2.0;
        else v = (f[(size_t)(i + 1) * stride] - f[(size_t)(i - 1) * stride]) / 2.0;
        out[(size_t)i * ostride] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// field conversion
// ------------------------------------------------------------------------------------------------------------------
static int convert_scalar(Arena& A, const cb2_scalar_field& f, const cb2_axisym* ax, double scale, DevScalar& o) {
    memset(&o, 0, sizeof o);
    o.kind = f.kind;
    switch (f.kind) {
    case CB2_FIELD_CONSTANT: o.c[0] = (float)(f.c[0] * scale); break;
    case CB2_FIELD_GAUSSIAN_VOLUME:
        if (f.c[2] <= 0) return cb2_fail(CB2_ERR_VALUE, "GaussianVolume sigma must be positive");
        o.c[0] = (float)(f.c[0] * scale);
        o.c[1] = (float)(f.c[1] * scale);
        o.c[2] = (float)(-1.0 / (2.0 * f.c[2] * f.c[2]) * 1.4426950408889634);  // exponent coefficient for exp2
        o.c[3] = (float)f.c[3];
        o.c[4] = (float)f.c[4];
        o.c[5] = (float)f.c[5];
        break;
    case CB2_FIELD_SLAB_ION:
        o.c[0] = (float)(f.c[0] * scale);
        o.c[1] = (float)(f.c[1] * scale);
        o.c[2] = (float)f.c[2];
        o.c[3] = (float)f.c[3];
        o.c[4] = (float)(1.0 / f.c[4]);
        break;
    case CB2_FIELD_SLAB_NEUTRAL:
        o.c[0] = (float)(f.c[0] * scale);
        o.c[1] = (float)(-1.0 / (2.0 * f.c[1] * f.c[1]) * 1.4426950408889634);
        break;
    case CB2_FIELD_AXISYM_BLEND: {
        if (!ax) return cb2_fail(CB2_ERR_RUNTIME, "AXISYM_BLEND field without an axisym context");
        if (f.edge) {
            std::vector<float> e(ax->n_triangles);
            for (int i = 0; i < ax->n_triangles; i++) e[i] = (float)(f.edge[i] * scale);
            o.edge = A.upload(e);
        }
        if (f.core) o.core = make_coef1d(A, ax->core_psin, f.core, ax->n_core, scale);
        break;
    }
    default: return cb2_fail(CB2_ERR_TYPE, "unsupported scalar field kind %d", f.kind);
    }
    return A.rc;
}

static int convert_vector(Arena& A, const cb2_vector_field& f, const cb2_axisym* ax, DevVector& o) {
    memset(&o, 0, sizeof o);
    o.kind = f.kind;
    o.c[0] = (float)f.c[0];
    o.c[1] = (float)f.c[1];
    o.c[2] = (float)f.c[2];
    if (f.kind == CB2_FIELD_CONSTANT) return CB2_OK;
    if (f.kind != CB2_FIELD_AXISYM_BLEND) return cb2_fail(CB2_ERR_TYPE, "unsupported vector field kind %d", f.kind);
    if (!ax) return cb2_fail(CB2_ERR_RUNTIME, "AXISYM_BLEND field without an axisym context");
    if (f.core_vtor) o.vtor = make_coef1d(A, ax->core_psin, f.core_vtor, ax->n_core, 1.0);
    if (f.core_vpol) o.vpol = make_coef1d(A, ax->core_psin, f.core_vpol, ax->n_core, 1.0);
    if (f.core_vnorm) o.vnorm = make_coef1d(A, ax->core_psin, f.core_vnorm, ax->n_core, 1.0);
    return A.rc;
}

static bool vec_has_pol(const cb2_vector_field& f, const cb2_axisym* ax) {
    if (f.kind != CB2_FIELD_AXISYM_BLEND || !ax) return false;
    for (int which = 0; which < 2; which++) {
        const double* p = which ? f.core_vnorm : f.core_vpol;
        if (!p) continue;
        for (int i = 0; i < ax->n_core; i++)
            if (p[i] != 0.0) return true;
    }
    return false;
}

static int convert_axisym(Arena& A, const cb2_axisym& ax, DevAxisym& o) {
    memset(&o, 0, sizeof o);
    o.present = 1;
    const cb2_equilibrium& e = ax.eq;
    const int nr = e.nr, nz = e.nz;
    if (nr < 2 || nz < 2) return cb2_fail(CB2_ERR_VALUE, "equilibrium grid must be at least 2x2");
    std::vector<double> psin((size_t)nr * nz), dr((size_t)nr * nz), dz((size_t)nr * nz), dr_di(nr), dz_di(nz);
    for (size_t k = 0; k < psin.size(); k++) psin[k] = (e.psi[k] - e.psi_axis) / (e.psi_lcfs - e.psi_axis);
    np_gradient_unit(e.r, nr, 1, dr_di.data(), 1);
    np_gradient_unit(e.z, nz, 1, dz_di.data(), 1);
    for (int j = 0; j < nz; j++) np_gradient_unit(e.psi + j, nr, nz, dr.data() + j, nz);
    for (int i = 0; i < nr; i++) np_gradient_unit(e.psi + (size_t)i * nz, nz, 1, dz.data() + (size_t)i * nz, 1);
    for (int i = 0; i < nr; i++)
        for (int j = 0; j < nz; j++) {
            dr[(size_t)i * nz + j] *= 1.0 / dr_di[i];
            dz[(size_t)i * nz + j] *= 1.0 / dz_di[j];
        }
    o.psin = make_table2d(A, e.r, e.z, psin.data(), nr, nz);
    o.dpsi_dr = make_table2d(A, e.r, e.z, dr.data(), nr, nz);
    o.dpsi_dz = make_table2d(A, e.r, e.z, dz.data(), nr, nz);
    o.core = make_knots1d(A, ax.core_psin, ax.n_core);
    o.fprof = make_knots1d(A, e.f_psin, e.n_f);
    o.fprof_coef = make_coef1d(A, e.f_psin, e.f_value, e.n_f, 1.0);
    o.b_vac = (float)(e.b_vacuum_magnitude * e.b_vacuum_radius);
    // polygon edges (even-odd crossing test, mask.pyx:53-67)
    o.n_poly = e.n_lcfs;
    std::vector<float4> edges(std::max(e.n_lcfs, 1));
    double pxmin = INFINITY, pxmax = -INFINITY, pymin = INFINITY, pymax = -INFINITY;
    for (int i = 0, j = e.n_lcfs - 1; i < e.n_lcfs; j = i++) {
        const double xi = e.lcfs_polygon[2 * i], yi = e.lcfs_polygon[2 * i + 1];
        const double xj = e.lcfs_polygon[2 * j], yj = e.lcfs_polygon[2 * j + 1];
        const double slope = (yj != yi) ? (xj - xi) / (yj - yi) : 0.0;
        edges[i] = make_float4((float)xi, (float)yi, (float)yj, (float)slope);
        pxmin = fmin(pxmin, xi); pxmax = fmax(pxmax, xi); pymin = fmin(pymin, yi); pymax = fmax(pymax, yi);
    }
    o.poly = A.upload(edges);
    o.poly_xmin = (float)pxmin; o.poly_xmax = (float)pxmax; o.poly_ymin = (float)pymin; o.poly_ymax = (float)pymax;
    {
        // classification grid: cells overlapped by the bounding box of any edge (grown by one cell) are boundary cells,
        // the others take the exact even-odd result of their centre
        const int gx = 192, gy = 320;
        o.pgx = gx; o.pgy = gy;
        const double icx = gx / (pxmax - pxmin), icy = gy / (pymax - pymin);
        o.p_icx = (float)icx; o.p_icy = (float)icy;
        std::vector<unsigned char> cls((size_t)gx * gy, 0);
        for (int i = 0; i < gx; i++)
            for (int j = 0; j < gy; j++) {
                const double cx = pxmin + (i + 0.5) / icx, cy = pymin + (j + 0.5) / icy;
                bool in = false;
                for (int a = 0, b = e.n_lcfs - 1; a < e.n_lcfs; b = a++) {
                    const double xi = e.lcfs_polygon[2 * a], yi = e.lcfs_polygon[2 * a + 1];
                    const double xj = e.lcfs_polygon[2 * b], yj = e.lcfs_polygon[2 * b + 1];
                    if (((yi > cy) != (yj > cy)) && (cx < (xj - xi) * (cy - yi) / (yj - yi) + xi)) in = !in;
                }
                cls[(size_t)i * gy + j] = in ? 1 : 0;
            }
        for (int a = 0, b = e.n_lcfs - 1; a < e.n_lcfs; b = a++) {
            const double xi = e.lcfs_polygon[2 * a], yi = e.lcfs_polygon[2 * a + 1];
            const double xj = e.lcfs_polygon[2 * b], yj = e.lcfs_polygon[2 * b + 1];
            int i0 = (int)floor((fmin(xi, xj) - pxmin) * icx) - 1, i1 = (int)floor((fmax(xi, xj) - pxmin) * icx) + 1;
            int j0 = (int)floor((fmin(yi, yj) - pymin) * icy) - 1, j1 = (int)floor((fmax(yi, yj) - pymin) * icy) + 1;
            i0 = std::max(i0, 0); j0 = std::max(j0, 0); i1 = std::min(i1, gx - 1); j1 = std::min(j1, gy - 1);
            for (int i = i0; i <= i1; i++)
                for (int j = j0; j <= j1; j++) cls[(size_t)i * gy + j] = 2;
        }
        o.poly_cls = A.upload(cls);
        // per grid row the edges whose y-span reaches it (grown by one row against float32 rounding of the row index): only these
        // can cross the horizontal ray of a point in that row — the same edge test on ~5 instead of all n_lcfs edges
        std::vector<int> rstart(gy + 1, 0);
        std::vector<float4> redges;
        std::vector<double2> redges_d;
        for (int j = 0; j < gy; j++) {
            rstart[j] = (int)redges.size();
            for (int a = 0, b = e.n_lcfs - 1; a < e.n_lcfs; b = a++) {
                const double yi = e.lcfs_polygon[2 * a + 1], yj = e.lcfs_polygon[2 * b + 1];
                const int j0 = (int)floor((fmin(yi, yj) - pymin) * icy) - 1, j1 = (int)floor((fmax(yi, yj) - pymin) * icy) + 1;
                if (j >= j0 && j <= j1) {
                    redges.push_back(edges[a]);
                    redges_d.push_back(make_double2(e.lcfs_polygon[2 * a], yi));
                    redges_d.push_back(make_double2(e.lcfs_polygon[2 * b], yj));
                }
            }
        }
        rstart[gy] = (int)redges.size();
        o.poly_row_start = A.upload(rstart);
        o.poly_row_edges = A.upload(redges);
        o.poly_row_edges_d = A.upload(redges_d);
    }
    if (ax.n_mask < 1 || ax.n_mask > 8) return cb2_fail(CB2_ERR_VALUE, "blend mask table must have 1..8 points");
    o.n_mask = ax.n_mask;
    for (int i = 0; i < ax.n_mask; i++) { o.mask_x[i] = (float)ax.mask_x[i]; o.mask_y[i] = (float)ax.mask_y[i]; }
    // mesh buckets
    o.n_tri = ax.n_triangles;
    if (ax.n_triangles > 0) {
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int i = 0; i < ax.n_vertices; i++) {
            xmin = fmin(xmin, ax.vertices[2 * i]); xmax = fmax(xmax, ax.vertices[2 * i]);
            ymin = fmin(ymin, ax.vertices[2 * i + 1]); ymax = fmax(ymax, ax.vertices[2 * i + 1]);
        }
        const double aspect = (ymax - ymin) / (xmax - xmin);
        // ~16 cells per triangle: edge meshes are made of slivers (Generomak: median 3.4 cm x 0.6 cm), so the lists are filled
        // by an exact triangle / cell overlap test instead of bounding boxes — 1.9 instead of 6.2 triangles per occupied cell
        int gx = (int)ceil(sqrt(16.0 * ax.n_triangles / aspect));
        gx = std::min(std::max(gx, 8), 1024);
        int gy = std::min(std::max((int)ceil(gx * aspect), 8), 2048);
        o.gx = gx; o.gy = gy;
        o.mx0 = (float)xmin; o.my0 = (float)ymin;
        o.mx0_d = xmin; o.my0_d = ymin;
        const double icx = gx / (xmax - xmin) * (1.0 - 1e-7), icy = gy / (ymax - ymin) * (1.0 - 1e-7);
        o.inv_cx = (float)icx; o.inv_cy = (float)icy;
        o.inv_cx_d = icx; o.inv_cy_d = icy;
        std::vector<int> count((size_t)gx * gy + 1, 0), start((size_t)gx * gy + 1, 0), tris;
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {
                int acc = 0;
                for (size_t k = 0; k < (size_t)gx * gy; k++) { start[k] = acc; acc += count[k]; count[k] = 0; }
                start[(size_t)gx * gy] = acc;
                tris.resize(std::max(acc, 1));
            }
            for (int t = 0; t < ax.n_triangles; t++) {
                const int32_t* tr = ax.triangles + 3 * t;
                double txmin = INFINITY, txmax = -INFINITY, tymin = INFINITY, tymax = -INFINITY;
                for (int k = 0; k < 3; k++) {
                    if (tr[k] < 0 || tr[k] >= ax.n_vertices) return cb2_fail(CB2_ERR_VALUE, "triangle vertex index out of range");
                    const double vx = ax.vertices[2 * tr[k]], vy = ax.vertices[2 * tr[k] + 1];
                    txmin = fmin(txmin, vx); txmax = fmax(txmax, vx); tymin = fmin(tymin, vy); tymax = fmax(tymax, vy);
                }
                // the device computes the bucket index with the same float64 expression (monotone in the coordinate), so a
                // point inside the triangle's bounding box lands in this cell range: no safety margin needed.  Triangles are
                // visited in ascending order, which keeps every bucket list sorted (first hit = lowest triangle id).
                // (the box is grown by 1e-9 of the mesh extent for points the rounded containment test accepts on an edge)
                const double ex = 1e-9 * (xmax - xmin), ey = 1e-9 * (ymax - ymin);
                int i0 = (int)floor((txmin - ex - xmin) * icx), i1 = (int)floor((txmax + ex - xmin) * icx);
                int j0 = (int)floor((tymin - ey - ymin) * icy), j1 = (int)floor((tymax + ey - ymin) * icy);
                i0 = std::max(i0, 0); j0 = std::max(j0, 0); i1 = std::min(i1, gx - 1); j1 = std::min(j1, gy - 1);
                double px[3], py[3];
                for (int k = 0; k < 3; k++) { px[k] = ax.vertices[2 * tr[k]]; py[k] = ax.vertices[2 * tr[k] + 1]; }
                for (int i = i0; i <= i1; i++)
                    for (int j = j0; j <= j1; j++) {
                        // separating-axis test of the triangle against the cell grown by the same margin (edge normals; the
                        // coordinate axes are covered by the index range): keeps a superset of the truly overlapping cells
                        const double cx0 = xmin + i / icx - ex, cx1 = xmin + (i + 1) / icx + ex, cy0 = ymin + j / icy - ey, cy1 = ymin + (j + 1) / icy + ey;
                        const double ccx = 0.5 * (cx0 + cx1), ccy = 0.5 * (cy0 + cy1), hx = 0.5 * (cx1 - cx0), hy = 0.5 * (cy1 - cy0);
                        bool apart = false;
                        for (int k = 0; k < 3 && !apart; k++) {
                            const double nx = py[(k + 1) % 3] - py[k], ny = px[k] - px[(k + 1) % 3];
                            double lo = INFINITY, hi = -INFINITY;
                            for (int q = 0; q < 3; q++) { const double d = nx * (px[q] - ccx) + ny * (py[q] - ccy); lo = fmin(lo, d); hi = fmax(hi, d); }
                            const double rr = hx * fabs(nx) + hy * fabs(ny);
                            const double slack = 1e-12 * (fabs(nx) + fabs(ny)) * (fabs(ccx) + fabs(ccy) + 1.0);
                            apart = lo > rr + slack || hi < -rr - slack;
                        }
                        if (apart) continue;
                        const size_t cell = (size_t)i * gy + j;
                        if (pass == 1) tris[start[cell] + count[cell]] = t;
                        count[cell]++;
                    }
            }
        }
        std::vector<double2> tv((size_t)ax.n_triangles * 3);
        for (int t = 0; t < ax.n_triangles; t++)
            for (int k = 0; k < 3; k++) {
                const int v = ax.triangles[3 * t + k];
                tv[(size_t)t * 3 + k] = make_double2(ax.vertices[2 * v], ax.vertices[2 * v + 1]);
            }
        // float32 copy of the vertices for the sign filter in front of the exact test: (ax, ay, bx, by), (cx, cy, -, -)
        std::vector<float4> tf((size_t)ax.n_triangles * 2);
        for (int t = 0; t < ax.n_triangles; t++) {
            const double2 *q = &tv[(size_t)t * 3];
            tf[2 * (size_t)t] = make_float4((float)q[0].x, (float)q[0].y, (float)q[1].x, (float)q[1].y);
            tf[2 * (size_t)t + 1] = make_float4((float)q[2].x, (float)q[2].y, 0.f, 0.f);
        }
        o.cell_start = A.upload(start);
        o.cell_tris = A.upload(tris);
        o.tri = A.upload(tv);
        o.trif = A.upload(tf);
    }
    return A.rc;
}

// ------------------------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------------------------
static const double SPEED_OF_LIGHT = 299792458.0, ATOMIC_MASS = 1.66053906660e-27, ELEMENTARY_CHARGE = 1.602176634e-19;
static const double PLANCK_CONSTANT = 6.62607015e-34, ELECTRON_REST_MASS = 9.1093837015e-31, VACUUM_PERMITTIVITY = 8.8541878128e-12;
static const double HC_EV_NM = 1239.8419738620933;

static void set_comp(DevScene& S, int slot, double lambda, int type) {
    const double c0 = (lambda - S.min_wavelength_d) / S.delta_d;
    const double ci = floor(c0);
    // bin ranges are stored relative to c0_int as packed int16 pairs: keep |c0_int| <= 28000 (a line that far outside
    // the window only matters if it is thousands of bins wide; the fraction then absorbs the remainder)
    const double cic = fmin(fmax(ci, -28000.0), 28000.0);
    S.comps[slot].c0_int = (int)cic;
    S.comps[slot].c0_frac = (float)(c0 - cic);
    S.comps[slot].dlambda = 0.f;
    S.comps[slot].type = type;
}


// ------------------------------------------------------------------------------------------------------------------
// Beam: host-side fp64 rate evaluation and the SingleRayAttenuator axis table (singleray.pyx:182-313)
// ------------------------------------------------------------------------------------------------------------------
static double host_cubic1d(const double* x, const double* f, int n, double v, bool quadratic = false) {
    // Interpolator1DArray 'cubic' restated (local Hermite, 3-point knot derivatives); outside the knots 'nearest', or with
    // `quadratic` the parabola through the edge value with the spline's slopes at both knots of the edge interval (the oracle's
    // interp1d_cubic_quadratic)
    if (n == 1) return f[0];
    if (quadratic && (v < x[0] || v > x[n - 1])) {
        const int i = v < x[0] ? 0 : n - 2;
        const double h = x[i + 1] - x[i], t = (v - x[i]) / h;
        const double d0 = d1(x, f, n, 1, i) * h, dd1 = d1(x, f, n, 1, i + 1) * h;
        const double a2 = v < x[0] ? f[i] : f[i + 1] - 0.5 * d0 - 0.5 * dd1;
        return (0.5 * (dd1 - d0) * t + d0) * t + a2;
    }
    v = fmin(fmax(v, x[0]), x[n - 1]);
    int i = (int)(std::upper_bound(x, x + n, v) - x) - 1;
    i = std::min(std::max(i, 0), n - 2);
    const double h = x[i + 1] - x[i], t = (v - x[i]) / h;
    double a[4];
    hermite_coef(f[i], f[i + 1], d1(x, f, n, 1, i) * h, d1(x, f, n, 1, i + 1) * h, a);
    return a[0] + t * (a[1] + t * (a[2] + t * a[3]));
}

static double host_cubic2d(const double* x, const double* y, const std::vector<double>& coef, int nx, int ny, double vx, double vy,
                           bool linear = false) {
    const double px = vx, py = vy;
    vx = fmin(fmax(vx, x[0]), x[nx - 1]);
    vy = fmin(fmax(vy, y[0]), y[ny - 1]);
    if (linear && (px != vx || py != vy)) {
        // 'linear' extrapolation: value, gradient and cross derivative of the patch at the nearest boundary point
        int i = (int)(std::upper_bound(x, x + nx, vx) - x) - 1, j = (int)(std::upper_bound(y, y + ny, vy) - y) - 1;
        i = std::min(std::max(i, 0), nx - 2);
        j = std::min(std::max(j, 0), ny - 2);
        const double hx = x[i + 1] - x[i], hy = y[j + 1] - y[j], t = (vx - x[i]) / hx, u = (vy - y[j]) / hy;
        const double* c = &coef[((size_t)i * (ny - 1) + j) * 16];
        double p = 0.0, pt = 0.0, pu = 0.0, ptu = 0.0;
        for (int k = 3; k >= 0; k--) {
            const double v = c[4 * k] + u * (c[4 * k + 1] + u * (c[4 * k + 2] + u * c[4 * k + 3]));
            const double dv = c[4 * k + 1] + u * (2.0 * c[4 * k + 2] + 3.0 * u * c[4 * k + 3]);
            ptu = ptu * t + pu; pt = pt * t + p; pu = pu * t + dv; p = p * t + v;
        }
        const double dx = (px - vx) / hx, dy = (py - vy) / hy;
        return p + pt * dx + pu * dy + ptu * dx * dy;
    }
    int i = (int)(std::upper_bound(x, x + nx, vx) - x) - 1, j = (int)(std::upper_bound(y, y + ny, vy) - y) - 1;
    i = std::min(std::max(i, 0), nx - 2);
    j = std::min(std::max(j, 0), ny - 2);
    const double t = (vx - x[i]) / (x[i + 1] - x[i]), u = (vy - y[j]) / (y[j + 1] - y[j]);
    const double* c = &coef[((size_t)i * (ny - 1) + j) * 16];
    double v = 0.0;
    for (int p = 3; p >= 0; p--) v = v * t + (c[4 * p] + u * (c[4 * p + 1] + u * (c[4 * p + 2] + u * c[4 * p + 3])));
    return v;
}

struct HostBeamRate {     // BeamStoppingRate in log space (openadas/rates/beam.pyx:62-103)
    bool constant;
    bool extrapolate = false;   // 'linear' / 'quadratic' beyond the tables (beam.pyx:73-84), else clamped
    double value;
    std::vector<double> le, ln, lt, lsen, lst, coef;
    double eval(double energy, double density, double temperature) const {
        if (constant) return value;
        if (energy <= 0 || density <= 0 || temperature <= 0) return 0.0;
        const double e = log10(energy), n = log10(density), t = log10(temperature);
        double a;
        const int ne = (int)le.size(), nn = (int)ln.size();
        if (ne == 1 && nn == 1) a = lsen[0];
        else if (ne == 1) a = host_cubic1d(ln.data(), lsen.data(), nn, n, extrapolate);
        else if (nn == 1) a = host_cubic1d(le.data(), lsen.data(), ne, e, extrapolate);
        else a = host_cubic2d(le.data(), ln.data(), coef, ne, nn, e, n, extrapolate);
        const double b = lt.size() > 1 ? host_cubic1d(lt.data(), lst.data(), (int)lt.size(), t, extrapolate) : lst[0];
        return pow(10.0, a + b);
    }
};

static void mat_mul_point(const double m[12], const double p[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = m[4 * i] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2] + m[4 * i + 3];
}

// called once the scene's tables are on the device: plasma state along the beam axis comes from the device's own field
// evaluation (sample_state_kernel), the stopping sum, cumulative trapezium and exponential are fp64 on the host
static int build_beam(Arena& A, const cb2_scene_desc& d, cb2_scene* sc) {
    DevScene& S = sc->host;
    const cb2_beam_desc& b = *d.beam;
    DevBeam& o = S.beam;
    memset(&o, 0, sizeof o);
    if (!(b.energy > 0)) return cb2_fail(CB2_ERR_VALUE, "Beam energy must be positive");
    if (!(b.sigma > 0) || !(b.length > 0) || !(b.attenuator_step > 0) || !(b.clamp_sigma > 0))
        return cb2_fail(CB2_ERR_VALUE, "Beam sigma, length, attenuator step and clamp_sigma must be positive");
    if (b.n_stopping < 0 || b.n_stopping > CB2_MAX_SPECIES) return cb2_fail(CB2_ERR_VALUE, "invalid number of stopping species");
    o.present = 1;
    for (int k = 0; k < 12; k++) o.l2p[k] = b.beam_to_plasma[k];
    const double evamu = 2.0 * ELEMENTARY_CHARGE / ATOMIC_MASS;       // EvAmuToMS.conversion_factor (conversion.py:31)
    const double speed = sqrt(b.energy * evamu);
    o.speed = (float)speed;
    o.length = (float)b.length;
    o.sigma2 = (float)(b.sigma * b.sigma);
    o.tanx = (float)tan(b.divergence_x * M_PI / 180.0);
    o.tany = (float)tan(b.divergence_y * M_PI / 180.0);
    o.clamp_to_zero = b.clamp_to_zero;
    o.clamp2 = (float)(b.clamp_sigma * b.clamp_sigma);
    int n = 1 + (int)ceil(b.length / b.attenuator_step);
    if (n < 4) n = 4;
    o.n_axis = n;
    o.inv_dz = (float)((n - 1) / b.length);
    // stopping rates
    std::vector<HostBeamRate> rates(b.n_stopping);
    for (int k = 0; k < b.n_stopping; k++) {
        const cb2_beam_rate& r = b.stopping_rates[k];
        if (b.stopping_species[k] < 0 || b.stopping_species[k] >= d.n_species) return cb2_fail(CB2_ERR_VALUE, "stopping species index out of range");
        HostBeamRate& h = rates[k];
        h.constant = r.n_e <= 0;
        h.extrapolate = r.extrapolate != 0;
        h.value = r.constant;
        if (h.constant) continue;
        if (r.n_e < 1 || r.n_n < 1 || r.n_t < 1 || !(r.sref > 0)) return cb2_fail(CB2_ERR_VALUE, "invalid beam stopping table");
        h.le.resize(r.n_e); h.ln.resize(r.n_n); h.lt.resize(r.n_t); h.lsen.resize((size_t)r.n_e * r.n_n); h.lst.resize(r.n_t);
        for (int i = 0; i < r.n_e; i++) h.le[i] = log10(r.e[i]);
        for (int i = 0; i < r.n_n; i++) h.ln[i] = log10(r.n[i]);
        for (int i = 0; i < r.n_t; i++) { h.lt[i] = log10(r.t[i]); h.lst[i] = log10(r.st[i] / r.sref); }
        for (size_t i = 0; i < h.lsen.size(); i++) h.lsen[i] = log10(r.sen[i]);
        if (r.n_e > 1 && r.n_n > 1) build_coef2d(h.le.data(), h.ln.data(), h.lsen.data(), r.n_e, r.n_n, h.coef);
    }
    // plasma state on the axis points (plasma space), evaluated by the device
    std::vector<double> pts((size_t)n * 3), z(n);
    for (int k = 0; k < n; k++) {
        z[k] = (k == n - 1) ? b.length : b.length * k / (n - 1);     // np.linspace(0, length, nbeam)
        const double pb[3] = {0.0, 0.0, z[k]};
        mat_mul_point(b.beam_to_plasma, pb, &pts[3 * k]);
    }
    const int w = 2 + 5 * S.n_species + 3;
    std::vector<double> state((size_t)n * w);
    {
        double *dp = nullptr, *ds = nullptr;
        CB2_CUDA(cudaMalloc((void**)&dp, pts.size() * sizeof(double)));
        int rc = cb2_cuda_check(cudaMalloc((void**)&ds, state.size() * sizeof(double)), "cudaMalloc(axis state)");
        if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpy(dp, pts.data(), pts.size() * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy(axis)");
        if (rc == CB2_OK) rc = cb2_launch_sample_state(sc, dp, n, ds, 1, 0);
        if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpy(state.data(), ds, state.size() * sizeof(double), cudaMemcpyDeviceToHost), "cudaMemcpy(axis state)");
        cudaFree(dp);
        if (ds) cudaFree(ds);
        if (rc != CB2_OK) return rc;
    }
    // beam velocity in plasma space: BEAM_AXIS transformed as a vector, normalised, times the speed
    double ax[3] = {b.beam_to_plasma[2], b.beam_to_plasma[6], b.beam_to_plasma[10]};
    const double al = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (int k = 0; k < 3; k++) ax[k] = ax[k] / al * speed;
    std::vector<double> stop(n, 0.0);
    for (int k = 0; k < n; k++) {
        const double* st = &state[(size_t)k * w];
        double density_sum = 0.0;
        for (int q = 0; q < b.n_stopping; q++) {
            const int sp = b.stopping_species[q];
            density_sum += (double)d.species[sp].charge * d.species[sp].charge * st[2 + 5 * sp];
        }
        for (int q = 0; q < b.n_stopping; q++) {
            const int sp = b.stopping_species[q], zc = d.species[sp].charge;
            if (zc == 0) continue;                                   // no beam stopping data for neutrals (SURVEY A.7 caveat)
            const double ne_t = st[2 + 5 * sp] * zc, ti = st[3 + 5 * sp];
            const double iv[3] = {ax[0] - st[4 + 5 * sp], ax[1] - st[5 + 5 * sp], ax[2] - st[6 + 5 * sp]};
            const double sp2 = iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2];
            stop[k] += ne_t * rates[q].eval(sp2 / evamu, density_sum / zc, ti);
        }
    }
    const double n0 = b.power / (b.energy * b.atomic_weight * ELEMENTARY_CHARGE) / speed;
    std::vector<float> dens(n);
    double cum = 0.0;
    for (int k = 0; k < n; k++) {
        if (k > 0) cum += 0.5 * (stop[k] + stop[k - 1]) * (z[k] - z[k - 1]);   // cumulative_trapezoid(..., initial=0)
        dens[k] = (float)(n0 * exp(-cum / speed) * CB2_DENSITY_SCALE);
    }
    o.axis_density = A.upload(dens);
    return A.rc;
}

// BeamCXPEC tables of a BEAM_CX_LINE model (openadas/rates/cx.pyx:66-103)
static int convert_cx(Arena& A, const cb2_cx_rate& r, double wavelength, DevCXRate& e) {
    e.extrapolate = r.extrapolate != 0;
    if (r.n_eb <= 0) {
        e.is_const = 1;
        e.lconst = r.constant > 0 ? (float)(log10(r.constant) + CB2_PEC_LOG_OFFSET) : -INFINITY;
        return CB2_OK;
    }
    if (!(r.qref > 0)) return cb2_fail(CB2_ERR_VALUE, "BeamCXPEC qref must be positive");
    const int nn[5] = {r.n_eb, r.n_ti, r.n_ni, r.n_z, r.n_b};
    const double* xs[5] = {r.eb, r.ti, r.ni, r.z, r.b};
    const double* qs[5] = {r.qeb, r.qti, r.qni, r.qz, r.qb};
    const double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9;
    for (int k = 0; k < 5; k++) {
        const int n = nn[k];
        if (n < 1) return cb2_fail(CB2_ERR_VALUE, "BeamCXPEC grids need at least one point");
        std::vector<double> x(n), f(n);
        for (int i = 0; i < n; i++) {
            if (k == 0) {
                if (!(qs[k][i] > 0)) return cb2_fail(CB2_ERR_VALUE, "BeamCXPEC qeb must be positive (log10 interpolation)");
                x[i] = log10(xs[k][i]);
                f[i] = log10(qs[k][i] / wavelength * conv) + CB2_PEC_LOG_OFFSET;
            } else {
                x[i] = xs[k][i] * (k == 2 ? CB2_DENSITY_SCALE : 1.0);     // n_ion knots in 1e19 m^-3
                f[i] = qs[k][i] / r.qref;
            }
            if (i > 0 && !(x[i] > x[i - 1])) return cb2_fail(CB2_ERR_VALUE, "BeamCXPEC grids must be increasing");
        }
        e.n[k] = n;
        e.single[k] = (float)f[0];
        if (n > 1) {
            e.t[k] = make_knots1d(A, x.data(), n);
            e.c[k] = make_coef1d(A, x.data(), f.data(), n, 1.0);
        }
    }
    return A.rc;
}

// BeamPopulationRate table (beam.pyx:143-170): log10 sen on (log10 E, log10 n), log10(st / sref) on log10 T
static int convert_population(Arena& A, const cb2_beam_rate& r, DevPopRate& e) {
    memset(&e, 0, sizeof e);
    e.extrapolate = r.extrapolate != 0;
    if (r.n_e <= 0) {
        e.is_const = 1;
        e.lconst = r.constant > 0 ? (float)log10(r.constant) : -INFINITY;
        return CB2_OK;
    }
    if (r.n_e < 2 || r.n_n < 2 || r.n_t < 2 || !(r.sref > 0))
        return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "beam population tables need at least 2 points per axis on the device path");
    std::vector<double> le(r.n_e), ln(r.n_n), lt(r.n_t), lsen((size_t)r.n_e * r.n_n), lst(r.n_t);
    for (int i = 0; i < r.n_e; i++) le[i] = log10(r.e[i]);
    for (int i = 0; i < r.n_n; i++) ln[i] = log10(r.n[i]);
    for (int i = 0; i < r.n_t; i++) { lt[i] = log10(r.t[i]); lst[i] = log10(r.st[i] / r.sref); }
    for (size_t i = 0; i < lsen.size(); i++) {
        if (!(r.sen[i] > 0)) return cb2_fail(CB2_ERR_VALUE, "beam population table values must be positive (log10 interpolation)");
        lsen[i] = log10(r.sen[i]);
    }
    e.a = make_table2d(A, le.data(), ln.data(), lsen.data(), r.n_e, r.n_n);
    e.tk = make_knots1d(A, lt.data(), r.n_t);
    e.tc = make_coef1d(A, lt.data(), lst.data(), r.n_t, 1.0);
    return A.rc;
}

static int convert_model(Arena& A, const cb2_scene_desc& d, const cb2_model& m, DevScene& S, DevModel& o) {
    memset(&o, 0, sizeof o);
    o.kind = m.kind;
    o.species = m.species;
    o.shape = m.shape.kind;
    o.polarisation = m.shape.polarisation;
    if (m.kind == CB2_MODEL_BREMSSTRAHLUNG) return CB2_OK;
    if (m.kind == CB2_MODEL_TOTAL_RADIATED_POWER) {
        // total_radiated_power.pyx:120-163: resolved species and the three power coefficients, log-log cubic in (ne, te)
        const cb2_model_ext* x = m.ext;
        if (!x) return cb2_fail(CB2_ERR_RUNTIME, "TotalRadiatedPower needs its resolved species and rates");
        if (x->line_rad_species < 0 || x->line_rad_species >= d.n_species || x->recom_species < 0 || x->recom_species >= d.n_species)
            return cb2_fail(CB2_ERR_RUNTIME, "The plasma object does not contain the required ion species for calculating total radiated power");
        if (x->n_hydrogen < 0 || x->n_hydrogen > 3) return cb2_fail(CB2_ERR_VALUE, "at most three hydrogen isotopes");
        DevModelExt e;
        memset(&e, 0, sizeof e);
        e.line_rad = x->line_rad_species; e.recom = x->recom_species; e.n_hyd = x->n_hydrogen;
        for (int k = 0; k < x->n_hydrogen; k++) {
            if (x->hydrogen_species[k] < 0 || x->hydrogen_species[k] >= d.n_species) return cb2_fail(CB2_ERR_VALUE, "hydrogen species index out of range");
            e.hyd[k] = x->hydrogen_species[k];
        }
        const cb2_rate2d* r[3] = {&x->plt, &x->prb, &x->prc};
        const int has[3] = {x->has_plt, x->has_prb, x->has_prc};
        for (int k = 0; k < 3; k++) {
            e.has[k] = has[k];
            if (!has[k]) continue;
            e.extrapolate[k] = r[k]->extrapolate;
            if (r[k]->n_ne <= 0) {
                e.is_const[k] = 1;
                e.lconst[k] = r[k]->constant > 0 ? (float)(log10(r[k]->constant) + CB2_PEC_LOG_OFFSET) : -INFINITY;
                continue;
            }
            const int nn = r[k]->n_ne, nt = r[k]->n_te;
            if (nn < 2 || nt < 2) return cb2_fail(CB2_ERR_VALUE, "rate tables need at least 2x2 points");
            std::vector<double> lne(nn), lte(nt), lr((size_t)nn * nt);
            for (int i = 0; i < nn; i++) lne[i] = log10(r[k]->ne[i]);
            for (int j = 0; j < nt; j++) lte[j] = log10(r[k]->te[j]);
            for (size_t q = 0; q < lr.size(); q++) {
                if (!(r[k]->rate[q] > 0)) return cb2_fail(CB2_ERR_VALUE, "rate table values must be positive (log10 interpolation)");
                lr[q] = log10(r[k]->rate[q]) + CB2_PEC_LOG_OFFSET;
            }
            e.tab[k] = make_table2d(A, lne.data(), lte.data(), lr.data(), nn, nt);
        }
        o.ext = A.upload(std::vector<DevModelExt>(1, e));
        S.has_flat = 1;
        return A.rc;
    }
    if (m.kind == CB2_MODEL_BEAM_EMISSION_LINE) {
        // beam_emission.pyx:178-216 + mse.pyx:62-135: nine Gaussians around the Doppler-shifted line, split by the motional
        // Stark effect, broadened by the BEAM temperature
        const cb2_model_ext* x = m.ext;
        if (!d.beam) return cb2_fail(CB2_ERR_RUNTIME, "The emission model is not connected to a beam object.");
        if (!x) return cb2_fail(CB2_ERR_RUNTIME, "BeamEmissionLine needs its resolved rates");
        if (x->n_bes < 0 || x->n_bes > CB2_MAX_SPECIES) return cb2_fail(CB2_ERR_VALUE, "invalid number of beam emission rates");
        if (!(m.wavelength > 0)) return cb2_fail(CB2_ERR_VALUE, "line wavelength must be positive");
        o.wavelength = (float)m.wavelength;
        o.inv_delta = (float)(1.0 / S.delta_d);
        o.inv_c = (float)(1.0 / SPEED_OF_LIGHT);
        o.shape = CB2_SHAPE_GAUSSIAN;
        o.pec_const = 1;
        o.pec_value = -INFINITY;
        o.pec_grid = -1;
        DevModelExt e;
        memset(&e, 0, sizeof e);
        e.n_bes = x->n_bes;
        const double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9 / m.wavelength;     // PhotonToJ (beam.pyx:238)
        for (int k = 0; k < x->n_bes; k++) {
            const cb2_beam_rate& r = x->bes_rates[k];
            const int sp = x->bes_species[k];
            if (sp < 0 || sp >= d.n_species) return cb2_fail(CB2_ERR_VALUE, "beam emission species index out of range");
            e.bes_species[k] = sp;
            e.bes_charge[k] = d.species[sp].charge;
            e.bes_extrapolate[k] = r.extrapolate != 0;
            if (r.n_e <= 0) {
                e.bes_const[k] = 1;
                e.bes_lconst[k] = r.constant > 0 ? (float)(log10(r.constant) + CB2_PEC_LOG_OFFSET) : -INFINITY;
                continue;
            }
            if (r.n_e < 2 || r.n_n < 2 || r.n_t < 2 || !(r.sref > 0))
                return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "beam emission tables need at least 2 points per axis on the device path");
            std::vector<double> le(r.n_e), ln(r.n_n), lt(r.n_t), lsen((size_t)r.n_e * r.n_n), lst(r.n_t);
            for (int i = 0; i < r.n_e; i++) le[i] = log10(r.e[i]);
            for (int i = 0; i < r.n_n; i++) ln[i] = log10(r.n[i]);
            for (int i = 0; i < r.n_t; i++) { lt[i] = log10(r.t[i]); lst[i] = log10(r.st[i] / r.sref); }
            for (size_t i = 0; i < lsen.size(); i++) lsen[i] = log10(r.sen[i] * conv) + CB2_PEC_LOG_OFFSET;
            e.bes_a[k] = make_table2d(A, le.data(), ln.data(), lsen.data(), r.n_e, r.n_n);
            e.bes_tk[k] = make_knots1d(A, lt.data(), r.n_t);
            e.bes_tc[k] = make_coef1d(A, lt.data(), lst.data(), r.n_t, 1.0);
        }
        const double s2p = x->mse_ratios[0], s1s0 = x->mse_ratios[1], p23 = x->mse_ratios[2], p43 = x->mse_ratios[3];
        const double dd = 1 / (1 + s2p), isig = s2p * dd, ipi = 0.5 * dd, is0 = 1 / (s1s0 + 1), is1 = 0.5 * s1s0 * is0;
        const double ip3 = 1 / (1 + p23 + p43), ip2 = p23 * ip3, ip4 = p43 * ip3;
        const double amp[9] = {isig * is0, isig * is1, isig * is1, ipi * ip2, ipi * ip2, ipi * ip3, ipi * ip3, ipi * ip4, ipi * ip4};
        for (int k = 0; k < 9; k++) e.mse_amp[k] = (float)amp[k];
        if (x->n_mse > 1) {
            if (!x->mse_ratio_tab || !(x->mse_dlne > 0)) return cb2_fail(CB2_ERR_VALUE, "MSE ratio table missing or knot spacing not positive");
            std::vector<float4> tab(x->n_mse);
            for (int q = 0; q < x->n_mse; q++)
                tab[q] = make_float4((float)x->mse_ratio_tab[q], (float)x->mse_ratio_tab[x->n_mse + q], (float)x->mse_ratio_tab[2 * x->n_mse + q],
                                     (float)x->mse_ratio_tab[3 * x->n_mse + q]);
            e.mse_n = x->n_mse;
            e.mse_lne0 = (float)x->mse_lne0;
            e.mse_inv_dlne = (float)(1.0 / x->mse_dlne);
            e.mse_tab = A.upload(tab);
        }
        e.mse_sigma_b = (float)(sqrt(d.beam->temperature * ELEMENTARY_CHARGE / (d.beam->atomic_weight * ATOMIC_MASS)) * m.wavelength /
                                SPEED_OF_LIGHT / S.delta_d);
        o.ext = A.upload(std::vector<DevModelExt>(1, e));
        o.comp0 = S.n_comp;
        o.ncomp = 9;
        if (S.n_comp + 9 > CB2_MAX_COMP) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "too many line components (max %d)", CB2_MAX_COMP);
        S.n_comp += 9;
        for (int k = 0; k < 9; k++) set_comp(S, o.comp0 + k, m.wavelength, 0);
        return A.rc;
    }
    if (m.kind != CB2_MODEL_EXCITATION_LINE && m.kind != CB2_MODEL_RECOMBINATION_LINE && m.kind != CB2_MODEL_THERMAL_CX_LINE &&
        m.kind != CB2_MODEL_BEAM_CX_LINE)
        return cb2_fail(CB2_ERR_TYPE, "unsupported model kind %d", m.kind);
    if (m.kind == CB2_MODEL_BEAM_CX_LINE && !d.beam) return cb2_fail(CB2_ERR_RUNTIME, "The emission model is not connected to a beam object.");
    if (m.species < 0 || m.species >= d.n_species)
        return cb2_fail(CB2_ERR_RUNTIME, "The plasma object does not contain the ion species for the specified line");
    if (!(m.wavelength > 0)) return cb2_fail(CB2_ERR_VALUE, "line wavelength must be positive");
    o.wavelength = (float)m.wavelength;
    o.inv_delta = (float)(1.0 / S.delta_d);
    o.inv_c = (float)(1.0 / SPEED_OF_LIGHT);
    o.sigma_coef = (float)(sqrt(ELEMENTARY_CHARGE / (m.atomic_weight * ATOMIC_MASS)) * m.wavelength / SPEED_OF_LIGHT / S.delta_d);
    for (int k = 0; k < 3; k++) o.param[k] = (float)m.shape.param[k];
    // rate table: log10(PhotonToJ(rate, wavelength)) + 38 on (log10 ne, log10 te)   (pec.pyx:59-68)
    o.pec_grid = -1;
    if (m.kind == CB2_MODEL_BEAM_CX_LINE) {
        // charge_exchange.pyx:311-361: effective emission coefficients per donor metastable + the populations of the excited ones
        const cb2_model_ext* x = m.ext;
        if (!x || x->n_cx < 1) return cb2_fail(CB2_ERR_RUNTIME, "BeamCXLine needs its resolved CX rates");
        if (x->n_cx > CB2_MAX_META) return cb2_fail(CB2_ERR_VALUE, "at most %d donor metastables are supported", CB2_MAX_META);
        if (x->n_cx > 1 && !x->cx_population) return cb2_fail(CB2_ERR_RUNTIME, "excited donor metastables need their beam population rates");
        DevModelExt e;
        memset(&e, 0, sizeof e);
        e.n_cx = x->n_cx;
        for (int k = 0; k < x->n_cx; k++) {
            const int rc = convert_cx(A, x->cx[k], m.wavelength, e.cx[k]);
            if (rc != CB2_OK) return rc;
        }
        if (x->n_cx > 1) {
            std::vector<DevPopRate> pop((size_t)(x->n_cx - 1) * d.n_species);
            for (size_t q = 0; q < pop.size(); q++) {
                const int rc = convert_population(A, x->cx_population[q], pop[q]);
                if (rc != CB2_OK) return rc;
            }
            e.pop = A.upload(pop);
        }
        o.ext = A.upload(std::vector<DevModelExt>(1, e));
        o.pec_const = 1;
        o.pec_value = -INFINITY;
    } else if (m.kind == CB2_MODEL_THERMAL_CX_LINE) {
        // thermal_cx.pyx:140-148: one rate per donor species, constant or tabulated
        const cb2_model_ext* x = m.ext;
        if (!x) return cb2_fail(CB2_ERR_RUNTIME, "ThermalCXLine needs its resolved donors");
        if (x->n_donors < 0 || x->n_donors > CB2_MAX_SPECIES) return cb2_fail(CB2_ERR_VALUE, "too many CX donors");
        DevModelExt e;
        memset(&e, 0, sizeof e);
        e.n_donors = x->n_donors;
        for (int k = 0; k < x->n_donors; k++) {
            if (x->donor_species[k] < 0 || x->donor_species[k] >= d.n_species) return cb2_fail(CB2_ERR_VALUE, "donor species index out of range");
            const cb2_rate3d& r3 = x->donor_rates[k];
            e.donor_species[k] = x->donor_species[k];
            e.donor_lrate[k] = r3.constant > 0 ? (float)(log10(r3.constant) + CB2_PEC_LOG_OFFSET) : -INFINITY;
            if (r3.n_ne > 0) {
                // ThermalCXPEC (pec.pyx:153-184): log10(PhotonToJ(rate)) on (log10 ne, log10 te, log10 td)
                if (r3.n_ne < 2 || r3.n_te < 2 || r3.n_td < 2) return cb2_fail(CB2_ERR_VALUE, "thermal CX rate tables need at least 2 knots per axis");
                if (!r3.ne || !r3.te || !r3.td || !r3.rate) return cb2_fail(CB2_ERR_VALUE, "thermal CX rate table pointers missing");
                std::vector<double> lx(r3.n_ne), ly(r3.n_te), lz(r3.n_td), lr((size_t)r3.n_ne * r3.n_te * r3.n_td);
                for (int q = 0; q < r3.n_ne; q++) lx[q] = log10(r3.ne[q]);
                for (int q = 0; q < r3.n_te; q++) ly[q] = log10(r3.te[q]);
                for (int q = 0; q < r3.n_td; q++) lz[q] = log10(r3.td[q]);
                for (int q = 0; q + 1 < r3.n_ne; q++) if (!(lx[q + 1] > lx[q])) return cb2_fail(CB2_ERR_VALUE, "rate table ne grid must be increasing");
                for (int q = 0; q + 1 < r3.n_te; q++) if (!(ly[q + 1] > ly[q])) return cb2_fail(CB2_ERR_VALUE, "rate table te grid must be increasing");
                for (int q = 0; q + 1 < r3.n_td; q++) if (!(lz[q + 1] > lz[q])) return cb2_fail(CB2_ERR_VALUE, "rate table td grid must be increasing");
                const double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9;
                for (size_t q = 0; q < lr.size(); q++) {
                    if (!(r3.rate[q] > 0)) return cb2_fail(CB2_ERR_VALUE, "rate table values must be positive (log10 interpolation)");
                    lr[q] = log10(r3.rate[q] / m.wavelength * conv) + CB2_PEC_LOG_OFFSET;
                }
                e.donor_tab[k] = 1;
                e.donor_extrapolate[k] = r3.extrapolate != 0;
                e.donor_t3[k] = make_table3d(A, lx.data(), ly.data(), lz.data(), lr.data(), r3.n_ne, r3.n_te, r3.n_td);
            }
        }
        o.ext = A.upload(std::vector<DevModelExt>(1, e));
        o.pec_const = 1;
        o.pec_value = -INFINITY;
    } else if (m.pec.n_ne <= 0) {
        o.pec_const = 1;
        o.pec_value = (m.pec.constant > 0) ? (float)(log10(m.pec.constant) + CB2_PEC_LOG_OFFSET) : -INFINITY;
    } else {
        const int nn = m.pec.n_ne, nt = m.pec.n_te;
        if (nn < 2 || nt < 2) return cb2_fail(CB2_ERR_VALUE, "rate tables need at least 2x2 points");
        std::vector<double> lne(nn), lte(nt), lr((size_t)nn * nt);
        for (int i = 0; i < nn; i++) lne[i] = log10(m.pec.ne[i]);
        for (int j = 0; j < nt; j++) lte[j] = log10(m.pec.te[j]);
        const double conv = PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9;
        for (size_t k = 0; k < lr.size(); k++) {
            if (!(m.pec.rate[k] > 0)) return cb2_fail(CB2_ERR_VALUE, "rate table values must be positive (log10 interpolation)");
            lr[k] = log10(m.pec.rate[k] / m.wavelength * conv) + CB2_PEC_LOG_OFFSET;
        }
        for (int i = 0; i + 1 < nn; i++)
            if (!(lne[i + 1] > lne[i])) return cb2_fail(CB2_ERR_VALUE, "rate table ne grid must be increasing");
        for (int j = 0; j + 1 < nt; j++)
            if (!(lte[j + 1] > lte[j])) return cb2_fail(CB2_ERR_VALUE, "rate table te grid must be increasing");
        o.pec = make_table2d(A, lne.data(), lte.data(), lr.data(), nn, nt);
        o.pec_extrapolate = m.pec.extrapolate;
        // share the cell search between models tabulated on identical knots (ADF15 blocks of one file usually are)
        const int self = (int)(&m - d.models);
        o.pec_grid = self;
        for (int p = 0; p < self; p++) {
            const cb2_rate2d& q = d.models[p].pec;
            if (d.models[p].kind == CB2_MODEL_BREMSSTRAHLUNG || q.n_ne != nn || q.n_te != nt) continue;
            if (!memcmp(q.ne, m.pec.ne, sizeof(double) * nn) && !memcmp(q.te, m.pec.te, sizeof(double) * nt)) { o.pec_grid = S.models[p].pec_grid; break; }
        }
    }
    // component slots
    o.comp0 = S.n_comp;
    int n = 0;
    switch (m.shape.kind) {
    case CB2_SHAPE_GAUSSIAN: n = 1; break;
    case CB2_SHAPE_MULTIPLET: n = m.shape.n_components; break;
    case CB2_SHAPE_ZEEMAN_TRIPLET:
    case CB2_SHAPE_PARAM_ZEEMAN: n = 3; break;
    case CB2_SHAPE_ZEEMAN_MULTIPLET: n = m.shape.n_pi + m.shape.n_sigma_plus + m.shape.n_sigma_minus; break;
    case CB2_SHAPE_STARK: n = 6; break;   /* pi, sigma+, sigma-: Gaussian slots 0-2, modified-Lorentzian slots 3-5 */
    default: return cb2_fail(CB2_ERR_TYPE, "unsupported line shape kind %d", m.shape.kind);
    }
    if (n < 1) return cb2_fail(CB2_ERR_VALUE, "line shape has no components");
    if (S.n_comp + n > CB2_MAX_COMP) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "too many line components (max %d)", CB2_MAX_COMP);
    o.ncomp = n;
    S.n_comp += n;
    for (int k = 0; k < n; k++) set_comp(S, o.comp0 + k, m.wavelength, (m.shape.kind == CB2_SHAPE_STARK && k >= 3) ? 1 : 0);
    if (m.shape.kind == CB2_SHAPE_MULTIPLET) {
        std::vector<float> ratio(n), lam(n);
        for (int k = 0; k < n; k++) {
            set_comp(S, o.comp0 + k, m.shape.multiplet[k], 0);
            lam[k] = (float)m.shape.multiplet[k];
            ratio[k] = (float)m.shape.multiplet[n + k];
        }
        o.n_mult = n;
        o.mult_ratio = A.upload(ratio);
        o.mult_lambda = A.upload(lam);
    }
    if (m.shape.kind == CB2_SHAPE_ZEEMAN_MULTIPLET) {
        const int nb = m.shape.n_b;
        if (nb < 2) return cb2_fail(CB2_ERR_VALUE, "ZeemanStructure tables need at least two |B| points");
        if (!is_uniform(m.shape.b_grid, nb)) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "ZeemanStructure |B| grid must be uniform");
        o.n_b = nb; o.n_pi = m.shape.n_pi; o.n_sp = m.shape.n_sigma_plus; o.n_sm = m.shape.n_sigma_minus;
        o.b0 = (float)m.shape.b_grid[0];
        o.inv_db = (float)((nb - 1) / (m.shape.b_grid[nb - 1] - m.shape.b_grid[0]));
        std::vector<float> dl((size_t)n * nb), ra((size_t)n * nb);
        for (size_t k = 0; k < dl.size(); k++) {
            dl[k] = (float)(m.shape.zeeman_wavelength[k] - m.wavelength);
            ra[k] = (float)m.shape.zeeman_ratio[k];
        }
        o.zee_dlambda = A.upload(dl);
        o.zee_ratio = A.upload(ra);
    }
    if (m.shape.kind == CB2_SHAPE_STARK) {
        // c_ij * ne^a / te^b with ne in units of 1e19 m^-3 on device
        o.param[0] = (float)(m.shape.param[0] * pow(1.0 / CB2_DENSITY_SCALE, m.shape.param[1]));
    }
    return A.rc;
}

// Gauss-Legendre nodes on [-1,1] for n = 1..4
static void gl_nodes(int n, double* x, double* w) {
    switch (n) {
    case 1: x[0] = 0; w[0] = 2; break;
    case 2: x[0] = -0.5773502691896257; x[1] = -x[0]; w[0] = w[1] = 1; break;
    case 3: x[0] = -0.7745966692414834; x[1] = 0; x[2] = -x[0]; w[0] = w[2] = 5.0 / 9; w[1] = 8.0 / 9; break;
    default:
        x[0] = -0.8611363115940526; x[1] = -0.3399810435848563; x[2] = -x[1]; x[3] = -x[0];
        w[0] = w[3] = 0.3478548451374538; w[1] = w[2] = 0.6521451548625461;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Bremsstrahlung moment formulation: host-side table (fp64) — see DevBrems in cb2_internal.h and DESIGN.md
// ------------------------------------------------------------------------------------------------------------------
// range of the positive values a scalar field can take; false if it cannot be bounded away from zero / infinity
static bool scalar_bounds(const cb2_scalar_field& f, const cb2_axisym* ax, double& lo, double& hi) {
    lo = INFINITY; hi = -INFINITY;
    auto add = [&](double v) { if (v > 0) { lo = fmin(lo, v); hi = fmax(hi, v); } };
    switch (f.kind) {
    case CB2_FIELD_CONSTANT: add(f.c[0]); break;
    case CB2_FIELD_GAUSSIAN_VOLUME:
        if (!(f.c[0] > 0) || f.c[1] < 0) return false;      // decays to the bias: needs a positive bias
        add(f.c[0]); add(f.c[0] + f.c[1]);
        break;
    case CB2_FIELD_SLAB_ION:
        if (!(f.c[0] > 0) || !(f.c[1] > 0)) return false;
        add(f.c[0]); add(f.c[1]);
        break;
    case CB2_FIELD_AXISYM_BLEND: {
        if (!ax) return false;
        if (f.edge)
            for (int i = 0; i < ax->n_triangles; i++) {
                if (!(f.edge[i] > 0)) return false;           // a zero edge value blends to arbitrarily small positives
                add(f.edge[i]);
            }
        if (f.core) {
            const int n = ax->n_core;
            for (int i = 0; i + 1 < n; i++) {
                const double h = ax->core_psin[i + 1] - ax->core_psin[i];
                double a[4];
                hermite_coef(f.core[i], f.core[i + 1], d1(ax->core_psin, f.core, n, 1, i) * h, d1(ax->core_psin, f.core, n, 1, i + 1) * h, a);
                for (int q = 0; q <= 16; q++) {
                    const double t = q / 16.0, v = a[0] + t * (a[1] + t * (a[2] + t * a[3]));
                    if (!(v > 0)) return false;
                    add(v);
                }
            }
            if (n == 1) add(f.core[0]);
        }
        break;
    }
    default: return false;                                   // SLAB_NEUTRAL decays to zero
    }
    return lo <= hi;
}

struct HostGaunt {
    int nu, ng;
    std::vector<double> lu, lg, coef;
    double u_min, u_max, g_min, g_max;
    // InterpolatedFreeFreeGauntFactor.evaluate (gaunt.pyx:109-140), fp64
    double eval(double z, double te, double wavelength) const {
        if (z == 0) return 0.0;
        const double gamma2 = z * z * 13.605693122994 / te, u = HC_EV_NM / (te * wavelength);
        if (u >= u_max || gamma2 >= g_max) return 1.0;
        if (u < u_min || gamma2 < g_min) return sqrt(3.0) / M_PI * (log(4.0 / u) - 0.5772156649015329);
        const double x = log10(u), y = log10(gamma2);
        int i = (int)(std::upper_bound(lu.begin(), lu.end(), x) - lu.begin()) - 1;
        int j = (int)(std::upper_bound(lg.begin(), lg.end(), y) - lg.begin()) - 1;
        i = std::min(std::max(i, 0), nu - 2);
        j = std::min(std::max(j, 0), ng - 2);
        const double t = (x - lu[i]) / (lu[i + 1] - lu[i]), w = (y - lg[j]) / (lg[j + 1] - lg[j]);
        const double* c = &coef[((size_t)i * (ng - 1) + j) * 16];
        double v = 0.0;
        for (int p = 3; p >= 0; p--) v = v * t + (c[4 * p] + w * (c[4 * p + 1] + w * (c[4 * p + 2] + w * c[4 * p + 3])));
        return v;
    }
};

// s = ln(tau) + tau/tau_c: logarithmic where the Gaunt factor varies with ln(T), linear in 1/T where exp(-x/T) does
static double node_map(double tau, double tau_c) { return log(tau) + tau / tau_c; }
static double node_map_inverse(double s, double tau_c) {
    double tau = exp(fmin(s, log(tau_c)));
    for (int it = 0; it < 100; it++) {
        const double d = (node_map(tau, tau_c) - s) / (1.0 / tau + 1.0 / tau_c);
        tau = fmax(tau - d, 1e-300);
        if (fabs(d) <= 1e-15 * tau) break;
    }
    return tau;
}

// Decide whether the moment formulation applies to this scene and, if so, build its table.
static int build_brems_moments(Arena& A, const cb2_scene_desc& d, DevScene& S) {
    DevBrems& b = S.brems;
    b.mode = 0;
    const char* env = getenv("CB2_BREMS_MODE");
    const bool force_direct = d.brems_quadrature > 0 || (env && !strcmp(env, "direct"));
    const bool force_moments = d.brems_quadrature < 0;
    if (force_direct) return CB2_OK;
    const cb2_gaunt& g = d.gaunt;
    const double lmin = d.grid.min_wavelength, lmax = d.grid.max_wavelength, delta = S.delta_d;
    const char* why = nullptr;
    // distinct charges
    int zval[CB2_MAX_BREMS_Z], nz = 0;
    for (int s = 0; s < b.n_charged && !why; s++) {
        const int z = d.species[b.charged[s]].charge;
        int k = 0;
        while (k < nz && zval[k] != z) k++;
        if (k == nz) {
            if (nz == CB2_MAX_BREMS_Z) { why = "more than 8 distinct ion charges"; break; }
            zval[nz++] = z;
        }
        b.zidx[s] = k;
    }
    if (!why && nz == 0) why = "no charged species";
    if (!why) {
        int cursor = 0;
        for (int k = 0; k < nz; k++) {
            b.zstart[k] = cursor;
            for (int s = 0; s < b.n_charged; s++)
                if (b.zidx[s] == k) b.zlist[cursor++] = b.charged[s];
        }
        for (int k = nz; k <= CB2_MAX_BREMS_Z; k++) b.zstart[k] = cursor;
    }
    double tlo = 0, thi = 0;
    if (!why && !scalar_bounds(d.electron_temperature, d.axisym, tlo, thi)) why = "electron temperature field is not bounded away from zero";
    if (!why) {
        // the node interpolation must not straddle the discontinuities of the reference's Gaunt factor (classical limit
        // u >= u_max or gamma2 >= g2_max, Born approximation u < u_min or gamma2 < g2_min; gaunt.pyx:129-134)
        const double xmin = HC_EV_NM / lmax, xmax = HC_EV_NM / lmin;
        int zmin = zval[0], zmax = zval[0];
        for (int k = 1; k < nz; k++) { zmin = std::min(zmin, zval[k]); zmax = std::max(zmax, zval[k]); }
        const double ry = 13.605693122994;
        double safe_hi = fmin(xmin / g.u[0], zmin * (double)zmin * ry / g.gamma2[0]);
        double safe_lo = fmax(xmax / g.u[g.n_u - 1], zmax * (double)zmax * ry / g.gamma2[g.n_gamma2 - 1]);
        tlo *= 0.98; thi *= 1.02;                               // fp32 evaluation of the profiles on the device
        if (!(thi < 0.98 * safe_hi) || !(tlo > 1.02 * safe_lo)) why = "temperature range reaches a Gaunt-factor limit switch";
        else if (tlo < 0.02) why = "electron temperature below 0.02 eV";
    }
    // node spacing: ds in s resolves the Gaunt factor's dependence on ln(T) (error ~1e-5 at 0.06, tools/proto_brems_moments.py);
    // in the cold, linear-in-tau part the spacing ds*tau_c must resolve exp(-(x - x_ref) tau): ((x - x_ref) ds tau_c)^4/24 <= 1e-5
    double ds = 0.06;
    if (const char* e = getenv("CB2_BREMS_DS")) { const double v = atof(e); if (v > 1e-3 && v < 1.0) ds = v; }
    const double dx_max = 0.5 * (HC_EV_NM / lmin - HC_EV_NM / lmax);
    const double tau_c = fmin(4.0, 0.125 / (ds * fmax(dx_max, 1e-6)));
    int M = 0;
    double s0 = 0;
    if (!why) {
        const double sa = node_map(1.0 / thi, tau_c), sb = node_map(1.0 / tlo, tau_c);
        M = (int)ceil((sb - sa) / ds) + 4;                      // stencil i-1..i+2 with 1 <= i <= M-3
        s0 = sa - ds;
        if ((long)M * nz > 6144) why = "temperature range needs too many nodes";
    }
    if (why) {
        if (force_moments) return cb2_fail(CB2_ERR_VALUE, "Bremsstrahlung moment formulation not applicable: %s", why);
        return CB2_OK;
    }
    HostGaunt hg;
    hg.nu = g.n_u; hg.ng = g.n_gamma2;
    hg.lu.resize(g.n_u); hg.lg.resize(g.n_gamma2);
    for (int i = 0; i < g.n_u; i++) hg.lu[i] = log10(g.u[i]);
    for (int i = 0; i < g.n_gamma2; i++) hg.lg[i] = log10(g.gamma2[i]);
    hg.u_min = g.u[0]; hg.u_max = g.u[g.n_u - 1]; hg.g_min = g.gamma2[0]; hg.g_max = g.gamma2[g.n_gamma2 - 1];
    build_coef2d(hg.lu.data(), hg.lg.data(), g.gaunt, g.n_u, g.n_gamma2, hg.coef);

    const int bins = S.bins;
    const int n_pad = (bins + 127) / 128 * 128, k_pad = (nz * M + 15) / 16 * 16;
    const double x_ref = 0.5 * (HC_EV_NM / lmin + HC_EV_NM / lmax);
    static const double gx[5] = {-0.9061798459386640, -0.5384693101056831, 0.0, 0.5384693101056831, 0.9061798459386640};
    static const double gw[5] = {0.2369268850561891, 0.4786286704993665, 0.5688888888888889, 0.4786286704993665, 0.2369268850561891};
    std::vector<float> phi((size_t)k_pad * n_pad, 0.f);
    std::vector<double> lam((size_t)bins * 5), geom((size_t)bins * 5);
    for (int j = 0; j < bins; j++)
        for (int q = 0; q < 5; q++) {
            const double l = lmin + delta * (j + 0.5 + 0.5 * gx[q]);
            lam[(size_t)j * 5 + q] = l;
            geom[(size_t)j * 5 + q] = 0.5 * gw[q] / (l * l);
        }
    for (int m = 0; m < M; m++) {
        const double tau = node_map_inverse(s0 + m * ds, tau_c), te = 1.0 / tau;
        for (int k = 0; k < nz; k++) {
            float* row = &phi[(size_t)(k * M + m) * n_pad];
            const double z = zval[k], z2 = z * z;
            for (int j = 0; j < bins; j++) {
                double v = 0.0;
                for (int q = 0; q < 5; q++) {
                    const double l = lam[(size_t)j * 5 + q];
                    v += geom[(size_t)j * 5 + q] * hg.eval(z, te, l) * exp(-(HC_EV_NM / l - x_ref) * tau);
                }
                row[j] = (float)(v * z2);
            }
        }
    }
    b.phi = A.upload(phi);
    b.mode = 3;
    b.n_z = nz;
    b.n_nodes = M;
    b.k_pad = k_pad;
    b.n_pad = n_pad;
    b.s0 = (float)s0;
    b.inv_ds = (float)(1.0 / ds);
    b.inv_tau_c = (float)(1.0 / tau_c);
    b.x_ref = (float)x_ref;
    b.te_lo = (float)(1.0 / node_map_inverse(s0 + (M - 2) * ds, tau_c));
    b.te_hi = (float)(1.0 / node_map_inverse(s0 + 1 * ds, tau_c));
    return A.rc;
}

static int convert_brems(Arena& A, const cb2_scene_desc& d, DevScene& S) {
    DevBrems& b = S.brems;
    memset(&b, 0, sizeof b);
    bool present = false;
    for (int m = 0; m < d.n_models; m++) present |= d.models[m].kind == CB2_MODEL_BREMSSTRAHLUNG;
    if (!present) return CB2_OK;
    const cb2_gaunt& g = d.gaunt;
    if (g.n_u < 2 || g.n_gamma2 < 2) return cb2_fail(CB2_ERR_RUNTIME, "Bremsstrahlung needs a free-free Gaunt factor table");
    b.present = 1;
    const double lmin = d.grid.min_wavelength, lmax = d.grid.max_wavelength, delta = S.delta_d;
    if (!(lmin > 0)) return cb2_fail(CB2_ERR_VALUE, "Bremsstrahlung needs min_wavelength > 0");
    if (lmax / lmin >= 100.0) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "spectral windows spanning more than two decades are not supported");
    // quadrature order from a midpoint/Gauss error bound at Te = 0.1 eV, lambda = lambda_min (kernel K1)
    int nq = d.brems_quadrature;
    if (nq <= 0) {
        const double a = HC_EV_NM / 0.1, dlog = fabs(a / (lmin * lmin) - 2.0 / lmin) + 0.05 / lmin;
        const double h = 0.5 * delta * dlog;
        nq = 4;
        const double bound[5] = {0, h * h / 6.0, pow(h, 4) / 270.0, pow(h, 6) / 31500.0, 0.0};
        for (int n = 1; n <= 3; n++)
            if (bound[n] < 2e-6) { nq = n; break; }
    }
    if (nq > 4) nq = 4;
    // per-sample choice on the device: the one-point (bin centre) rule wherever its error model is below 2e-6,
    // otherwise this nq-point Gauss-Legendre rule (at least 2 points unless the caller forces 1)
    if (d.brems_quadrature <= 0 && nq < 2) nq = 2;
    b.nq = nq;
    b.mid_c1 = (float)(0.5 * delta * HC_EV_NM / (lmin * lmin));
    b.mid_c0 = (float)(0.5 * delta * 2.0 / lmin);
    double gx[4], gw[4];
    gl_nodes(nq, gx, gw);
    const double lref = log10(0.5 * (lmin + lmax));
    // the tables are padded to the largest CTA tile (8 warps x 32 lanes x 16 bins) so that they do not depend on the
    // launch configuration, which is chosen after the Bremsstrahlung mode is known
    const int tab_bins = std::max(4096, (S.bins + 4095) / 4096 * 4096);
    std::vector<float4> tab((size_t)tab_bins * nq);
    for (int i = 0; i < tab_bins; i++)
        for (int q = 0; q < nq; q++) {
            const double lam = lmin + delta * (i + 0.5 + 0.5 * gx[q]);
            const double rho = 1.0 / lam;
            tab[(size_t)i * nq + q] = make_float4((float)rho, (float)(2.0 * log2(rho)), (float)(log10(lam) - lref), (float)(0.5 * gw[q]));
        }
    b.bin_tab = A.upload(tab);
    {
        std::vector<float4> tab1((size_t)tab_bins);
        for (int i = 0; i < tab_bins; i++) {
            const double lam = lmin + delta * (i + 0.5), rho = 1.0 / lam;
            tab1[i] = make_float4((float)rho, (float)(2.0 * log2(rho)), (float)(log10(lam) - lref), 1.0f);
        }
        b.bin_tab1 = A.upload(tab1);
    }
    b.lref = (float)lref;
    b.log_hc = (float)log10(HC_EV_NM);
    b.exp_coef = (float)(PLANCK_CONSTANT * SPEED_OF_LIGHT * 1e9 / ELEMENTARY_CHARGE * 1.4426950408889634);
    double bc = pow(ELEMENTARY_CHARGE * ELEMENTARY_CHARGE / (4.0 * M_PI) / VACUUM_PERMITTIVITY, 3);
    bc *= 32 * M_PI * M_PI / (3 * sqrt(3.0) * ELECTRON_REST_MASS * ELECTRON_REST_MASS * SPEED_OF_LIGHT * SPEED_OF_LIGHT * SPEED_OF_LIGHT);
    bc *= sqrt(2 * ELECTRON_REST_MASS / (M_PI * ELEMENTARY_CHARGE));
    bc *= SPEED_OF_LIGHT * 1e9 / (4.0 * M_PI);
    b.pref = (float)(bc / (CB2_DENSITY_SCALE * CB2_DENSITY_SCALE));
    b.rho_min = (float)(1.0 / lmax);
    b.rho_max = (float)(1.0 / lmin);
    b.lp_min = (float)(log10(lmin) - lref);
    b.lp_max = (float)(log10(lmax) - lref);
    std::vector<double> lu(g.n_u), lg(g.n_gamma2);
    for (int i = 0; i < g.n_u; i++) lu[i] = log10(g.u[i]);
    for (int i = 0; i < g.n_gamma2; i++) lg[i] = log10(g.gamma2[i]);
    for (int i = 0; i + 1 < g.n_u; i++)
        if (!(lu[i + 1] > lu[i])) return cb2_fail(CB2_ERR_VALUE, "Gaunt table u grid must be increasing");
    for (int i = 0; i + 1 < g.n_gamma2; i++)
        if (!(lg[i + 1] > lg[i])) return cb2_fail(CB2_ERR_VALUE, "Gaunt table gamma2 grid must be increasing");
    b.gaunt = make_table2d(A, lu.data(), lg.data(), g.gaunt, g.n_u, g.n_gamma2);
    b.lu_min = (float)lu[0]; b.lu_max = (float)lu[g.n_u - 1];
    b.lg_min = (float)lg[0]; b.lg_max = (float)lg[g.n_gamma2 - 1];
    b.n_charged = 0;
    for (int i = 0; i < d.n_species; i++)
        if (d.species[i].charge > 0) b.charged[b.n_charged++] = i;
    if (A.rc != CB2_OK) return A.rc;
    return build_brems_moments(A, d, S);
}

// Universal cumulative profile of the modified Lorentzian of StarkBroadenedLine (stark.pyx:52-81) in u = (x - x0)/FWHM:
// s(u) = K / (|u|^2.5 + A), K = 0.5^1.5 / C, A = 0.5^2.5, C = 4*50*2F1(0.4, 1, 1.4, -(100)^2.5) so that the integral over
// [-50, 50] is 1.  Phi(u) = int_0^u s; table on [0, 4] (step 1/512, composite 16-point Gauss-Legendre per interval).
#define CB2_STARK_NORM 2.641279471021934
#define CB2_LORENTZ_KNOTS 2048
#define CB2_LORENTZ_UMAX 4.0

static double lorentz_density(double u) {
    const double K = pow(0.5, 1.5) / CB2_STARK_NORM, A = pow(0.5, 2.5);
    return K / (pow(fabs(u), 2.5) + A);
}

// tail mass int_u^inf s(v) dv for u >= 4: K sum_n (-A)^n u^-(2.5 n + 1.5) / (2.5 n + 1.5)
static double lorentz_tail(double u) {
    const double K = pow(0.5, 1.5) / CB2_STARK_NORM, A = pow(0.5, 2.5);
    const double r = pow(u, -2.5);
    double term = pow(u, -1.5), sum = 0.0;
    for (int n = 0; n < 12; n++) {
        sum += term / (2.5 * n + 1.5);
        term *= -A * r;
    }
    return K * sum;
}

static int build_lorentz(Arena& A, DevScene& S) {
    static const double gx[8] = {0.0950125098376374, 0.2816035507792589, 0.4580167776572274, 0.6178762444026438,
                                 0.7554044083550030, 0.8656312023878318, 0.9445750230732326, 0.9894009349916499};
    static const double gw[8] = {0.1894506104550685, 0.1826034150449236, 0.1691565193950025, 0.1495959888165767,
                                 0.1246289712555339, 0.0951585116824928, 0.0622535239386479, 0.0271524594117541};
    std::vector<double2> tab(CB2_LORENTZ_KNOTS + 1);
    const double h = CB2_LORENTZ_UMAX / CB2_LORENTZ_KNOTS;
    double acc = 0.0;
    tab[0] = make_double2(0.0, lorentz_density(0.0));
    for (int k = 0; k < CB2_LORENTZ_KNOTS; k++) {
        // the integrand has a |u|^2.5 cusp at 0: split the first intervals finer
        const int sub = k < 8 ? 64 : 1;
        for (int q = 0; q < sub; q++) {
            const double a = (k + (double)q / sub) * h, b = (k + (double)(q + 1) / sub) * h, c = 0.5 * (a + b), d = 0.5 * (b - a);
            double v = 0.0;
            for (int i = 0; i < 8; i++) v += gw[i] * (lorentz_density(c + d * gx[i]) + lorentz_density(c - d * gx[i]));
            acc += v * d;
        }
        tab[k + 1] = make_double2(acc, lorentz_density((k + 1) * h));
    }
    S.has_lorentz = 1;
    S.lorentz_tab = A.upload(tab);
    S.lorentz_phi_inf = acc + lorentz_tail(CB2_LORENTZ_UMAX);
    // self-check of the normalisation: the profile integrates to 1 over [-50, 50] FWHM (stark.pyx:62-66)
    if (fabs(2.0 * (S.lorentz_phi_inf - lorentz_tail(50.0)) - 1.0) > 1e-9)
        return cb2_fail(CB2_ERR_RUNTIME, "internal error: modified-Lorentzian table is not normalised");
    return A.rc;
}

// ------------------------------------------------------------------------------------------------------------------
// scene create / destroy
// ------------------------------------------------------------------------------------------------------------------
extern "C" int cb2_scene_create(const cb2_scene_desc* d, int device, cb2_scene** out) {
    if (!d || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    *out = nullptr;
    if (d->abi_version != CB2_ABI_VERSION) return cb2_fail(CB2_ERR_VALUE, "abi_version mismatch (header %d, descriptor %d)", CB2_ABI_VERSION, d->abi_version);
    if (d->n_species > CB2_MAX_SPECIES) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "too many species (max %d)", CB2_MAX_SPECIES);
    if (d->n_models > CB2_MAX_MODELS) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "too many models (max %d)", CB2_MAX_MODELS);
    if (d->grid.bins < 1 || !(d->grid.max_wavelength > d->grid.min_wavelength)) return cb2_fail(CB2_ERR_VALUE, "invalid spectral grid");
    if (!(d->step > 0)) return cb2_fail(CB2_ERR_VALUE, "Numerical integration step size can not be less than or equal to zero");
    if (d->min_samples < 2) return cb2_fail(CB2_ERR_VALUE, "At least two samples are required to perform the numerical integration.");
    int ndev = cb2_device_count();
    if (ndev <= 0) return ndev < 0 ? CB2_ERR_CUDA : cb2_fail(CB2_ERR_CUDA, "no CUDA device visible (libcherab_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return cb2_fail(CB2_ERR_VALUE, "device %d out of range (%d visible)", device, ndev);
    CB2_CUDA(cudaSetDevice(device));

    cb2_scene* sc = (cb2_scene*)calloc(1, sizeof(cb2_scene));
    if (!sc) return cb2_fail(CB2_ERR_MEMORY, "out of host memory");
    sc->device = device;
    DevScene& S = sc->host;
    Arena A;
    int rc = CB2_OK;
    do {
        S.n_species = d->n_species;
        S.n_models = d->n_models;
        S.bins = d->grid.bins;
        S.min_wavelength_d = d->grid.min_wavelength;
        S.delta_d = (d->grid.max_wavelength - d->grid.min_wavelength) / d->grid.bins;
        S.min_wavelength = (float)S.min_wavelength_d;
        S.delta = (float)S.delta_d;
        S.step = d->step;
        S.min_samples = d->min_samples;
        S.inv_range = (float)(1.0 / (d->grid.max_wavelength - d->grid.min_wavelength));
        for (int k = 0; k < 12; k++) S.w2p[k] = d->world_to_plasma[k];
        if ((rc = convert_scalar(A, d->electron_density, d->axisym, CB2_DENSITY_SCALE, S.ne)) != CB2_OK) break;
        if ((rc = convert_scalar(A, d->electron_temperature, d->axisym, 1.0, S.te)) != CB2_OK) break;
        bool need_pol = false;
        for (int i = 0; i < d->n_species && rc == CB2_OK; i++) {
            S.species[i].charge = d->species[i].charge;
            S.species[i].z2 = (float)(d->species[i].charge * (double)d->species[i].charge);
            if ((rc = convert_scalar(A, d->species[i].density, d->axisym, CB2_DENSITY_SCALE, S.species[i].density)) != CB2_OK) break;
            if ((rc = convert_scalar(A, d->species[i].temperature, d->axisym, 1.0, S.species[i].temperature)) != CB2_OK) break;
            if ((rc = convert_vector(A, d->species[i].velocity, d->axisym, S.species[i].velocity)) != CB2_OK) break;
            need_pol |= vec_has_pol(d->species[i].velocity, d->axisym);
        }
        if (rc != CB2_OK) break;
        if (d->axisym && (rc = convert_axisym(A, *d->axisym, S.ax)) != CB2_OK) break;
        S.b_kind = d->b_field_kind;
        if (S.b_kind == 1 && !d->axisym) { rc = cb2_fail(CB2_ERR_RUNTIME, "EFIT b_field needs an axisym context"); break; }
        for (int k = 0; k < 3; k++) S.b_const[k] = (float)d->b_field[k];
        S.n_comp = 0;
        bool need_b = false;
        for (int m = 0; m < d->n_models; m++) {
            if ((rc = convert_model(A, *d, d->models[m], S, S.models[m])) != CB2_OK) break;
            const int sh = d->models[m].shape.kind, kd = d->models[m].kind;
            const bool is_line = kd == CB2_MODEL_EXCITATION_LINE || kd == CB2_MODEL_RECOMBINATION_LINE || kd == CB2_MODEL_THERMAL_CX_LINE ||
                                 kd == CB2_MODEL_BEAM_CX_LINE;
            if (is_line && sh != CB2_SHAPE_GAUSSIAN && sh != CB2_SHAPE_MULTIPLET) need_b = true;
            if (kd == CB2_MODEL_BEAM_CX_LINE || kd == CB2_MODEL_BEAM_EMISSION_LINE) need_b = true;   // q_eff(|B|); v x B
            if ((kd == CB2_MODEL_BEAM_CX_LINE || kd == CB2_MODEL_BEAM_EMISSION_LINE) != (d->beam != nullptr)) { rc = cb2_fail(CB2_ERR_TYPE, "a beam scene renders beam models only, and beam models need a beam"); break; }
        }
        if (rc != CB2_OK) break;
        S.need_b = need_b;
        S.need_pol = need_pol || (need_b && S.b_kind == 1);
        bool has_brems = false;
        for (int m = 0; m < d->n_models; m++) has_brems |= d->models[m].kind == CB2_MODEL_BREMSSTRAHLUNG;
        S.brems.present = has_brems;
        {
            bool ax_only = d->axisym && d->electron_density.kind == CB2_FIELD_AXISYM_BLEND && d->electron_temperature.kind == CB2_FIELD_AXISYM_BLEND;
            for (int i = 0; i < d->n_species; i++)
                ax_only = ax_only && d->species[i].density.kind == CB2_FIELD_AXISYM_BLEND && d->species[i].temperature.kind == CB2_FIELD_AXISYM_BLEND;
            sc->ax_only = ax_only;
            sc->feat = d->beam != nullptr;
            for (int m = 0; m < d->n_models; m++)
                sc->feat |= d->models[m].kind == CB2_MODEL_THERMAL_CX_LINE || d->models[m].kind == CB2_MODEL_TOTAL_RADIATED_POWER ||
                            d->models[m].kind == CB2_MODEL_BEAM_CX_LINE || d->models[m].kind == CB2_MODEL_BEAM_EMISSION_LINE;
        }
        if ((rc = convert_brems(A, *d, S)) != CB2_OK) break;  // decides the Bremsstrahlung mode (direct / moments)
        if ((rc = cb2_emission_config(sc)) != CB2_OK) break;  // sets nw, bpl, bins_padded (needs n_comp and the mode for the shared-memory budget)
        if (S.brems.present && S.brems.mode == 3 && (rc = cb2_contract_tc_init(sc, S.brems.phi, S.brems.k_pad, S.brems.n_pad)) != CB2_OK) break;
        bool any_stark = false;
        for (int m = 0; m < d->n_models; m++)
            any_stark |= d->models[m].kind != CB2_MODEL_BREMSSTRAHLUNG && d->models[m].kind != CB2_MODEL_TOTAL_RADIATED_POWER &&
                         d->models[m].kind != CB2_MODEL_BEAM_EMISSION_LINE && d->models[m].shape.kind == CB2_SHAPE_STARK;
        if (any_stark && (rc = build_lorentz(A, S)) != CB2_OK) break;
        if ((rc = A.rc) != CB2_OK) break;
        void* p = nullptr;
        if ((rc = cb2_cuda_check(cudaMalloc(&p, sizeof(DevScene)), "cudaMalloc(scene)")) != CB2_OK) break;
        A.ptrs.push_back(p);
        sc->dev = (DevScene*)p;
        if ((rc = cb2_cuda_check(cudaMemcpy(p, &S, sizeof(DevScene), cudaMemcpyHostToDevice), "cudaMemcpy(scene)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMalloc(&p, sizeof(cb2_stats)), "cudaMalloc(stats)")) != CB2_OK) break;
        A.ptrs.push_back(p);
        sc->stats_dev = (unsigned long long*)p;
        if ((rc = cb2_memo_build(sc)) != CB2_OK) break;
        if (d->beam) {
            // the attenuation table needs the plasma state on the beam axis, which the device evaluates from the tables just uploaded
            if ((rc = build_beam(A, *d, sc)) != CB2_OK) break;
            if ((rc = cb2_cuda_check(cudaMemcpy(sc->dev, &S, sizeof(DevScene), cudaMemcpyHostToDevice), "cudaMemcpy(scene)")) != CB2_OK) break;
        }
    } while (0);
    if (rc != CB2_OK) {
        A.release();
        free(sc);
        return rc;
    }
    sc->n_allocs = (int)A.ptrs.size();
    sc->allocs = (void**)malloc(sizeof(void*) * A.ptrs.size());
    memcpy(sc->allocs, A.ptrs.data(), sizeof(void*) * A.ptrs.size());
    *out = sc;
    return CB2_OK;
}

static void free_stage(void** stage, size_t* bytes) {
    for (int i = 0; i < 8; i++) {
        if (stage[i]) cudaFree(stage[i]);
        stage[i] = nullptr;
        bytes[i] = 0;
    }
}

extern "C" int cb2_scene_destroy(cb2_scene* sc) {
    if (!sc) return CB2_OK;
    cudaSetDevice(sc->device);
    for (int i = 0; i < sc->n_allocs; i++) cudaFree(sc->allocs[i]);
    free(sc->allocs);
    free_stage(sc->stage, sc->stage_bytes);
    if (sc->mom) cudaFree(sc->mom);
    cb2_contract_tc_destroy(sc);
    if (sc->gbase) cudaFree(sc->gbase);
    if (sc->gmask) cudaFree(sc->gmask);
    if (sc->gblend) cudaFree(sc->gblend);
    if (sc->memo.core) cudaFree((void*)sc->memo.core);
    if (sc->memo.edge) cudaFree((void*)sc->memo.edge);
    if (sc->rec) cudaFree(sc->rec);
    if (sc->flat) cudaFree(sc->flat);
    if (sc->total_host) cudaFreeHost(sc->total_host);
    for (int i = 0; i < 10; i++)
        if (sc->prof_ev[i]) cudaEventDestroy(sc->prof_ev[i]);
    if (sc->copy_ev) cudaEventDestroy(sc->copy_ev);
    if (sc->copy_stream) cudaStreamDestroy(sc->copy_stream);
    if (sc->copy_stream2) cudaStreamDestroy(sc->copy_stream2);
    free(sc);
    return CB2_OK;
}

// grow-only device staging buffer
// ------------------------------------------------------------------------------------------------------------------
// Device -> pageable host memory, large transfers: cudaMemcpy into pageable memory runs at the speed of one CPU thread copying out of
// the driver's staging buffer and faulting the destination pages in (measured 3.3 GB/s: 935 ms for the 3 GB CSR of C4).  Here the
// transfer is cut into chunks that go through a ring of pinned buffers at PCIe speed while worker threads copy finished chunks to
// their destination in parallel (each thread a slice of every chunk, so the first-touch faults are spread too).  A pinned
// (cudaHostAlloc / cudaHostRegister) destination takes one plain asynchronous copy.  The stream is synchronised on return.
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int D2H_RING = 4;
constexpr size_t D2H_CHUNK = (size_t)32 << 20;
struct D2HRing {
    std::mutex lock;
    void* buf[D2H_RING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[D2H_RING] = {nullptr, nullptr, nullptr, nullptr};
};
D2HRing g_d2h;
}  // namespace

int cb2_d2h(void* dst, const void* src_dev, size_t bytes, cudaStream_t st) {
    if (!bytes) return CB2_OK;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    static const bool chunked = !(getenv("CB2_D2H_CHUNKED") && atoi(getenv("CB2_D2H_CHUNKED")) == 0);
    if (pinned || bytes < 2 * D2H_CHUNK || !chunked) {
        CB2_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        CB2_CUDA(cudaStreamSynchronize(st));
        return CB2_OK;
    }
    {
        // fresh destination memory is faulted in page by page as the workers write it (732 000 faults for the 3 GB CSR of C4): ask for
        // transparent huge pages on the part of the range that is 2 MB aligned (a hint; ignored where THP is off or the pages exist)
        const uintptr_t a = ((uintptr_t)dst + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), e = ((uintptr_t)dst + bytes) & ~(((uintptr_t)2 << 20) - 1);
        if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
    }
    std::lock_guard<std::mutex> guard(g_d2h.lock);
    for (int b = 0; b < D2H_RING; b++) {
        if (!g_d2h.buf[b]) CB2_CUDA(cudaHostAlloc(&g_d2h.buf[b], D2H_CHUNK, cudaHostAllocDefault));
        if (!g_d2h.ev[b]) CB2_CUDA(cudaEventCreateWithFlags(&g_d2h.ev[b], cudaEventDisableTiming));
    }
    const size_t n_chunks = (bytes + D2H_CHUNK - 1) / D2H_CHUNK;
    const int T = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<size_t> issued(0);                       // chunks whose copy has been enqueued and whose event has been recorded
    std::vector<std::atomic<int>> done(n_chunks);        // worker slices finished per chunk
    for (auto& d : done) d.store(0);
    std::atomic<int> failed(0);
    int device = 0;
    cudaGetDevice(&device);
    auto worker = [&](int t) {
        cudaSetDevice(device);
        for (size_t i = 0; i < n_chunks; i++) {
            while (issued.load(std::memory_order_acquire) <= i && !failed.load()) std::this_thread::yield();
            if (failed.load()) return;
            if (cudaEventSynchronize(g_d2h.ev[i % D2H_RING]) != cudaSuccess) { failed.store(1); return; }
            const size_t len = std::min(D2H_CHUNK, bytes - i * D2H_CHUNK);
            const size_t a = len * (size_t)t / T / 64 * 64, b = (t + 1 == T) ? len : len * (size_t)(t + 1) / T / 64 * 64;
            if (b > a) memcpy((char*)dst + i * D2H_CHUNK + a, (const char*)g_d2h.buf[i % D2H_RING] + a, b - a);
            done[i].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) pool.emplace_back(worker, t);
    cudaError_t err = cudaSuccess;
    for (size_t i = 0; i < n_chunks && err == cudaSuccess; i++) {
        if (i >= (size_t)D2H_RING)                       // the ring slot is free once every worker has copied its previous content out
            while (done[i - D2H_RING].load(std::memory_order_acquire) < T && !failed.load()) std::this_thread::yield();
        if (failed.load()) break;
        const size_t len = std::min(D2H_CHUNK, bytes - i * D2H_CHUNK);
        err = cudaMemcpyAsync(g_d2h.buf[i % D2H_RING], (const char*)src_dev + i * D2H_CHUNK, len, cudaMemcpyDeviceToHost, st);
        if (err == cudaSuccess) err = cudaEventRecord(g_d2h.ev[i % D2H_RING], st);
        if (err == cudaSuccess) issued.store(i + 1, std::memory_order_release);
    }
    if (err != cudaSuccess) failed.store(1);
    for (auto& th : pool) th.join();
    if (err != cudaSuccess) return cb2_cuda_check(err, "chunked device -> host copy");
    if (failed.load()) return cb2_fail(CB2_ERR_CUDA, "chunked device -> host copy failed");
    CB2_CUDA(cudaStreamSynchronize(st));
    return CB2_OK;
}

// Rows src_dev[i] (i < n, row_bytes each) to dst_base + dest_row[i] * row_bytes, asynchronously on `st`: consecutive destination
// rows form a run; runs of equal length at a constant pitch form one cudaMemcpy2DAsync (a 16 x 16 tile of a pixel-ordered frame:
// 16 runs of 16 rows, pitch = ny rows).
// the copy that starts at ray i: `len` consecutive destination rows, repeated `h` times at a constant `pitch` (in rows)
static void rows_next_op(const int64_t* dest_row, int64_t n, int64_t i, int64_t& len, int64_t& h, int64_t& pitch) {
    len = 1;
    while (i + len < n && dest_row[i + len] == dest_row[i] + len) len++;
    h = 1; pitch = 0;
    while (i + (h + 1) * len <= n) {
        const int64_t j = i + h * len;
        const int64_t p = dest_row[j] - dest_row[j - len];
        if (h == 1) { if (p <= len) break; pitch = p; } else if (p != pitch) break;
        bool run = true;
        for (int64_t k = 1; k < len && run; k++) run = dest_row[j + k] == dest_row[j] + k;
        if (!run) break;
        h++;
    }
}

// The copy plan of cb2_emission_render_rows for a destination-row list, without a device: plan[4 k .. 4 k + 3] = (first ray, run
// length, repeats, pitch in rows) of copy k.  Returns the number of copies (the plan is filled up to `capacity` of them).
extern "C" int64_t cb2_rows_plan(const int64_t* dest_row, int64_t n, int64_t* plan, int64_t capacity) {
    int64_t k = 0;
    for (int64_t i = 0; i < n; k++) {
        int64_t len, h, pitch;
        rows_next_op(dest_row, n, i, len, h, pitch);
        if (plan && k < capacity) { plan[4 * k] = i; plan[4 * k + 1] = len; plan[4 * k + 2] = h; plan[4 * k + 3] = pitch; }
        i += h * len;
    }
    return k;
}

int cb2_d2h_rows(void* dst_base, const int64_t* dest_row, int64_t n, const void* src_dev, size_t row_bytes, cudaStream_t st, cudaStream_t st2) {
    int64_t i = 0;
    int turn = 0;
    while (i < n) {
        cudaStream_t cs = (st2 && (turn++ & 1)) ? st2 : st;
        int64_t len, h, pitch;
        rows_next_op(dest_row, n, i, len, h, pitch);
        char* d = (char*)dst_base + (size_t)dest_row[i] * row_bytes;
        const char* s = (const char*)src_dev + (size_t)i * row_bytes;
        if (h == 1) CB2_CUDA(cudaMemcpyAsync(d, s, (size_t)len * row_bytes, cudaMemcpyDeviceToHost, cs));
        else CB2_CUDA(cudaMemcpy2DAsync(d, (size_t)pitch * row_bytes, s, (size_t)len * row_bytes, (size_t)len * row_bytes, (size_t)h,
                                        cudaMemcpyDeviceToHost, cs));
        i += h * len;
    }
    return CB2_OK;
}

static int stage_reserve(void** stage, size_t* bytes, int slot, size_t need) {
    if (bytes[slot] >= need && stage[slot]) return CB2_OK;
    if (stage[slot]) cudaFree(stage[slot]);
    stage[slot] = nullptr;
    bytes[slot] = 0;
    size_t cap = need + need / 4 + 256;
    CB2_CUDA(cudaMalloc(&stage[slot], cap));
    bytes[slot] = cap;
    return CB2_OK;
}

static int check_rays(const cb2_rays* r) {
    if (!r) return cb2_fail(CB2_ERR_VALUE, "null rays");
    if (r->n_rays < 0 || r->n_segments < 0) return cb2_fail(CB2_ERR_VALUE, "negative ray count");
    if (r->n_rays > 0 && (!r->origin || !r->direction || !r->seg_offset)) return cb2_fail(CB2_ERR_VALUE, "null ray arrays");
    return CB2_OK;
}

// upload host rays into staging slots 0..4
static int upload_rays(void** stage, size_t* bytes, const cb2_rays* r, DevRays& dr, cudaStream_t st) {
    const size_t n = (size_t)r->n_rays, ns = (size_t)r->n_segments;
    int rc;
    // the kernels index seg_t0 / seg_t1 with these: a list that is not monotone inside [0, n_segments] would read out of bounds
    if (r->seg_offset[0] < 0 || r->seg_offset[n] > r->n_segments) return cb2_fail(CB2_ERR_VALUE, "seg_offset leaves [0, n_segments]");
    for (size_t i = 0; i < n; i++)
        if (r->seg_offset[i + 1] < r->seg_offset[i]) return cb2_fail(CB2_ERR_VALUE, "seg_offset is not monotone at ray %lld", (long long)i);
    if (ns && (!r->seg_t0 || !r->seg_t1)) return cb2_fail(CB2_ERR_VALUE, "null segment arrays");
    if ((rc = stage_reserve(stage, bytes, 0, n * 3 * sizeof(double))) != CB2_OK) return rc;
    if ((rc = stage_reserve(stage, bytes, 1, n * 3 * sizeof(double))) != CB2_OK) return rc;
    if ((rc = stage_reserve(stage, bytes, 2, (n + 1) * sizeof(int64_t))) != CB2_OK) return rc;
    if ((rc = stage_reserve(stage, bytes, 3, ns * sizeof(double))) != CB2_OK) return rc;
    if ((rc = stage_reserve(stage, bytes, 4, ns * sizeof(double))) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpyAsync(stage[0], r->origin, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CB2_CUDA(cudaMemcpyAsync(stage[1], r->direction, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CB2_CUDA(cudaMemcpyAsync(stage[2], r->seg_offset, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (ns) {
        CB2_CUDA(cudaMemcpyAsync(stage[3], r->seg_t0, ns * sizeof(double), cudaMemcpyHostToDevice, st));
        CB2_CUDA(cudaMemcpyAsync(stage[4], r->seg_t1, ns * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    dr.n_rays = r->n_rays;
    dr.origin = (const double*)stage[0];
    dr.direction = (const double*)stage[1];
    dr.seg_offset = (const int64_t*)stage[2];
    dr.seg_t0 = (const double*)stage[3];
    dr.seg_t1 = (const double*)stage[4];
    return CB2_OK;
}

static DevRays as_dev_rays(const cb2_rays* r) {
    DevRays dr;
    dr.n_rays = r->n_rays;
    dr.origin = r->origin;
    dr.direction = r->direction;
    dr.seg_offset = r->seg_offset;
    dr.seg_t0 = r->seg_t0;
    dr.seg_t1 = r->seg_t1;
    return dr;
}

extern "C" int cb2_emission_render_device(cb2_scene* sc, const cb2_rays* rays, void* out, int out_f64, double scale,
                                          int accumulate, cb2_stats* stats_dev, void* stream) {
    if (!sc || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    int rc = check_rays(rays);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaSetDevice(sc->device));
    if (rays->n_rays == 0) return CB2_OK;
    return cb2_launch_emission(sc, as_dev_rays(rays), out, out_f64, scale, accumulate, (unsigned long long*)stats_dev, (cudaStream_t)stream);
}

static int emission_render_host(cb2_scene* sc, const cb2_rays* rays, const int64_t* dest_row, void* out, int out_f64, double scale,
                                int accumulate, cb2_stats* stats) {
    if (!sc || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    int rc = check_rays(rays);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaSetDevice(sc->device));
    if (stats) memset(stats, 0, sizeof *stats);
    if (rays->n_rays == 0) return CB2_OK;
    if (dest_row)
        for (int64_t i = 0; i < rays->n_rays; i++)
            if (dest_row[i] < 0) return cb2_fail(CB2_ERR_VALUE, "negative destination row for ray %lld", (long long)i);
    cudaStream_t st = 0;
    DevRays dr;
    if ((rc = upload_rays(sc->stage, sc->stage_bytes, rays, dr, st)) != CB2_OK) return rc;
    const size_t esz = out_f64 ? sizeof(double) : sizeof(float);
    const size_t obytes = (size_t)rays->n_rays * sc->host.bins * esz;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 5, obytes)) != CB2_OK) return rc;
    if (accumulate) CB2_CUDA(cudaMemcpyAsync(sc->stage[5], out, obytes, cudaMemcpyHostToDevice, st));
    CB2_CUDA(cudaMemsetAsync(sc->stats_dev, 0, sizeof(cb2_stats), st));
    static const bool overlap = !(getenv("CB2_D2H_OVERLAP") && atoi(getenv("CB2_D2H_OVERLAP")) == 0);
    if (sc->warp_kernel && overlap) {
        // the two-kernel path works in ray batches: each batch's rows go back to the host while the next batch computes
        if (!sc->copy_stream) CB2_CUDA(cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking));
        if (!sc->copy_stream2 && !(getenv("CB2_D2H_ONE_STREAM") && atoi(getenv("CB2_D2H_ONE_STREAM")))) CB2_CUDA(cudaStreamCreateWithFlags(&sc->copy_stream2, cudaStreamNonBlocking));
        if (!sc->copy_ev) CB2_CUDA(cudaEventCreateWithFlags(&sc->copy_ev, cudaEventDisableTiming));
        sc->d2h_host = out;
        sc->d2h_rows = dest_row;
        rc = cb2_launch_emission(sc, dr, sc->stage[5], out_f64, scale, accumulate, sc->stats_dev, st);
        sc->d2h_host = nullptr;
        sc->d2h_rows = nullptr;
        if (rc != CB2_OK) { cudaStreamSynchronize(sc->copy_stream); if (sc->copy_stream2) cudaStreamSynchronize(sc->copy_stream2); return rc; }
        CB2_CUDA(cudaStreamSynchronize(sc->copy_stream));
        if (sc->copy_stream2) CB2_CUDA(cudaStreamSynchronize(sc->copy_stream2));
    } else {
        if ((rc = cb2_launch_emission(sc, dr, sc->stage[5], out_f64, scale, accumulate, sc->stats_dev, st)) != CB2_OK) return rc;
        if (dest_row) {
            if ((rc = cb2_d2h_rows(out, dest_row, rays->n_rays, sc->stage[5], (size_t)sc->host.bins * esz, st)) != CB2_OK) return rc;
        } else if ((rc = cb2_d2h(out, sc->stage[5], obytes, st)) != CB2_OK) return rc;
    }
    if (stats) CB2_CUDA(cudaMemcpyAsync(stats, sc->stats_dev, sizeof(cb2_stats), cudaMemcpyDeviceToHost, st));
    CB2_CUDA(cudaStreamSynchronize(st));
    return CB2_OK;
}

extern "C" int cb2_emission_render(cb2_scene* sc, const cb2_rays* rays, void* out, int out_f64, double scale, int accumulate,
                                   cb2_stats* stats) {
    return emission_render_host(sc, rays, nullptr, out, out_f64, scale, accumulate, stats);
}

extern "C" int cb2_emission_render_rows(cb2_scene* sc, const cb2_rays* rays, const int64_t* dest_row, void* out, int out_f64, double scale,
                                        cb2_stats* stats) {
    if (!dest_row) return cb2_fail(CB2_ERR_VALUE, "null argument");
    return emission_render_host(sc, rays, dest_row, out, out_f64, scale, 0, stats);
}

extern "C" int64_t cb2_scene_info(const cb2_scene* sc, int key) {
    if (!sc) return -1;
    const DevBrems& b = sc->host.brems;
    switch (key) {
    case 0: return sc->nw;
    case 1: return sc->bpl;
    case 2: return !b.present ? 0 : (b.mode == 3 ? 3 : 1);
    case 3: return b.present && b.mode == 3 ? b.k_pad : 0;
    case 4: return b.present && b.mode == 3 ? b.n_nodes : 0;
    case 5: return b.present && b.mode == 3 ? b.n_z : 0;
    case 6: return sc->warp_kernel ? cb2_warp_batch_rays(sc) : 0;
    case 7: return sc->warp_kernel;
    case 8: return sc->contract_tc;
    case 9: return sc->memo.enabled ? (sc->memo.core_n > 0 ? sc->memo.core_n : -2) : 0;   // state tables: psi_n intervals (-2: edge table only)
    case 11: return (int64_t)(sc->prof_fixup_ms * 1e3);                                   // fix-up pass, microseconds of the last profile
    case 10: return (int64_t)(sc->memo_err * 1e9f);                                         // accepted table's mid-interval error, 1e-9 units
    }
    return -1;
}

extern "C" int cb2_scene_profile(cb2_scene* sc, int enable, double* ms_out, int64_t* launches_out) {
    if (!sc) return cb2_fail(CB2_ERR_VALUE, "null argument");
    CB2_CUDA(cudaSetDevice(sc->device));
    if (enable) {
        for (int i = 0; i < 4; i++) { sc->prof_ms[i] = 0.0; sc->prof_launches[i] = 0; }
        sc->prof_fixup_ms = 0.0;
        for (int i = 0; i < 10; i++)
            if (!sc->prof_ev[i]) CB2_CUDA(cudaEventCreate(&sc->prof_ev[i]));
        sc->prof_on = 1;
        return CB2_OK;
    }
    sc->prof_on = 0;
    for (int i = 0; i < 4; i++) {
        if (ms_out) ms_out[i] = sc->prof_ms[i];
        if (launches_out) launches_out[i] = sc->prof_launches[i];
    }
    return CB2_OK;
}

extern "C" int cb2_beam_sample(cb2_scene* sc, const double* beam_points, int64_t n, double* out) {
    if (!sc || !beam_points || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (!sc->host.beam.present) return cb2_fail(CB2_ERR_VALUE, "the scene has no beam");
    if (n <= 0) return CB2_OK;
    CB2_CUDA(cudaSetDevice(sc->device));
    int rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 6, (size_t)n * 3 * sizeof(double))) != CB2_OK) return rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 7, (size_t)n * 4 * sizeof(double))) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpy(sc->stage[6], beam_points, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    if ((rc = cb2_launch_beam_sample(sc, (const double*)sc->stage[6], n, (double*)sc->stage[7], 0)) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpy(out, sc->stage[7], (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return CB2_OK;
}

extern "C" int cb2_state_width(const cb2_scene* sc) { return sc ? 2 + 5 * sc->host.n_species + 3 : 0; }

extern "C" int cb2_sample_state(cb2_scene* sc, const double* points, int64_t n, double* out) {
    if (!sc || !points || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (n <= 0) return CB2_OK;
    CB2_CUDA(cudaSetDevice(sc->device));
    const int w = cb2_state_width(sc);
    int rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 6, (size_t)n * 3 * sizeof(double))) != CB2_OK) return rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 7, (size_t)n * w * sizeof(double))) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpy(sc->stage[6], points, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    if ((rc = cb2_launch_sample_state(sc, (const double*)sc->stage[6], n, (double*)sc->stage[7], 0, 0)) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpy(out, sc->stage[7], (size_t)n * w * sizeof(double), cudaMemcpyDeviceToHost));
    return CB2_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// ray transfer
// ------------------------------------------------------------------------------------------------------------------
extern "C" int cb2_rt_create(const cb2_rt_desc* d, int device, cb2_rt_scene** out) {
    if (!d || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    *out = nullptr;
    if (d->abi_version != CB2_ABI_VERSION) return cb2_fail(CB2_ERR_VALUE, "abi_version mismatch");
    if (d->kind != CB2_RT_CYLINDRICAL && d->kind != CB2_RT_CARTESIAN) return cb2_fail(CB2_ERR_TYPE, "unsupported ray-transfer grid kind %d", d->kind);
    if (d->integrator != 0 && d->integrator != 1) return cb2_fail(CB2_ERR_TYPE, "unsupported ray-transfer integrator %d", d->integrator);
    for (int k = 0; k < 3; k++) {
        if (d->grid_shape[k] < 1) return cb2_fail(CB2_ERR_VALUE, "Number of grid cells must be > 0.");
        if (!(d->grid_steps[k] > 0)) return cb2_fail(CB2_ERR_VALUE, "Grid steps must be > 0.");
    }
    if (!(d->step > 0)) return cb2_fail(CB2_ERR_VALUE, "Numerical integration step size can not be less than or equal to zero.");
    if (d->min_samples < 2) return cb2_fail(CB2_ERR_VALUE, "At least two samples are required to perform the numerical integration.");
    if (!d->voxel_map || d->bins < 0) return cb2_fail(CB2_ERR_VALUE, "voxel_map missing");
    int ndev = cb2_device_count();
    if (ndev <= 0) return ndev < 0 ? CB2_ERR_CUDA : cb2_fail(CB2_ERR_CUDA, "no CUDA device visible (libcherab_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return cb2_fail(CB2_ERR_VALUE, "device %d out of range (%d visible)", device, ndev);
    CB2_CUDA(cudaSetDevice(device));
    cb2_rt_scene* sc = (cb2_rt_scene*)calloc(1, sizeof(cb2_rt_scene));
    if (!sc) return cb2_fail(CB2_ERR_MEMORY, "out of host memory");
    sc->device = device;
    DevRT& r = sc->rt;
    r.kind = d->kind;
    r.n0 = d->grid_shape[0]; r.n1 = d->grid_shape[1]; r.n2 = d->grid_shape[2];
    r.min_samples = d->min_samples;
    r.bins = d->bins;
    r.s0 = d->grid_steps[0]; r.s1 = d->grid_steps[1]; r.s2 = d->grid_steps[2];
    r.rmin = d->rmin; r.period = d->period; r.step = d->step;
    r.trapezium = d->integrator == 1;
    for (int k = 0; k < 12; k++) r.w2l[k] = d->world_to_local[k];
    const size_t ncell = (size_t)r.n0 * r.n1 * r.n2;
    for (size_t k = 0; k < ncell; k++)
        if (d->voxel_map[k] >= d->bins) { free(sc); return cb2_fail(CB2_ERR_VALUE, "voxel_map entry exceeds bins"); }
    int rc = CB2_OK;
    do {
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&sc->voxel_map_dev, ncell * sizeof(int32_t)), "cudaMalloc(voxel_map)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMemcpy(sc->voxel_map_dev, d->voxel_map, ncell * sizeof(int32_t), cudaMemcpyHostToDevice), "cudaMemcpy(voxel_map)")) != CB2_OK) break;
        r.voxel_map = sc->voxel_map_dev;
        cudaDeviceProp prop;
        if ((rc = cb2_cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) != CB2_OK) break;
        sc->n_warps = prop.multiProcessorCount * 32;
        // distinct sources one ray can touch: bounded by the cell-boundary crossings of a straight line (each radial
        // boundary at most twice, z and phi are monotone along the ray) and by bins
        size_t cross = (size_t)2 * r.n0 + r.n2 + 4;
        if (r.kind == CB2_RT_CYLINDRICAL) cross += (size_t)r.n1 * (size_t)ceil(360.0 / r.period) + 2;
        else cross = (size_t)r.n0 + r.n1 + r.n2 + 4;
        sc->touch_cap = (int)std::min<size_t>(std::max<size_t>(r.bins, 1), cross + 4);
        if (sc->touch_cap > 32768) { rc = cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "ray-transfer grids with more than 32768 distinct sources per ray are not supported"); break; }
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&sc->stats_dev, sizeof(cb2_stats)), "cudaMalloc(stats)")) != CB2_OK) break;
    } while (0);
    if (rc != CB2_OK) {
        cb2_rt_destroy(sc);
        return rc;
    }
    *out = sc;
    return CB2_OK;
}

extern "C" int cb2_rt_destroy(cb2_rt_scene* sc) {
    if (!sc) return CB2_OK;
    cudaSetDevice(sc->device);
    if (sc->voxel_map_dev) cudaFree(sc->voxel_map_dev);
    if (sc->scratch) cudaFree(sc->scratch);
    if (sc->touched) cudaFree(sc->touched);
    if (sc->row_cols) cudaFree(sc->row_cols);
    if (sc->row_len) cudaFree(sc->row_len);
    if (sc->stats_dev) cudaFree(sc->stats_dev);
    free_stage(sc->stage, sc->stage_bytes);
    free(sc);
    return CB2_OK;
}

extern "C" int cb2_rt_render_dense(cb2_rt_scene* sc, const cb2_rays* rays, double* out, int accumulate, cb2_stats* stats) {
    if (!sc || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    int rc = check_rays(rays);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaSetDevice(sc->device));
    if (stats) memset(stats, 0, sizeof *stats);
    if (rays->n_rays == 0 || sc->rt.bins == 0) return CB2_OK;
    cudaStream_t st = 0;
    DevRays dr;
    if ((rc = upload_rays(sc->stage, sc->stage_bytes, rays, dr, st)) != CB2_OK) return rc;
    const size_t obytes = (size_t)rays->n_rays * sc->rt.bins * sizeof(double);
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 5, obytes)) != CB2_OK) return rc;
    if (accumulate) CB2_CUDA(cudaMemcpyAsync(sc->stage[5], out, obytes, cudaMemcpyHostToDevice, st));
    else CB2_CUDA(cudaMemsetAsync(sc->stage[5], 0, obytes, st));
    CB2_CUDA(cudaMemsetAsync(sc->stats_dev, 0, sizeof(cb2_stats), st));
    if ((rc = cb2_launch_rt(sc, dr, 0, (double*)sc->stage[5], accumulate, nullptr, nullptr, nullptr, sc->stats_dev, st)) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpyAsync(out, sc->stage[5], obytes, cudaMemcpyDeviceToHost, st));
    if (stats) CB2_CUDA(cudaMemcpyAsync(stats, sc->stats_dev, sizeof(cb2_stats), cudaMemcpyDeviceToHost, st));
    CB2_CUDA(cudaStreamSynchronize(st));
    return CB2_OK;
}

extern "C" int cb2_rt_render_csr_device(cb2_rt_scene* sc, const cb2_rays* rays, int64_t* row_offset, int32_t* columns,
                                        double* lengths, int64_t capacity, int64_t* nnz_host, cb2_stats* stats_dev, void* stream) {
    if (!sc || !row_offset || !nnz_host) return cb2_fail(CB2_ERR_VALUE, "null argument");
    int rc = check_rays(rays);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaSetDevice(sc->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DevRays dr = as_dev_rays(rays);
    *nnz_host = 0;
    if (rays->n_rays == 0) {
        CB2_CUDA(cudaMemsetAsync(row_offset, 0, sizeof(int64_t), st));
        return CB2_OK;
    }
    // Single traversal when the strided scratch rows fit (n_rays x touch_cap x 12 B; C4: 5 GB of the 180 GB): every ray's row is
    // built once in its scratch row, the counts are scanned and the rows move to their places — the grid is walked once, not twice.
    const bool one_pass = !(getenv("CB2_RT_TWO_PASS") && atoi(getenv("CB2_RT_TWO_PASS")) != 0);
    const size_t row_bytes = (size_t)rays->n_rays * sc->touch_cap * (sizeof(int32_t) + sizeof(double));
    bool have_rows = one_pass && row_bytes <= ((size_t)24 << 30);
    if (have_rows && sc->row_cap_rays < (size_t)rays->n_rays) {
        CB2_CUDA(cudaStreamSynchronize(st));
        if (sc->row_cols) cudaFree(sc->row_cols);
        if (sc->row_len) cudaFree(sc->row_len);
        sc->row_cols = nullptr; sc->row_len = nullptr; sc->row_cap_rays = 0;
        if (cudaMalloc((void**)&sc->row_cols, (size_t)rays->n_rays * sc->touch_cap * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc((void**)&sc->row_len, (size_t)rays->n_rays * sc->touch_cap * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();                                   // not enough memory for the scratch rows: two traversals
            if (sc->row_cols) cudaFree(sc->row_cols);
            sc->row_cols = nullptr; sc->row_len = nullptr;
            have_rows = false;
        } else sc->row_cap_rays = (size_t)rays->n_rays;
    }
    if (have_rows) {
        if ((rc = cb2_launch_rt(sc, dr, 3, nullptr, 0, row_offset, sc->row_cols, sc->row_len, (unsigned long long*)stats_dev, st)) != CB2_OK) return rc;
        if ((rc = cb2_launch_scan(row_offset, rays->n_rays + 1, st)) != CB2_OK) return rc;
        CB2_CUDA(cudaMemcpyAsync(nnz_host, row_offset + rays->n_rays, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CB2_CUDA(cudaStreamSynchronize(st));
        if (*nnz_host > capacity) return cb2_fail(CB2_ERR_OVERFLOW, "CSR capacity %lld too small, need %lld", (long long)capacity, (long long)*nnz_host);
        if (*nnz_host == 0) return CB2_OK;
        if (!columns || !lengths) return cb2_fail(CB2_ERR_VALUE, "null CSR arrays");
        return cb2_launch_rt_compact(rays->n_rays, sc->touch_cap, row_offset, sc->row_cols, sc->row_len, columns, lengths, st);
    }
    // pass 1: distinct sources per ray -> row_offset[r]; exclusive scan; pass 2: fill
    if ((rc = cb2_launch_rt(sc, dr, 1, nullptr, 0, row_offset, nullptr, nullptr, nullptr, st)) != CB2_OK) return rc;
    if ((rc = cb2_launch_scan(row_offset, rays->n_rays + 1, st)) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpyAsync(nnz_host, row_offset + rays->n_rays, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB2_CUDA(cudaStreamSynchronize(st));
    if (*nnz_host > capacity) return cb2_fail(CB2_ERR_OVERFLOW, "CSR capacity %lld too small, need %lld", (long long)capacity, (long long)*nnz_host);
    if (*nnz_host == 0) return CB2_OK;
    if (!columns || !lengths) return cb2_fail(CB2_ERR_VALUE, "null CSR arrays");
    return cb2_launch_rt(sc, dr, 2, nullptr, 0, row_offset, columns, lengths, (unsigned long long*)stats_dev, st);
}

extern "C" int cb2_rt_render_csr(cb2_rt_scene* sc, const cb2_rays* rays, int64_t* row_offset, int32_t* columns, double* lengths,
                                 int64_t capacity, cb2_stats* stats) {
    if (!sc || !row_offset) return cb2_fail(CB2_ERR_VALUE, "null argument");
    int rc = check_rays(rays);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaSetDevice(sc->device));
    if (stats) memset(stats, 0, sizeof *stats);
    if (rays->n_rays == 0) { row_offset[0] = 0; return CB2_OK; }
    cudaStream_t st = 0;
    DevRays dr;
    if ((rc = upload_rays(sc->stage, sc->stage_bytes, rays, dr, st)) != CB2_OK) return rc;
    const size_t n = (size_t)rays->n_rays;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 5, (n + 1) * sizeof(int64_t))) != CB2_OK) return rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 6, (size_t)std::max<int64_t>(capacity, 1) * sizeof(int32_t))) != CB2_OK) return rc;
    if ((rc = stage_reserve(sc->stage, sc->stage_bytes, 7, (size_t)std::max<int64_t>(capacity, 1) * sizeof(double))) != CB2_OK) return rc;
    CB2_CUDA(cudaMemsetAsync(sc->stats_dev, 0, sizeof(cb2_stats), st));
    cb2_rays drays;
    drays.n_rays = rays->n_rays; drays.n_segments = rays->n_segments;
    drays.origin = dr.origin; drays.direction = dr.direction; drays.seg_offset = dr.seg_offset; drays.seg_t0 = dr.seg_t0; drays.seg_t1 = dr.seg_t1;
    int64_t nnz = 0;
    rc = cb2_rt_render_csr_device(sc, &drays, (int64_t*)sc->stage[5], (int32_t*)sc->stage[6], (double*)sc->stage[7], capacity, &nnz,
                                  (cb2_stats*)sc->stats_dev, st);
    if (rc == CB2_ERR_OVERFLOW) { row_offset[n] = nnz; return rc; }
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpyAsync(row_offset, sc->stage[5], (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (nnz > 0) {
        if ((rc = cb2_d2h(columns, sc->stage[6], (size_t)nnz * sizeof(int32_t), st)) != CB2_OK) return rc;
        if ((rc = cb2_d2h(lengths, sc->stage[7], (size_t)nnz * sizeof(double), st)) != CB2_OK) return rc;
    }
    if (stats) CB2_CUDA(cudaMemcpyAsync(stats, sc->stats_dev, sizeof(cb2_stats), cudaMemcpyDeviceToHost, st));
    CB2_CUDA(cudaStreamSynchronize(st));
    return CB2_OK;
}
