// cb2_device.cuh — device helpers shared by the emission kernels (cb2_emission.cu, cb2_emission_warp.cu): table lookups,
// the per-sample context of the axisymmetric function tree, field evaluation, Bremsstrahlung moments.
#pragma once
#include <float.h>
#include <limits.h>
#include <math.h>

#include "cb2_internal.h"

#define FULL 0xffffffffu
#define L2E 1.4426950408889634f
#define RECIP_4_PI 0.07957747154594767f
#define INV_SQRT_PI 0.5641895835477563f
#define H_SERIES_MAX 0.36f
#define RYDBERG_EV 13.605693122994f
#define BOHR_MAGNETON 5.78838180123e-5f
#define HC_EV_NM_F 1239.8419738620933f

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1/2 erfc(a), a >= 0, relative error ~5e-7: s Q(s) exp(-a^2) with s = 1/(1 + a/2)
__device__ __forceinline__ float half_erfc(float a) {
    const float s = rcp_approx(fmaf(0.5f, a, 1.0f));
    float q = 0.01022576f;
    q = fmaf(q, s, -0.08354225f);
    q = fmaf(q, s, 0.26529264f);
    q = fmaf(q, s, -0.40173006f);
    q = fmaf(q, s, 0.26512444f);
    q = fmaf(q, s, -0.07779776f);
    q = fmaf(q, s, 0.12243541f);
    q = fmaf(q, s, 0.11730774f);
    q = fmaf(q, s, 0.1416635f);
    q = fmaf(q, s, 0.14102058f);
    return q * s * ex2_approx(a * a * -L2E);
}

// Packed float32 pairs (sm_100a: add / mul / fma .f32x2 = SASS FADD2 / FMUL2 / FFMA2, one issue slot for two lanes' worth of
// arithmetic).  bin_kernel is issue-bound, not pipe-bound (r1c ncu: issue-active 81 %, FMA pipe 47 %, XU 51 %): pairing two
// bins per instruction takes the series evaluation from ~10 to ~6 issue slots per bin and leaves MUFU.EX2 as the bound.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ float horner4(const float4 c, float t) { return fmaf(fmaf(fmaf(c.w, t, c.z), t, c.y), t, c.x); }

// ------------------------------------------------------------------------------------------------------------------
// table lookups
// ------------------------------------------------------------------------------------------------------------------
struct Cell2 {
    int i, j;
    float t, u;
    bool inside;
};

__device__ __forceinline__ int search_knots(const float* __restrict__ x, int n, float v) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ Cell2 locate2d(const DevTable2D& T, float x, float y) {
    Cell2 c;
    c.inside = (x >= T.xmin) && (x <= T.xmax) && (y >= T.ymin) && (y <= T.ymax);
    x = fminf(fmaxf(x, T.xmin), T.xmax);
    y = fminf(fmaxf(y, T.ymin), T.ymax);
    if (T.uniform) {
        const float fx = (x - T.x0) * T.inv_dx, fy = (y - T.y0) * T.inv_dy;
        c.i = min(max((int)fx, 0), T.nx - 2);
        c.j = min(max((int)fy, 0), T.ny - 2);
        c.t = fx - (float)c.i;
        c.u = fy - (float)c.j;
    } else {
        c.i = search_knots(T.x, T.nx, x);
        c.j = search_knots(T.y, T.ny, y);
        c.t = (x - __ldg(T.x + c.i)) * __ldg(T.inv_wx + c.i);
        c.u = (y - __ldg(T.y + c.j)) * __ldg(T.inv_wy + c.j);
    }
    return c;
}

__device__ __forceinline__ float eval2d(const DevTable2D& T, const Cell2& c) {
    const float4* q = T.coef + ((size_t)c.i * (T.ny - 1) + c.j) * 4;
    const float p0 = horner4(__ldg(q), c.u), p1 = horner4(__ldg(q + 1), c.u);
    const float p2 = horner4(__ldg(q + 2), c.u), p3 = horner4(__ldg(q + 3), c.u);
    return fmaf(fmaf(fmaf(p3, c.t, p2), c.t, p1), c.t, p0);
}

// tricubic table lookup, arguments clamped to the table ('nearest'); inside = all three arguments within the table
__device__ __forceinline__ float eval3d(const DevTable3D& T, float x, float y, float z, bool& inside) {
    inside = (x >= T.xmin) && (x <= T.xmax) && (y >= T.ymin) && (y <= T.ymax) && (z >= T.zmin) && (z <= T.zmax);
    x = fminf(fmaxf(x, T.xmin), T.xmax);
    y = fminf(fmaxf(y, T.ymin), T.ymax);
    z = fminf(fmaxf(z, T.zmin), T.zmax);
    const int i = search_knots(T.x, T.nx, x), j = search_knots(T.y, T.ny, y), k = search_knots(T.z, T.nz, z);
    const float t = (x - __ldg(T.x + i)) * __ldg(T.inv_wx + i);
    const float u = (y - __ldg(T.y + j)) * __ldg(T.inv_wy + j);
    const float w = (z - __ldg(T.z + k)) * __ldg(T.inv_wz + k);
    const float4* q = T.coef + (((size_t)i * (T.ny - 1) + j) * (T.nz - 1) + k) * 16;
    float r = 0.f;
#pragma unroll
    for (int p = 3; p >= 0; p--) {
        const float e0 = horner4(__ldg(q + 4 * p), w), e1 = horner4(__ldg(q + 4 * p + 1), w);
        const float e2 = horner4(__ldg(q + 4 * p + 2), w), e3 = horner4(__ldg(q + 4 * p + 3), w);
        r = fmaf(r, t, fmaf(fmaf(fmaf(e3, u, e2), u, e1), u, e0));
    }
    return r;
}

// coefficients e_p(u) of t^p for a fixed second coordinate (used to turn the Gaunt bicubic into a cubic in log10 u)
__device__ __forceinline__ float4 eval2d_rows(const DevTable2D& T, int i, int j, float u) {
    const float4* q = T.coef + ((size_t)i * (T.ny - 1) + j) * 4;
    return make_float4(horner4(__ldg(q), u), horner4(__ldg(q + 1), u), horner4(__ldg(q + 2), u), horner4(__ldg(q + 3), u));
}

__device__ __forceinline__ void locate1d(const DevTable1D& T, float x, int& i, float& t) {
    x = fminf(fmaxf(x, T.xmin), T.xmax);
    if (T.uniform) {
        const float f = (x - T.x0) * T.inv_dx;
        i = min(max((int)f, 0), T.n - 2);
        t = f - (float)i;
        return;
    }
    i = (T.n > 1) ? search_knots(T.x, T.n, x) : 0;
    t = (T.n > 1) ? (x - __ldg(T.x + i)) * __ldg(T.inv_w + i) : 0.f;
}

// Interpolator2DArray 'cubic' with 'linear' extrapolation (beam.pyx:75,84; restated like the oracle's interp2d_cubic_linear): the
// patch's value, gradient and cross derivative at the nearest boundary point.  Rare path (a sample outside a beam table).
static __device__ __noinline__ float eval2d_linear(const DevTable2D& T, float x, float y) {
    const Cell2 c = locate2d(T, x, y);
    const float xc = fminf(fmaxf(x, T.xmin), T.xmax), yc = fminf(fmaxf(y, T.ymin), T.ymax);
    const float iwx = T.uniform ? T.inv_dx : __ldg(T.inv_wx + c.i), iwy = T.uniform ? T.inv_dy : __ldg(T.inv_wy + c.j);
    const float4* q = T.coef + ((size_t)c.i * (T.ny - 1) + c.j) * 4;
    float p = 0.f, pt = 0.f, pu = 0.f, ptu = 0.f;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        const float4 r = __ldg(q + k);
        const float v = horner4(r, c.u), dv = fmaf(fmaf(3.0f * r.w, c.u, 2.0f * r.z), c.u, r.y);
        ptu = fmaf(ptu, c.t, pu);          // d/dt of the Horner recurrences, before they advance
        pt = fmaf(pt, c.t, p);
        pu = fmaf(pu, c.t, dv);
        p = fmaf(p, c.t, v);
    }
    const float dx = (x - xc) * iwx, dy = (y - yc) * iwy;
    return p + pt * dx + pu * dy + ptu * dx * dy;
}

// Interpolator1DArray 'cubic' with 'quadratic' extrapolation (beam.pyx:76, cx.pyx:96; oracle: interp1d_cubic_quadratic): outside the
// knots a parabola in the edge interval's normalised coordinate through the edge value with the spline's slopes at both knots
static __device__ __noinline__ float eval1d_quadratic(const DevTable1D& T, const float4* __restrict__ coef, float x) {
    if (T.n < 2) return __ldg(coef).x;
    if (x >= T.xmin && x <= T.xmax) {
        int i; float t;
        locate1d(T, x, i, t);
        return horner4(__ldg(coef + i), t);
    }
    const bool low = x < T.xmin;
    const int i = low ? 0 : T.n - 2;
    const float xi = T.uniform ? T.x0 + (float)i / T.inv_dx : __ldg(T.x + i), iw = T.uniform ? T.inv_dx : __ldg(T.inv_w + i);
    const float t = (x - xi) * iw;
    const float4 a = __ldg(coef + i);
    const float d0 = a.y, d1 = a.y + 2.0f * a.z + 3.0f * a.w;
    const float a2 = low ? a.x : (a.x + a.y + a.z + a.w) - 0.5f * (d0 + d1);
    return fmaf(fmaf(0.5f * (d1 - d0), t, d0), t, a2);
}

// ------------------------------------------------------------------------------------------------------------------
// per-sample shared context of the axisymmetric (Generomak-type) function tree — SURVEY Appendix C
// ------------------------------------------------------------------------------------------------------------------
struct AxCtx {
    float R, Z, cphi, sphi;
    float m;        // blend weight (plasma.py:610)
    int tri;        // edge-mesh triangle or -1
    int ci;         // core psi_n interval
    float ct;
    float psi;
    bool in_lcfs;
    float br, bt, bz;
    // branch-free Blend2D: value = we * edge[tri_c] + wc * core(ci, ct)
    float we, wc;
    int tri_c;
    bool b_outside;  // the B-field was taken from the clamped psi grid (counted as out of domain only if B is actually used)
};

// the even-odd crossing test in float64 with the oracle's expression (cherab/core/math/mask.pyx:53-67 as restated there): taken by
// the few samples whose float32 test comes within rounding of a vertex height or of an edge — the blend weight jumps at the LCFS
// polygon (psi_n is not exactly 1 on its chords), so a sample on the wrong side shows in the pixel (found by the full C1 frame)
static __device__ __noinline__ bool polygon_contains_exact(const DevAxisym& A, int gj, double px, double py) {
    int crossings = 0;
    const int k1 = __ldg(A.poly_row_start + gj + 1);
    for (int k = __ldg(A.poly_row_start + gj); k < k1; k++) {
        const double2 a = __ldg(A.poly_row_edges_d + 2 * k), b = __ldg(A.poly_row_edges_d + 2 * k + 1);   // (xi, yi), (xj, yj)
        if (((a.y > py) != (b.y > py)) && (px < __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(b.x, a.x), __dsub_rn(py, a.y)), __dsub_rn(b.y, a.y)), a.x)))
            crossings++;
    }
    return crossings & 1;
}

__device__ __forceinline__ bool polygon_contains(const DevAxisym& A, float px, float py, double pxd, double pyd) {
    if (px < A.poly_xmin || px > A.poly_xmax || py < A.poly_ymin || py > A.poly_ymax) return false;
    // coarse grid: only boundary cells need the edge loop (even-odd crossing test, mask.pyx:53-67)
    const int gi = min((int)((px - A.poly_xmin) * A.p_icx), A.pgx - 1), gj = min((int)((py - A.poly_ymin) * A.p_icy), A.pgy - 1);
    const int cls = __ldg(A.poly_cls + gi * A.pgy + gj);
    if (cls != 2) return cls == 1;
    int crossings = 0;
    bool unsure = false;
    const int k1 = __ldg(A.poly_row_start + gj + 1);
    for (int k = __ldg(A.poly_row_start + gj); k < k1; k++) {
        const float4 e = __ldg(A.poly_row_edges + k);  // (xi, yi, yj, slope)
        // float32 coordinates carry ~1.2e-7 m of rounding each: a decision inside 1e-6 m is left to the float64 test
        unsure = unsure || fabsf(e.y - py) < 1e-6f || fabsf(e.z - py) < 1e-6f;
        if ((e.y > py) != (e.z > py)) {
            const float xi = fmaf(py - e.y, e.w, e.x);
            unsure = unsure || fabsf(px - xi) < 1e-6f * (1.0f + fabsf(e.w));
            if (px < xi) crossings++;
        }
    }
    if (unsure) return polygon_contains_exact(A, gj, pxd, pyd);
    return crossings & 1;
}

// Discrete2DMesh lookup in float64 with the reference's operation order (no FMA contraction): the edge data is
// piecewise constant, so a sample that lands on the other side of a triangle edge changes a pixel visibly.
__device__ __forceinline__ int mesh_locate(const DevAxisym& A, double r, double z) {
    if (A.n_tri <= 0) return -1;
    const double fx = (r - A.mx0_d) * A.inv_cx_d, fy = (z - A.my0_d) * A.inv_cy_d;
    if (!(fx >= 0.0) || !(fy >= 0.0)) return -1;
    const int i = (int)fx, j = (int)fy;
    if (i >= A.gx || j >= A.gy) return -1;
    const int cell = i * A.gy + j;
    // the bucket lists are in ascending triangle order, so the first hit is the lowest-numbered triangle containing the
    // point (the tie rule for points on shared edges)
    const int k1 = __ldg(A.cell_start + cell + 1);
    const float rf = (float)r, zf = (float)z;
    for (int k = __ldg(A.cell_start + cell); k < k1; k++) {
        const int t = __ldg(A.cell_tris + k);
        // float32 sign filter: the three edge functions from rounded coordinates, each with a bound on what the roundings can
        // have changed (inputs: half an ulp of a coordinate < 4 m each, 2.4e-7; the arithmetic adds ~1e-7 relative).  A point
        // clearly outside (one function certainly negative, one certainly positive) is rejected, a point clearly inside
        // (all three certain, one sign) is inside this triangle only, so the tie rule cannot matter; everything else — a
        // band of ~1e-6 m around the edge lines — takes the exact float64 test.
        {
            const float4 p = __ldg(A.trif + 2 * t), q = __ldg(A.trif + 2 * t + 1);
            const float u1 = rf - p.z, v1 = p.y - p.w, u2 = p.x - p.z, v2 = zf - p.w;      // d1: (r - bx)(ay - by) - (ax - bx)(z - by)
            const float u3 = rf - q.x, v3 = p.w - q.y, u4 = p.z - q.x, v4 = zf - q.y;      // d2: (r - cx)(by - cy) - (bx - cx)(z - cy)
            const float u5 = rf - p.x, v5 = q.y - p.y, u6 = q.x - p.x, v6 = zf - p.y;      // d3: (r - ax)(cy - ay) - (cx - ax)(z - ay)
            const float f1 = u1 * v1 - u2 * v2, f2 = u3 * v3 - u4 * v4, f3 = u5 * v5 - u6 * v6;
            const float e1 = 1.5e-6f * (fabsf(u1) + fabsf(v1) + fabsf(u2) + fabsf(v2));
            const float e2 = 1.5e-6f * (fabsf(u3) + fabsf(v3) + fabsf(u4) + fabsf(v4));
            const float e3 = 1.5e-6f * (fabsf(u5) + fabsf(v5) + fabsf(u6) + fabsf(v6));
            const bool n1 = f1 < -e1, n2 = f2 < -e2, n3 = f3 < -e3, p1 = f1 > e1, p2 = f2 > e2, p3 = f3 > e3;
            if ((n1 || n2 || n3) && (p1 || p2 || p3)) continue;
            if ((n1 && n2 && n3) || (p1 && p2 && p3)) return t;
        }
        const double2 a = __ldg(A.tri + 3 * t), b = __ldg(A.tri + 3 * t + 1), c = __ldg(A.tri + 3 * t + 2);
        const double d1 = __dsub_rn(__dmul_rn(__dsub_rn(r, b.x), __dsub_rn(a.y, b.y)), __dmul_rn(__dsub_rn(a.x, b.x), __dsub_rn(z, b.y)));
        const double d2 = __dsub_rn(__dmul_rn(__dsub_rn(r, c.x), __dsub_rn(b.y, c.y)), __dmul_rn(__dsub_rn(b.x, c.x), __dsub_rn(z, c.y)));
        const double d3 = __dsub_rn(__dmul_rn(__dsub_rn(r, a.x), __dsub_rn(c.y, a.y)), __dmul_rn(__dsub_rn(c.x, a.x), __dsub_rn(z, a.y)));
        const bool neg = (d1 < 0) || (d2 < 0) || (d3 < 0), pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
        if (!(neg && pos)) return t;
    }
    return -1;
}

__device__ __forceinline__ void ax_setup(const DevScene& S, double xd, double yd, double zd, AxCtx& c, unsigned& ood) {
    const DevAxisym& A = S.ax;
    const float x = (float)xd, y = (float)yd, z = (float)zd;
    // R in float64 exactly as AxisymmetricMapper computes it (mappers.pyx:264): it feeds the exact triangle test
    const double r64 = A.present ? __dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd))) : 0.0;
    c.R = A.present ? (float)r64 : sqrtf(x * x + y * y);
    c.Z = z;
    const float inv_r = c.R > 0.f ? 1.0f / c.R : 0.f;
    c.cphi = c.R > 0.f ? x * inv_r : 1.f;
    c.sphi = y * inv_r;
    c.m = 0.f; c.tri = -1; c.ci = 0; c.ct = 0.f; c.psi = 0.f; c.in_lcfs = false; c.b_outside = false;
    c.br = c.bt = c.bz = 0.f;
    if (!A.present) return;
    const bool in_poly = polygon_contains(A, c.R, c.Z, r64, zd);
    Cell2 cell;
    bool have_cell = false;
    if (in_poly) {
        cell = locate2d(A.psin, c.R, c.Z);
        have_cell = true;
        if (!cell.inside) ood++;
        c.psi = fmaxf(eval2d(A.psin, cell), 0.f);   // ClampOutput2D(min=0), efit.pyx:116
        c.in_lcfs = c.psi <= 1.0f;                  // EFITLCFSMask, efit.pyx:405-410
    }
    if (c.in_lcfs) {
        // Interpolator1DArray(mask_x, mask_y, 'linear') of psi_n (plasma.py:610)
        float m = A.mask_y[0];
        const float p = fminf(fmaxf(c.psi, A.mask_x[0]), A.mask_x[A.n_mask - 1]);
        for (int k = 0; k + 1 < A.n_mask; k++)
            if (p >= A.mask_x[k] && p <= A.mask_x[k + 1]) {
                m = A.mask_y[k] + (p - A.mask_x[k]) / (A.mask_x[k + 1] - A.mask_x[k]) * (A.mask_y[k + 1] - A.mask_y[k]);
                break;
            }
        c.m = m;
    }
    if (c.m < 1.0f) c.tri = mesh_locate(A, r64, zd);
    if (c.m > 0.0f) locate1d(A.core, c.psi, c.ci, c.ct);
    c.tri_c = max(c.tri, 0);
    c.we = (c.m < 1.0f && c.tri >= 0) ? (c.m <= 0.f ? 1.0f : 1.0f - c.m) : 0.f;
    c.wc = c.m > 0.f ? (c.m >= 1.0f ? 1.0f : c.m) : 0.f;
    const bool want_pol = (S.need_pol && c.m > 0.f) || S.need_b;
    if (want_pol) {
        if (!have_cell) {
            cell = locate2d(A.psin, c.R, c.Z);
            c.b_outside = !cell.inside;
        }
        c.br = -eval2d(A.dpsi_dz, cell) * inv_r;   // MagneticField.evaluate, efit.pyx:443-445
        c.bz = eval2d(A.dpsi_dr, cell) * inv_r;
        if (S.need_b) {
            if (c.in_lcfs) {
                int fi; float ft;
                locate1d(A.fprof, c.psi, fi, ft);
                c.bt = horner4(__ldg(A.fprof_coef + fi), ft) * inv_r;
            } else {
                c.bt = A.b_vac * inv_r;
            }
        }
    }
}

// Blend2D(edge, map2d(core), mask) without branches: m <= 0 -> edge, m >= 1 -> core, else (1-m) edge + m core (plasma.py:622-636)
__device__ __forceinline__ float eval_blend(const DevScalar& f, const AxCtx& c) {
    const float e = f.edge ? __ldg(f.edge + c.tri_c) : 0.f;
    float cv = 0.f;
    if (f.core) cv = horner4(__ldg(f.core + c.ci), c.ct);
    return fmaf(c.wc, cv, c.we * e);
}

__device__ __forceinline__ float eval_scalar(const DevScalar& f, const AxCtx& c, float x, float y, float z) {
    switch (f.kind) {
    case CB2_FIELD_CONSTANT: return f.c[0];
    case CB2_FIELD_GAUSSIAN_VOLUME: {
        const float dx = x - f.c[3], dy = y - f.c[4], dz = z - f.c[5];
        return fmaf(f.c[1], exp2f((dx * dx + dy * dy + dz * dz) * f.c[2]), f.c[0]);
    }
    case CB2_FIELD_SLAB_ION: {
        const float xn = x * f.c[4];
        if (xn >= 0.f && xn <= 1.f) return (f.c[0] - f.c[1]) * powf(1.f - powf(1.f - xn, f.c[2]), f.c[3]) + f.c[1];
        return xn >= 1.f ? f.c[0] : 0.f;
    }
    case CB2_FIELD_SLAB_NEUTRAL: return x >= 0.f ? f.c[0] * exp2f(x * x * f.c[1]) : f.c[0];
    case CB2_FIELD_AXISYM_BLEND: {
        // Blend2D(edge, map2d(core), mask): m <= 0 -> edge, m >= 1 -> core, else (1-m) edge + m core
        float edge = 0.f, core = 0.f;
        if (c.m < 1.f && c.tri >= 0 && f.edge) edge = __ldg(f.edge + c.tri);
        if (c.m > 0.f && f.core) core = horner4(__ldg(f.core + c.ci), c.ct);
        if (c.m <= 0.f) return edge;
        if (c.m >= 1.f) return core;
        return fmaf(c.m, core, (1.f - c.m) * edge);
    }
    }
    return 0.f;
}

// AXONLY: every scalar field of the scene is an AXISYM_BLEND (decided on the host), so the kind switch disappears
template <int AXONLY>
__device__ __forceinline__ float eval_scalar_t(const DevScalar& f, const AxCtx& c, float x, float y, float z) {
    if (AXONLY) return eval_blend(f, c);
    return eval_scalar(f, c, x, y, z);
}

// cartesian velocity in plasma space
__device__ __forceinline__ float3 eval_vector(const DevVector& f, const AxCtx& c) {
    if (f.kind == CB2_FIELD_CONSTANT) return make_float3(f.c[0], f.c[1], f.c[2]);
    float vr = f.c[0], vp = f.c[1], vz = f.c[2];  // edge vector (R, phi, Z)
    if (c.m > 0.f) {
        // FluxCoordToCartesian.evaluate (efit.pyx:521-546) inside the LCFS
        float cr = 0.f, cp = 0.f, cz = 0.f;
        if (c.in_lcfs) {
            cp = f.vtor ? horner4(__ldg(f.vtor + c.ci), c.ct) : 0.f;
            if (!(c.br == 0.f && c.bz == 0.f)) {
                const float inv = rsqrtf(c.br * c.br + c.bz * c.bz);
                const float vpol = f.vpol ? horner4(__ldg(f.vpol + c.ci), c.ct) : 0.f;
                const float vnorm = f.vnorm ? horner4(__ldg(f.vnorm + c.ci), c.ct) : 0.f;
                cr = (c.br * vpol - c.bz * vnorm) * inv;
                cz = (c.bz * vpol + c.br * vnorm) * inv;
            }
        }
        if (c.m >= 1.f) { vr = cr; vp = cp; vz = cz; }
        else { vr = fmaf(c.m, cr, (1.f - c.m) * vr); vp = fmaf(c.m, cp, (1.f - c.m) * vp); vz = fmaf(c.m, cz, (1.f - c.m) * vz); }
    }
    // VectorAxisymmetricMapper: rotate by phi about z (mappers.pyx:302-312)
    return make_float3(vr * c.cphi - vp * c.sphi, vr * c.sphi + vp * c.cphi, vz);
}

__device__ __forceinline__ float3 eval_b_field(const DevScene& S, const AxCtx& c) {
    if (S.b_kind == 0) return make_float3(S.b_const[0], S.b_const[1], S.b_const[2]);
    return make_float3(c.br * c.cphi - c.bt * c.sphi, c.br * c.sphi + c.bt * c.cphi, c.bz);
}

// ImpactExcitationPEC.evaluate (pec.pyx:70-77) in log space: returns log10(PEC [W m^3]) + 38
__device__ __forceinline__ float eval_pec_log(const DevModel& M, float lne, float lte, unsigned& ood) {
    if (M.pec_const) return M.pec_value;
    const Cell2 c = locate2d(M.pec, lne, lte);
    if (!c.inside && !M.pec_extrapolate) ood++;
    return eval2d(M.pec, c);
}


#define SQRT_L2E 1.2011224087864498f
#define INV_L2E 0.6931471805599453f

struct SampleIn {
    float x, y, z;       // plasma-space position
    float dx, dy, dz;    // unit ray direction in plasma space
    float weight;        // trapezium weight (h or h/2), metres
    // beam scenes: donor density (1e19 m^-3) and donor velocity (m/s, plasma frame) at the sample
    float donor, bvx, bvy, bvz;
};

// Beam.density -> SingleRayAttenuator.density (beam/node.pyx:214-234, singleray.pyx:152-168), beam coordinates, 1e19 m^-3
__device__ __forceinline__ float beam_density(const DevBeam& B, float x, float y, float z) {
    if (z < 0.f || z > B.length) return 0.f;
    const float sx2 = fmaf(z * B.tanx, z * B.tanx, B.sigma2), sy2 = fmaf(z * B.tany, z * B.tany, B.sigma2);
    const float nr2 = x * x / sx2 + y * y / sy2;
    if (B.clamp_to_zero && nr2 > B.clamp2) return 0.f;
    const float g = __expf(-0.5f * nr2) * 0.15915494309189535f * rsqrtf(sx2 * sy2);
    // Interpolator1DArray(beam_z, beam_density, 'linear', 'nearest')
    const float f = fminf(fmaxf(z * B.inv_dz, 0.f), (float)(B.n_axis - 1));
    const int i = min((int)f, B.n_axis - 2);
    const float t = f - (float)i;
    const float a = __ldg(B.axis_density + i), b = __ldg(B.axis_density + i + 1);
    return fmaf(t, b - a, a) * g;
}

// Beam.direction (beam/node.pyx:236-279), beam coordinates
__device__ __forceinline__ float3 beam_direction(const DevBeam& B, float x, float y, float z) {
    if (z <= 0.f) return make_float3(0.f, 0.f, 1.f);
    const float ztx = z * z * B.tanx * B.tanx, zty = z * z * B.tany * B.tany;
    const float ex = x * ztx / (B.sigma2 + ztx), ey = y * zty / (B.sigma2 + zty);
    const float inv = rsqrtf(ex * ex + ey * ey + z * z);
    return make_float3(ex * inv, ey * inv, z * inv);
}

__device__ __forceinline__ double xform_row(const double* m, double x, double y, double z, bool point) {
    // ((m0 x + m1 y) + m2 z) (+ m3), separate roundings like the reference's compiled C
    double v = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), __dmul_rn(m[2], z));
    return point ? __dadd_rn(v, m[3]) : v;
}

// Phi(u) = int_0^u s(v) dv of the modified Lorentzian (odd in u), float64: cubic Hermite on the [0, 4] table, asymptotic
// tail series beyond (K sum_n (-A)^n u^-(2.5 n + 1.5)/(2.5 n + 1.5)).
__device__ __forceinline__ double lorentz_cdf(const double2* __restrict__ tab, double phi_inf, double u) {
    const double au = fabs(u);
    double v;
    if (au < 4.0) {
        const double f = au * 512.0;
        const int i = min((int)f, 2047);
        const double t = f - (double)i, h = 1.0 / 512.0;
        const double2 p0 = __ldg(tab + i), p1 = __ldg(tab + i + 1);
        const double d0 = p0.y * h, d1 = p1.y * h;
        const double a2 = 3.0 * (p1.x - p0.x) - 2.0 * d0 - d1, a3 = 2.0 * (p0.x - p1.x) + d0 + d1;
        v = p0.x + t * (d0 + t * (a2 + t * a3));
    } else {
        const double K = 0.13385686538368502, A = 0.1767766952966369;     // 0.5^1.5 / C, 0.5^2.5
        const double sq = sqrt(au), r = 1.0 / (au * au * sq);              // u^-2.5
        double term = 1.0 / (au * sq), sum = 0.0;                          // u^-1.5
#pragma unroll
        for (int n = 0; n < 8; n++) { sum += term * (1.0 / (2.5 * n + 1.5)); term *= -A * r; }
        v = phi_inf - K * sum;
    }
    return u < 0.0 ? -v : v;
}

// Bremsstrahlung, moment formulation (DevBrems mode 3): instead of evaluating the continuum at every (sample, bin) the
// sample adds W_s N_{z,s} L_m(Te_s) to the ray's moments on the temperature-node grid (4 nodes per distinct charge z);
// the spectrum follows from one dense contraction mom . phi after the ray is finished (cb2_contract.cu).
// Lanes = samples.  Consecutive samples mostly share the node interval, so the warp loops over the distinct intervals it
// holds and transpose-reduces the 4 n_z (<= 32) values of each; lane 4 z + k then owns node (i - 1 + k) of charge z.
// state part (per unit trapezium weight): node coordinate f (0: the sample does not emit) and U_z = pref ne/sqrt(Te) e^{-x_ref/Te} N_z
template <int AXONLY>
__device__ __forceinline__ void brems_state(const DevScene& S, const SampleIn& in, const AxCtx& ctx, float ne, float te, float& f,
                                            float (&U)[CB2_MAX_BREMS_Z], unsigned& ood) {
    const DevBrems& B = S.brems;
    f = 0.f;
#pragma unroll
    for (int z = 0; z < CB2_MAX_BREMS_Z; z++) U[z] = 0.f;
    if (!(ne > 0.f && te > 0.f)) return;
    float tc = te;
    if (tc < B.te_lo || tc > B.te_hi) { ood++; tc = fminf(fmaxf(tc, B.te_lo), B.te_hi); }
    const float tau = 1.0f / tc;
    const float sv = fmaf(tau, B.inv_tau_c, -logf(tc));                 // ln(tau) + tau/tau_c
    f = fminf(fmaxf((sv - B.s0) * B.inv_ds, 1.0f), (float)(B.n_nodes - 3) + 0.9999f);
    const float W = B.pref * ne * rsqrtf(te) * __expf(-B.x_ref * tau);
    // per distinct charge: N_z = sum of the (positive) densities of the species with that charge; the species are grouped by
    // charge on the host, so the sums need no per-species select chain
#pragma unroll
    for (int z = 0; z < CB2_MAX_BREMS_Z; z++) {
        float a = 0.f;
        for (int s = B.zstart[z]; s < B.zstart[z + 1]; s++) {
            const float ni = eval_scalar_t<AXONLY>(S.species[B.zlist[s]].density, ctx, in.x, in.y, in.z);
            a += fmaxf(ni, 0.f);
        }
        U[z] = W * a;
    }
}

// scatter part: the warp's samples (weight w, node coordinate f >= 1, U_z) go to the ray's moments.  Lanes = samples.
// Consecutive samples mostly share the node interval, so the warp loops over the distinct intervals it holds and
// transpose-reduces the 4 n_z (<= 32) values of each; lane 4 z + k then owns node (i - 1 + k) of charge z.
template <typename MomT>
__device__ __forceinline__ void brems_scatter(const DevBrems& B, bool live, float w, float f, const float (&U)[CB2_MAX_BREMS_Z],
                                              MomT* __restrict__ mom, int lane) {
    // Element e = 4 z + k of a lane is w U_z L_k.  The transpose-reduction runs in the lane-permuted order of bin_kernel's windows —
    // register r of lane l holds element r ^ l, so `part[i] += shfl_xor(part[i + o], o)` needs no selects and leaves element l on
    // lane l — and since r ^ l = 4 ((r >> 2) ^ (l >> 2)) + ((r & 3) ^ (l & 3)), the permuted elements are products of the lane's own
    // U and L arrays permuted once by XOR swaps (32 selects per group instead of 94 per distinct node).
    int node = -1;
    float L[4] = {0.f, 0.f, 0.f, 0.f}, V[CB2_MAX_BREMS_Z];
#pragma unroll
    for (int z = 0; z < CB2_MAX_BREMS_Z; z++) V[z] = 0.f;
    if (live) {
        node = (int)f;
        const float t = f - (float)node;
        // cubic Lagrange weights on the nodes -1, 0, 1, 2
        const float tm1 = t - 1.0f, tm2 = t - 2.0f, tp1 = t + 1.0f;
        L[0] = -t * tm1 * tm2 * (1.0f / 6.0f); L[1] = tp1 * tm1 * tm2 * 0.5f; L[2] = -tp1 * t * tm2 * 0.5f; L[3] = tp1 * t * tm1 * (1.0f / 6.0f);
#pragma unroll
        for (int z = 0; z < CB2_MAX_BREMS_Z; z++) V[z] = w * U[z];
    }
    static_assert(CB2_MAX_BREMS_Z == 8, "the XOR permutation below is written for 8 charges x 4 nodes = 32 elements");
#pragma unroll
    for (int bit = 0; bit < 2; bit++) {                // L'[k] = L[k ^ (lane & 3)]
        const bool sw = (lane >> bit) & 1;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (!(k & (1 << bit))) {
                const float x = L[k], y = L[k | (1 << bit)];
                L[k] = sw ? y : x; L[k | (1 << bit)] = sw ? x : y;
            }
    }
#pragma unroll
    for (int bit = 0; bit < 3; bit++) {                // V'[z] = V[z ^ (lane >> 2)]
        const bool sw = (lane >> (bit + 2)) & 1;
#pragma unroll
        for (int z = 0; z < 8; z++)
            if (!(z & (1 << bit))) {
                const float x = V[z], y = V[z | (1 << bit)];
                V[z] = sw ? y : x; V[z | (1 << bit)] = sw ? x : y;
            }
    }
    unsigned todo = __ballot_sync(FULL, live);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int nd = __shfl_sync(FULL, node, leader);
        const bool mine = node == nd;
        todo &= ~__ballot_sync(FULL, mine);
        float Lm[4];
#pragma unroll
        for (int k = 0; k < 4; k++) Lm[k] = mine ? L[k] : 0.f;
        // registers (2 j, 2 j + 1) as one packed pair: 16 FMUL2 for the products, FADD2 in the butterfly
        const f32x2 L01 = pack2(Lm[0], Lm[1]), L23 = pack2(Lm[2], Lm[3]);
        f32x2 part[16];
#pragma unroll
        for (int j = 0; j < 16; j++) part[j] = mul2(pack2(V[j >> 1], V[j >> 1]), (j & 1) ? L23 : L01);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < o; j++) {
                float a, b;
                unpack2(part[j + o], a, b);
                part[j] = add2(part[j], pack2(__shfl_xor_sync(FULL, a, 2 * o), __shfl_xor_sync(FULL, b, 2 * o)));
            }
        float p0, p1;
        unpack2(part[0], p0, p1);
        p0 += __shfl_xor_sync(FULL, p1, 1);
        const int z = lane >> 2, k = lane & 3;
        if (z < B.n_z && p0 != 0.f) atomicAdd(&mom[z * B.n_nodes + nd - 1 + k], (MomT)p0);
    }
}

template <typename CntT, int AXONLY = 0, typename MomT = double>
__device__ __forceinline__ void sample_brems_moments(const DevScene& S, const SampleIn& in, const AxCtx& ctx, float ne, float te,
                                                     MomT* __restrict__ mom, int lane, CntT& brems_evals, unsigned& ood) {
    const bool live = ne > 0.f && te > 0.f && in.weight > 0.f;
    float f = 0.f, U[CB2_MAX_BREMS_Z];
#pragma unroll
    for (int z = 0; z < CB2_MAX_BREMS_Z; z++) U[z] = 0.f;
    if (live) {
        brems_state<AXONLY>(S, in, ctx, ne, te, f, U, ood);
        brems_evals += (CntT)S.bins;
    }
    brems_scatter(S.brems, live, in.weight, f, U, mom, lane);
}
