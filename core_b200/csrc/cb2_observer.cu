// Observer front-end on the device: pinhole-camera rays and their chords through the plasma / beam /
// ray-transfer primitive, written straight into the device-resident cb2_rays layout the render calls take, so a frame's
// 10^6..10^7 rays never exist in host memory.
//
// Restates raysect's PinholeCamera ray generation (raysect 0.8.1 is not vendored: parity unpinned) and
// the analytic hit intervals of the primitives the BASELINE configs use — Subtract(Cylinder, Cylinder)
// (cherab/generomak/plasma/plasma.py:673-681, cherab/tools/raytransfer/raytransfer.py:198), Sphere, Box
// (cherab/tools/plasmas/slab.pyx:255, raytransfer.py:266) — with the arithmetic of the host mirror core_b200/geometry.py,
// in float64.
#include <cuda_runtime.h>

#include <cmath>

#include "cb2_internal.h"

namespace {

struct Interval { double t0, t1; };

__device__ __forceinline__ Interval cyl_interval(const double o[3], const double d[3], double radius) {
    const double a = d[0] * d[0] + d[1] * d[1];
    const double b = 2.0 * (o[0] * d[0] + o[1] * d[1]);
    const double c = o[0] * o[0] + o[1] * o[1] - radius * radius;
    Interval r = {INFINITY, -INFINITY};
    if (a < 1e-300) {                       // parallel to the axis: inside for every t or never
        if (c < 0) { r.t0 = -INFINITY; r.t1 = INFINITY; }
        return r;
    }
    const double disc = b * b - 4.0 * a * c;
    if (disc > 0) {
        const double sq = sqrt(disc);
        r.t0 = (-b - sq) / (2 * a);
        r.t1 = (-b + sq) / (2 * a);
    }
    return r;
}

__device__ __forceinline__ Interval slab_interval(double o, double d, double lo, double hi) {
    Interval r;
    if (d == 0) {                           // a ray lying in a face plane counts as inside
        const bool inside = o >= lo && o <= hi;
        r.t0 = inside ? -INFINITY : INFINITY;
        r.t1 = inside ? INFINITY : -INFINITY;
        return r;
    }
    const double ta = (lo - o) / d, tb = (hi - o) / d;
    r.t0 = fmin(ta, tb);
    r.t1 = fmax(ta, tb);
    return r;
}

// up to two chords [t0, t1] of the ray o + t d (primitive-local) through the primitive, ordered along the ray
__device__ __forceinline__ int primitive_intervals(const cb2_primitive& P, const double o[3], const double d[3], Interval out[2]) {
    int n = 0;
    auto keep = [&](double t0, double t1) {
        if (t1 > t0 && isfinite(t0) && isfinite(t1)) { out[n].t0 = t0; out[n].t1 = t1; n++; }
    };
    if (P.kind == CB2_PRIM_HOLLOW_CYLINDER) {
        const Interval a = cyl_interval(o, d, P.p[1]);
        const Interval s = slab_interval(o[2], d[2], P.p[2], P.p[3]);
        const double a0 = fmax(fmax(a.t0, s.t0), 0.0), a1 = fmin(a.t1, s.t1);
        if (P.p[0] <= 0) { keep(a0, a1); return n; }
        const Interval b = cyl_interval(o, d, P.p[0]);
        const bool hit = b.t1 > b.t0;
        keep(a0, hit ? fmin(a1, b.t0) : a1);
        keep(hit ? fmax(a0, b.t1) : INFINITY, a1);
    } else if (P.kind == CB2_PRIM_SPHERE) {
        const double b = 2.0 * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        const double c = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - P.p[0] * P.p[0];
        const double disc = b * b - 4.0 * c;
        if (disc > 0) {
            const double sq = sqrt(disc);
            keep(fmax((-b - sq) / 2, 0.0), (-b + sq) / 2);
        }
    } else {
        double t0 = 0.0, t1 = INFINITY;
        for (int ax = 0; ax < 3; ax++) {
            const Interval s = slab_interval(o[ax], d[ax], P.p[ax], P.p[3 + ax]);
            t0 = fmax(t0, s.t0);
            t1 = fmin(t1, s.t1);
        }
        keep(t0, t1);
    }
    return n;
}

// pass 0: origin, direction, chord count (into seg_offset) and the chords parked at [2i], [2i+1] of the scratch arrays
__global__ void pinhole_rays_kernel(cb2_pinhole cam, cb2_primitive prim, const int64_t* __restrict__ pixel_index, int64_t n,
                                    double sub_x, double sub_y, double* __restrict__ origin, double* __restrict__ direction,
                                    int64_t* __restrict__ seg_offset, double* __restrict__ park_t0, double* __restrict__ park_t1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { seg_offset[n] = 0; return; }
    const int64_t p = pixel_index ? pixel_index[i] : i;
    const int64_t ix = p / cam.ny, iy = p % cam.ny;
    // image plane at z = 1 in camera space, width 2 tan(fov / 2), pixels counted from the +x, +y corner
    const double delta = cam.width / cam.nx;
    const double x = 0.5 * cam.width - delta * ((double)ix + sub_x);
    const double y = 0.5 * delta * cam.ny - delta * ((double)iy + sub_y);
    const double inv = 1.0 / sqrt(x * x + y * y + 1.0);
    const double dc[3] = {x * inv, y * inv, inv};
    const double* m = cam.to_world;
    double d[3], o[3];
    for (int r = 0; r < 3; r++) {
        d[r] = m[4 * r] * dc[0] + m[4 * r + 1] * dc[1] + m[4 * r + 2] * dc[2];
        o[r] = m[4 * r + 3];
    }
    const double dn = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);   // the host mirror renormalises the world direction
    for (int r = 0; r < 3; r++) d[r] *= dn;
    const double* w = prim.world_to_local;
    double ol[3], dl[3];
    for (int r = 0; r < 3; r++) {
        ol[r] = w[4 * r] * o[0] + w[4 * r + 1] * o[1] + w[4 * r + 2] * o[2] + w[4 * r + 3];
        dl[r] = w[4 * r] * d[0] + w[4 * r + 1] * d[1] + w[4 * r + 2] * d[2];
    }
    Interval iv[2];
    const int cnt = primitive_intervals(prim, ol, dl, iv);
    for (int r = 0; r < 3; r++) { origin[3 * i + r] = o[r]; direction[3 * i + r] = d[r]; }
    seg_offset[i] = cnt;
    for (int k = 0; k < cnt; k++) { park_t0[2 * i + k] = iv[k].t0; park_t1[2 * i + k] = iv[k].t1; }
}

// pass 1 (after the exclusive scan of the counts): compact the parked chords
__global__ void compact_segments_kernel(int64_t n, const int64_t* __restrict__ seg_offset, const double* __restrict__ park_t0,
                                        const double* __restrict__ park_t1, double* __restrict__ seg_t0, double* __restrict__ seg_t1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = seg_offset[i], cnt = seg_offset[i + 1] - s;
    for (int64_t k = 0; k < cnt; k++) { seg_t0[s + k] = park_t0[2 * i + k]; seg_t1[s + k] = park_t1[2 * i + k]; }
}

}  // namespace

extern "C" int cb2_pinhole_rays_device(const cb2_pinhole* cam, const cb2_primitive* prim, const int64_t* pixel_index_dev, int64_t n,
                                       double sub_x, double sub_y, cb2_rays* out_dev, void* stream) {
    if (!cam || !prim || !out_dev) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (cam->nx < 1 || cam->ny < 1 || !(cam->width > 0)) return cb2_fail(CB2_ERR_VALUE, "pinhole camera needs pixels >= 1 and a positive field of view");
    if (prim->kind < CB2_PRIM_HOLLOW_CYLINDER || prim->kind > CB2_PRIM_BOX) return cb2_fail(CB2_ERR_TYPE, "unsupported primitive kind %d", prim->kind);
    if (n < 0 || (!pixel_index_dev && n != (int64_t)cam->nx * cam->ny)) return cb2_fail(CB2_ERR_VALUE, "n must equal nx * ny when no pixel list is given");
    if (!out_dev->origin || !out_dev->direction || !out_dev->seg_offset || !out_dev->seg_t0 || !out_dev->seg_t1)
        return cb2_fail(CB2_ERR_VALUE, "output ray arrays missing");
    cudaStream_t st = (cudaStream_t)stream;
    out_dev->n_rays = n;
    out_dev->n_segments = 0;
    if (n == 0) return CB2_OK;
    double* park = nullptr;
    CB2_CUDA(cudaMallocAsync((void**)&park, (size_t)4 * n * sizeof(double), st));
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + 1 + threads - 1) / threads);
    pinhole_rays_kernel<<<blocks, threads, 0, st>>>(*cam, *prim, pixel_index_dev, n, sub_x, sub_y, (double*)out_dev->origin,
                                                    (double*)out_dev->direction, (int64_t*)out_dev->seg_offset, park, park + 2 * n);
    int rc = cb2_launch_scan((int64_t*)out_dev->seg_offset, n + 1, st);
    if (rc == CB2_OK) {
        compact_segments_kernel<<<blocks, threads, 0, st>>>(n, out_dev->seg_offset, park, park + 2 * n, (double*)out_dev->seg_t0,
                                                            (double*)out_dev->seg_t1);
        rc = cb2_cuda_check(cudaGetLastError(), "pinhole ray kernels");
    }
    int64_t total = 0;
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpyAsync(&total, out_dev->seg_offset + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st), "segment count");
    cudaFreeAsync(park, st);
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaStreamSynchronize(st), "pinhole ray generation");
    out_dev->n_segments = total;
    return rc;
}
