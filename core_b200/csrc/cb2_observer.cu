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
#include <vector>

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

// world-space ray i: origin, direction, chord count (into seg_offset) and the chords parked at [2i], [2i+1] of the scratch arrays
__device__ __forceinline__ void park_ray(const cb2_primitive& prim, int64_t i, const double o[3], const double d[3], double* __restrict__ origin,
                                         double* __restrict__ direction, int64_t* __restrict__ seg_offset, double* __restrict__ park_t0,
                                         double* __restrict__ park_t1) {
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

// pass 0 of the pinhole camera
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
    park_ray(prim, i, o, d, origin, direction, seg_offset, park_t0, park_t1);
}

// 0-D observers: ray i belongs to the observer whose [ray_offset[k], ray_offset[k + 1]) holds it; sample j of n — the formulas of
// core_b200/observers.py::FibreOptic.rays, float64
__global__ void observer0d_rays_kernel(const cb2_observer0d* __restrict__ obs, const int64_t* __restrict__ ray_offset, int64_t n_obs,
                                       cb2_primitive prim, int64_t n, double* __restrict__ origin, double* __restrict__ direction,
                                       int64_t* __restrict__ seg_offset, double* __restrict__ park_t0, double* __restrict__ park_t1,
                                       double* __restrict__ weight) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { seg_offset[n] = 0; return; }
    int64_t lo = 0, hi = n_obs;                       // last k with ray_offset[k] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ray_offset[mid] <= i) lo = mid; else hi = mid;
    }
    const cb2_observer0d& ob = obs[lo];
    double ol[3] = {0.0, 0.0, 0.0}, dl[3] = {0.0, 0.0, 1.0}, cos_t = 1.0;
    if (ob.radius > 0.0) {
        const double golden = 3.14159265358979323846 * (3.0 - sqrt(5.0));
        const double nn = (double)ob.samples, k = (double)(i - ray_offset[lo]) + 0.5;
        const double cos_max = cos(ob.acceptance_angle * (3.14159265358979323846 / 180.0));
        cos_t = 1.0 - (1.0 - cos_max) * k / nn;                                   // uniform in solid angle on the cap
        const double sin_t = sqrt(fmax(0.0, 1.0 - cos_t * cos_t));
        const double phi = golden * k;
        dl[0] = sin_t * cos(phi); dl[1] = sin_t * sin(phi); dl[2] = cos_t;
        const double rr = ob.radius * sqrt(k / nn);                               // uniform in area on the disc
        const double psi = golden * k * 7.0 + 1.0;
        ol[0] = rr * cos(psi); ol[1] = rr * sin(psi);
    }
    const double* m = ob.to_world;
    double o[3], d[3];
    for (int r = 0; r < 3; r++) {
        o[r] = m[4 * r] * ol[0] + m[4 * r + 1] * ol[1] + m[4 * r + 2] * ol[2] + m[4 * r + 3];
        d[r] = m[4 * r] * dl[0] + m[4 * r + 1] * dl[1] + m[4 * r + 2] * dl[2];
    }
    weight[i] = cos_t;
    park_ray(prim, i, o, d, origin, direction, seg_offset, park_t0, park_t1);
}

// one CTA per (observer, 256-bin chunk): thread = bin, the observer's rays in order (fixed summation order, coalesced rows)
template <typename T>
__global__ void observer0d_reduce_kernel(const T* __restrict__ spectra, const double* __restrict__ weight, const int64_t* __restrict__ ray_offset,
                                         const double* __restrict__ etendue, int bins, double* __restrict__ radiance, double* __restrict__ power) {
    const int64_t ob = blockIdx.x;
    const int bin = blockIdx.y * blockDim.x + threadIdx.x;
    if (bin >= bins) return;
    const int64_t r0 = ray_offset[ob], r1 = ray_offset[ob + 1];
    double num = 0.0, den = 0.0;
    for (int64_t r = r0; r < r1; r++) {
        const double w = weight[r];
        num += w * (double)spectra[(size_t)r * bins + bin];
        den += w;
    }
    if (radiance) radiance[(size_t)ob * bins + bin] = r1 > r0 ? num / den : 0.0;
    if (power) power[(size_t)ob * bins + bin] = r1 > r0 ? num / (double)(r1 - r0) * etendue[ob] : 0.0;
}

// pass 1 (after the exclusive scan of the counts): compact the parked chords
__global__ void compact_segments_kernel(int64_t n, const int64_t* __restrict__ seg_offset, const double* __restrict__ park_t0,
                                        const double* __restrict__ park_t1, double* __restrict__ seg_t0, double* __restrict__ seg_t1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = seg_offset[i], cnt = seg_offset[i + 1] - s;
    for (int64_t k = 0; k < cnt; k++) { seg_t0[s + k] = park_t0[2 * i + k]; seg_t1[s + k] = park_t1[2 * i + k]; }
}

// exclusive scan of the chord counts, compaction of the parked chords, segment total (one stream synchronisation); frees `park`
int finish_rays(cb2_rays* out_dev, int64_t n, double* park, cudaStream_t st) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + 1 + threads - 1) / threads);
    int rc = cb2_cuda_check(cudaGetLastError(), "ray generation kernel launch");
    if (rc == CB2_OK) rc = cb2_launch_scan((int64_t*)out_dev->seg_offset, n + 1, st);
    if (rc == CB2_OK) {
        compact_segments_kernel<<<blocks, threads, 0, st>>>(n, out_dev->seg_offset, park, park + 2 * n, (double*)out_dev->seg_t0,
                                                            (double*)out_dev->seg_t1);
        rc = cb2_cuda_check(cudaGetLastError(), "compact_segments_kernel launch");
    }
    int64_t total = 0;
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpyAsync(&total, out_dev->seg_offset + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st), "segment count");
    cudaFreeAsync(park, st);
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaStreamSynchronize(st), "ray generation");
    out_dev->n_segments = total;
    return rc;
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
    return finish_rays(out_dev, n, park, st);
}

extern "C" int cb2_observer0d_rays_device(const cb2_observer0d* observers, int64_t n_observers, const cb2_primitive* prim, cb2_rays* out_dev,
                                          double* weight_dev, void* stream) {
    if (!observers || !prim || !out_dev || !weight_dev) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (n_observers < 1) return cb2_fail(CB2_ERR_VALUE, "The group has no observers.");
    if (prim->kind < CB2_PRIM_HOLLOW_CYLINDER || prim->kind > CB2_PRIM_BOX) return cb2_fail(CB2_ERR_TYPE, "unsupported primitive kind %d", prim->kind);
    if (!out_dev->origin || !out_dev->direction || !out_dev->seg_offset || !out_dev->seg_t0 || !out_dev->seg_t1)
        return cb2_fail(CB2_ERR_VALUE, "output ray arrays missing");
    std::vector<int64_t> offs((size_t)n_observers + 1, 0);
    for (int64_t k = 0; k < n_observers; k++) {
        const cb2_observer0d& ob = observers[k];
        if (ob.samples < 1) return cb2_fail(CB2_ERR_VALUE, "The fibre radius and the number of pixel samples must be positive.");
        if (ob.radius < 0 || (ob.radius > 0 && !(ob.acceptance_angle > 0 && ob.acceptance_angle <= 90)))
            return cb2_fail(CB2_ERR_VALUE, "Acceptance angle must be in the range (0, 90] degrees.");
        if (ob.radius == 0 && ob.samples != 1) return cb2_fail(CB2_ERR_VALUE, "a sight line has one ray");
        offs[k + 1] = offs[k] + ob.samples;
    }
    const int64_t n = offs[n_observers];
    cudaStream_t st = (cudaStream_t)stream;
    out_dev->n_rays = n;
    out_dev->n_segments = 0;
    char* scratch = nullptr;
    const size_t park_bytes = (size_t)4 * n * sizeof(double), obs_bytes = (size_t)n_observers * sizeof(cb2_observer0d),
                 off_bytes = ((size_t)n_observers + 1) * sizeof(int64_t);
    CB2_CUDA(cudaMallocAsync((void**)&scratch, park_bytes + obs_bytes + off_bytes, st));
    double* park = (double*)scratch;
    cb2_observer0d* obs_dev = (cb2_observer0d*)(scratch + park_bytes);
    int64_t* off_dev = (int64_t*)(scratch + park_bytes + obs_bytes);
    int rc = cb2_cuda_check(cudaMemcpyAsync(obs_dev, observers, obs_bytes, cudaMemcpyHostToDevice, st), "observer upload");
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpyAsync(off_dev, offs.data(), off_bytes, cudaMemcpyHostToDevice, st), "observer upload");
    if (rc != CB2_OK) { cudaFreeAsync(scratch, st); return rc; }
    const int threads = 128;
    observer0d_rays_kernel<<<(unsigned)((n + 1 + threads - 1) / threads), threads, 0, st>>>(
        obs_dev, off_dev, n_observers, *prim, n, (double*)out_dev->origin, (double*)out_dev->direction, (int64_t*)out_dev->seg_offset, park,
        park + 2 * n, weight_dev);
    return finish_rays(out_dev, n, park, st);                      // (synchronises the stream: the host staging above is consumed)
}

extern "C" int cb2_observer0d_reduce_device(const void* spectra_dev, int spectra_f64, const double* weight_dev, const int64_t* ray_offset,
                                            const double* etendue, int64_t n_observers, int32_t bins, double* radiance_dev, double* power_dev,
                                            void* stream) {
    if (!spectra_dev || !weight_dev || !ray_offset || (!radiance_dev && !power_dev) || (power_dev && !etendue))
        return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (n_observers < 1 || bins < 1) return cb2_fail(CB2_ERR_VALUE, "The group has no observers.");
    cudaStream_t st = (cudaStream_t)stream;
    char* scratch = nullptr;
    const size_t off_bytes = ((size_t)n_observers + 1) * sizeof(int64_t), et_bytes = (size_t)n_observers * sizeof(double);
    CB2_CUDA(cudaMallocAsync((void**)&scratch, off_bytes + et_bytes, st));
    int rc = cb2_cuda_check(cudaMemcpyAsync(scratch, ray_offset, off_bytes, cudaMemcpyHostToDevice, st), "ray offsets upload");
    if (rc == CB2_OK && etendue) rc = cb2_cuda_check(cudaMemcpyAsync(scratch + off_bytes, etendue, et_bytes, cudaMemcpyHostToDevice, st), "etendue upload");
    if (rc == CB2_OK) {
        const dim3 grid((unsigned)n_observers, (unsigned)((bins + 255) / 256));
        if (spectra_f64)
            observer0d_reduce_kernel<double><<<grid, 256, 0, st>>>((const double*)spectra_dev, weight_dev, (const int64_t*)scratch,
                                                                   (const double*)(scratch + off_bytes), bins, radiance_dev, power_dev);
        else
            observer0d_reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)spectra_dev, weight_dev, (const int64_t*)scratch,
                                                                  (const double*)(scratch + off_bytes), bins, radiance_dev, power_dev);
        rc = cb2_cuda_check(cudaGetLastError(), "observer0d_reduce_kernel launch");
    }
    cudaFreeAsync(scratch, st);
    // the host arrays were staged from pageable memory: the copies above have completed on return (CUDA stages them synchronously)
    return rc;
}
