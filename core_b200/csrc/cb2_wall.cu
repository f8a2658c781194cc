// cb2_wall.cu — first-wall occlusion: first opaque hit of every ray against a triangle soup.
//
// Replaces what the wall meshes of cherab/generomak/machine/first_wall.py:120-184 do inside Raysect's tracer: the ray ends at its
// first opaque hit, so the plasma's volume integral stops there.  Layout in HBM:
//   * nodes  [n_nodes] 32 B: float3 lower corner, int a, float3 upper corner, int b — internal node: children a and b (b > 0);
//     leaf: triangles [a, a - b) (b < 0).  Boxes are float32, grown outward by their rounding; built on the host by median splits
//     of the triangle centroids along the widest axis (leaves of <= 4 triangles);
//   * tri    [n_tri][9] float64 vertex coordinates in BVH leaf order (the exact test decides; the boxes only prune).
// One thread per ray walks the tree with a private stack, nearer child first, and keeps the smallest t of the float64
// Moeller-Trumbore test (the oracle's brute-force loop uses the same expression).  wall_clip_kernel then cuts the ray's segments.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "cb2_internal.h"

struct WallNode {
    float lo[3];
    int a;
    float hi[3];
    int b;
};
static_assert(sizeof(WallNode) == 32, "two float4 per node");

struct cb2_wall {
    int device;
    int64_t n_tri;
    int n_nodes;
    WallNode* nodes;
    double* tri;
    void* stage[3];
    size_t stage_bytes[3];
};

namespace {

// ------------------------------------------------------------------------------------------------------------------
// host: BVH construction
// ------------------------------------------------------------------------------------------------------------------
struct Builder {
    const double* v;                       // [n][9]
    std::vector<float> clo, chi;           // per-triangle bounds [n][3] (float32, outward)
    std::vector<float> cen;                // centroids [n][3]
    std::vector<int> order;
    std::vector<WallNode> nodes;

    static float down(double x) { float f = (float)x; return (double)f > x ? nextafterf(f, -INFINITY) : f; }
    static float up(double x) { float f = (float)x; return (double)f < x ? nextafterf(f, INFINITY) : f; }

    int build(int first, int count) {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, clo3[3] = {FLT_MAX, FLT_MAX, FLT_MAX},
              chi3[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = first; k < first + count; k++) {
            const int t = order[k];
            for (int a = 0; a < 3; a++) {
                lo[a] = std::min(lo[a], clo[3 * (size_t)t + a]); hi[a] = std::max(hi[a], chi[3 * (size_t)t + a]);
                clo3[a] = std::min(clo3[a], cen[3 * (size_t)t + a]); chi3[a] = std::max(chi3[a], cen[3 * (size_t)t + a]);
            }
        }
        for (int a = 0; a < 3; a++) {
            // a margin for the float32 slab arithmetic of the traversal (origin and reciprocal direction are rounded too)
            const float m = 1e-5f + 4e-6f * std::max(fabsf(lo[a]), fabsf(hi[a]));
            nodes[id].lo[a] = lo[a] - m;
            nodes[id].hi[a] = hi[a] + m;
        }
        int axis = 0;
        for (int a = 1; a < 3; a++)
            if (chi3[a] - clo3[a] > chi3[axis] - clo3[axis]) axis = a;
        if (count <= 4 || !(chi3[axis] > clo3[axis])) {
            nodes[id].a = first;
            nodes[id].b = -count;
            return id;
        }
        const int mid = first + count / 2;
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                         [&](int p, int q) { return cen[3 * (size_t)p + axis] < cen[3 * (size_t)q + axis]; });
        const int l = build(first, mid - first);
        const int r = build(mid, first + count - mid);
        nodes[id].a = l;
        nodes[id].b = r;
        return id;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// device: traversal
// ------------------------------------------------------------------------------------------------------------------
// Moeller-Trumbore, float64: t of the hit of o + t d with triangle (v0, v1, v2), or +inf.  t must exceed t_min.
__host__ __device__ inline double tri_hit(const double* __restrict__ v, double ox, double oy, double oz, double dx, double dy, double dz,
                                         double t_min) {
    const double e1x = v[3] - v[0], e1y = v[4] - v[1], e1z = v[5] - v[2];
    const double e2x = v[6] - v[0], e2y = v[7] - v[1], e2z = v[8] - v[2];
    const double px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
    const double det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.0) return INFINITY;
    const double inv = 1.0 / det;
    const double sx = ox - v[0], sy = oy - v[1], sz = oz - v[2];
    const double u = (sx * px + sy * py + sz * pz) * inv;
    if (u < 0.0 || u > 1.0) return INFINITY;
    const double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
    const double w = (dx * qx + dy * qy + dz * qz) * inv;
    if (w < 0.0 || u + w > 1.0) return INFINITY;
    const double t = (e2x * qx + e2y * qy + e2z * qz) * inv;
    return t > t_min ? t : INFINITY;
}

__device__ __forceinline__ bool box_hit(const WallNode& n, float ox, float oy, float oz, float ix, float iy, float iz, float t_best, float& t_near) {
    const float ax = (n.lo[0] - ox) * ix, bx = (n.hi[0] - ox) * ix;
    const float ay = (n.lo[1] - oy) * iy, by = (n.hi[1] - oy) * iy;
    const float az = (n.lo[2] - oz) * iz, bz = (n.hi[2] - oz) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    t_near = tn;
    return tn <= tf * 1.00001f + 1e-6f && tn <= t_best;
}

__device__ double first_hit(const WallNode* __restrict__ nodes, const double* __restrict__ tri, double ox, double oy, double oz, double dx,
                            double dy, double dz) {
    const float fox = (float)ox, foy = (float)oy, foz = (float)oz;
    auto rcp = [](double d) { return (float)(1.0 / (fabs(d) > 1e-30 ? d : (d < 0 ? -1e-30 : 1e-30))); };
    const float ix = rcp(dx), iy = rcp(dy), iz = rcp(dz);
    double best = INFINITY;
    float best_f = FLT_MAX;
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const int id = stack[--sp];
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(nodes + id)), q1 = __ldg(reinterpret_cast<const float4*>(nodes + id) + 1);
        WallNode n;
        n.lo[0] = q0.x; n.lo[1] = q0.y; n.lo[2] = q0.z; n.a = __float_as_int(q0.w);
        n.hi[0] = q1.x; n.hi[1] = q1.y; n.hi[2] = q1.z; n.b = __float_as_int(q1.w);
        float tn;
        if (!box_hit(n, fox, foy, foz, ix, iy, iz, best_f, tn)) continue;
        if (n.b < 0) {
            for (int k = n.a; k < n.a - n.b; k++) {
                const double t = tri_hit(tri + 9 * (size_t)k, ox, oy, oz, dx, dy, dz, 1e-9);
                if (t < best) { best = t; best_f = (float)t * 1.00001f + 1e-6f; }
            }
            continue;
        }
        // nearer child last onto the stack
        const float4 l0 = __ldg(reinterpret_cast<const float4*>(nodes + n.a)), l1 = __ldg(reinterpret_cast<const float4*>(nodes + n.a) + 1);
        const float cl = ((0.5f * (l0.x + l1.x) - fox) * (float)dx + (0.5f * (l0.y + l1.y) - foy) * (float)dy + (0.5f * (l0.z + l1.z) - foz) * (float)dz);
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(nodes + n.b)), r1 = __ldg(reinterpret_cast<const float4*>(nodes + n.b) + 1);
        const float cr = ((0.5f * (r0.x + r1.x) - fox) * (float)dx + (0.5f * (r0.y + r1.y) - foy) * (float)dy + (0.5f * (r0.z + r1.z) - foz) * (float)dz);
        if (sp + 2 > 64) continue;                                   // cannot happen: depth <= log2(n / 4) + a few
        if (cl < cr) { stack[sp++] = n.b; stack[sp++] = n.a; }
        else { stack[sp++] = n.a; stack[sp++] = n.b; }
    }
    return best;
}

__global__ void wall_hit_kernel(const WallNode* __restrict__ nodes, const double* __restrict__ tri, const double* __restrict__ origin,
                                const double* __restrict__ direction, int64_t n, double* __restrict__ t_hit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    t_hit[i] = first_hit(nodes, tri, origin[3 * i], origin[3 * i + 1], origin[3 * i + 2], direction[3 * i], direction[3 * i + 1], direction[3 * i + 2]);
}

__global__ void wall_clip_kernel(const WallNode* __restrict__ nodes, const double* __restrict__ tri, DevRays rays, double* __restrict__ seg_t0,
                                 double* __restrict__ seg_t1, double* __restrict__ t_hit) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rays.n_rays) return;
    const int64_t s0 = rays.seg_offset[i], s1 = rays.seg_offset[i + 1];
    double t = INFINITY;
    if (s1 > s0 || t_hit)
        t = first_hit(nodes, tri, rays.origin[3 * i], rays.origin[3 * i + 1], rays.origin[3 * i + 2], rays.direction[3 * i],
                      rays.direction[3 * i + 1], rays.direction[3 * i + 2]);
    if (t_hit) t_hit[i] = t;
    for (int64_t s = s0; s < s1; s++) seg_t1[s] = fmax(seg_t0[s], fmin(seg_t1[s], t));
}

int reserve(cb2_wall* w, int slot, size_t need) {
    if (w->stage_bytes[slot] >= need && w->stage[slot]) return CB2_OK;
    if (w->stage[slot]) cudaFree(w->stage[slot]);
    w->stage[slot] = nullptr;
    w->stage_bytes[slot] = 0;
    CB2_CUDA(cudaMalloc(&w->stage[slot], need + need / 4 + 256));
    w->stage_bytes[slot] = need + need / 4 + 256;
    return CB2_OK;
}

}  // namespace

extern "C" int cb2_wall_create(const cb2_wall_desc* d, int device, cb2_wall** out) {
    if (!d || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    *out = nullptr;
    if (d->abi_version != CB2_ABI_VERSION) return cb2_fail(CB2_ERR_VALUE, "abi_version mismatch (header %d, descriptor %d)", CB2_ABI_VERSION, d->abi_version);
    if (d->n_triangles < 1 || !d->vertices) return cb2_fail(CB2_ERR_VALUE, "a wall needs at least one triangle");
    if (d->n_triangles > 0x3fffffffLL) return cb2_fail(CB2_ERR_VALUE, "too many triangles");
    const int ndev = cb2_device_count();
    if (ndev <= 0) return ndev < 0 ? CB2_ERR_CUDA : cb2_fail(CB2_ERR_CUDA, "no CUDA device visible (libcherab_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return cb2_fail(CB2_ERR_VALUE, "device %d out of range (%d visible)", device, ndev);
    CB2_CUDA(cudaSetDevice(device));
    const int n = (int)d->n_triangles;
    Builder B;
    B.v = d->vertices;
    B.clo.resize(3 * (size_t)n); B.chi.resize(3 * (size_t)n); B.cen.resize(3 * (size_t)n); B.order.resize(n);
    for (int t = 0; t < n; t++) {
        const double* v = d->vertices + 9 * (size_t)t;
        for (int a = 0; a < 3; a++) {
            const double lo = fmin(v[a], fmin(v[3 + a], v[6 + a])), hi = fmax(v[a], fmax(v[3 + a], v[6 + a]));
            if (!(lo == lo) || !(hi - lo < 1e30)) return cb2_fail(CB2_ERR_VALUE, "triangle %d has a non-finite vertex", t);
            B.clo[3 * (size_t)t + a] = Builder::down(lo);
            B.chi[3 * (size_t)t + a] = Builder::up(hi);
            B.cen[3 * (size_t)t + a] = (float)((v[a] + v[3 + a] + v[6 + a]) / 3.0);
        }
    }
    std::iota(B.order.begin(), B.order.end(), 0);
    B.nodes.reserve((size_t)n);
    B.build(0, n);
    std::vector<double> tri(9 * (size_t)n);
    for (int k = 0; k < n; k++) memcpy(&tri[9 * (size_t)k], d->vertices + 9 * (size_t)B.order[k], 9 * sizeof(double));
    cb2_wall* w = (cb2_wall*)calloc(1, sizeof(cb2_wall));
    if (!w) return cb2_fail(CB2_ERR_MEMORY, "out of host memory");
    w->device = device;
    w->n_tri = n;
    w->n_nodes = (int)B.nodes.size();
    int rc = cb2_cuda_check(cudaMalloc((void**)&w->nodes, B.nodes.size() * sizeof(WallNode)), "cudaMalloc(wall nodes)");
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMalloc((void**)&w->tri, tri.size() * sizeof(double)), "cudaMalloc(wall triangles)");
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpy(w->nodes, B.nodes.data(), B.nodes.size() * sizeof(WallNode), cudaMemcpyHostToDevice), "cudaMemcpy(wall nodes)");
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaMemcpy(w->tri, tri.data(), tri.size() * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy(wall triangles)");
    if (rc != CB2_OK) { cb2_wall_destroy(w); return rc; }
    *out = w;
    return CB2_OK;
}

extern "C" int cb2_wall_destroy(cb2_wall* w) {
    if (!w) return CB2_OK;
    cudaSetDevice(w->device);
    if (w->nodes) cudaFree(w->nodes);
    if (w->tri) cudaFree(w->tri);
    for (int i = 0; i < 3; i++)
        if (w->stage[i]) cudaFree(w->stage[i]);
    free(w);
    return CB2_OK;
}

extern "C" int cb2_wall_hit(cb2_wall* w, const double* origin, const double* direction, int64_t n, double* t_hit) {
    if (!w || !origin || !direction || !t_hit) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (n <= 0) return CB2_OK;
    CB2_CUDA(cudaSetDevice(w->device));
    int rc;
    for (int k = 0; k < 2; k++)
        if ((rc = reserve(w, k, (size_t)n * 3 * sizeof(double))) != CB2_OK) return rc;
    if ((rc = reserve(w, 2, (size_t)n * sizeof(double))) != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpy(w->stage[0], origin, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CB2_CUDA(cudaMemcpy(w->stage[1], direction, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    wall_hit_kernel<<<(unsigned)((n + 127) / 128), 128>>>(w->nodes, w->tri, (const double*)w->stage[0], (const double*)w->stage[1], n, (double*)w->stage[2]);
    CB2_CUDA(cudaGetLastError());
    CB2_CUDA(cudaMemcpy(t_hit, w->stage[2], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return CB2_OK;
}

extern "C" int cb2_wall_clip_device(cb2_wall* w, const cb2_rays* r, double* t_hit_dev, void* stream) {
    if (!w || !r) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (r->n_rays < 0) return cb2_fail(CB2_ERR_VALUE, "negative ray count");
    if (r->n_rays == 0) return CB2_OK;
    if (!r->origin || !r->direction || !r->seg_offset) return cb2_fail(CB2_ERR_VALUE, "null ray arrays");
    CB2_CUDA(cudaSetDevice(w->device));
    DevRays dr;
    dr.n_rays = r->n_rays; dr.origin = r->origin; dr.direction = r->direction; dr.seg_offset = r->seg_offset; dr.seg_t0 = r->seg_t0; dr.seg_t1 = r->seg_t1;
    wall_clip_kernel<<<(unsigned)((r->n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(w->nodes, w->tri, dr, (double*)r->seg_t0, (double*)r->seg_t1, t_hit_dev);
    return cb2_cuda_check(cudaGetLastError(), "wall_clip_kernel launch");
}
