// cb2_raytransfer.cu — ray-transfer (geometry matrix) path-length sampler on sm_100a.
//
// Restates CylindricalRayTransferIntegrator.integrate / CartesianRayTransferIntegrator.integrate
// (cherab/tools/raytransfer/emitters.pyx:88-224): fixed-step midpoint sampling t = (it + 1/2) dt with
// n = max(min_samples, int(L/step)), cell index by C truncation, `res += dt` per step spent in a mapped cell.
// The per-source result is dt * (number of steps whose cell maps to that source), so the sequential run detection of
// the reference becomes a histogram: one warp per ray, 32 consecutive steps per iteration, lanes that start a run of
// equal sources add run_length * dt to that source with one atomic.
//
// Index decisions must match the reference bit for bit (a flipped step moves a whole dt between voxels),
// so the position -> index arithmetic is IEEE float64 with explicit _rn intrinsics in the reference's expression
// order (no FMA contraction: the reference is compiled for x86-64 without FMA).
//
// Output modes: 0 dense rows (small `bins`, rt_kernel), 1 count distinct sources per ray, 2 fill CSR rows (rt_csr_kernel).
// The CSR modes deduplicate a ray's sources in a per-warp open-addressing hash table in SHARED memory (key -> slot in the
// ray's CSR row; the number of distinct sources is bounded by the cell-boundary crossings of a straight line); the lengths
// accumulate in the row itself (fp64 atomics on a few KB that stay in L2).  The first version kept a dense double[bins]
// scratch row per warp in HBM and was latency-bound on it (ncu: long-scoreboard 50 stall cycles per issue, issue-active 7 %).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "cb2_internal.h"

#define FULL 0xffffffffu

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// m[0]*x + m[1]*y + m[2]*z + m[3], left to right
__device__ __forceinline__ double row_point(const double* m, double x, double y, double z) {
    return add_rn(add_rn(add_rn(mul_rn(m[0], x), mul_rn(m[1], y)), mul_rn(m[2], z)), m[3]);
}

__device__ __forceinline__ int rt_source(const DevRT& R, const double* st, const double* dir, double dt, int it) {
    // RayTransferIntegrator: midpoints (it + 1/2) dt (emitters.pyx:118); NumericalIntegrator: the trapezium nodes it * h
    const double t = R.trapezium ? mul_rn((double)it, dt) : mul_rn(add_rn((double)it, 0.5), dt);
    const double x = add_rn(st[0], mul_rn(dir[0], t));
    const double y = add_rn(st[1], mul_rn(dir[1], t));
    const double z = add_rn(st[2], mul_rn(dir[2], t));
    int i0, i1, i2;
    if (R.kind == CB2_RT_CYLINDRICAL) {
        i2 = (int)__ddiv_rn(z, R.s2);
        const double r = __dsqrt_rn(add_rn(mul_rn(x, x), mul_rn(y, y)));
        i0 = (int)__ddiv_rn(add_rn(r, -R.rmin), R.s0);
        if (R.n1 == 1) i1 = 0;
        else {
            double phi = mul_rn(180.0 / M_PI, atan2(y, x));
            phi = fmod(add_rn(phi, 360.0), R.period);
            i1 = (int)__ddiv_rn(phi, R.s1);
        }
    } else {
        i0 = (int)__ddiv_rn(x, R.s0);
        i1 = (int)__ddiv_rn(y, R.s1);
        i2 = (int)__ddiv_rn(z, R.s2);
    }
    // the reference would raise IndexError outside the grid; such steps are skipped
    if (i0 < 0 || i0 >= R.n0 || i1 < 0 || i1 >= R.n1 || i2 < 0 || i2 >= R.n2) return -1;
    return __ldg(R.voxel_map + ((size_t)i0 * R.n1 + i1) * R.n2 + i2);
}

__global__ void __launch_bounds__(256)
rt_kernel(DevRT R, DevRays rays, int mode, double* __restrict__ dense, int64_t* __restrict__ row_offset,
          int32_t* __restrict__ columns, double* __restrict__ lengths, double* __restrict__ scratch_all,
          int32_t* __restrict__ touched_all, int touch_cap, unsigned long long* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double* scratch = scratch_all ? scratch_all + (size_t)gw * R.bins : nullptr;
    int32_t* touched = touched_all ? touched_all + (size_t)gw * touch_cap : nullptr;
    unsigned long long steps = 0, overflow = 0;

    for (int64_t ray = gw; ray < rays.n_rays; ray += nwarps) {
        double* row = (mode == 0) ? dense + (size_t)ray * R.bins : scratch;
        int count = 0;
        const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
        const double dx = rays.direction[3 * ray], dy = rays.direction[3 * ray + 1], dz = rays.direction[3 * ray + 2];
        for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
            const double t0 = rays.seg_t0[sg], t1 = rays.seg_t1[sg];
            // start_point = far end, end_point = near end (Raysect convention), both to local space
            const double swx = add_rn(ox, mul_rn(t1, dx)), swy = add_rn(oy, mul_rn(t1, dy)), swz = add_rn(oz, mul_rn(t1, dz));
            const double ewx = add_rn(ox, mul_rn(t0, dx)), ewy = add_rn(oy, mul_rn(t0, dy)), ewz = add_rn(oz, mul_rn(t0, dz));
            double st[3], en[3], dir[3];
            for (int k = 0; k < 3; k++) {
                st[k] = row_point(R.w2l + 4 * k, swx, swy, swz);
                en[k] = row_point(R.w2l + 4 * k, ewx, ewy, ewz);
                dir[k] = add_rn(en[k], -st[k]);
            }
            const double length = __dsqrt_rn(add_rn(add_rn(mul_rn(dir[0], dir[0]), mul_rn(dir[1], dir[1])), mul_rn(dir[2], dir[2])));
            int n;
            double dt;
            if (R.trapezium) {
                // NumericalIntegrator [raysect]: intervals = max(min_samples - 1, ceil(L / step)), samples at k L / intervals, k = 0 .. intervals
                if (!(length > 0.0)) continue;
                for (int k = 0; k < 3; k++) dir[k] = __ddiv_rn(dir[k], length);
                int iv = (int)ceil(__ddiv_rn(length, R.step));
                if (iv < R.min_samples - 1) iv = R.min_samples - 1;
                if (iv < 1) iv = 1;
                dt = __ddiv_rn(length, (double)iv);
                n = iv + 1;
            } else {
                if (length < mul_rn(0.1, R.step)) continue;            // emitters.pyx:104-105
                for (int k = 0; k < 3; k++) dir[k] = __ddiv_rn(dir[k], length);
                n = (int)__ddiv_rn(length, R.step);
                if (n < R.min_samples) n = R.min_samples;
                dt = __ddiv_rn(length, (double)n);
            }
            // a run of `run` samples starting at sample `first` weighs (2 run - [first == 0] - [first + run == n]) half steps under
            // the trapezium rule, 2 run under the midpoint rule
            const double half = mul_rn(0.5, dt);
            const int trap = R.trapezium;
            if (lane == 0) steps += (unsigned long long)n;
            for (int it0 = 0; it0 < n; it0 += 32) {
                const int it = it0 + lane;
                const int src = (it < n) ? rt_source(R, st, dir, dt, it) : -2;
                int prev = __shfl_up_sync(FULL, src, 1);
                const bool head = (lane == 0) || (src != prev);
                const unsigned heads = __ballot_sync(FULL, head);
                bool first = false;
                if (head && src >= 0) {
                    const unsigned higher = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
                    const int next = higher ? (__ffs(higher) - 1) : 32;
                    const int run = next - lane, first_it = it0 + lane;
                    const int halves = 2 * run - (trap ? ((first_it == 0) + (first_it + run == n)) : 0);
                    const double add = mul_rn((double)halves, half);
                    const double old = atomicAdd(row + src, add);
                    first = (mode != 0) && (old == 0.0);
                }
                if (mode != 0) {
                    const unsigned fm = __ballot_sync(FULL, first);
                    if (first) {
                        const int pos = count + __popc(fm & ((1u << lane) - 1u));
                        if (pos < touch_cap) touched[pos] = src; else overflow++;
                    }
                    count += __popc(fm);
                }
            }
        }
        if (mode != 0) {
            __threadfence();
            __syncwarp();
            const int cnt = min(count, touch_cap);
            const int64_t off = (mode == 2) ? row_offset[ray] : 0;
            for (int pos = lane; pos < cnt; pos += 32) {
                const int src = touched[pos];
                const double v = __ldcg(scratch + src);
                __stcg(scratch + src, 0.0);
                if (mode == 2) { columns[off + pos] = src; lengths[off + pos] = v; }
            }
            if (mode == 1 && lane == 0) row_offset[ray] = cnt;
            __syncwarp();
        }
    }
    if (mode == 1 && gw == 0 && lane == 0) row_offset[rays.n_rays] = 0;
    if (stats) {
        for (int off = 16; off > 0; off >>= 1) overflow += __shfl_down_sync(FULL, overflow, off);
        if (lane == 0) {
            if (steps) atomicAdd(stats + 4, steps);
            if (overflow) atomicAdd(stats + 5, overflow);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// CSR modes: per-warp shared-memory hash (mode 1 count, mode 2 fill)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned rt_hash(int src, int hbits) { return ((unsigned)src * 2654435761u) >> (32 - hbits); }

// PACKED: one 32-bit word per table position, (source << 12) | slot, for grids with fewer than 2^20 sources and rays that can
// touch fewer than 4095 of them (every BASELINE configuration): 4 bytes per position instead of 6 + a used-position list, i.e.
// about twice the resident warps per SM for a latency-bound kernel.  The whole table is cleared after every ray.
#define RT_PENDING 0xFFFu
#ifndef RT_STEPS_PER_LANE
#define RT_STEPS_PER_LANE 4
#endif
#define RT_CSR_MAX_WARPS 9          // 288 threads x 3 CTAs per SM: 72 registers per thread, 27 resident warps
template <bool PACKED>
__global__ void __launch_bounds__(RT_CSR_MAX_WARPS * 32, 3)
rt_csr_kernel(DevRT R, DevRays rays, int mode, int64_t* __restrict__ row_offset, int32_t* __restrict__ columns,
              double* __restrict__ lengths, int hbits, int cap, int64_t row_stride, unsigned long long* __restrict__ stats) {
    extern __shared__ int rt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int H = 1 << hbits;
    // per warp: keys int32[H], slots uint16[H], used uint16[cap]   |   PACKED: entries uint32[H]
    const size_t per_warp = PACKED ? (size_t)H * 4 : (size_t)H * 4 + (size_t)H * 2 + (((size_t)cap * 2 + 3) & ~(size_t)3);
    unsigned char* base = reinterpret_cast<unsigned char*>(rt_smem) + per_warp * warp;
    int* keys = reinterpret_cast<int*>(base);
    unsigned short* slots = reinterpret_cast<unsigned short*>(base + (size_t)H * 4);
    unsigned short* used = reinterpret_cast<unsigned short*>(base + (size_t)H * 6);
    for (int i = lane; i < H; i += 32) keys[i] = -1;
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp, nwarps = (int64_t)gridDim.x * wpc;
    unsigned long long steps = 0, overflow = 0;

    for (int64_t ray = gw; ray < rays.n_rays; ray += nwarps) {
        int count = 0;
        // mode 2: the ray's row of the final CSR (offsets from the count pass); mode 3 (single traversal): its row of the strided
        // scratch [ray][row_stride], compacted afterwards
        const bool fill = mode >= 2;
        const int64_t off = (mode == 2) ? row_offset[ray] : (mode == 3 ? ray * row_stride : 0);
        const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
        const double dx = rays.direction[3 * ray], dy = rays.direction[3 * ray + 1], dz = rays.direction[3 * ray + 2];
        for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
            const double t0 = rays.seg_t0[sg], t1 = rays.seg_t1[sg];
            // start_point = far end, end_point = near end (Raysect convention), both to local space
            const double swx = add_rn(ox, mul_rn(t1, dx)), swy = add_rn(oy, mul_rn(t1, dy)), swz = add_rn(oz, mul_rn(t1, dz));
            const double ewx = add_rn(ox, mul_rn(t0, dx)), ewy = add_rn(oy, mul_rn(t0, dy)), ewz = add_rn(oz, mul_rn(t0, dz));
            double st[3], en[3], dir[3];
            for (int k = 0; k < 3; k++) {
                st[k] = row_point(R.w2l + 4 * k, swx, swy, swz);
                en[k] = row_point(R.w2l + 4 * k, ewx, ewy, ewz);
                dir[k] = add_rn(en[k], -st[k]);
            }
            const double length = __dsqrt_rn(add_rn(add_rn(mul_rn(dir[0], dir[0]), mul_rn(dir[1], dir[1])), mul_rn(dir[2], dir[2])));
            int n;
            double dt;
            if (R.trapezium) {
                // NumericalIntegrator [raysect]: intervals = max(min_samples - 1, ceil(L / step)), samples at k L / intervals, k = 0 .. intervals
                if (!(length > 0.0)) continue;
                for (int k = 0; k < 3; k++) dir[k] = __ddiv_rn(dir[k], length);
                int iv = (int)ceil(__ddiv_rn(length, R.step));
                if (iv < R.min_samples - 1) iv = R.min_samples - 1;
                if (iv < 1) iv = 1;
                dt = __ddiv_rn(length, (double)iv);
                n = iv + 1;
            } else {
                if (length < mul_rn(0.1, R.step)) continue;            // emitters.pyx:104-105
                for (int k = 0; k < 3; k++) dir[k] = __ddiv_rn(dir[k], length);
                n = (int)__ddiv_rn(length, R.step);
                if (n < R.min_samples) n = R.min_samples;
                dt = __ddiv_rn(length, (double)n);
            }
            // a run of `run` samples starting at sample `first` weighs (2 run - [first == 0] - [first + run == n]) half steps under
            // the trapezium rule, 2 run under the midpoint rule
            const double half = mul_rn(0.5, dt);
            const int trap = R.trapezium;
            if (lane == 0) steps += (unsigned long long)n;
            // 32 K consecutive steps per iteration: lane l owns steps it0 + K l + j, j < K.  The bookkeeping below — run heads, merge,
            // hash probe, slot numbering — is warp-wide work per iteration, not per step: K steps per lane divide it by K (it was
            // ~ 2/3 of the kernel's instructions with one step per lane; C4: 37.3 ms at K = 1, 29.8 at K = 2, 27.5 at K = 4, 29.8 at K = 8).  Tried and
            // dropped: index decisions by reciprocal multiplication and squared cell faces with the IEEE division / square-root chain
            // as the fallback near a face — bit-identical, but no faster (31.9 ms at K = 4): the chain is not what the kernel waits for.
            constexpr int K = RT_STEPS_PER_LANE;
            for (int it0 = 0; it0 < n; it0 += 32 * K) {
                int sj[K];
                unsigned mj[K];
                bool hj[K];
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int it = it0 + K * lane + j;
                    sj[j] = (it < n) ? rt_source(R, st, dir, dt, it) : -2;
                }
                const int prev = __shfl_up_sync(FULL, sj[K - 1], 1);
                unsigned many = 0u;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    hj[j] = j == 0 ? ((lane == 0) || (sj[0] != prev)) : (sj[j] != sj[j - 1]);
                    mj[j] = __ballot_sync(FULL, hj[j]);
                    many |= mj[j];
                }
                // first head position (in steps within the iteration, 0 .. 32 K - 1) after this lane's steps
                const unsigned later = (lane == 31) ? 0u : (many & ~((2u << lane) - 1u));
                int next = 32 * K;
                if (later) {
                    const int nl = __ffs(later) - 1;
                    int first = K - 1;
#pragma unroll
                    for (int j = K - 2; j >= 0; j--)
                        if ((mj[j] >> nl) & 1u) first = j;
                    next = K * nl + first;
                }
                // run length of every head of this lane: up to the lane's next head, else up to `next`
                int runj[K];
                int live = 0;                                      // bit j: head j is live (a mapped source starts a run there)
                {
                    int end = next - K * lane;                     // position of the next head relative to this lane's first step
#pragma unroll
                    for (int j = K - 1; j >= 0; j--) {
                        const int run = end - j, first_it = it0 + K * lane + j;
                        runj[j] = 2 * run - (trap ? ((first_it == 0) + (first_it + run >= n)) : 0);     // in half steps
                        if (hj[j]) { end = j; if (sj[j] >= 0) live |= 1 << j; }
                    }
                }
                // a lane with several live heads (runs shorter than K steps) takes them in further passes, in step order
                for (int pass = 0; pass < K; pass++) {
                    const bool head_act = live != 0;
                    const unsigned ham = __ballot_sync(FULL, head_act);
                    if (!ham) break;
                    int src = -1, run = 0;
                    if (head_act) {
                        const int j0 = __ffs(live) - 1;
                        live &= live - 1;
#pragma unroll
                        for (int j = 0; j < K; j++)
                            if (j == j0) { src = sj[j]; run = runj[j]; }
                    }
                    // run heads of this pass that carry the same source (a ray can leave a cell and come back within 64 steps) are
                    // merged: the lowest such lane speaks for the source with the sum of their run lengths — slot numbers (= order
                    // of first visit) and the order of the additions do not depend on which lane wins a race
                    bool act = head_act;
                    if (head_act && mode != 1) {                   // (the count of distinct sources does not depend on who claims)
                        const unsigned peers = __match_any_sync(ham, src);
                        run = (int)__reduce_add_sync(peers, (unsigned)run);
                        act = lane == __ffs(peers) - 1;
                    }
                    // phase 1: find or claim the key's table position
                    int h = 0;
                    bool is_new = false;
                    if (act) {
                        h = (int)rt_hash(src, hbits);
                        if (PACKED) {
                            const int claim = (int)(((unsigned)src << 12) | RT_PENDING);
                            for (;;) {
                                const int old = atomicCAS(&keys[h], -1, claim);
                                if (old == -1) { is_new = true; break; }
                                if (((unsigned)old >> 12) == (unsigned)src) break;
                                h = (h + 1) & (H - 1);
                            }
                        } else {
                            for (;;) {
                                const int old = atomicCAS(&keys[h], -1, src);
                                if (old == -1) { is_new = true; break; }
                                if (old == src) break;
                                h = (h + 1) & (H - 1);
                            }
                        }
                    }
                    // phase 2: new keys take consecutive slots of the ray's row
                    const unsigned nm = __ballot_sync(FULL, is_new);
                    if (is_new) {
                        const int slot = count + __popc(nm & ((1u << lane) - 1u));
                        if (slot < cap) {
                            if (PACKED) keys[h] = (int)(((unsigned)src << 12) | (unsigned)slot);
                            else { slots[h] = (unsigned short)slot; used[slot] = (unsigned short)h; }
                            if (fill) { columns[off + slot] = src; lengths[off + slot] = 0.0; }
                        } else overflow++;
                    }
                    count += __popc(nm);
                    __syncwarp();
                    // phase 3: run length into the row
                    if (fill && act) {
                        const int slot = PACKED ? (int)((unsigned)keys[h] & 0xFFFu) : (int)slots[h];
                        if (slot < cap) atomicAdd(lengths + off + slot, mul_rn((double)run, half));
                    }
                }
            }
        }
        // reset the touched table positions for the next ray
        __syncwarp();
        const int cnt = min(count, cap);
        if (PACKED || count > cap) {                      // (after a slot overflow the used-position list is incomplete: clear everything)
            if (count > 0)
                for (int i = lane; i < H; i += 32) keys[i] = -1;
        } else {
            for (int pos = lane; pos < cnt; pos += 32) keys[used[pos]] = -1;
        }
        if ((mode == 1 || mode == 3) && lane == 0) row_offset[ray] = cnt;
        __syncwarp();
    }
    if ((mode == 1 || mode == 3) && gw == 0 && lane == 0) row_offset[rays.n_rays] = 0;
    if (stats) {
        for (int o = 16; o > 0; o >>= 1) overflow += __shfl_down_sync(FULL, overflow, o);
        if (lane == 0) {
            if (steps) atomicAdd(stats + 4, steps);
            if (overflow) atomicAdd(stats + 5, overflow);
        }
    }
}

// single-traversal CSR, last step: the rays' rows move from the strided scratch to their places in the CSR (one warp per ray)
__global__ void __launch_bounds__(256)
rt_compact_kernel(int64_t n_rays, int64_t row_stride, const int64_t* __restrict__ row_offset, const int32_t* __restrict__ scratch_cols,
                  const double* __restrict__ scratch_len, int32_t* __restrict__ columns, double* __restrict__ lengths) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= n_rays) return;
    const int64_t o = row_offset[ray], cnt = row_offset[ray + 1] - o, s = ray * row_stride;
    for (int64_t i = lane; i < cnt; i += 32) {
        columns[o + i] = __ldcs(scratch_cols + s + i);
        lengths[o + i] = __ldcs(scratch_len + s + i);
    }
}

int cb2_launch_rt_compact(int64_t n_rays, int64_t row_stride, const int64_t* row_offset, const int32_t* scratch_cols, const double* scratch_len,
                          int32_t* columns, double* lengths, cudaStream_t st) {
    rt_compact_kernel<<<(unsigned)((n_rays * 32 + 255) / 256), 256, 0, st>>>(n_rays, row_stride, row_offset, scratch_cols, scratch_len, columns, lengths);
    return cb2_cuda_check(cudaGetLastError(), "rt_compact_kernel launch");
}

int cb2_launch_rt(const cb2_rt_scene* sc, const DevRays& rays, int mode, double* dense_out, int accumulate, int64_t* row_offset,
                  int32_t* columns, double* lengths, unsigned long long* stats_dev, cudaStream_t st) {
    (void)accumulate;
    if (mode == 0) {
        const int threads = 256;
        int64_t warps = rays.n_rays;
        if (warps > (int64_t)sc->n_warps * 4) warps = (int64_t)sc->n_warps * 4;
        int64_t blocks = (warps * 32 + threads - 1) / threads;
        if (blocks < 1) blocks = 1;
        rt_kernel<<<(unsigned)blocks, threads, 0, st>>>(sc->rt, rays, mode, dense_out, row_offset, columns, lengths, nullptr, nullptr,
                                                        sc->touch_cap, stats_dev);
        return cb2_cuda_check(cudaGetLastError(), "rt_kernel launch");
    }
    // CSR: per-warp hash table in shared memory sized from the bound on distinct sources per ray
    const int cap = sc->touch_cap;
    int hbits = 8;
    while ((1 << hbits) < cap + cap / 4 + 8 && hbits < 16) hbits++;     // load factor <= 0.8 even for a ray that reaches the bound
    const bool packed = sc->rt.bins < (1 << 20) && cap < (int)RT_PENDING && getenv("CB2_RT_UNPACKED") == nullptr;
    const size_t per_warp = packed ? ((size_t)4 << hbits) : ((size_t)6 << hbits) + (((size_t)cap * 2 + 3) & ~(size_t)3);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device);
    cudaFuncAttributes fa;
    CB2_CUDA(cudaFuncGetAttributes(&fa, packed ? rt_csr_kernel<true> : rt_csr_kernel<false>));
    const int reg_warps = fa.numRegs > 0 ? 65536 / (((fa.numRegs + 7) & ~7) * 32) : 16;   // warps per SM the register file allows (8-register granules)
    // the kernel is latency-bound: take the CTA shape that keeps the most warps resident (shared memory and registers permitting)
    int wpc = 0, ctas_per_sm = 1;
    for (int c = 1; c <= 4; c++) {
        int w = (int)std::min<size_t>(RT_CSR_MAX_WARPS, ((size_t)220 * 1024 / c) / per_warp);
        w = std::min(w, reg_warps / c);
        if (w >= 1 && w * c > wpc * ctas_per_sm) { wpc = w; ctas_per_sm = c; }
    }
    if (const char* shape = getenv("CB2_RT_SHAPE")) {     // "ctas,warps" — tuning experiments only
        int c = 0, w = 0;
        if (sscanf(shape, "%d,%d", &c, &w) == 2 && c >= 1 && w >= 1 && w <= RT_CSR_MAX_WARPS && (size_t)c * w * per_warp <= (size_t)220 * 1024) { ctas_per_sm = c; wpc = w; }
    }
    if (wpc < 1) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "ray-transfer grid too large for the shared-memory source table (%d distinct sources per ray)", cap);
    const size_t smem = per_warp * wpc;
    int64_t blocks = std::min<int64_t>((rays.n_rays + wpc - 1) / wpc, (int64_t)sms * ctas_per_sm);
    if (blocks < 1) blocks = 1;
    if (packed) {
        CB2_CUDA(cudaFuncSetAttribute(rt_csr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rt_csr_kernel<true><<<(unsigned)blocks, wpc * 32, smem, st>>>(sc->rt, rays, mode, row_offset, columns, lengths, hbits, cap, (int64_t)cap, stats_dev);
    } else {
        CB2_CUDA(cudaFuncSetAttribute(rt_csr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rt_csr_kernel<false><<<(unsigned)blocks, wpc * 32, smem, st>>>(sc->rt, rays, mode, row_offset, columns, lengths, hbits, cap, (int64_t)cap, stats_dev);
    }
    return cb2_cuda_check(cudaGetLastError(), "rt_csr_kernel launch");
}

// in-place exclusive scan of int64 counts (single block; n is at most a few million rays)
__global__ void __launch_bounds__(1024) scan_kernel(int64_t* __restrict__ a, int64_t n) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + tid;
        const int64_t v = (i < n) ? a[i] : 0;
        int64_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(FULL, w, off);
                if (lane >= off) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t incl = x + (warp ? warp_sums[warp - 1] : 0) + carry;
        if (i < n) a[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
}

int cb2_launch_scan(int64_t* counts_inout, int64_t n, cudaStream_t st) {
    scan_kernel<<<1, 1024, 0, st>>>(counts_inout, n);
    return cb2_cuda_check(cudaGetLastError(), "scan_kernel launch");
}
