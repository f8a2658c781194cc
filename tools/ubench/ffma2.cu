// Microbenchmark: does fma.rn.f32x2 (SASS FFMA2) free issue slots next to MUFU.EX2 on sm_100a?
// Variants per loop iteration and thread (8 independent chains):
//   0: 8 x (MUFU.EX2 + 7 FFMA)        — the instruction mix of bin_kernel's series evaluation, scalar
//   1: 4 x (2 MUFU.EX2 + 7 FFMA2)     — the same work on f32x2 pairs
//   2: 8 x 7 FFMA                      3: 4 x 7 FFMA2          4: 8 x MUFU.EX2
// Prints evaluations (one ex2 + 7 fma = one "bin evaluation") per second.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int V>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = a * (threadIdx.x + i);
    u64 p[4];
    for (int i = 0; i < 4; i++) p[i] = pack(x[2 * i], x[2 * i + 1]);
    const u64 ab = pack(a, a), bb = pack(b, b);
    for (int it = 0; it < iters; it++) {
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float e = ex2(x[i]);
                float y = x[i];
#pragma unroll
                for (int j = 0; j < 6; j++) y = fmaf(y, a, b);
                x[i] = fmaf(e, y, b);
            }
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float lo, hi;
                unpack(p[i], lo, hi);
                u64 e = pack(ex2(lo), ex2(hi));
                u64 y = p[i];
#pragma unroll
                for (int j = 0; j < 6; j++) y = fma2(y, ab, bb);
                p[i] = fma2(e, y, bb);
            }
        } else if (V == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float y = x[i];
#pragma unroll
                for (int j = 0; j < 7; j++) y = fmaf(y, a, b);
                x[i] = y;
            }
        } else if (V == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                u64 y = p[i];
#pragma unroll
                for (int j = 0; j < 7; j++) y = fma2(y, ab, bb);
                p[i] = y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = ex2(x[i]);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    for (int i = 0; i < 4; i++) { float lo, hi; unpack(p[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, float* out, int sms) {
    const int iters = 20000, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<grid, 256>>>(out, 100, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double evals = (double)grid * 256 * iters * 8;
    printf("%-28s %8.3f ms  %7.3f T evaluations/s  (%.2f cycles per warp-evaluation per SMSP at 1.965 GHz)\n", name, ms, evals / ms * 1e-9,
           1.965e9 * ms * 1e-3 * sms * 4 / (evals / 32));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    run<0>("ex2 + 7 FFMA", out, p.multiProcessorCount);
    run<1>("2 ex2 + 7 FFMA2 (pairs)", out, p.multiProcessorCount);
    run<2>("7 FFMA", out, p.multiProcessorCount);
    run<3>("7 FFMA2 (pairs)", out, p.multiProcessorCount);
    run<4>("ex2", out, p.multiProcessorCount);
    return 0;
}
