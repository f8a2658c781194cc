// cb2_peaks.cu — microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
// FP32 FMA issue rate and MUFU.EX2 (SFU) rate of the device, measured live in the same process as the bench
// ("builder must measure FP32 FMA ... MUFU ... peaks with microbenchmarks in the same run").
#include "cb2_internal.h"

template <int ILP>
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456f) out[0] = s;   // never true: keeps the chain alive
}

template <int ILP>
__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (float)(threadIdx.x + i) * 1e-4f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456) out[0] = s;
}

// FP64 FMA issue rate (the ray-transfer kernel's index arithmetic is float64): TFLOP/s
extern "C" int cb2_measure_peak_fp64(int device, double* fp64_tflops) {
    CB2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB2_CUDA(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CB2_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CB2_CUDA(cudaEventCreate(&e0));
    CB2_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    constexpr int ILP = 8;
    float ms = 0.f;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CB2_CUDA(cudaEventRecord(e0));
        dfma_peak_kernel<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-7);
        CB2_CUDA(cudaEventRecord(e1));
        CB2_CUDA(cudaEventSynchronize(e1));
        CB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * blocks * threads * (double)iters * ILP / (ms * 1e-3) * 1e-12;
        if (rep > 0 && fl > best) best = fl;
    }
    CB2_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (fp64_tflops) *fp64_tflops = best;
    return CB2_OK;
}

extern "C" int cb2_measure_peaks(int device, double* fp32_tflops, double* sfu_tops, double* sm_clock_mhz) {
    CB2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB2_CUDA(cudaGetDeviceProperties(&prop, device));
    float* d = nullptr;
    CB2_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CB2_CUDA(cudaEventCreate(&e0));
    CB2_CUDA(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    constexpr int ILP = 8;
    float ms = 0.f;
    double best_fma = 0, best_sfu = 0;
    for (int rep = 0; rep < 6; rep++) {
        CB2_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<ILP><<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
        CB2_CUDA(cudaEventRecord(e1));
        CB2_CUDA(cudaEventSynchronize(e1));
        CB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * blocks * threads * (double)iters * ILP / (ms * 1e-3) * 1e-12;
        if (rep > 0 && fl > best_fma) best_fma = fl;
        CB2_CUDA(cudaEventRecord(e0));
        mufu_peak_kernel<ILP><<<blocks, threads>>>(d, iters);
        CB2_CUDA(cudaEventRecord(e1));
        CB2_CUDA(cudaEventSynchronize(e1));
        CB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * threads * (double)iters * ILP / (ms * 1e-3) * 1e-12;
        if (rep > 0 && ops > best_sfu) best_sfu = ops;
    }
    CB2_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (fp32_tflops) *fp32_tflops = best_fma;
    if (sfu_tops) *sfu_tops = best_sfu;
    if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate * 1e-3;
    return CB2_OK;
}
