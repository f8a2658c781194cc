// cb2_contract.cu — K2: the Bremsstrahlung contraction  out[ray][bin] += scale * sum_k mom[ray][k] * phi[k][bin].
//
// mom  [n_rays][k_pad] fp32  per-ray moments on the (charge, temperature-node) grid, written by emission_kernel<.,.,3>
// phi  [k_pad][n_pad]  fp32  per-scene table (cb2_api.cu::build_brems_moments), k_pad % 16 == 0, n_pad % 128 == 0
// out  [n_rays][bins]  fp32 or fp64 frame, read-modify-write
//
// The continuum needs fp32-level accuracy per term (the acceptance is 1e-4 relative per bin and the Lagrange weights
// change sign), so this runs on the FP32 pipe, not on reduced-precision tensor-core formats: a 128x128x16 tiled FFMA
// kernel, 256 threads, 8x8 outputs per thread in two 4-wide groups per axis (conflict-free 128-bit shared loads),
// register-staged double buffering (global loads of tile k+1 overlap the FMAs of tile k, one barrier per tile).
// Bound: FP32 issue — 2 * n_rays * k_pad * n_pad flop; operands come from L2 (phi is a few MB, a moment tile is reused
// by the n_pad/128 CTAs that are adjacent in launch order).
#include "cb2_internal.h"

#define CT_BM 128
#define CT_BN 128
#define CT_BK 16
#define CT_AS (CT_BM + 4)   // padded row stride of the transposed A tile

template <int OUT_F64>
__global__ void __launch_bounds__(256, 2)
contract_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t M, int K, int N, int bins, void* __restrict__ out,
                float scale_f, double scale_d) {
    __shared__ __align__(16) float As[2][CT_BK][CT_AS];
    __shared__ __align__(16) float Bs[2][CT_BK][CT_BN];
    const int tid = threadIdx.x;
    const int n_tiles = N / CT_BN;
    const int64_t tile_m = blockIdx.x / n_tiles;
    const int tile_n = blockIdx.x % n_tiles;
    const int64_t m0 = tile_m * CT_BM;
    const int n0 = tile_n * CT_BN;

    // global -> register staging: A tile 128 rows x 16 k (each thread two float4 along k), B tile 16 k x 128 n
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;
    const int64_t ar0 = min(m0 + a_row, M - 1), ar1 = min(m0 + a_row + 64, M - 1);
    const float* a_ptr0 = A + ar0 * K + a_kq;
    const float* a_ptr1 = A + ar1 * K + a_kq;
    const float* b_ptr0 = B + (size_t)b_k * N + n0 + b_n;
    const float* b_ptr1 = B + (size_t)(b_k + 8) * N + n0 + b_n;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    float4 ra0 = __ldg(reinterpret_cast<const float4*>(a_ptr0)), ra1 = __ldg(reinterpret_cast<const float4*>(a_ptr1));
    float4 rb0 = __ldg(reinterpret_cast<const float4*>(b_ptr0)), rb1 = __ldg(reinterpret_cast<const float4*>(b_ptr1));
    auto stage = [&](int buf) {
        As[buf][a_kq + 0][a_row] = ra0.x; As[buf][a_kq + 1][a_row] = ra0.y; As[buf][a_kq + 2][a_row] = ra0.z; As[buf][a_kq + 3][a_row] = ra0.w;
        As[buf][a_kq + 0][a_row + 64] = ra1.x; As[buf][a_kq + 1][a_row + 64] = ra1.y; As[buf][a_kq + 2][a_row + 64] = ra1.z; As[buf][a_kq + 3][a_row + 64] = ra1.w;
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb0;
        *reinterpret_cast<float4*>(&Bs[buf][b_k + 8][b_n]) = rb1;
    };
    stage(0);
    __syncthreads();

    const int nk = K / CT_BK;
    for (int kt = 0; kt < nk; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            const size_t ko = (size_t)(kt + 1) * CT_BK;
            ra0 = __ldg(reinterpret_cast<const float4*>(a_ptr0 + ko));
            ra1 = __ldg(reinterpret_cast<const float4*>(a_ptr1 + ko));
            rb0 = __ldg(reinterpret_cast<const float4*>(b_ptr0 + ko * N));
            rb1 = __ldg(reinterpret_cast<const float4*>(b_ptr1 + ko * N));
        }
#pragma unroll
        for (int k = 0; k < CT_BK; k++) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stage(buf ^ 1);      // the other buffer was last read before the previous barrier
            __syncthreads();
        }
    }

    // epilogue: read-modify-write of the frame rows (128-bit when the row pitch allows it)
    const bool vec_ok = OUT_F64 ? ((bins & 1) == 0) : ((bins & 3) == 0);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int c = n0 + h * 64 + tx * 4;
            if (c >= bins) continue;
            const size_t idx = (size_t)r * bins + c;
            if (OUT_F64) {
                double* p = (double*)out + idx;
                if (vec_ok && c + 3 < bins) {
                    double2 v0 = *reinterpret_cast<double2*>(p), v1 = *reinterpret_cast<double2*>(p + 2);
                    v0.x += scale_d * (double)acc[i][4 * h]; v0.y += scale_d * (double)acc[i][4 * h + 1];
                    v1.x += scale_d * (double)acc[i][4 * h + 2]; v1.y += scale_d * (double)acc[i][4 * h + 3];
                    *reinterpret_cast<double2*>(p) = v0; *reinterpret_cast<double2*>(p + 2) = v1;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (c + j < bins) p[j] += scale_d * (double)acc[i][4 * h + j];
                }
            } else {
                float* p = (float*)out + idx;
                if (vec_ok && c + 3 < bins) {
                    float4 v = *reinterpret_cast<float4*>(p);
                    v.x = fmaf(scale_f, acc[i][4 * h], v.x); v.y = fmaf(scale_f, acc[i][4 * h + 1], v.y);
                    v.z = fmaf(scale_f, acc[i][4 * h + 2], v.z); v.w = fmaf(scale_f, acc[i][4 * h + 3], v.w);
                    *reinterpret_cast<float4*>(p) = v;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (c + j < bins) p[j] = fmaf(scale_f, acc[i][4 * h + j], p[j]);
                }
            }
        }
    }
}

int cb2_launch_contract(const float* mom, const float* phi, int64_t n_rays, int k_pad, int n_pad, int bins, void* out, int out_f64,
                        double scale, cudaStream_t st) {
    if (n_rays <= 0) return CB2_OK;
    if (k_pad % CT_BK || n_pad % CT_BN) return cb2_fail(CB2_ERR_RUNTIME, "internal error: contraction operands are not padded");
    const int64_t m_tiles = (n_rays + CT_BM - 1) / CT_BM;
    const int64_t blocks = m_tiles * (n_pad / CT_BN);
    if (blocks > 0x7fffffffLL) return cb2_fail(CB2_ERR_VALUE, "too many rays for one contraction launch");
    if (out_f64) contract_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(mom, phi, n_rays, k_pad, n_pad, bins, out, (float)scale, scale);
    else contract_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(mom, phi, n_rays, k_pad, n_pad, bins, out, (float)scale, scale);
    return cb2_cuda_check(cudaGetLastError(), "contract_kernel launch");
}
