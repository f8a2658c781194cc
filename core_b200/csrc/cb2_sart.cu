// SART inversion on the device-resident geometry matrix.
//
// Replaces cherab/tools/inversions/sart.pyx:26-155 (invert_sart), :161-302 (invert_constrained_sart) and the OpenCL
// solver cherab/tools/inversions/opencl/sart_opencl.py:33-318.  Not a translation of either: the reference walks a dense
// matrix cell by cell; here the matrix is sparse (what the ray-transfer kernel produces), stored once as CSR and once as
// CSC, and one iteration is two HBM streams over it,
//     backward (CSC, one warp per source):   x_j <- max(0, x_j + omega/W_+j * sum_i W_ij w_i - beta (L x)_j)
//     forward  (CSR, one warp per detector): y_i = sum_j W_ij x_j ;  w_i = (m_i - y_i) / W_i+ ;  sum_i y_i^2
// for FR measurement frames at a time, so the bytes streamed per iteration do not grow with the number of frames.
// The stop test of the reference (sart.pyx:147-153) runs on the device in the last CTA of the forward pass; once a frame
// has stopped, later launches leave its solution untouched, so the host only reads the flags back every few iterations.
// All reductions have a fixed order (warp butterflies, per-CTA partials summed by one warp): results are reproducible.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "cb2_internal.h"

extern "C" int cb2_device_count(void);

namespace {

constexpr int SART_THREADS = 256;              // 8 warps per CTA
constexpr int SART_WARPS = SART_THREADS / 32;
constexpr int SART_CHECK_EVERY = 8;            // iterations between read-backs of the stop flags
constexpr int SART_FR = 4;                     // frames per pass

struct SartMatrix {                            // one sparse operand: CSR rows or CSC columns
    int64_t* offset = nullptr;                 // [n + 1]
    int32_t* index = nullptr;                  // [nnz]
    void* value = nullptr;                     // [nnz] float or double
};

template <typename VT>
__device__ __forceinline__ double load_value(const void* v, int64_t e) {
    return (double)__ldg(reinterpret_cast<const VT*>(v) + e);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// acc[f] = sum over the entries [e0, e1) of value * vec[index * FR + f]; lanes stride the entries, 4 loads in flight
template <typename VT, int FR>
__device__ __forceinline__ void sparse_dot(const int32_t* __restrict__ index, const void* __restrict__ value, int64_t e0, int64_t e1,
                                           const double* __restrict__ vec, int lane, double (&acc)[FR]) {
#pragma unroll
    for (int f = 0; f < FR; f++) acc[f] = 0.0;
    int64_t e = e0 + lane;
    for (; e + 96 < e1; e += 128) {
        int32_t c[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            c[u] = __ldg(index + e + 32 * u);
            v[u] = load_value<VT>(value, e + 32 * u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int f = 0; f < FR; f++) acc[f] = fma(v[u], vec[(int64_t)c[u] * FR + f], acc[f]);
    }
    for (; e < e1; e += 32) {
        const int32_t c = __ldg(index + e);
        const double v = load_value<VT>(value, e);
#pragma unroll
        for (int f = 0; f < FR; f++) acc[f] = fma(v, vec[(int64_t)c * FR + f], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < FR; f++) acc[f] = warp_sum(acc[f]);
}

// per-frame solver state on the device
struct SartFrames {
    double* x;            // [n_sources][FR]  current estimate
    double* w;            // [n_detectors][FR] (m - y_hat) / ray_length
    double* m;            // [n_detectors][FR] measurements
    double* gp;           // [n_sources][FR]  beta * L x
    double* partial;      // [grid][FR] per-CTA sums of y_hat^2
    double* conv;         // [FR][max_iterations]
    double* m_sq;         // [FR]
    int32_t* stopped;     // [FR] 0 while iterating, else the iteration count at which the frame stopped
    int32_t* flags;       // [0] ticket, [1] iteration index k of the running pass, [2] number of frames still iterating
};

// gp = beta * L x   (sart.pyx:255-256)
template <int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_penalty_kernel(int64_t n_sources, const int64_t* __restrict__ off,
                                                                    const int32_t* __restrict__ idx, const void* __restrict__ val,
                                                                    SartFrames fr, double beta) {
    if (fr.flags[2] == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * SART_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * SART_WARPS + (threadIdx.x >> 5); j < n_sources; j += n_warps) {
        double acc[FR];
        sparse_dot<double, FR>(idx, val, off[j], off[j + 1], fr.x, lane, acc);
        if (lane < FR) {
            double a = acc[0];
#pragma unroll
            for (int f = 1; f < FR; f++) a = lane == f ? acc[f] : a;
            fr.gp[j * FR + lane] = beta * a;
        }
    }
}

// x_j <- max(0, x_j + omega / W_+j * sum_i W_ij w_i - gp_j)   (sart.pyx:118-142, :258-285)
template <typename VT, int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_backward_kernel(int64_t n_sources, const int64_t* __restrict__ off,
                                                                     const int32_t* __restrict__ idx, const void* __restrict__ val,
                                                                     const double* __restrict__ density, SartFrames fr,
                                                                     double relaxation, int with_penalty) {
    if (fr.flags[2] == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * SART_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * SART_WARPS + (threadIdx.x >> 5); j < n_sources; j += n_warps) {
        const double dens = density[j];
        double acc[FR];
        if (dens > 0.0) {
            sparse_dot<VT, FR>(idx, val, off[j], off[j + 1], fr.w, lane, acc);
        } else {
#pragma unroll
            for (int f = 0; f < FR; f++) acc[f] = 0.0;
        }
        if (lane < FR && fr.stopped[lane] == 0) {
            double a = acc[0];
#pragma unroll
            for (int f = 1; f < FR; f++) a = lane == f ? acc[f] : a;
            double xn = fr.x[j * FR + lane];
            if (dens > 0.0) xn += relaxation / dens * a;
            if (with_penalty) xn -= fr.gp[j * FR + lane];
            fr.x[j * FR + lane] = xn < 0.0 ? 0.0 : xn;
        }
    }
}

// y_hat = W x, w = (m - y_hat)/W_i+, |y_hat|^2; the last CTA records the convergence and applies the stop test
// (sart.pyx:96-97, :144-153).  record == 0: the initial projection before the first iteration.
template <typename VT, int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_forward_kernel(int64_t n_detectors, const int64_t* __restrict__ off,
                                                                    const int32_t* __restrict__ idx, const void* __restrict__ val,
                                                                    const double* __restrict__ inv_length, SartFrames fr, int record,
                                                                    int max_iterations, double conv_tol) {
    if (fr.flags[2] == 0) return;
    __shared__ double s_sq[SART_WARPS][FR];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * SART_WARPS;
    double sq = 0.0;   // lane f < FR carries frame f
    for (int64_t i = (int64_t)blockIdx.x * SART_WARPS + warp; i < n_detectors; i += n_warps) {
        double acc[FR];
        sparse_dot<VT, FR>(idx, val, off[i], off[i + 1], fr.x, lane, acc);
        if (lane < FR) {
            double y = acc[0];
#pragma unroll
            for (int f = 1; f < FR; f++) y = lane == f ? acc[f] : y;
            fr.w[i * FR + lane] = (fr.m[i * FR + lane] - y) * inv_length[i];   // inv_length is 0 for rays of zero length (sart.pyx:127-128)
            sq = fma(y, y, sq);
        }
    }
    if (!record) return;
    if (lane < FR) s_sq[warp][lane] = sq;
    __syncthreads();
    if (threadIdx.x < FR) {
        double t = 0.0;
        for (int k = 0; k < SART_WARPS; k++) t += s_sq[k][threadIdx.x];
        fr.partial[(int64_t)blockIdx.x * FR + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&fr.flags[0], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    const int k = fr.flags[1];
    int still = 0;
    for (int f = 0; f < FR; f++) {
        double t = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(fr.partial + (int64_t)b * FR + f);
        t = warp_sum(t);
        if (lane == 0 && fr.stopped[f] == 0) {
            const double msq = fr.m_sq[f];
            const double c = (msq - t) / msq;
            fr.conv[(int64_t)f * max_iterations + k] = c;
            int stop = k + 1 == max_iterations;
            if (k > 0 && fabs(c - fr.conv[(int64_t)f * max_iterations + k - 1]) < conv_tol) stop = 1;
            if (stop) fr.stopped[f] = k + 1; else still++;
        }
    }
    if (lane == 0) {
        fr.flags[0] = 0;
        fr.flags[1] = k + 1;
        fr.flags[2] = still;
    }
}

template <typename T>
__global__ void gather_values_kernel(int64_t nnz, const int32_t* __restrict__ perm, const double* __restrict__ src, T* __restrict__ dst) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) dst[e] = (T)src[perm[e]];
}
__global__ void gather_rows_kernel(int64_t nnz, const int32_t* __restrict__ perm, const int32_t* __restrict__ rows, int32_t* __restrict__ dst) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) dst[e] = rows[perm[e]];
}
template <typename T>
__global__ void convert_values_kernel(int64_t nnz, const double* __restrict__ src, T* __restrict__ dst) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) dst[e] = (T)src[e];
}
// rows[e] = row of entry e; iota
__global__ void expand_rows_kernel(int64_t n_rows, const int64_t* __restrict__ off, int32_t* __restrict__ rows, int32_t* __restrict__ iota) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += n_warps)
        for (int64_t e = off[r] + lane; e < off[r + 1]; e += 32) { rows[e] = (int32_t)r; iota[e] = (int32_t)e; }
}
// column offsets from the sorted column keys: off[c] = first entry with key >= c
__global__ void column_offsets_kernel(int64_t nnz, int64_t n_cols, const int32_t* __restrict__ keys, int64_t* __restrict__ off) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e <= nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = e == 0 ? 0 : (int64_t)keys[e - 1] + 1;
        const int64_t hi = e == nnz ? n_cols : (int64_t)keys[e];
        for (int64_t c = lo; c <= hi; c++) off[c] = e;
    }
}
// sums of the rows of a sparse operand (W_i+ from CSR, W_+j from CSC); inverse != 0 stores 1/sum, 0 for empty sums
template <typename VT>
__global__ void row_sums_kernel(int64_t n, const int64_t* __restrict__ off, const void* __restrict__ val, double* __restrict__ out, int inverse) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += n_warps) {
        double s = 0.0;
        for (int64_t e = off[r] + lane; e < off[r + 1]; e += 32) s += load_value<VT>(val, e);
        s = warp_sum(s);
        if (lane == 0) out[r] = inverse ? (s == 0.0 ? 0.0 : 1.0 / s) : s;
    }
}

}  // namespace

struct cb2_sart {
    int device = 0, value_f64 = 1, sm_count = 148;
    int64_t n_det = 0, n_src = 0, nnz = 0, lap_nnz = 0;
    SartMatrix csr, csc, lap;
    double* density = nullptr;      // W_+j
    double* inv_length = nullptr;   // 1 / W_i+
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;
    int64_t last_iterations = 0;
};

namespace {

void free_matrix(SartMatrix& m) {
    cudaFree(m.offset); cudaFree(m.index); cudaFree(m.value);
    m = SartMatrix();
}

int grid_for(const cb2_sart* s, int64_t rows) {
    int64_t need = (rows + SART_WARPS - 1) / SART_WARPS;
    int64_t cap = (int64_t)s->sm_count * 8;     // 8 CTAs of 256 threads per SM: every warp slot filled, fixed grid => fixed reduction order
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

// dense row-major host matrix -> host CSR (exact zeros dropped: they add nothing to any sum of the reference loop)
template <typename T>
void dense_to_csr(const T* a, int64_t rows, int64_t cols, std::vector<int64_t>& off, std::vector<int32_t>& idx, std::vector<double>& val) {
    off.assign(rows + 1, 0);
    for (int64_t r = 0; r < rows; r++) {
        int64_t n = 0;
        for (int64_t c = 0; c < cols; c++) n += a[r * cols + c] != (T)0;
        off[r + 1] = off[r] + n;
    }
    idx.resize(off[rows]); val.resize(off[rows]);
    for (int64_t r = 0; r < rows; r++) {
        int64_t e = off[r];
        for (int64_t c = 0; c < cols; c++)
            if (a[r * cols + c] != (T)0) { idx[e] = (int32_t)c; val[e] = (double)a[r * cols + c]; e++; }
    }
}

int upload_laplacian(cb2_sart* s, const double* dense, const int64_t* off, const int32_t* cols, const double* vals) {
    free_matrix(s->lap);
    s->lap_nnz = 0;
    if (!dense && !off) return CB2_OK;
    std::vector<int64_t> h_off; std::vector<int32_t> h_idx; std::vector<double> h_val;
    if (dense) {
        dense_to_csr<double>(dense, s->n_src, s->n_src, h_off, h_idx, h_val);
        off = h_off.data(); cols = h_idx.data(); vals = h_val.data();
    } else if (!cols || !vals) {
        return cb2_fail(CB2_ERR_VALUE, "Laplacian CSR needs row_offset, columns and values");
    }
    const int64_t nnz = off[s->n_src];
    for (int64_t e = 0; e < nnz; e++)
        if (cols[e] < 0 || cols[e] >= s->n_src) return cb2_fail(CB2_ERR_VALUE, "Laplacian column index out of range");
    CB2_CUDA(cudaMalloc((void**)&s->lap.offset, (s->n_src + 1) * sizeof(int64_t)));
    CB2_CUDA(cudaMalloc((void**)&s->lap.index, (nnz ? nnz : 1) * sizeof(int32_t)));
    CB2_CUDA(cudaMalloc(&s->lap.value, (nnz ? nnz : 1) * sizeof(double)));
    CB2_CUDA(cudaMemcpy(s->lap.offset, off, (s->n_src + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CB2_CUDA(cudaMemcpy(s->lap.index, cols, nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
    CB2_CUDA(cudaMemcpy(s->lap.value, vals, nnz * sizeof(double), cudaMemcpyHostToDevice));
    s->lap_nnz = nnz;
    return CB2_OK;
}

// CSR (device, fp64 values in `val64`) -> stored CSR + CSC in the solver's value type, W_+j and 1/W_i+
template <typename VT>
int build_operands(cb2_sart* s, const double* val64) {
    const int64_t nnz = s->nnz, n_alloc = nnz ? nnz : 1;
    cudaStream_t st = s->stream;
    CB2_CUDA(cudaMalloc(&s->csr.value, n_alloc * sizeof(VT)));
    CB2_CUDA(cudaMalloc((void**)&s->csc.offset, (s->n_src + 1) * sizeof(int64_t)));
    CB2_CUDA(cudaMalloc((void**)&s->csc.index, n_alloc * sizeof(int32_t)));
    CB2_CUDA(cudaMalloc(&s->csc.value, n_alloc * sizeof(VT)));
    CB2_CUDA(cudaMalloc((void**)&s->density, (s->n_src ? s->n_src : 1) * sizeof(double)));
    CB2_CUDA(cudaMalloc((void**)&s->inv_length, (s->n_det ? s->n_det : 1) * sizeof(double)));
    const int blocks = s->sm_count * 8;
    convert_values_kernel<VT><<<blocks, 256, 0, st>>>(nnz, val64, (VT*)s->csr.value);
    // transpose: a stable radix sort of (column key, entry id) keeps the rows of every column in ascending order
    int32_t *rows = nullptr, *iota = nullptr, *keys_out = nullptr, *perm = nullptr;
    void* tmp = nullptr;
    int rc = CB2_OK;
    do {
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&rows, n_alloc * sizeof(int32_t)), "cudaMalloc(rows)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&iota, n_alloc * sizeof(int32_t)), "cudaMalloc(iota)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&keys_out, n_alloc * sizeof(int32_t)), "cudaMalloc(keys)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&perm, n_alloc * sizeof(int32_t)), "cudaMalloc(perm)")) != CB2_OK) break;
        expand_rows_kernel<<<blocks, 256, 0, st>>>(s->n_det, s->csr.offset, rows, iota);
        int end_bit = 1;
        while (end_bit < 31 && ((int64_t)1 << end_bit) < s->n_src) end_bit++;
        size_t tmp_bytes = 0;
        if ((rc = cb2_cuda_check(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s->csr.index, keys_out, iota, perm, nnz, 0, end_bit, st),
                                 "cub::SortPairs(size)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1), "cudaMalloc(sort scratch)")) != CB2_OK) break;
        if (nnz > 0 && (rc = cb2_cuda_check(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, s->csr.index, keys_out, iota, perm, nnz, 0, end_bit, st),
                                            "cub::SortPairs")) != CB2_OK) break;
        column_offsets_kernel<<<blocks, 256, 0, st>>>(nnz, s->n_src, keys_out, s->csc.offset);
        gather_values_kernel<VT><<<blocks, 256, 0, st>>>(nnz, perm, val64, (VT*)s->csc.value);
        gather_rows_kernel<<<blocks, 256, 0, st>>>(nnz, perm, rows, s->csc.index);
        row_sums_kernel<VT><<<blocks, 256, 0, st>>>(s->n_det, s->csr.offset, s->csr.value, s->inv_length, 1);
        row_sums_kernel<VT><<<blocks, 256, 0, st>>>(s->n_src, s->csc.offset, s->csc.value, s->density, 0);
        if ((rc = cb2_cuda_check(cudaGetLastError(), "SART operand kernels")) != CB2_OK) break;
        rc = cb2_cuda_check(cudaStreamSynchronize(st), "SART operand build");
    } while (0);
    cudaFree(rows); cudaFree(iota); cudaFree(keys_out); cudaFree(perm); cudaFree(tmp);
    return rc;
}

int create_impl(cb2_sart* s, const cb2_sart_desc* d) {
    cudaDeviceProp prop;
    CB2_CUDA(cudaGetDeviceProperties(&prop, s->device));
    s->sm_count = prop.multiProcessorCount;
    CB2_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CB2_CUDA(cudaEventCreate(&s->ev0));
    CB2_CUDA(cudaEventCreate(&s->ev1));
    std::vector<int64_t> h_off; std::vector<int32_t> h_idx; std::vector<double> h_val;
    const int64_t* off = d->row_offset; const int32_t* cols = d->columns; const double* vals = d->values;
    const bool on_device = d->memory == 1 && !d->dense;
    if (d->dense) {
        if (d->dense_f64) dense_to_csr<double>((const double*)d->dense, s->n_det, s->n_src, h_off, h_idx, h_val);
        else dense_to_csr<float>((const float*)d->dense, s->n_det, s->n_src, h_off, h_idx, h_val);
        off = h_off.data(); cols = h_idx.data(); vals = h_val.data();
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (on_device) CB2_CUDA(cudaMemcpy(&s->nnz, off + s->n_det, sizeof(int64_t), cudaMemcpyDeviceToHost));
    else s->nnz = off[s->n_det];
    if (s->nnz < 0 || s->nnz > (int64_t)0x7fffffff) return cb2_fail(CB2_ERR_VALUE, "geometry matrix with %lld stored entries is not supported (limit 2^31-1)", (long long)s->nnz);
    if (!on_device)
        for (int64_t e = 0; e < s->nnz; e++)
            if (cols[e] < 0 || cols[e] >= s->n_src) return cb2_fail(CB2_ERR_VALUE, "geometry matrix column index out of range");
    const int64_t n_alloc = s->nnz ? s->nnz : 1;
    double* val64 = nullptr;
    CB2_CUDA(cudaMalloc((void**)&s->csr.offset, (s->n_det + 1) * sizeof(int64_t)));
    CB2_CUDA(cudaMalloc((void**)&s->csr.index, n_alloc * sizeof(int32_t)));
    CB2_CUDA(cudaMalloc((void**)&val64, n_alloc * sizeof(double)));
    int rc = CB2_OK;
    do {
        if ((rc = cb2_cuda_check(cudaMemcpy(s->csr.offset, off, (s->n_det + 1) * sizeof(int64_t), kind), "cudaMemcpy(row_offset)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMemcpy(s->csr.index, cols, s->nnz * sizeof(int32_t), kind), "cudaMemcpy(columns)")) != CB2_OK) break;
        if ((rc = cb2_cuda_check(cudaMemcpy(val64, vals, s->nnz * sizeof(double), kind), "cudaMemcpy(values)")) != CB2_OK) break;
        rc = s->value_f64 ? build_operands<double>(s, val64) : build_operands<float>(s, val64);
    } while (0);
    cudaFree(val64);
    if (rc != CB2_OK) return rc;
    return upload_laplacian(s, d->laplacian_dense, d->lap_row_offset, d->lap_columns, d->lap_values);
}

}  // namespace

extern "C" int cb2_sart_destroy(cb2_sart* s) {
    if (!s) return CB2_OK;
    cudaSetDevice(s->device);
    free_matrix(s->csr); free_matrix(s->csc); free_matrix(s->lap);
    cudaFree(s->density); cudaFree(s->inv_length);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return CB2_OK;
}

extern "C" int cb2_sart_create(const cb2_sart_desc* d, int device, cb2_sart** out) {
    if (!d || !out) return cb2_fail(CB2_ERR_VALUE, "null argument");
    *out = nullptr;
    if (d->abi_version != CB2_ABI_VERSION) return cb2_fail(CB2_ERR_VALUE, "abi_version mismatch");
    if (d->n_detectors < 1 || d->n_sources < 1) return cb2_fail(CB2_ERR_VALUE, "geometry matrix must have shape (N_d, N_s) with N_d, N_s > 0");
    if (d->n_sources > (int64_t)0x7fffffff || d->n_detectors > (int64_t)0x7fffffff) return cb2_fail(CB2_ERR_VALUE, "geometry matrix dimensions exceed int32");
    if (!d->dense && !(d->row_offset && d->columns && d->values)) return cb2_fail(CB2_ERR_VALUE, "geometry matrix missing: give `dense` or the CSR arrays");
    if (d->dense && d->memory == 1) return cb2_fail(CB2_ERR_VALUE, "a dense geometry matrix must be in host memory");
    int ndev = cb2_device_count();
    if (ndev <= 0) return ndev < 0 ? CB2_ERR_CUDA : cb2_fail(CB2_ERR_CUDA, "no CUDA device visible (libcherab_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return cb2_fail(CB2_ERR_VALUE, "device %d out of range (%d visible)", device, ndev);
    CB2_CUDA(cudaSetDevice(device));
    cb2_sart* s = new (std::nothrow) cb2_sart();
    if (!s) return cb2_fail(CB2_ERR_MEMORY, "out of host memory");
    s->device = device;
    s->value_f64 = d->value_f64 != 0;
    s->n_det = d->n_detectors;
    s->n_src = d->n_sources;
    int rc = create_impl(s, d);
    if (rc != CB2_OK) { cb2_sart_destroy(s); return rc; }
    *out = s;
    return CB2_OK;
}

extern "C" int cb2_sart_set_laplacian(cb2_sart* s, const double* dense, const int64_t* row_offset, const int32_t* columns, const double* values) {
    if (!s) return cb2_fail(CB2_ERR_VALUE, "null argument");
    CB2_CUDA(cudaSetDevice(s->device));
    return upload_laplacian(s, dense, row_offset, columns, values);
}

extern "C" double cb2_sart_info(const cb2_sart* s, int what) {
    if (!s) return -1.0;
    const double vb = s->value_f64 ? 8.0 : 4.0;
    switch (what) {
        case 0: return (double)s->nnz;
        case 1: return (double)s->lap_nnz;
        case 2: return 2.0 * (double)s->nnz * (vb + 4.0) + 8.0 * (double)(s->n_det + s->n_src + 2);
        case 3: return s->last_ms;
        case 4: return (double)s->last_iterations;
    }
    return -1.0;
}

namespace {

template <typename VT>
int solve_group(cb2_sart* s, const double* meas, int64_t n_frames, int64_t f0, const double* guess, double initial_value,
                int max_iterations, double relaxation, double beta, double conv_tol, double* solution, double* convergence,
                int32_t* n_iterations) {
    constexpr int FR = SART_FR;
    const int nf = (int)((n_frames - f0) < FR ? (n_frames - f0) : FR);
    const int64_t nd = s->n_det, ns = s->n_src;
    cudaStream_t st = s->stream;
    const int g_fwd = grid_for(s, nd), g_bwd = grid_for(s, ns);
    const bool penal = s->lap.offset != nullptr;
    // host staging, frames innermost
    std::vector<double> h_m((size_t)nd * FR, 0.0), h_x((size_t)ns * FR, 0.0), h_msq(FR, 1.0);
    std::vector<int32_t> h_stop(FR, 1);
    for (int f = 0; f < nf; f++) {
        const double* mv = meas + (f0 + f) * nd;
        double q = 0.0;
        for (int64_t i = 0; i < nd; i++) { h_m[(size_t)i * FR + f] = mv[i]; q += mv[i] * mv[i]; }
        h_msq[f] = q;
        for (int64_t j = 0; j < ns; j++) h_x[(size_t)j * FR + f] = guess ? guess[(f0 + f) * ns + j] : initial_value;
        h_stop[f] = 0;
    }
    SartFrames fr{};
    int32_t h_flags[3] = {0, 0, nf};
    int rc = CB2_OK;
    do {
#define SART_TRY(call) if ((rc = cb2_cuda_check((call), #call)) != CB2_OK) break
        SART_TRY(cudaMalloc((void**)&fr.x, (size_t)ns * FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.gp, (size_t)ns * FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.w, (size_t)nd * FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.m, (size_t)nd * FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.partial, (size_t)g_fwd * FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.conv, (size_t)FR * max_iterations * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.m_sq, FR * sizeof(double)));
        SART_TRY(cudaMalloc((void**)&fr.stopped, FR * sizeof(int32_t)));
        SART_TRY(cudaMalloc((void**)&fr.flags, 3 * sizeof(int32_t)));
        SART_TRY(cudaMemcpyAsync(fr.x, h_x.data(), h_x.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.m, h_m.data(), h_m.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.m_sq, h_msq.data(), FR * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.stopped, h_stop.data(), FR * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemsetAsync(fr.conv, 0, (size_t)FR * max_iterations * sizeof(double), st));
        SART_TRY(cudaMemsetAsync(fr.gp, 0, (size_t)ns * FR * sizeof(double), st));
        SART_TRY(cudaEventRecord(s->ev0, st));
        sart_forward_kernel<VT, FR><<<g_fwd, SART_THREADS, 0, st>>>(nd, s->csr.offset, s->csr.index, s->csr.value, s->inv_length, fr, 0,
                                                                    max_iterations, conv_tol);
        int launched = 0;
        while (launched < max_iterations) {
            const int chunk = (max_iterations - launched) < SART_CHECK_EVERY ? (max_iterations - launched) : SART_CHECK_EVERY;
            for (int c = 0; c < chunk; c++) {
                if (penal) sart_penalty_kernel<FR><<<g_bwd, SART_THREADS, 0, st>>>(ns, s->lap.offset, s->lap.index, s->lap.value, fr, beta);
                sart_backward_kernel<VT, FR><<<g_bwd, SART_THREADS, 0, st>>>(ns, s->csc.offset, s->csc.index, s->csc.value, s->density, fr,
                                                                             relaxation, penal ? 1 : 0);
                sart_forward_kernel<VT, FR><<<g_fwd, SART_THREADS, 0, st>>>(nd, s->csr.offset, s->csr.index, s->csr.value, s->inv_length, fr, 1,
                                                                            max_iterations, conv_tol);
            }
            launched += chunk;
            SART_TRY(cudaGetLastError());
            SART_TRY(cudaMemcpyAsync(h_flags, fr.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
            SART_TRY(cudaStreamSynchronize(st));
            if (h_flags[2] == 0) break;
        }
        if (rc != CB2_OK) break;
        SART_TRY(cudaEventRecord(s->ev1, st));
        SART_TRY(cudaMemcpyAsync(h_x.data(), fr.x, h_x.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        SART_TRY(cudaMemcpyAsync(h_stop.data(), fr.stopped, FR * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        std::vector<double> h_conv((size_t)FR * max_iterations);
        SART_TRY(cudaMemcpyAsync(h_conv.data(), fr.conv, h_conv.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        SART_TRY(cudaStreamSynchronize(st));
        float ms = 0.f;
        SART_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->last_ms += ms;
        s->last_iterations = launched > s->last_iterations ? launched : s->last_iterations;
        for (int f = 0; f < nf; f++) {
            for (int64_t j = 0; j < ns; j++) solution[(f0 + f) * ns + j] = h_x[(size_t)j * FR + f];
            const int nit = h_stop[f];
            if (n_iterations) n_iterations[f0 + f] = nit;
            if (convergence) memcpy(convergence + (f0 + f) * max_iterations, h_conv.data() + (size_t)f * max_iterations, (size_t)nit * sizeof(double));
        }
#undef SART_TRY
    } while (0);
    cudaFree(fr.x); cudaFree(fr.gp); cudaFree(fr.w); cudaFree(fr.m); cudaFree(fr.partial); cudaFree(fr.conv); cudaFree(fr.m_sq);
    cudaFree(fr.stopped); cudaFree(fr.flags);
    return rc;
}

}  // namespace

extern "C" int cb2_sart_solve(cb2_sart* s, const double* measurements, int64_t n_frames, const double* initial_guess,
                              double initial_value, int max_iterations, double relaxation, double beta_laplace, double conv_tol,
                              double* solution, double* convergence, int32_t* n_iterations) {
    if (!s || !measurements || !solution) return cb2_fail(CB2_ERR_VALUE, "null argument");
    if (n_frames < 1) return cb2_fail(CB2_ERR_VALUE, "n_frames must be >= 1");
    if (max_iterations < 1) return cb2_fail(CB2_ERR_VALUE, "max_iterations must be >= 1");
    CB2_CUDA(cudaSetDevice(s->device));
    s->last_ms = 0.0;
    s->last_iterations = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += SART_FR) {
        int rc = s->value_f64 ? solve_group<double>(s, measurements, n_frames, f0, initial_guess, initial_value, max_iterations, relaxation,
                                                    beta_laplace, conv_tol, solution, convergence, n_iterations)
                              : solve_group<float>(s, measurements, n_frames, f0, initial_guess, initial_value, max_iterations, relaxation,
                                                   beta_laplace, conv_tol, solution, convergence, n_iterations);
        if (rc != CB2_OK) return rc;
    }
    return CB2_OK;
}
