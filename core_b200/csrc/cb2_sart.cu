// SART inversion on the device-resident geometry matrix.
//
// Replaces cherab/tools/inversions/sart.pyx:26-155 (invert_sart), :161-302 (invert_constrained_sart) and the OpenCL
// solver cherab/tools/inversions/opencl/sart_opencl.py:33-318.  Not a translation of either: the reference walks a dense
// matrix cell by cell; here the matrix is sparse (what the ray-transfer kernel produces), stored once as CSR and once as
// CSC, and one iteration is two HBM streams over it,
//     backward (CSC):  x_j <- max(0, x_j + omega/W_+j * sum_i W_ij w_i - beta (L x)_j)
//     forward  (CSR):  y_i = sum_j W_ij x_j ;  w_i = (m_i - y_i) / W_i+ ;  sum_i y_i^2
// for up to four measurement frames at a time, so the bytes streamed per iteration do not grow with the number of frames.
// Each pass is sart_dot_kernel (one warp per chunk of <= 2048 stored entries: a pinhole's own cell is crossed by every ray
// of the frame, so whole rows / columns cannot be the unit of work) followed by a small finishing kernel that adds the chunk
// sums of every row in order.  The stop test of the reference (sart.pyx:147-153) runs on the device in the last CTA of the
// forward finishing kernel; once a frame has stopped, later launches leave its solution untouched, so the host only reads
// the flags back every few iterations.  All sums have a fixed order (lane-strided entries, warp butterflies, chunk sums in
// chunk order, per-CTA partials summed by one warp): results are reproducible and do not depend on how frames are grouped.
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
constexpr int SART_CHUNK = 2048;               // stored entries per warp task: long rows / columns are cut into chunks
constexpr int SART_FIN_GRID = 296;             // fixed grid of the finishing kernels => fixed summation order of |y_hat|^2

// One sparse operand (CSR rows of W, CSC columns of W): rows are cut into chunks of at most SART_CHUNK entries so
// that one warp task is bounded whatever the row length (the cell holding a pinhole is crossed by every ray of the frame).
struct SartMatrix {
    int64_t* offset = nullptr;                 // [n + 1]
    int32_t* index = nullptr;                  // [nnz]
    void* value = nullptr;                     // [nnz] float or double
    int64_t* chunk_off = nullptr;              // [n + 1] first chunk of every row
    int32_t* chunk_row = nullptr;              // [n_chunks]
    int64_t n_chunks = 0;
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

// part[c][f] = sum over the entries of chunk c of value * vec[index][f]; lanes stride the entries with 4 loads in flight,
// then a butterfly: the order of the additions depends on nothing but the chunk.
template <typename VT, int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_dot_kernel(SartMatrix mat, const double* __restrict__ vec, double* __restrict__ part,
                                                                const int32_t* __restrict__ flags) {
    if (flags[2] == 0) return;                 // every frame has stopped
    const int32_t* __restrict__ index = mat.index;
    const void* __restrict__ value = mat.value;
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * SART_WARPS;
    for (int64_t c = (int64_t)blockIdx.x * SART_WARPS + (threadIdx.x >> 5); c < mat.n_chunks; c += n_warps) {
        const int32_t r = __ldg(mat.chunk_row + c);
        const int64_t e0 = __ldg(mat.offset + r) + (c - __ldg(mat.chunk_off + r)) * SART_CHUNK;
        const int64_t row_end = __ldg(mat.offset + r + 1);
        const int64_t e1 = e0 + SART_CHUNK < row_end ? e0 + SART_CHUNK : row_end;
        double acc[FR];
#pragma unroll
        for (int f = 0; f < FR; f++) acc[f] = 0.0;
        int64_t e = e0 + lane;
        for (; e + 96 < e1; e += 128) {
            int32_t ci[4];
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                ci[u] = __ldg(index + e + 32 * u);
                v[u] = load_value<VT>(value, e + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int f = 0; f < FR; f++) acc[f] = fma(v[u], vec[(int64_t)ci[u] * FR + f], acc[f]);
        }
        for (; e < e1; e += 32) {
            const int32_t ci = __ldg(index + e);
            const double v = load_value<VT>(value, e);
#pragma unroll
            for (int f = 0; f < FR; f++) acc[f] = fma(v, vec[(int64_t)ci * FR + f], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < FR; f++) acc[f] = warp_sum(acc[f]);
        if (lane < FR) {
            double a = acc[0];
#pragma unroll
            for (int f = 1; f < FR; f++) a = lane == f ? acc[f] : a;
            part[c * FR + lane] = a;
        }
    }
}

// per-frame solver state on the device
struct SartFrames {
    double* x;            // [n_sources][FR]  current estimate
    double* x_new;        // [n_sources][FR]  next estimate (swapped with x after every backward pass)
    double* w;            // [n_detectors][FR] (m - y_hat) / ray_length
    double* m;            // [n_detectors][FR] measurements
    double* part;         // [max chunks][FR] chunk sums of the running pass
    double* partial;      // [SART_FIN_GRID][FR] per-CTA sums of y_hat^2
    double* conv;         // [FR][max_iterations]
    double* m_sq;         // [FR]
    int32_t* stopped;     // [FR] 0 while iterating, else the iteration count at which the frame stopped
    int32_t* flags;       // [0] ticket, [1] iteration index k of the running pass, [2] number of frames still iterating
};

// x_new_j = max(0, x_j + omega / W_+j * sum_i W_ij w_i - beta (L x)_j)   (sart.pyx:118-142, :255-285); one thread per source
template <int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_update_kernel(int64_t n_sources, const int64_t* __restrict__ chunk_off,
                                                                   const double* __restrict__ density, SartMatrix lap, SartFrames fr,
                                                                   double relaxation, double beta) {
    if (fr.flags[2] == 0) return;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_sources; j += (int64_t)gridDim.x * blockDim.x) {
        const double dens = density[j];
        double a[FR], pen[FR];
#pragma unroll
        for (int f = 0; f < FR; f++) { a[f] = 0.0; pen[f] = 0.0; }
        for (int64_t c = chunk_off[j]; c < chunk_off[j + 1]; c++)
#pragma unroll
            for (int f = 0; f < FR; f++) a[f] += fr.part[c * FR + f];
        if (lap.offset) {
            const double* lv = reinterpret_cast<const double*>(lap.value);
            for (int64_t e = lap.offset[j]; e < lap.offset[j + 1]; e++) {
                const double l = lv[e];
                const int64_t col = lap.index[e];
#pragma unroll
                for (int f = 0; f < FR; f++) pen[f] = fma(l, fr.x[col * FR + f], pen[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < FR; f++) {
            const double xo = fr.x[j * FR + f];
            double xn = xo;
            if (dens > 0.0) xn += relaxation / dens * a[f];
            if (lap.offset) xn -= beta * pen[f];
            xn = xn < 0.0 ? 0.0 : xn;
            fr.x_new[j * FR + f] = fr.stopped[f] ? xo : xn;     // a stopped frame keeps its solution
        }
    }
}

// y_hat_i from the chunk sums, w_i = (m_i - y_hat_i) / W_i+, |y_hat|^2; the last CTA records the convergence and applies the
// stop test (sart.pyx:96-97, :144-153).  One thread per detector; record == 0: the projection before the first iteration.
template <int FR>
__global__ void __launch_bounds__(SART_THREADS) sart_residual_kernel(int64_t n_detectors, const int64_t* __restrict__ chunk_off,
                                                                     const double* __restrict__ inv_length, SartFrames fr, int record,
                                                                     int max_iterations, double conv_tol) {
    if (fr.flags[2] == 0) return;
    __shared__ double s_sq[SART_WARPS][FR];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sq[FR];
#pragma unroll
    for (int f = 0; f < FR; f++) sq[f] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_detectors; i += (int64_t)gridDim.x * blockDim.x) {
        double y[FR];
#pragma unroll
        for (int f = 0; f < FR; f++) y[f] = 0.0;
        for (int64_t c = chunk_off[i]; c < chunk_off[i + 1]; c++)
#pragma unroll
            for (int f = 0; f < FR; f++) y[f] += fr.part[c * FR + f];
        const double il = inv_length[i];                         // 0 for rays of zero length (sart.pyx:127-128)
#pragma unroll
        for (int f = 0; f < FR; f++) {
            fr.w[i * FR + f] = (fr.m[i * FR + f] - y[f]) * il;
            sq[f] = fma(y[f], y[f], sq[f]);
        }
    }
    if (!record) return;
#pragma unroll
    for (int f = 0; f < FR; f++) {
        const double t = warp_sum(sq[f]);
        if (lane == 0) s_sq[warp][f] = t;
    }
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

// chunk table of a sparse operand
__global__ void chunk_count_kernel(int64_t n, const int64_t* __restrict__ off, int64_t* __restrict__ chunk_off) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t len = r < n ? off[r + 1] - off[r] : 0;
        chunk_off[r] = r < n ? (len > SART_CHUNK ? (len + SART_CHUNK - 1) / SART_CHUNK : 1) : 0;
    }
}
__global__ void chunk_fill_kernel(int64_t n, const int64_t* __restrict__ chunk_off, int32_t* __restrict__ chunk_row) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        for (int64_t c = chunk_off[r]; c < chunk_off[r + 1]; c++) chunk_row[c] = (int32_t)r;
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
    char* work = nullptr;           // the frames' work arrays of solve_group: one grow-only allocation kept by the handle
    size_t work_bytes = 0;
};

namespace {

void free_matrix(SartMatrix& m) {
    cudaFree(m.offset); cudaFree(m.index); cudaFree(m.value); cudaFree(m.chunk_off); cudaFree(m.chunk_row);
    m = SartMatrix();
}

int dot_grid(const cb2_sart* s, int64_t n_chunks) {
    int64_t need = (n_chunks + SART_WARPS - 1) / SART_WARPS;
    int64_t cap = (int64_t)s->sm_count * 32;      // a few waves of 8-warp CTAs; chunk results do not depend on the grid
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

// chunk table (chunk_off, chunk_row) of an operand whose offset array is on the device
int build_chunks(cb2_sart* s, SartMatrix& m, int64_t n) {
    cudaStream_t st = s->stream;
    CB2_CUDA(cudaMalloc((void**)&m.chunk_off, (n + 1) * sizeof(int64_t)));
    const int blocks = s->sm_count * 4;
    chunk_count_kernel<<<blocks, 256, 0, st>>>(n, m.offset, m.chunk_off);
    int rc = cb2_launch_scan(m.chunk_off, n + 1, st);
    if (rc != CB2_OK) return rc;
    CB2_CUDA(cudaMemcpyAsync(&m.n_chunks, m.chunk_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CB2_CUDA(cudaStreamSynchronize(st));
    CB2_CUDA(cudaMalloc((void**)&m.chunk_row, (m.n_chunks ? m.n_chunks : 1) * sizeof(int32_t)));
    chunk_fill_kernel<<<blocks, 256, 0, st>>>(n, m.chunk_off, m.chunk_row);
    return cb2_cuda_check(cudaGetLastError(), "SART chunk table");
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
        if ((rc = build_chunks(s, s->csr, s->n_det)) != CB2_OK) break;
        if ((rc = build_chunks(s, s->csc, s->n_src)) != CB2_OK) break;
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
    cudaFree(s->work);
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

template <typename VT, int FR>
int solve_group(cb2_sart* s, const double* meas, int64_t n_frames, int64_t f0, const double* guess, double initial_value,
                int max_iterations, double relaxation, double beta, double conv_tol, double* solution, double* convergence,
                int32_t* n_iterations) {
    const int nf = (int)((n_frames - f0) < FR ? (n_frames - f0) : FR);
    const int64_t nd = s->n_det, ns = s->n_src;
    cudaStream_t st = s->stream;
    const int g_fwd = dot_grid(s, s->csr.n_chunks), g_bwd = dot_grid(s, s->csc.n_chunks);
    const int g_upd = (int)((ns + SART_THREADS - 1) / SART_THREADS < (int64_t)s->sm_count * 8 ? (ns + SART_THREADS - 1) / SART_THREADS : (int64_t)s->sm_count * 8);
    const int g_res = (int)((nd + SART_THREADS - 1) / SART_THREADS < SART_FIN_GRID ? (nd + SART_THREADS - 1) / SART_THREADS : SART_FIN_GRID);
    const int64_t max_chunks = s->csr.n_chunks > s->csc.n_chunks ? s->csr.n_chunks : s->csc.n_chunks;
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
        // the work arrays live in one allocation owned by the handle (grow-only): a solve allocates nothing after the first
        const size_t sizes[10] = {(size_t)ns * FR * sizeof(double), (size_t)ns * FR * sizeof(double), (size_t)nd * FR * sizeof(double),
                                  (size_t)nd * FR * sizeof(double), (size_t)(max_chunks ? max_chunks : 1) * FR * sizeof(double),
                                  (size_t)SART_FIN_GRID * FR * sizeof(double), (size_t)FR * max_iterations * sizeof(double),
                                  FR * sizeof(double), FR * sizeof(int32_t), 3 * sizeof(int32_t)};
        size_t off[11] = {0};
        for (int k = 0; k < 10; k++) off[k + 1] = off[k] + ((sizes[k] + 255) & ~(size_t)255);
        if (s->work_bytes < off[10]) {
            SART_TRY(cudaStreamSynchronize(st));
            cudaFree(s->work);
            s->work = nullptr; s->work_bytes = 0;
            SART_TRY(cudaMalloc((void**)&s->work, off[10]));
            s->work_bytes = off[10];
        }
        fr.x = (double*)(s->work + off[0]); fr.x_new = (double*)(s->work + off[1]); fr.w = (double*)(s->work + off[2]);
        fr.m = (double*)(s->work + off[3]); fr.part = (double*)(s->work + off[4]); fr.partial = (double*)(s->work + off[5]);
        fr.conv = (double*)(s->work + off[6]); fr.m_sq = (double*)(s->work + off[7]); fr.stopped = (int32_t*)(s->work + off[8]);
        fr.flags = (int32_t*)(s->work + off[9]);
        SART_TRY(cudaMemcpyAsync(fr.x, h_x.data(), h_x.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.m, h_m.data(), h_m.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.m_sq, h_msq.data(), FR * sizeof(double), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.stopped, h_stop.data(), FR * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemcpyAsync(fr.flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, st));
        SART_TRY(cudaMemsetAsync(fr.conv, 0, (size_t)FR * max_iterations * sizeof(double), st));
        SART_TRY(cudaEventRecord(s->ev0, st));
        // y_hat of the initial guess (sart.pyx:96-97)
        sart_dot_kernel<VT, FR><<<g_fwd, SART_THREADS, 0, st>>>(s->csr, fr.x, fr.part, fr.flags);
        sart_residual_kernel<FR><<<g_res, SART_THREADS, 0, st>>>(nd, s->csr.chunk_off, s->inv_length, fr, 0, max_iterations, conv_tol);
        int launched = 0;
        while (launched < max_iterations) {
            const int chunk = (max_iterations - launched) < SART_CHECK_EVERY ? (max_iterations - launched) : SART_CHECK_EVERY;
            for (int c = 0; c < chunk; c++) {
                sart_dot_kernel<VT, FR><<<g_bwd, SART_THREADS, 0, st>>>(s->csc, fr.w, fr.part, fr.flags);
                sart_update_kernel<FR><<<g_upd, SART_THREADS, 0, st>>>(ns, s->csc.chunk_off, s->density, s->lap, fr, relaxation, beta);
                double* t = fr.x; fr.x = fr.x_new; fr.x_new = t;
                sart_dot_kernel<VT, FR><<<g_fwd, SART_THREADS, 0, st>>>(s->csr, fr.x, fr.part, fr.flags);
                sart_residual_kernel<FR><<<g_res, SART_THREADS, 0, st>>>(nd, s->csr.chunk_off, s->inv_length, fr, 1, max_iterations, conv_tol);
            }
            launched += chunk;
            SART_TRY(cudaGetLastError());
            SART_TRY(cudaMemcpyAsync(h_flags, fr.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
            SART_TRY(cudaStreamSynchronize(st));
            if (h_flags[2] == 0) break;
        }
        if (rc != CB2_OK) break;
        SART_TRY(cudaEventRecord(s->ev1, st));
        SART_TRY(cudaMemcpyAsync(h_stop.data(), fr.stopped, FR * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        std::vector<double> h_conv((size_t)FR * max_iterations);
        SART_TRY(cudaMemcpyAsync(h_conv.data(), fr.conv, h_conv.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        SART_TRY(cudaStreamSynchronize(st));
        // the host swapped the buffers once per launched iteration, the device only wrote x_new while a frame was still
        // iterating (h_flags[1] iterations): undo the swaps of the launches that returned at once
        if ((launched - h_flags[1]) & 1) { double* t = fr.x; fr.x = fr.x_new; fr.x_new = t; }
        SART_TRY(cudaMemcpyAsync(h_x.data(), fr.x, h_x.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
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
    return rc;
}

template <typename VT>
int solve_all(cb2_sart* s, const double* meas, int64_t n_frames, const double* guess, double initial_value, int max_iterations,
              double relaxation, double beta, double conv_tol, double* solution, double* convergence, int32_t* n_iterations) {
    int64_t f0 = 0;
    while (f0 < n_frames) {
        // groups of four frames share every pass over the matrix; a last single frame runs with one-wide vectors
        const bool four = n_frames - f0 >= 2;
        int rc = four ? solve_group<VT, 4>(s, meas, n_frames, f0, guess, initial_value, max_iterations, relaxation, beta, conv_tol, solution,
                                           convergence, n_iterations)
                      : solve_group<VT, 1>(s, meas, n_frames, f0, guess, initial_value, max_iterations, relaxation, beta, conv_tol, solution,
                                           convergence, n_iterations);
        if (rc != CB2_OK) return rc;
        f0 += four ? 4 : 1;
    }
    return CB2_OK;
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
    return s->value_f64 ? solve_all<double>(s, measurements, n_frames, initial_guess, initial_value, max_iterations, relaxation, beta_laplace,
                                            conv_tol, solution, convergence, n_iterations)
                        : solve_all<float>(s, measurements, n_frames, initial_guess, initial_value, max_iterations, relaxation, beta_laplace,
                                           conv_tol, solution, convergence, n_iterations);
}
