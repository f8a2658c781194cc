// cb2_contract_tc.cu — K2 on the 5th-generation tensor cores: the Bremsstrahlung contraction
//     out[ray][bin] += scale * sum_k mom[ray][k] phi[k][bin]        (bremsstrahlung.pyx:199-206 in the moment formulation)
// as a hand-written tcgen05 kernel for sm_100a.  float32 accuracy is required (1e-4 parity over ~1e3-term sums whose
// Lagrange weights cancel), so the product is the error-compensated 3xTF32 form  A_hi B_hi + A_lo B_hi + A_hi B_lo  with
// X_hi = tf32(X), X_lo = tf32(X - X_hi): every operand is exactly representable in TF32, the tensor cores accumulate in
// float32 in TMEM, the dropped A_lo B_lo term is 2^-22 relative.
//
// Shape of the kernel (one launch per ray batch, persistent):
//   * CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256 x 256 output tile per pair and step — UMMA M = 256 (128 rays
//     per CTA), N = 256 (each CTA stages 128 of the 256 phi columns), K = 8 per instruction;
//   * operands are K-major 128-byte-swizzled tiles of 32 k x 128 rows, loaded by TMA (cp.async.bulk.tensor.2d, 4 tiles per
//     CTA and stage: A_hi, A_lo, B_hi, B_lo = 64 KB), 3 stages; the peer CTA's loads complete on the leader's mbarrier;
//   * warp 0 = TMA producer (one lane), warp 1 = MMA issuer (leader CTA, one lane: 12 tcgen05.mma per stage — 4 k-steps x
//     the three hi/lo products into the same accumulator), warp 2 = TMEM allocator, warps 4..7 = epilogue;
//   * two 256-column accumulator stages in TMEM (all 512 columns): the epilogue of tile i (tcgen05.ld -> registers ->
//     read-modify-write of the frame rows, float32 or float64, scale applied) overlaps the MMAs of tile i + 1;
//   * phi is transposed and split once per scene ([bins_pad][k_pad] hi / lo, 20 MB on C3: L2-resident), the moment rows are
//     split per batch by split_tf32_kernel.
// Algorithmic work 2 rays k_pad bins_pad flop; issued tensor work 3x that.  Operand traffic per pair and stage 128 KB for
// 12.6 MFLOP (98 flop/B): L2 -> SM bandwidth, not HBM, is the companion bound.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cb2_internal.h"

namespace {

constexpr int TILE_ROWS = 128;               // rows of one operand tile (rays per CTA, phi columns per CTA)
constexpr int TILE_K = 32;                   // tf32 elements per tile row: 128 bytes, one swizzle atom
constexpr int TILE_BYTES = TILE_ROWS * TILE_K * 4;   // 16 KB
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 4 * TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo
constexpr int UMMA_M = 256, UMMA_N = 256, UMMA_K = 8;
constexpr int ACC_COLS = 256, TMEM_COLS = 512;
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format F32 (1) at bit 4, a/b format TF32 (2) at
// bits 7 / 10, both K-major, n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UMMA_N >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
        : "memory");
}
// a phase that does not complete within ~2 s is a protocol error: trap instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0;; spins++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 26)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// 2-D tile load into this CTA's shared memory; the bytes complete on the barrier at `bar`'s offset in the LEADER CTA
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b32 rb;\n\t"
        "mapa.shared::cluster.u32 rb, %2, 0;\n\t"
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [rb];\n\t}" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// shared-memory matrix descriptor (SmemDescriptor): K-major tile, 128-byte swizzle — start address >> 4, stride between the
// 8-row groups 1024 B (>> 4 = 64) at bit 32, version 1 at bit 46, layout SWIZZLE_128B (2) at bit 61; the leading-dimension
// offset is not used by swizzled K-major layouts
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
// all MMAs issued so far by this thread arrive (once) on the barrier at this offset in both CTAs of the pair when they retire
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

struct TcParams {
    int64_t n_rays;
    int k_pad, n_pad, bins, out_f64;
    float scale;
    void* out;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;                   // swizzle-128B tiles want 1024-byte alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int m_tiles = (int)((P.n_rays + UMMA_M - 1) / UMMA_M), n_tiles = (P.n_pad + UMMA_N - 1) / UMMA_N;
    const int n_work = m_tiles * n_tiles, pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const int k_blocks = (P.k_pad + TILE_K - 1) / TILE_K;

    if (warp == 0) {
        // ===== TMA producer: both CTAs stage their halves (rays m0 + 128 rank .., phi columns n0 + 128 rank ..) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = pair; w < n_work; w += pairs) {
                const int m0 = (w / n_tiles) * UMMA_M + (int)rank * TILE_ROWS, n0 = (w % n_tiles) * UMMA_N + (int)rank * TILE_ROWS;
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                    const uint32_t dst = base + stage * STAGE_BYTES;
                    tma_load_2d(dst, &map_a_hi, kb * TILE_K, m0, full_bar(stage));
                    tma_load_2d(dst + TILE_BYTES, &map_a_lo, kb * TILE_K, m0, full_bar(stage));
                    tma_load_2d(dst + 2 * TILE_BYTES, &map_b_hi, kb * TILE_K, n0, full_bar(stage));
                    tma_load_2d(dst + 3 * TILE_BYTES, &map_b_lo, kb * TILE_K, n0, full_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one lane of the leader CTA drives the tensor cores of both SMs =====
        if (leader && lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int w = pair; w < n_work; w += pairs) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);                  // both epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + stage * STAGE_BYTES;
                    const uint64_t a_hi = umma_desc(st), a_lo = umma_desc(st + TILE_BYTES), b_hi = umma_desc(st + 2 * TILE_BYTES),
                                   b_lo = umma_desc(st + 3 * TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TILE_K / UMMA_K; kk++) {
                        const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);   // 32 bytes along K inside the swizzle atom
                        umma_tf32_2sm(tmem_d, a_hi + adv, b_hi + adv, (kb | kk) != 0);
                        umma_tf32_2sm(tmem_d, a_lo + adv, b_hi + adv, 1);
                        umma_tf32_2sm(tmem_d, a_hi + adv, b_lo + adv, 1);
                    }
                    umma_commit_2sm(empty_bar(stage));                      // frees the stage in both CTAs when these MMAs retire
                    if (kb == k_blocks - 1) umma_commit_2sm(tfull_bar(acc));  // accumulator complete: both epilogues may read
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> frame rows (read-modify-write); warp q owns TMEM lanes 32 q .. 32 q + 31 =====
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool vec4 = (P.bins & 3) == 0 && !P.out_f64;
        for (int w = pair; w < n_work; w += pairs) {
            const int64_t row = (int64_t)(w / n_tiles) * UMMA_M + rank * TILE_ROWS + q * 32 + lane;
            const int n0 = (w % n_tiles) * UMMA_N;
            mbar_wait(tfull_bar(acc), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
            const bool row_ok = row < P.n_rays;
            for (int c = 0; c < UMMA_N; c += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int col = n0 + c;
                if (!row_ok || col >= P.bins) continue;
                if (vec4 && col + 16 <= P.bins) {
                    float4* p = reinterpret_cast<float4*>((float*)P.out + row * P.bins + col);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float4 o = p[j];
                        o.x = fmaf(P.scale, __uint_as_float(v[4 * j]), o.x); o.y = fmaf(P.scale, __uint_as_float(v[4 * j + 1]), o.y);
                        o.z = fmaf(P.scale, __uint_as_float(v[4 * j + 2]), o.z); o.w = fmaf(P.scale, __uint_as_float(v[4 * j + 3]), o.w);
                        p[j] = o;
                    }
                } else if (P.out_f64) {
                    double* p = (double*)P.out + row * P.bins + col;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (col + j < P.bins) p[j] += (double)P.scale * (double)__uint_as_float(v[j]);
                } else {
                    float* p = (float*)P.out + row * P.bins + col;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (col + j < P.bins) p[j] = fmaf(P.scale, __uint_as_float(v[j]), p[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);         // the MMA issuer waits on the leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    __syncwarp();                                                       // the single-lane roles rejoin their warps
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// phi [k_pad][n_pad] -> transposed tf32 hi / lo parts [n_pad][k_pad] (K-major B operand)
__global__ void transpose_split_kernel(const float* __restrict__ phi, int k_pad, int n_pad, float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j, n = n0 + threadIdx.x;
        tile[j][threadIdx.x] = (k < k_pad && n < n_pad) ? phi[(size_t)k * n_pad + n] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int n = n0 + j, k = k0 + threadIdx.x;
        if (n >= n_pad || k >= k_pad) continue;
        const float v = tile[threadIdx.x][j];
        unsigned h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hf));
        hi[(size_t)n * k_pad + k] = hf;
        lo[(size_t)n * k_pad + k] = __uint_as_float(l);
    }
}

typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_t g_encode = nullptr;
bool g_encode_tried = false;

bool load_encode() {
    if (g_encode_tried) return g_encode != nullptr;
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    g_encode = (encode_tiled_t)fn;
    return true;
}

// [rows][k_pad] float32, boxes of 32 k x 128 rows, 128-byte swizzle, zero fill outside
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int k_pad) {
    const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k_pad * sizeof(float)};
    const cuuint32_t box[2] = {TILE_K, TILE_ROWS}, estr[2] = {1, 1};
    const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cb2_fail(CB2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CB2_OK;
}

}  // namespace

// hi = tf32(x) (round to nearest, ties away: cvt.rna), lo = tf32(x - hi); both stored as float32 bit patterns
__global__ void split_tf32_kernel(int64_t n, const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        unsigned h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hf));
        hi[i] = hf;
        lo[i] = __uint_as_float(l);
    }
}

static int grow(float** p, size_t* have, size_t need) {
    if (need <= *have) return CB2_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    CB2_CUDA(cudaMalloc((void**)p, need));
    *have = need;
    return CB2_OK;
}

static int launch_tc(cb2_scene* sc, const float* mom_hi, const float* mom_lo, int64_t n_rays, int64_t rows_alloc, int k_pad, int n_pad, int bins,
                     void* out, int out_f64, double scale, cudaStream_t st) {
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_map(&ma_hi, mom_hi, rows_alloc, k_pad)) != CB2_OK) return rc;
    if ((rc = make_map(&ma_lo, mom_lo, rows_alloc, k_pad)) != CB2_OK) return rc;
    if ((rc = make_map(&mb_hi, sc->phi_hi, n_pad, k_pad)) != CB2_OK) return rc;
    if ((rc = make_map(&mb_lo, sc->phi_lo, n_pad, k_pad)) != CB2_OK) return rc;
    TcParams P;
    P.n_rays = n_rays; P.k_pad = k_pad; P.n_pad = n_pad; P.bins = bins; P.out_f64 = out_f64; P.scale = (float)scale; P.out = out;
    const int m_tiles = (int)((n_rays + UMMA_M - 1) / UMMA_M), n_tiles = (n_pad + UMMA_N - 1) / UMMA_N;
    int pairs = std::min(sc->tc_pairs, m_tiles * n_tiles);
    if (pairs < 1) pairs = 1;
    contract_tc_kernel<<<dim3(2 * pairs), dim3(THREADS), SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, P);
    return cb2_cuda_check(cudaGetLastError(), "contract_tc_kernel launch");
}

// called once per scene that uses the moment formulation; leaves sc->contract_tc = 0 when the tensor path is unavailable
// (CB2_CONTRACT=ffma forces the FFMA tile kernel of cb2_contract.cu).  The kernel is checked against the FFMA kernel on a
// small random problem before it is trusted.
int cb2_contract_tc_init(cb2_scene* sc, const float* phi, int k_pad, int n_pad) {
    sc->contract_tc = 0;
    const char* force = getenv("CB2_CONTRACT");
    if (force && strcmp(force, "ffma") == 0) return CB2_OK;
    if (!load_encode()) return CB2_OK;
    cudaDeviceProp prop;
    CB2_CUDA(cudaGetDeviceProperties(&prop, sc->device));
    if (prop.major != 10) return CB2_OK;                            // tcgen05 is sm_100-family only
    if ((k_pad & 3) != 0) return CB2_OK;                            // TMA: row pitch must be a multiple of 16 bytes
    sc->tc_pairs = prop.multiProcessorCount / 2;
    CB2_CUDA(cudaFuncSetAttribute(contract_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const size_t n = (size_t)k_pad * n_pad;
    CB2_CUDA(cudaMalloc((void**)&sc->phi_hi, n * sizeof(float)));
    CB2_CUDA(cudaMalloc((void**)&sc->phi_lo, n * sizeof(float)));
    transpose_split_kernel<<<dim3((n_pad + 31) / 32, (k_pad + 31) / 32), dim3(32, 8)>>>(phi, k_pad, n_pad, sc->phi_hi, sc->phi_lo);
    CB2_CUDA(cudaGetLastError());
    CB2_CUDA(cudaDeviceSynchronize());
    sc->contract_tc = 1;
    return CB2_OK;
}

void cb2_contract_tc_destroy(cb2_scene* sc) {
    cudaFree(sc->phi_hi); cudaFree(sc->phi_lo); cudaFree(sc->mom_split); cudaFree(sc->tmp32);
    sc->phi_hi = sc->phi_lo = sc->mom_split = sc->tmp32 = nullptr;
    sc->mom_split_bytes = sc->tmp32_bytes = 0;
}

int cb2_launch_contract_tc(cb2_scene* sc, const float* mom, int64_t n_rays, int k_pad, int n_pad, int bins, void* out, int out_f64,
                           double scale, cudaStream_t st) {
    if (n_rays <= 0) return CB2_OK;
    if (n_rays > 0x7fffffffLL) return cb2_fail(CB2_ERR_VALUE, "too many rays for one contraction launch");
    const int64_t n_mom = n_rays * (int64_t)k_pad;
    int rc = grow(&sc->mom_split, &sc->mom_split_bytes, (size_t)2 * n_mom * sizeof(float));
    if (rc != CB2_OK) return rc;
    float* mom_hi = sc->mom_split;
    float* mom_lo = sc->mom_split + n_mom;
    split_tf32_kernel<<<1184, 256, 0, st>>>(n_mom, mom, mom_hi, mom_lo);
    return launch_tc(sc, mom_hi, mom_lo, n_rays, n_rays, k_pad, n_pad, bins, out, out_f64, scale, st);
}
