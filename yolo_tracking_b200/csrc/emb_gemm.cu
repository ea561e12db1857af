// Appearance (cosine) cost of many streams at once, thresholded the way its callers use it:
//     out[b, t, d] = fill                          if gate[b, t, d] != 0
//                  = scale * max(0, 1 - cos(a_t, b_d))   if that value is <= thresh   (exact, fp64 on the fp32 inputs)
//                  = fill                          otherwise.
// Reference: embedding_distance (boxmot/utils/matching.py:145-167) followed by
//   BoT-SORT  bot_sort.py:304-306 / :363-365  (emb / 2; emb[emb > appearance_thresh] = 1; emb[iou mask] = 1)
//   StrongSORT strongsort/sort/linear_assignment.py:59-78 (cost > max_distance -> max_distance + 1e-5)
//   DeepOCSORT deep_ocsort.py:433 (dets_embs @ trk_embs.T).
//
// The cosine matrix is the one dense contraction of the tracking loop (2*T*D*F flop per stream), so it is
// the one place the tensor cores are used: the unit-normalised embeddings are rounded to bf16 and
// multiplied by tcgen05.mma (operands staged by TMA in 128B-swizzled shared memory, fp32 accumulators in
// TMEM).  bf16 can only be a PRE-FILTER - the value that survives the threshold enters an exact assignment
// problem - so the epilogue compares the approximate cosine against the threshold widened by the worst-case
// bf16 error (|cos_bf16 - cos| <= 2^-7 by Cauchy-Schwarz on unit vectors; band 1e-2) and every entry that
// may survive is recomputed in the same kernel, one warp per entry, in fp64 on the fp32 inputs exactly like
// scipy's cdist.  Entries that cannot survive get `fill` without ever being evaluated exactly.
//
// Kernel 1 (unit_bf16_kernel): fp32 rows -> unit-norm bf16 rows (one warp per row, HBM bound).
// Kernel 2 (appearance_cost_kernel): one CTA per (128-track tile, <=256-detection tile, stream):
//   warp 0: TMA producer (2-stage ring of 128x64 + Nx64 bf16 tiles; two CTAs share an SM), warp 1: TMEM allocation +
//   single-thread tcgen05.mma issue, warps 2-5: fill the output tile while the GEMM runs, then tcgen05.ld + threshold
//   test into a candidate bitmask; then all six warps run the fp64 recheck of the candidates.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "../../include/b200track.h"
#include "api_util.h"

namespace b200 {
namespace {

constexpr int BM = 128, BK = 64, STAGES = 2, BN_MAX = 256;       // 2 stages -> 97 KB: two CTAs per SM overlap each other's phases
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN_MAX * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr double BAND_COS = 1e-2;
constexpr uint32_t SPIN_LIMIT = 1u << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error becomes an error flag instead of a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t it = 0; it < SPIN_LIMIT; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- kernel 1: fp32 rows -> unit-norm bf16 rows ----------------------------------------------------
__global__ void __launch_bounds__(256) unit_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        long long rows, int dim) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* s = reinterpret_cast<const float4*>(src + row * dim);
    const int nv = dim >> 2;
    float acc = 0.f;
    for (int i = lane; i < nv; i += 32) { const float4 v = s[i]; acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    const float inv = acc > 0.f ? rsqrtf(acc) : 0.f;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(dst + row * dim);
    for (int i = lane; i < nv; i += 32) {
        const float4 v = s[i];
        o[2 * i] = __floats2bfloat162_rn(v.x * inv, v.y * inv);
        o[2 * i + 1] = __floats2bfloat162_rn(v.z * inv, v.w * inv);
    }
}

// ---- gallery maintenance: append one feature row per track to a device-resident gallery (fp32 row + unit-norm bf16
// row, the operand formats of gallery_cost_kernel); one warp per row.  NearestNeighborDistanceMetric.partial_fit keeps the
// last `budget` features of a track (matching.py:343-358): the caller passes the ring position.
__global__ void __launch_bounds__(256) gallery_append_kernel(int n, int dim, int budget, const float* __restrict__ rows,
                                                             const int* __restrict__ slot, const int* __restrict__ pos,
                                                             float* __restrict__ gal32, __nv_bfloat16* __restrict__ gal16) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const float4* s = reinterpret_cast<const float4*>(rows + (size_t)r * dim);
    const size_t dst = ((size_t)slot[r] * budget + pos[r]) * dim;
    float4* o32 = reinterpret_cast<float4*>(gal32 + dst);
    __nv_bfloat162* o16 = reinterpret_cast<__nv_bfloat162*>(gal16 + dst);
    const int nv = dim >> 2;
    float acc = 0.f;
    for (int i = lane; i < nv; i += 32) { const float4 v = s[i]; o32[i] = v; acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    const float inv = acc > 0.f ? rsqrtf(acc) : 0.f;
    for (int i = lane; i < nv; i += 32) {
        const float4 v = s[i];
        o16[2 * i] = __floats2bfloat162_rn(v.x * inv, v.y * inv);
        o16[2 * i + 1] = __floats2bfloat162_rn(v.z * inv, v.w * inv);
    }
}

struct CostArgs {
    int n_trk, n_det, dim, npad;          // npad: detections of one N tile rounded up to 16
    const float* trk;                     // [B, n_trk, dim]
    const float* det;                     // [B, n_det, dim]
    const uint8_t* gate;                  // [B, n_trk, n_det] or null
    double scale, thresh, fill;
    double* out;                          // [B, n_trk, n_det]
    unsigned long long* stats;            // [0] entries rechecked in fp64, [1] protocol errors
};

// ---- kernel 2 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 2)
appearance_cost_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const CostArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 128B swizzle atoms are 1024 B
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int fail_s;
    __shared__ uint32_t cand_bits[BM * (BN_MAX / 32)];     // one bit per entry of the tile: survives the bf16 pre-filter
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN_MAX, b = blockIdx.z;
    const int nk = g.dim / BK;
    const int ncols = min(g.n_det - n0, BN_MAX);          // valid detections of this tile
    const int npad = (ncols + 15) & ~15;                  // UMMA N

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        fail_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                       // TMEM: 256 fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer =====
            const uint32_t tx = (uint32_t)(A_BYTES + g.npad * BK * 2);       // the box is g.npad rows tall for every tile
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES && !mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1)) { fail_s = 1; break; }
                unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], tx);
                tma_load_3d(sa, &map_a, &full_bar[s], kb * BK, m0, b);
                tma_load_3d(sa + A_BYTES, &map_b, &full_bar[s], kb * BK, n0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            bool ok = true;
            for (int kb = 0; kb < nk && ok; ++kb) {
                const int s = kb % STAGES;
                ok = mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                if (!ok) { fail_s = 1; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sa + A_BYTES + k * 32), idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty_bar[s]);                // frees the stage when these MMAs have read it
            }
            umma_commit(&acc_bar);                         // accumulator complete
        }
    }
    // ===== epilogue warps 2..5 =====
    // (1) while the GEMM runs: the whole tile is filled with `fill`, coalesced (row by row, lanes across columns);
    //     entries that survive the pre-filter are overwritten with their exact value in (3).
    const int rows_valid = min(BM, g.n_trk - m0);
    if (warp >= 2) {
        for (int r = warp - 2; r < rows_valid; r += 4) {
            double* orow = g.out + ((size_t)b * g.n_trk + m0 + r) * g.n_det + n0;
            for (int c = lane; c < ncols; c += 32) orow[c] = g.fill;
        }
    }
    // (2) accumulator -> one candidate bit per entry: warp w owns TMEM lanes 32 * (w % 4) .. + 31, thread = row
    uint32_t (*cbits)[BN_MAX / 32] = reinterpret_cast<uint32_t (*)[BN_MAX / 32]>(cand_bits);
    bool acc_ok = mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!acc_ok) fail_s = 1;
    __syncthreads();
    const bool failed = fail_s != 0;
    if (warp >= 2 && !failed) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool row_ok = row < rows_valid;
        // candidate iff scale * (1 - cos) <= thresh + scale * band  <=>  cos >= 1 - thresh / scale - band (NaN -> candidate)
        const float cos_lim = (float)(1.0 - g.thresh / g.scale - BAND_COS) - 1e-6f;
        uint32_t word = 0;
        for (int c0 = 0; c0 < BN_MAX; c0 += 16) {
            if (c0 < npad) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);     // warp-collective
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < ncols && !(__uint_as_float(v[i]) < cos_lim)) word |= 1u << ((c0 + i) & 31);
            }
            if ((c0 & 16) != 0) { cbits[row][c0 >> 5] = row_ok ? word : 0u; word = 0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    if (failed) {
        if (threadIdx.x == 0 && g.stats) atomicAdd(&g.stats[1], 1ull);
        return;
    }
    // (3) exact re-evaluation, one warp per surviving entry (scipy cdist 'cosine' in double on the fp32 values)
    const int nv = g.dim >> 2;
    const int nwords = (ncols + 31) >> 5;
    int done = 0;
    for (int idx = warp; idx < rows_valid * nwords; idx += GEMM_THREADS / 32) {
        const int row = idx / nwords, wd = idx - row * nwords;
        uint32_t bits = cbits[row][wd];
        const int t = m0 + row;
        if (g.gate) {                                      // gated entries keep `fill`: one coalesced 32-byte read per word
            const int dl = wd * 32 + lane;
            const bool gt = dl < ncols && g.gate[((size_t)b * g.n_trk + t) * g.n_det + n0 + dl] != 0;
            bits &= ~__ballot_sync(0xffffffffu, gt);
        }
        const float4* a = reinterpret_cast<const float4*>(g.trk + ((size_t)b * g.n_trk + t) * g.dim);
        while (bits) {
            const int d = n0 + wd * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            const float4* bb = reinterpret_cast<const float4*>(g.det + ((size_t)b * g.n_det + d) * g.dim);
            double uv = 0.0, uu = 0.0, vv = 0.0;
            for (int i = lane; i < nv; i += 32) {
                const float4 x = a[i], y = bb[i];
                uv += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
                uu += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
                vv += (double)y.x * y.x + (double)y.y * y.y + (double)y.z * y.z + (double)y.w * y.w;
            }
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                uv += __shfl_xor_sync(0xffffffffu, uv, s); uu += __shfl_xor_sync(0xffffffffu, uu, s); vv += __shfl_xor_sync(0xffffffffu, vv, s);
            }
            ++done;
            if (lane == 0) {
                double c = uv / (sqrt(uu) * sqrt(vv));
                if (fabs(c) > 1.0) c = copysign(1.0, c);
                const double val = g.scale * fmax(0.0, 1.0 - c);
                if (!(val > g.thresh)) g.out[((size_t)b * g.n_trk + t) * g.n_det + d] = val;
            }
        }
    }
    if (lane == 0 && g.stats && done) atomicAdd(&g.stats[0], (unsigned long long)done);
}

// ---- kernel 3: StrongSORT's gallery distance -------------------------------------------------------
//     out[b, t, d] = min over the stored features g < count[b, t] of track t of  1 - a_g^ . b_d^      if that is <= thresh
//                  = fill                                                                            otherwise
// (NearestNeighborDistanceMetric.distance with _nn_cosine_distance, boxmot/utils/matching.py:247-308, :360-378, followed
// by the clip of strongsort/sort/linear_assignment.py:59-78: everything above max_distance becomes one value).
// This is the GEMM-shaped part of the tracking loop proper: per stream [T x G, F] . [F, D] with G up to 128 stored
// features per track (2 * T * G * D * F flop, 4.1 GFLOP per stream and frame at T = D = 200, G = 100, F = 512).
// One CTA per (track, stream): M = 128 gallery rows of that track (TMA zero-fills the rows past the budget), N = the
// stream's detections (<= 256), K = F through the same 2-stage TMA / tcgen05 ring as above; two CTAs share an SM so
// one's epilogue runs under the other's MMAs.  Epilogue: column maxima of the cosine tile over the 128 TMEM lanes
// (register butterfly inside a warp, shared memory across the four warps), threshold test on 1 - max with the bf16
// band, then the surviving (track, detection) pairs - about one per track - are re-evaluated exactly in float32 on
// the fp32 gallery, restricted to the rows whose bf16 cosine is within the band of the column maximum.
struct GalleryArgs {
    int n_trk, budget, n_det, dim, npad;
    const float* gallery;                 // [B, n_trk, budget, dim] fp32
    const int* count;                     // [B, n_trk]
    const float* det;                     // [B, n_det, dim] fp32
    double thresh, fill;
    double* out;                          // [B, n_trk, n_det]
    unsigned long long* stats;            // [0] exact row evaluations, [1] protocol errors, [2] candidate pairs
};

__global__ void __launch_bounds__(GEMM_THREADS, 2)
gallery_cost_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GalleryArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int fail_s, ncand_s;
    __shared__ float colmax[4][BN_MAX];                    // per epilogue warp: max cosine of every column over its 32 rows
    __shared__ int cand[BN_MAX];                           // surviving detections of this track
    __shared__ uint32_t rowbits[BN_MAX][4];                // per surviving detection: gallery rows worth an exact evaluation
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x, b = blockIdx.y;
    const int nk = g.dim / BK;
    const int ncols = g.n_det, npad = g.npad;
    const int cnt = min(g.count[(size_t)b * g.n_trk + t], min(g.budget, BM));
    double* orow = g.out + ((size_t)b * g.n_trk + t) * g.n_det;
    if (cnt <= 0) {                                        // no stored feature: nothing can match (uniform exit)
        for (int c = threadIdx.x; c < ncols; c += GEMM_THREADS) orow[c] = g.fill;
        return;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        fail_s = 0; ncand_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer =====
            const uint32_t tx = (uint32_t)(A_BYTES + npad * BK * 2);
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES && !mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1)) { fail_s = 1; break; }
                unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], tx);
                tma_load_3d(sa, &map_a, &full_bar[s], kb * BK, 0, b * g.n_trk + t);
                tma_load_3d(sa + A_BYTES, &map_b, &full_bar[s], kb * BK, 0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            bool ok = true;
            for (int kb = 0; kb < nk && ok; ++kb) {
                const int s = kb % STAGES;
                ok = mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                if (!ok) { fail_s = 1; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sa + A_BYTES + k * 32), idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&acc_bar);
        }
    }
    // ===== epilogue =====
    bool acc_ok = mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!acc_ok) fail_s = 1;
    __syncthreads();
    const bool failed = fail_s != 0;
    const float NEG = -3.0e38f;
    if (warp >= 2 && !failed) {
        const int q = warp & 3;                            // TMEM lanes 32 q .. 32 q + 31 = gallery rows
        const bool row_ok = q * 32 + lane < cnt;
        for (int c0 = 0; c0 < npad; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = row_ok ? __uint_as_float(v[i]) : NEG;
            // butterfly with halving: after the 16 / 8 / 4 / 2 steps a lane holds one column's maximum over 16 lanes
#pragma unroll
            for (int h = 8, d = 16; h >= 1; h >>= 1, d >>= 1) {
                const bool up = (lane & d) != 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < h) {
                        const float keep = up ? f[i + h] : f[i];
                        const float give = up ? f[i] : f[i + h];
                        f[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, give, d));
                    }
                }
            }
            f[0] = fmaxf(f[0], __shfl_xor_sync(0xffffffffu, f[0], 1));
            // column held by this lane: bit 4 of the lane chose +8, bit 3 +4, bit 2 +2, bit 1 +1
            const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0) colmax[q][col] = f[0];
        }
    }
    __syncthreads();
    // candidate iff 1 - max cos <= thresh + band  (NaN -> candidate); everything else is `fill`
    const float cos_lim = (float)(1.0 - g.thresh - BAND_COS) - 1e-6f;
    if (!failed) {
        for (int c = threadIdx.x; c < ncols; c += GEMM_THREADS) {
            const float m = fmaxf(fmaxf(colmax[0][c], colmax[1][c]), fmaxf(colmax[2][c], colmax[3][c]));
            if (!(m < cos_lim)) {
                const int k = atomicAdd(&ncand_s, 1);
                cand[k] = c;
                colmax[0][c] = m;                          // the overall maximum, read back by the row filter
                rowbits[k][0] = rowbits[k][1] = rowbits[k][2] = rowbits[k][3] = 0u;
            } else orow[c] = g.fill;
        }
    }
    __syncthreads();
    const int nc = ncand_s;
    // rows of a surviving column that can hold the exact maximum: bf16 cosine within twice the band of the column maximum
    if (warp >= 2 && !failed && nc > 0) {
        const int q = warp & 3;
        const bool row_ok = q * 32 + lane < cnt;
        for (int k = 0; k < nc; ++k) {
            const int c = cand[k];
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c & ~15), v);     // warp-collective
            float x = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) if (i == (c & 15)) x = __uint_as_float(v[i]);
            const bool keep = row_ok && !(x < colmax[0][c] - (float)(2.0 * BAND_COS));
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) rowbits[k][q] = m;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    if (failed) {
        if (threadIdx.x == 0 && g.stats) atomicAdd(&g.stats[1], 1ull);
        return;
    }
    // exact float32 re-evaluation (matching.py:247-267: rows normalised, 1 - dot).  A surviving pair needs up to `count`
    // gallery rows of 2 KB each and there is about one pair per CTA, so all six warps share a pair: warp w takes every
    // sixth kept row, two rows in flight at a time, and the minima meet in shared memory.
    __shared__ float wbest[GEMM_THREADS / 32];
    const int nv = g.dim >> 2;
    unsigned long long done = 0;
    const float* gbase = g.gallery + ((size_t)b * g.n_trk + t) * g.budget * g.dim;
    for (int k = 0; k < nc; ++k) {
        const int d = cand[k];
        const float4* bb = reinterpret_cast<const float4*>(g.det + ((size_t)b * g.n_det + d) * g.dim);
        float vv = 0.f;
        for (int i = lane; i < nv; i += 32) { const float4 y = bb[i]; vv += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w; }
#pragma unroll
        for (int s = 16; s; s >>= 1) vv += __shfl_xor_sync(0xffffffffu, vv, s);
        const float nb = sqrtf(vv);
        float best = 3.0e38f;
        auto eval2 = [&](int r0, int r1) {                 // r1 < 0: single row
            const float4* a0 = reinterpret_cast<const float4*>(gbase + (size_t)r0 * g.dim);
            const float4* a1 = reinterpret_cast<const float4*>(gbase + (size_t)(r1 < 0 ? r0 : r1) * g.dim);
            float uv0 = 0.f, uu0 = 0.f, uv1 = 0.f, uu1 = 0.f;
            for (int i = lane; i < nv; i += 32) {
                const float4 x0 = a0[i], x1 = a1[i], y = bb[i];
                uv0 += x0.x * y.x + x0.y * y.y + x0.z * y.z + x0.w * y.w;
                uu0 += x0.x * x0.x + x0.y * x0.y + x0.z * x0.z + x0.w * x0.w;
                uv1 += x1.x * y.x + x1.y * y.y + x1.z * y.z + x1.w * y.w;
                uu1 += x1.x * x1.x + x1.y * x1.y + x1.z * x1.z + x1.w * x1.w;
            }
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                uv0 += __shfl_xor_sync(0xffffffffu, uv0, s); uu0 += __shfl_xor_sync(0xffffffffu, uu0, s);
                uv1 += __shfl_xor_sync(0xffffffffu, uv1, s); uu1 += __shfl_xor_sync(0xffffffffu, uu1, s);
            }
            best = fminf(best, fminf(1.0f - uv0 / (sqrtf(uu0) * nb), 1.0f - uv1 / (sqrtf(uu1) * nb)));
            done += r1 < 0 ? 1 : 2;
        };
        int seen = 0, pend = -1;
        for (int w = 0; w < 4; ++w) {
            uint32_t bits = rowbits[k][w];
            while (bits) {
                const int r = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                if (seen++ % (GEMM_THREADS / 32) != warp) continue;
                if (pend < 0) pend = r; else { eval2(pend, r); pend = -1; }
            }
        }
        if (pend >= 0) eval2(pend, -1);
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = wbest[0];
#pragma unroll
            for (int w = 1; w < GEMM_THREADS / 32; ++w) m = fminf(m, wbest[w]);
            orow[d] = ((double)m > g.thresh) ? g.fill : (double)m;
        }
        __syncthreads();
    }
    if (lane == 0 && g.stats) {
        if (done) atomicAdd(&g.stats[0], done);
        if (warp == 0 && nc) atomicAdd(&g.stats[2], (unsigned long long)nc);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [batch, rows, dim] bf16, box = {64, box_rows, 1}, 128B swizzle, zero fill outside
bool make_map(CUtensorMap* m, const void* base, int batch, int rows, int dim, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)dim, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)dim * 2, (cuuint64_t)rows * dim * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t gallery_ws_bytes(int batch, int n_trk, int budget, int n_det, int dim, bool with_gallery) {
    const size_t a = with_gallery ? (((size_t)batch * n_trk * budget * dim * 2 + 255) & ~size_t(255)) : 0;
    const size_t b = ((size_t)batch * n_det * dim * 2 + 255) & ~size_t(255);
    return a + b;
}

size_t ws_bytes(int batch, int n_trk, int n_det, int dim) {
    const size_t a = ((size_t)batch * n_trk * dim * 2 + 255) & ~size_t(255);
    const size_t b = ((size_t)batch * n_det * dim * 2 + 255) & ~size_t(255);
    return a + b;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200track_appearance_cost_workspace(int32_t batch, int32_t n_tracks, int32_t n_dets, int32_t dim, uint64_t* h_bytes) {
    if (batch < 0 || n_tracks < 0 || n_dets < 0 || dim <= 0 || !h_bytes) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    *h_bytes = ws_bytes(batch, n_tracks, n_dets, dim);
    return 0;
}

extern "C" int b200track_appearance_cost(int32_t batch, int32_t n_tracks, int32_t n_dets, int32_t dim, const float* d_trk,
                                         const float* d_det, const uint8_t* d_gate, double scale, double thresh, double fill,
                                         double* d_out, void* d_workspace, uint64_t workspace_bytes, uint64_t* d_stats, void* st) {
    if (batch < 0 || n_tracks < 0 || n_dets < 0 || !d_out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (dim <= 0 || dim % 64) { set_error("dim must be a positive multiple of 64"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || n_tracks == 0 || n_dets == 0) return 0;
    if (!d_trk || !d_det || !d_workspace) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    if (workspace_bytes < ws_bytes(batch, n_tracks, n_dets, dim)) { set_error("workspace too small"); return B200TRACK_ERR_ARG; }
    if (batch > 65535) { set_error("batch > 65535"); return B200TRACK_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)st;
    __nv_bfloat16* wa = reinterpret_cast<__nv_bfloat16*>(d_workspace);
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(d_workspace) +
                                                         (((size_t)batch * n_tracks * dim * 2 + 255) & ~size_t(255)));
    const long long ra = (long long)batch * n_tracks, rb = (long long)batch * n_dets;
    unit_bf16_kernel<<<(unsigned)((ra + 7) / 8), 256, 0, stream>>>(d_trk, wa, ra, dim);
    unit_bf16_kernel<<<(unsigned)((rb + 7) / 8), 256, 0, stream>>>(d_det, wb, rb, dim);
    B200_CU_TRY(cudaGetLastError());
    const int box_n = std::min((n_dets + 15) & ~15, BN_MAX);
    CUtensorMap ma, mb;
    if (!make_map(&ma, wa, batch, n_tracks, dim, BM) || !make_map(&mb, wb, batch, n_dets, dim, box_n)) {
        set_error("cuTensorMapEncodeTiled failed"); return B200TRACK_ERR_CUDA; }
    CostArgs g;
    g.n_trk = n_tracks; g.n_det = n_dets; g.dim = dim; g.npad = box_n;
    g.trk = d_trk; g.det = d_det; g.gate = d_gate; g.scale = scale; g.thresh = thresh; g.fill = fill; g.out = d_out;
    g.stats = reinterpret_cast<unsigned long long*>(d_stats);
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    B200_CU_TRY(cudaFuncSetAttribute(appearance_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n_tracks + BM - 1) / BM, (n_dets + BN_MAX - 1) / BN_MAX, batch);
    appearance_cost_kernel<<<grid, GEMM_THREADS, smem, stream>>>(ma, mb, g);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int b200track_unit_bf16(int64_t rows, int32_t dim, const float* d_src, void* d_dst, void* st) {
    if (rows < 0 || dim <= 0 || dim % 4 || !d_dst || (!d_src && rows > 0)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (rows == 0) return 0;
    unit_bf16_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)st>>>(d_src, reinterpret_cast<__nv_bfloat16*>(d_dst), rows, dim);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int b200track_gallery_append(int32_t n, int32_t dim, int32_t budget, const float* d_rows, const int32_t* d_slot,
                                        const int32_t* d_pos, float* d_gallery, void* d_gallery_bf16, void* st) {
    if (n < 0 || dim <= 0 || dim % 4 || budget <= 0 || !d_gallery || !d_gallery_bf16 || (n > 0 && (!d_rows || !d_slot || !d_pos))) {
        set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    gallery_append_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)st>>>(n, dim, budget, d_rows, d_slot, d_pos, d_gallery,
                                                                      reinterpret_cast<__nv_bfloat16*>(d_gallery_bf16));
    B200_CU_TRY(cudaGetLastError());
    return 0;
}

extern "C" int b200track_gallery_cost_workspace(int32_t batch, int32_t n_tracks, int32_t budget, int32_t n_dets, int32_t dim,
                                                int32_t with_gallery, uint64_t* h_bytes) {
    if (batch < 0 || n_tracks < 0 || budget < 0 || n_dets < 0 || dim <= 0 || !h_bytes) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    *h_bytes = gallery_ws_bytes(batch, n_tracks, budget, n_dets, dim, with_gallery != 0);
    return 0;
}

extern "C" int b200track_gallery_cost(int32_t batch, int32_t n_tracks, int32_t budget, int32_t n_dets, int32_t dim,
                                      const float* d_gallery, const void* d_gallery_bf16, const int32_t* d_count, const float* d_det,
                                      double thresh, double fill, double* d_out, void* d_workspace, uint64_t workspace_bytes,
                                      uint64_t* d_stats, void* st) {
    if (batch < 0 || n_tracks < 0 || n_dets < 0 || budget < 0 || !d_out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (dim <= 0 || dim % 64) { set_error("dim must be a positive multiple of 64"); return B200TRACK_ERR_ARG; }
    if (budget > BM) { set_error("gallery budget above 128 rows per track"); return B200TRACK_ERR_CAPACITY; }
    if (n_dets > BN_MAX) { set_error("more than 256 detections per stream"); return B200TRACK_ERR_CAPACITY; }
    if (batch == 0 || n_tracks == 0 || n_dets == 0) return 0;
    if (batch > 65535) { set_error("batch > 65535"); return B200TRACK_ERR_ARG; }
    if (!d_gallery || !d_count || !d_det || !d_workspace || budget == 0) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    const bool own = d_gallery_bf16 == nullptr;
    if (workspace_bytes < gallery_ws_bytes(batch, n_tracks, budget, n_dets, dim, own)) { set_error("workspace too small"); return B200TRACK_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)st;
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
    const __nv_bfloat16* ga = reinterpret_cast<const __nv_bfloat16*>(d_gallery_bf16);
    if (own) {
        const long long ra = (long long)batch * n_tracks * budget;
        unit_bf16_kernel<<<(unsigned)((ra + 7) / 8), 256, 0, stream>>>(d_gallery, reinterpret_cast<__nv_bfloat16*>(ws), ra, dim);
        ga = reinterpret_cast<const __nv_bfloat16*>(ws);
        ws += ((size_t)ra * dim * 2 + 255) & ~size_t(255);
    }
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(ws);
    const long long rb = (long long)batch * n_dets;
    unit_bf16_kernel<<<(unsigned)((rb + 7) / 8), 256, 0, stream>>>(d_det, wb, rb, dim);
    B200_CU_TRY(cudaGetLastError());
    const int box_n = std::min((n_dets + 15) & ~15, BN_MAX);
    CUtensorMap ma, mb;
    if (!make_map(&ma, ga, batch * n_tracks, budget, dim, BM) || !make_map(&mb, wb, batch, n_dets, dim, box_n)) {
        set_error("cuTensorMapEncodeTiled failed"); return B200TRACK_ERR_CUDA; }
    GalleryArgs g;
    g.n_trk = n_tracks; g.budget = budget; g.n_det = n_dets; g.dim = dim; g.npad = box_n;
    g.gallery = d_gallery; g.count = d_count; g.det = d_det; g.thresh = thresh; g.fill = fill; g.out = d_out;
    g.stats = reinterpret_cast<unsigned long long*>(d_stats);
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    B200_CU_TRY(cudaFuncSetAttribute(gallery_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_tracks, batch, 1);
    gallery_cost_kernel<<<grid, GEMM_THREADS, smem, stream>>>(ma, mb, g);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}
