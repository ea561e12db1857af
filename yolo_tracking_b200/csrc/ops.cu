// Operator-level kernels: the reference's functional API, batched (include/b200track.h).
// Dense [n,8] / [n,8,8] arrays in the reference's layout, so these also accept covariances
// that are NOT block-sparse (e.g. after an externally applied camera-motion warp).
//
// Kalman kernels: a CTA stages a contiguous run of tracks (mean + covariance, 72 doubles
// each) in shared memory with coalesced loads, one thread then owns one track (row stride
// 73 doubles = conflict-free), results go back with coalesced stores.  HBM-bound:
// 2 x 576 B per track for predict / update.
#include <type_traits>

#include "api_util.h"
#include "boxes.cuh"
#include "kf.cuh"
#include "lap_dense_matrix.cuh"
#include "lap_sparse.cuh"

namespace b200 {
namespace {

constexpr int KF_TPB = 64;          // tracks (= threads) per CTA
constexpr int KF_STRIDE = 73;       // 8 mean + 64 cov + 1 pad

__device__ __forceinline__ void kf_stage_in(double* sm, const double* mean, const double* cov, int base, int cnt) {
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) sm[(i >> 3) * KF_STRIDE + (i & 7)] = mean[(size_t)base * 8 + i];
    for (int i = threadIdx.x; i < cnt * 64; i += blockDim.x) sm[(i >> 6) * KF_STRIDE + 8 + (i & 63)] = cov[(size_t)base * 64 + i];
    __syncthreads();
}
__device__ __forceinline__ void kf_stage_out(const double* sm, double* mean, double* cov, int base, int cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) mean[(size_t)base * 8 + i] = sm[(i >> 3) * KF_STRIDE + (i & 7)];
    for (int i = threadIdx.x; i < cnt * 64; i += blockDim.x) cov[(size_t)base * 64 + i] = sm[(i >> 6) * KF_STRIDE + 8 + (i & 63)];
}

// per-axis standard deviations (bytetrack_kf.py:76-85,107-116,143-148 / botsort_kf.py same lines)
template <int KIND>
__device__ __forceinline__ void kf_std(const double* ref4, double ps, double vs, double cp, double cv, double* std8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (KIND != KF_XYWH && i == 2) { std8[i] = cp; std8[i + 4] = cv; }
        else {
            const double r = kf_ref<KIND>(ref4, i);
            std8[i] = xmul(ps, r);
            std8[i + 4] = xmul(vs, r);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(KF_TPB) kf_initiate_kernel(int n, const double* __restrict__ z, double* mean, double* cov) {
    __shared__ double sm[KF_TPB * KF_STRIDE];
    const int base = blockIdx.x * KF_TPB, cnt = min(KF_TPB, n - base), t = threadIdx.x;
    if (t < cnt) {
        double* m = sm + t * KF_STRIDE;
        double* P = m + 8;
        double zz[4], sd[8];
        for (int i = 0; i < 4; ++i) zz[i] = z[(size_t)(base + t) * 4 + i];
        kf_std<KIND>(zz, 2 * KF_W_POS, 10 * KF_W_VEL, 1e-2, 1e-5, sd);
        for (int i = 0; i < 64; ++i) P[i] = 0.0;
        for (int i = 0; i < 4; ++i) { m[i] = zz[i]; m[i + 4] = 0.0; }
        for (int i = 0; i < 8; ++i) P[i * 9] = xmul(sd[i], sd[i]);
    }
    kf_stage_out(sm, mean, cov, base, cnt);
}

// One lane per covariance ROW (8 lanes per track, KF_TPB tracks per CTA): every lane computes its new row in
// registers from the staged tile, the group synchronises, then the rows are written back - 8x the threads of a
// thread-per-track mapping for the same shared memory, which is what a latency-bound streaming kernel needs.
constexpr int KF_LANES = 8;
constexpr int KF_THREADS = KF_TPB * KF_LANES;

template <int KIND>
__global__ void __launch_bounds__(KF_THREADS) kf_predict_kernel(int n, double* mean, double* cov) {
    __shared__ double sm[KF_TPB * KF_STRIDE];
    const int base = blockIdx.x * KF_TPB, cnt = min(KF_TPB, n - base);
    const int t = threadIdx.x / KF_LANES, r = threadIdx.x % KF_LANES;
    kf_stage_in(sm, mean, cov, base, cnt);
    double row[8], mr = 0.0;
    if (t < cnt) {
        const double* m = sm + t * KF_STRIDE;
        const double* P = m + 8;
        double sd[8];
        kf_std<KIND>(m, KF_W_POS, KF_W_VEL, 1e-2, 1e-5, sd);
        mr = r < 4 ? xadd(m[r], m[r + 4]) : m[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = r < 4 ? xadd(P[r * 8 + j], P[(r + 4) * 8 + j]) : P[r * 8 + j];      // (F P)[r, :]
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = xadd(row[j], row[j + 4]);                                          // ... F^T
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j == r) row[j] = xadd(row[j], xmul(sd[j], sd[j]));
    }
    __syncthreads();
    if (t < cnt) {
        double* m = sm + t * KF_STRIDE;
        m[r] = mr;
#pragma unroll
        for (int j = 0; j < 8; ++j) m[8 + r * 8 + j] = row[j];
    }
    kf_stage_out(sm, mean, cov, base, cnt);
}

// S = H P H^T + R (4x4) and its lower Cholesky factor
template <int KIND>
__device__ __forceinline__ void kf_innovation(const double* m, const double* P, double conf, double S[4][4]) {
    double sd[8];
    kf_std<KIND>(m, KF_W_POS, KF_W_VEL, 1e-1, 0.0, sd);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i][j] = P[i * 8 + j];
    for (int i = 0; i < 4; ++i) {
        double s = sd[i];
        if (KIND == KF_XYAH_CONF) s = xmul(xsub(1.0, conf), s);
        S[i][i] = xadd(S[i][i], xmul(s, s));
    }
}
template <int N>
__device__ __forceinline__ void chol_lower(double S[4][4], double L[4][4]) {
    for (int j = 0; j < N; ++j) {
        double d = S[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        L[j][j] = sqrt(d);
        for (int i = j + 1; i < N; ++i) {
            double v = S[i][j];
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v / L[j][j];
        }
    }
}

// project (bytetrack_kf.py:126-153) only touches mean[:4] and the top-left 4x4 block of P: five 32-byte sectors per
// track are read (160 B), 160 B written; 4 lanes per track (one per row of the block).
template <int KIND>
__global__ void __launch_bounds__(256) kf_project_kernel(int n, const double* __restrict__ mean, const double* __restrict__ cov,
                                                         const double* __restrict__ conf, double* __restrict__ pmean,
                                                         double* __restrict__ pcov) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int t = idx >> 2, r = idx & 3;
    if (t >= n) return;
    const double2 m01 = *reinterpret_cast<const double2*>(mean + (size_t)t * 8);
    const double2 m23 = *reinterpret_cast<const double2*>(mean + (size_t)t * 8 + 2);
    const double m4[4] = {m01.x, m01.y, m23.x, m23.y};
    const double2 p01 = *reinterpret_cast<const double2*>(cov + (size_t)t * 64 + r * 8);
    const double2 p23 = *reinterpret_cast<const double2*>(cov + (size_t)t * 64 + r * 8 + 2);
    double row[4] = {p01.x, p01.y, p23.x, p23.y};
    double sd[8];
    kf_std<KIND>(m4, KF_W_POS, KF_W_VEL, 1e-1, 0.0, sd);
    double sr = 0.0, mr = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j == r) { sr = sd[j]; mr = m4[j]; }
    if (KIND == KF_XYAH_CONF) sr = xmul(xsub(1.0, conf ? conf[t] : 0.0), sr);
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j == r) row[j] = xadd(row[j], xmul(sr, sr));
    pmean[(size_t)t * 4 + r] = mr;
    *reinterpret_cast<double2*>(pcov + (size_t)t * 16 + r * 4) = make_double2(row[0], row[1]);
    *reinterpret_cast<double2*>(pcov + (size_t)t * 16 + r * 4 + 2) = make_double2(row[2], row[3]);
}

// update (bytetrack_kf.py:194-226): S = L L^T; lane r solves its own gain row
// K[r, :] = P[r, :4] S^-1 (cho_solve), updates mean[r] and covariance row r.  P' = P - K S K^T is evaluated as
// P - K (H P): S K^T = S S^-1 (P H^T)^T = H P exactly in real arithmetic, and rounding-wise inside the 1e-9 bar.
constexpr int KFU_TPB = 32;                    // tracks per CTA of the update kernel (256 threads)
// MASKED (the batched StrongSORT step, strongsort_step.cu): the n tracks are the slots of all streams ([streams, T]); only
// the slots with sel[slot] >= 0 are updated, with measurement / confidence row sel[slot] of their stream's [D, 4] / [D]
// blocks; tiles without a selected slot leave at once.
template <int KIND, bool MASKED = false>
__global__ void __launch_bounds__(KFU_TPB * KF_LANES, 6) kf_update_kernel(int n, double* mean, double* cov, const double* __restrict__ z,
                                                                         const double* __restrict__ conf, const int* __restrict__ sel = nullptr,
                                                                         int T = 1, int D = 1) {
    __shared__ double sm[KFU_TPB * KF_STRIDE];
    __shared__ double sF[KFU_TPB][15];             // per track: L (10 lower-triangle entries), 1 / L[i][i] (4); stride 15 = conflict-free
    const int base = blockIdx.x * KFU_TPB, cnt = min(KFU_TPB, n - base);
    const int t = threadIdx.x / KF_LANES, r = threadIdx.x % KF_LANES;
    int zrow = base + t;                           // row of z / conf this thread's track reads
    bool mine = t < cnt;
    if (MASKED) {
        const int j = t < cnt ? sel[base + t] : -1;
        mine = j >= 0;
        zrow = ((base + t) / T) * D + j;
        if (!__syncthreads_or(mine)) return;
    }
    kf_stage_in(sm, mean, cov, base, cnt);
    // The 4x4 factorisation (4 sqrt, 6 divisions, 4 reciprocals: most of the kernel's fp64 instructions) is common to the
    // eight lanes of a track: the first warp factors all 32 tracks of the CTA, one per lane, instead of every lane for itself.
    int qrow = base + threadIdx.x;
    bool qsel = threadIdx.x < cnt;
    if (MASKED && qsel) { const int j = sel[base + threadIdx.x]; qsel = j >= 0; qrow = ((base + threadIdx.x) / T) * D + j; }
    if (qsel) {
        const int q = threadIdx.x;
        const double* m = sm + q * KF_STRIDE;
        double S[4][4], L[4][4];
        kf_innovation<KIND>(m, m + 8, conf ? conf[qrow] : 0.0, S);
        chol_lower<4>(S, L);
        int o = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) sF[q][o++] = L[i][j];
#pragma unroll
        for (int i = 0; i < 4; ++i) sF[q][10 + i] = 1.0 / L[i][i];
    }
    __syncthreads();
    double row[8], mr = 0.0;
    if (mine) {
        const double* m = sm + t * KF_STRIDE;
        const double* P = m + 8;
        double L[4][4], inv[4];
        {
            int o = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) L[i][j] = sF[t][o++];
#pragma unroll
            for (int i = 0; i < 4; ++i) inv[i] = sF[t][10 + i];
        }
        double y[4], k[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                 // L y = P[r, :4]
            double v = P[r * 8 + i];
#pragma unroll
            for (int q = 0; q < i; ++q) v -= L[i][q] * y[q];
            y[i] = v * inv[i];
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {                // L^T k = y
            double v = y[i];
#pragma unroll
            for (int q = i + 1; q < 4; ++q) v -= L[q][i] * k[q];
            k[i] = v * inv[i];
        }
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) a += xsub(z[(size_t)zrow * 4 + i], m[i]) * k[i];
        mr = xadd(m[r], a);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double b = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) b += k[i] * P[i * 8 + c];
            row[c] = xsub(P[r * 8 + c], b);
        }
    }
    __syncthreads();
    if (mine) {
        double* m = sm + t * KF_STRIDE;
        m[r] = mr;
#pragma unroll
        for (int c = 0; c < 8; ++c) m[8 + r * 8 + c] = row[c];
    }
    kf_stage_out(sm, mean, cov, base, cnt);
}

// multi_gmc (bot_sort.py:95-111): mean <- R8 mean (+ t on x, y), P <- R8 P R8^T with R8 = kron(I4, R), R = warp[:2, :2],
// t = warp[:2, 2].  One warp matrix per problem (warp_index[t] picks it, NULL = all tracks use warp 0).  Same staging as
// predict; lane r owns row r: A[r, :] = R[r&1, 0] P[r&~1, :] + R[r&1, 1] P[r|1, :], then P'[r, 2k+b] = A[r, 2k] R[b, 0] +
// A[r, 2k+1] R[b, 1].  The product breaks the 2x2 block sparsity the frame steps rely on, hence the dense layout.
__global__ void __launch_bounds__(KF_THREADS) kf_gmc_kernel(int n, double* mean, double* cov, const double* __restrict__ warp,
                                                            const int* __restrict__ warp_index) {
    __shared__ double sm[KF_TPB * KF_STRIDE];
    const int base = blockIdx.x * KF_TPB, cnt = min(KF_TPB, n - base);
    const int t = threadIdx.x / KF_LANES, r = threadIdx.x % KF_LANES;
    kf_stage_in(sm, mean, cov, base, cnt);
    double row[8], mr = 0.0;
    if (t < cnt) {
        const double* m = sm + t * KF_STRIDE;
        const double* P = m + 8;
        const double* w = warp + (size_t)(warp_index ? warp_index[base + t] : 0) * 6;
        const double R[2][2] = {{w[0], w[1]}, {w[3], w[4]}};
        const int a = r & 1, r0 = r & ~1;
        mr = xadd(xmul(R[a][0], m[r0]), xmul(R[a][1], m[r0 + 1]));
        if (r < 2) mr = xadd(mr, w[2 + 3 * r]);
        double A[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) A[c] = xadd(xmul(R[a][0], P[r0 * 8 + c]), xmul(R[a][1], P[(r0 + 1) * 8 + c]));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            row[2 * k + 0] = xadd(xmul(A[2 * k], R[0][0]), xmul(A[2 * k + 1], R[0][1]));
            row[2 * k + 1] = xadd(xmul(A[2 * k], R[1][0]), xmul(A[2 * k + 1], R[1][1]));
        }
    }
    __syncthreads();
    if (t < cnt) {
        double* m = sm + t * KF_STRIDE;
        m[r] = mr;
#pragma unroll
        for (int c = 0; c < 8; ++c) m[8 + r * 8 + c] = row[c];
    }
    kf_stage_out(sm, mean, cov, base, cnt);
}

// compute_aw_max_metric (association.py:79-108, DeepOCSORT): per row and per column of the embedding-similarity matrix
// the two largest entries decide a weight  w = 1 - max(second / first - bottom, 0) / (1 - bottom)  (0 if the largest
// entry is 0, untouched if the row / column has fewer than two entries); out = w_assoc * w_row * w_col * emb.
// One CTA per problem: warp per row for the row weights, thread per column for the column weights, then the product.
__global__ void __launch_bounds__(256) aw_max_metric_kernel(int R, int C, const double* __restrict__ emb, double w_assoc, double bottom,
                                                            double* __restrict__ out) {
    extern __shared__ double wts[];                        // [R] row weights, [C] column weights
    emb += (size_t)blockIdx.x * R * C;
    out += (size_t)blockIdx.x * R * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double NINF = __longlong_as_double(0xfff0000000000000LL);
    auto weight = [&](double first, double second) {
        if (first == 0.0) return 0.0;
        return xsub(1.0, xdiv(fmax(xsub(xdiv(second, first), bottom), 0.0), xsub(1.0, bottom)));
    };
    for (int i = warp; i < R; i += 8) {
        double a = NINF, b = NINF;                         // largest, second largest of the row
        for (int j = lane; j < C; j += 32) {
            const double v = emb[(size_t)i * C + j];
            if (v > a) { b = a; a = v; } else if (v > b) b = v;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, a, d), ob = __shfl_xor_sync(0xffffffffu, b, d);
            if (oa > a) { b = fmax(a, ob); a = oa; } else b = fmax(b, oa);
        }
        if (lane == 0) wts[i] = C < 2 ? 1.0 : weight(a, b);
    }
    for (int j = threadIdx.x; j < C; j += 256) {
        double a = NINF, b = NINF;
        for (int i = 0; i < R; ++i) {
            const double v = emb[(size_t)i * C + j];
            if (v > a) { b = a; a = v; } else if (v > b) b = v;
        }
        wts[R + j] = R < 2 ? 1.0 : weight(a, b);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < R * C; k += 256) {
        const int i = k / C, j = k - i * C;
        out[k] = xmul(xmul(xmul(w_assoc, wts[i]), wts[R + j]), emb[k]);
    }
}

// gating_distance: 32 tracks per CTA factor S once (one warp), the stream's measurements are staged planar in shared
// memory, then every thread sweeps (track, measurement) pairs with coalesced 8-byte stores.
constexpr int GD_TRACKS = 64;
template <int KIND>
__global__ void __launch_bounds__(256) kf_gating_kernel(int T, int D, const double* __restrict__ mean, const double* __restrict__ cov,
                                                        const double* __restrict__ meas, int only_position, int metric,
                                                        const double* __restrict__ conf, double* __restrict__ out,
                                                        double* __restrict__ cost = nullptr, int fuse = 0, double gate_thr = 0.0,
                                                        double lambda = 0.0) {
    __shared__ double sL[GD_TRACKS][16];
    __shared__ double sM[GD_TRACKS][4];
    extern __shared__ double sZ[];              // [4][D] measurements, planar
    {   // blockIdx.y = independent problem (stream): [T] tracks x [D] measurements each
        const size_t bi = blockIdx.y;
        mean += bi * T * 8; cov += bi * T * 64; meas += bi * D * 4;
        if (out) out += bi * T * D;
        if (cost) cost += bi * T * D;
        if (conf) conf += bi * T;
    }
    const int t0 = blockIdx.x * GD_TRACKS, cnt = min(GD_TRACKS, T - t0);
    const int nd = only_position ? 2 : 4;
    for (int i = threadIdx.x; i < D * 4; i += blockDim.x) sZ[(i & 3) * D + (i >> 2)] = meas[i];
    if (threadIdx.x < cnt) {
        const int t = t0 + threadIdx.x;
        double m[8], S[4][4], L[4][4];
        for (int i = 0; i < 8; ++i) m[i] = mean[(size_t)t * 8 + i];
        kf_innovation<KIND>(m, cov + (size_t)t * 64, conf ? conf[t] : 0.0, S);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) L[i][j] = 0.0;
        if (only_position) chol_lower<2>(S, L); else chol_lower<4>(S, L);
        for (int i = 0; i < 16; ++i) sL[threadIdx.x][i] = (i >> 2) == (i & 3) ? 1.0 / L[i >> 2][i & 3] : L[i >> 2][i & 3];   // reciprocal diagonal
        for (int i = 0; i < 4; ++i) sM[threadIdx.x][i] = m[i];
    }
    __syncthreads();
    // A warp owns 32 consecutive measurements (kept in registers) and walks the CTA's tracks: the factor and the mean of
    // a track are then the same address for every lane (broadcast shared-memory loads, one wavefront each) - with one
    // (track, measurement) pair per thread the 18 eight-byte loads per pair made the kernel shared-memory bound.
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int ch = threadIdx.x >> 5; ch * 32 < D; ch += nwarps) {
        const int j = ch * 32 + lane;
        const bool valid = j < D;
        double z[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < nd; ++i) z[i] = valid ? sZ[i * D + j] : 0.0;
#pragma unroll 2
        for (int tl = 0; tl < cnt; ++tl) {
            double d[4], zz[4];
            for (int i = 0; i < nd; ++i) d[i] = xsub(z[i], sM[tl][i]);
            double acc = 0.0;
            if (metric == 1) {
                for (int i = 0; i < nd; ++i) acc = i ? xadd(acc, xmul(d[i], d[i])) : xmul(d[i], d[i]);
            } else {
                for (int i = 0; i < nd; ++i) {               // solve_triangular(L, d)
                    double v = d[i];
                    for (int k = 0; k < i; ++k) v -= sL[tl][i * 4 + k] * zz[k];
                    zz[i] = v * sL[tl][i * 4 + i];
                    acc = i ? xadd(acc, xmul(zz[i], zz[i])) : xmul(zz[i], zz[i]);
                }
            }
            if (!valid) continue;
            const size_t o = (size_t)(t0 + tl) * D + j;
            if (cost) {
                // gate_cost_matrix / fuse_motion (matching.py:170-196) without the T x D distance matrix ever reaching HBM
                double c = cost[o];
                if (acc > gate_thr) c = __longlong_as_double(0x7ff0000000000000LL);
                if (fuse) c = xadd(xmul(lambda, c), xmul(xsub(1.0, lambda), acc));
                cost[o] = c;
            } else out[o] = acc;
        }
    }
}

// pairwise similarities / costs: thread per (i, j), j fastest (coalesced stores)
__global__ void box_similarity_kernel(int sim, int n, int m, const double* __restrict__ a, const double* __restrict__ b,
                                      double W, double H, const double* __restrict__ score, int as_distance,
                                      double* __restrict__ out) {
    const size_t total = (size_t)n * m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / m), j = (int)(idx - (size_t)i * m);
        Box A, Bx;
        A.x1 = a[i * 4]; A.y1 = a[i * 4 + 1]; A.x2 = a[i * 4 + 2]; A.y2 = a[i * 4 + 3];
        Bx.x1 = b[j * 4]; Bx.y1 = b[j * 4 + 1]; Bx.x2 = b[j * 4 + 2]; Bx.y2 = b[j * 4 + 3];
        double v;
        switch (sim) {
            case B200TRACK_SIM_GIOU: v = box_giou(A, Bx); break;
            case B200TRACK_SIM_DIOU: v = box_diou(A, Bx); break;
            case B200TRACK_SIM_CIOU: v = box_ciou(A, Bx); break;
            case B200TRACK_SIM_CENTROID: v = box_centroid(A, Bx, W, H); break;
            default: v = box_iou(A, Bx);
        }
        if (as_distance) v = score ? fused_cost(v, score[j]) : xsub(1.0, v);
        out[idx] = v;
    }
}

// embedding_distance (matching.py:145-167): fp32 features, double accumulation, max(0, 1 - cos).
// 16x16 output tile per CTA, K staged through shared memory in chunks of 64.
constexpr int ED_TILE = 16, ED_K = 64;
__global__ void __launch_bounds__(ED_TILE * ED_TILE) embedding_distance_kernel(int n, int m, int dim, const float* __restrict__ a,
                                                                               const float* __restrict__ b, double* __restrict__ out) {
    __shared__ float sa[ED_TILE][ED_K + 1], sb[ED_TILE][ED_K + 1];
    const int tx = threadIdx.x % ED_TILE, ty = threadIdx.x / ED_TILE;
    const int i0 = blockIdx.y * ED_TILE, j0 = blockIdx.x * ED_TILE;
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int k0 = 0; k0 < dim; k0 += ED_K) {
        for (int idx = threadIdx.x; idx < ED_TILE * ED_K; idx += ED_TILE * ED_TILE) {
            const int r = idx / ED_K, k = idx - r * ED_K;
            sa[r][k] = (i0 + r < n && k0 + k < dim) ? a[(size_t)(i0 + r) * dim + k0 + k] : 0.f;
            sb[r][k] = (j0 + r < m && k0 + k < dim) ? b[(size_t)(j0 + r) * dim + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < ED_K; ++k) {
            const double x = (double)sa[ty][k], y = (double)sb[tx][k];
            dot = fma(x, y, dot); na = fma(x, x, na); nb = fma(y, y, nb);
        }
        __syncthreads();
    }
    const int i = i0 + ty, j = j0 + tx;
    if (i < n && j < m) {
        double c = dot / (sqrt(na) * sqrt(nb));
        if (fabs(c) > 1.0) c = copysign(1.0, c);
        out[(size_t)i * m + j] = fmax(0.0, 1.0 - c);
    }
}

// ---- lapjv: one CTA per problem -------------------------------------------------------
struct GlobalCost {
    const double* c; int cols;
    __device__ __forceinline__ double operator()(int i, int j) const { return c[(size_t)i * cols + j]; }
};

// finite cost_limit: prune c > L exactly, then the sparse component solver
__global__ void __launch_bounds__(256) lapjv_sparse_kernel(int rows, int cols, int Rpad, int Cpad, const double* __restrict__ cost,
                                                           double limit, int* __restrict__ x, int* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int words = Cpad / 32;
    LapWork w;
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = raw + off; off = (off + bytes + 15) & ~size_t(15); return p; };
    w.Tmax = Rpad; w.Dmax = Cpad;
    w.u = (double*)take(8 * Rpad); w.v = (double*)take(8 * Cpad); w.dist = (double*)take(8 * Cpad);
    w.adj = (uint32_t*)take(4 * (size_t)words * Rpad);
    w.parent = (int*)take(4 * (Rpad + Cpad)); w.head = (int*)take(4 * Rpad);
    w.rnext = (short*)take(2 * Rpad); w.xr = (short*)take(2 * Rpad); w.yc = (short*)take(2 * Cpad);
    w.pred = (short*)take(2 * Cpad); w.nextc = (short*)take(2 * Cpad); w.mark = (short*)take(2 * Cpad); w.scn = (short*)take(2 * Cpad);
    w.coldeg = (int*)take(4 * Cpad); w.ncomplex = (int*)take(16);
    const double* c = cost + (size_t)blockIdx.x * rows * cols;
    GlobalCost gc{c, cols};
    lap_prepare<256>(w, rows, words);
    __syncthreads();
    for (int task = threadIdx.x; task < rows * words; task += blockDim.x) {
        const int t = task / words, wd = task - t * words;
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int j = wd * 32 + b;
            if (j < cols && c[(size_t)t * cols + j] <= limit) { bits |= 1u << b; atomicAdd(&w.coldeg[j], 1); }
        }
        w.adj[wd * Rpad + t] = bits;
    }
    __syncthreads();
    lap_sparse_solve<256>(w, rows, words, [limit](int) { return limit; }, gc);
    for (int t = threadIdx.x; t < rows; t += blockDim.x) x[(size_t)blockIdx.x * rows + t] = w.xr[t];
    for (int j = threadIdx.x; j < cols; j += blockDim.x) y[(size_t)blockIdx.x * cols + j] = w.yc[j];
}

// cost_limit = +inf (association.py:23): dense problem, see lap_dense.cuh
__global__ void __launch_bounds__(256) lapjv_dense_kernel(int rows, int cols, const double* __restrict__ cost,
                                                          int* __restrict__ x, int* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int NT = 256;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = raw + off; off = (off + bytes + 15) & ~size_t(15); return p; };
    DenseLapM w;
    w.u = (double*)take(8 * rows); w.v = (double*)take(8 * cols); w.dist = (double*)take(8 * cols);
    w.red_v = (double*)take(8 * 32); w.sh_d = (double*)take(8 * 4);
    w.red_i = (int*)take(4 * 32); w.sh_i = (int*)take(4 * 4);
    w.xr = (int*)take(4 * rows); w.claim = (int*)take(4 * rows); w.yc = (int*)take(4 * cols); w.pred = (int*)take(4 * cols);
    w.scn = take(cols);
    const double* c = cost + (size_t)blockIdx.x * rows * cols;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mx = -INF;
    for (int i = tid; i < rows * cols; i += NT) mx = fmax(mx, c[i]);
    for (int d = 16; d; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0) w.red_v[warp] = mx;
    __syncthreads();
    double m2 = w.red_v[0];
    for (int k = 1; k < NT / 32; ++k) m2 = fmax(m2, w.red_v[k]);
    const double lambda = 2.0 * (m2 + 1.0);
    __syncthreads();
    dense_lapm_init<NT>(w, c, cols, rows, cols, lambda);
    dense_lapm_augment<NT>(w, c, cols, rows, cols, lambda);
    for (int t = tid; t < rows; t += NT) x[(size_t)blockIdx.x * rows + t] = w.xr[t];
    for (int j = tid; j < cols; j += NT) y[(size_t)blockIdx.x * cols + j] = w.yc[j];
}

// NearestNeighborDistanceMetric.distance with the cosine metric (matching.py:247-308, :360-378): per track the smallest
// 1 - a_hat . b_hat over its gallery rows, float32 arithmetic like the reference (its np.dot is a float32 BLAS call, so
// the last bits depend on the summation order on both sides).  One thread per (track, detection).
__global__ void __launch_bounds__(128) nn_cosine_kernel(int T, int D, int dim, const float* __restrict__ gal, const int* __restrict__ seg,
                                                        const float* __restrict__ det, double* __restrict__ out) {
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= T * D) return;
    const int t = idx / D, d = idx - t * D;
    const float* b = det + (size_t)d * dim;
    float nb = 0.f;
    for (int i = 0; i < dim; ++i) nb = fmaf(b[i], b[i], nb);
    nb = sqrtf(nb);
    float best = __int_as_float(0x7f800000);
    for (int g = seg[t]; g < seg[t + 1]; ++g) {
        const float* a = gal + (size_t)g * dim;
        float na = 0.f;
        for (int i = 0; i < dim; ++i) na = fmaf(a[i], a[i], na);
        na = sqrtf(na);
        float dot = 0.f;
        for (int i = 0; i < dim; ++i) dot = fmaf(__fdiv_rn(a[i], na), __fdiv_rn(b[i], nb), dot);
        best = fminf(best, 1.0f - dot);
    }
    out[idx] = (double)best;
}

// Track.update's feature smoothing (strongsort/sort/track.py:166-172), float32 like the reference: f = det / |det|,
// s = alpha * trk + (1 - alpha) * f, trk = s / |s|; separately rounded operations, norms accumulated in double (the
// reference's come from BLAS, whose summation order is unspecified).  One warp per row.
__global__ void __launch_bounds__(256) ema_unit_kernel(int n, int dim, float* __restrict__ trk, const float* __restrict__ det, float alpha, float beta) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    float* a = trk + (size_t)r * dim;
    const float* b = det + (size_t)r * dim;
    double acc = 0.0;
    for (int i = lane; i < dim; i += 32) acc += (double)b[i] * b[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    const float nb = sqrtf((float)acc);
    acc = 0.0;
    for (int i = lane; i < dim; i += 32) {
        const float v = __fadd_rn(__fmul_rn(alpha, a[i]), __fmul_rn(beta, __fdiv_rn(b[i], nb)));
        a[i] = v;
        acc += (double)v * v;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    const float ns = sqrtf((float)acc);
    for (int i = lane; i < dim; i += 32) a[i] = __fdiv_rn(a[i], ns);
}

// rows /= |row| in float32 (a new StrongSORT track's first feature, tracker.py:170-172 + track.py:88-90)
__global__ void __launch_bounds__(256) unit_rows_kernel(int n, int dim, float* __restrict__ rows) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    float* a = rows + (size_t)r * dim;
    double acc = 0.0;
    for (int i = lane; i < dim; i += 32) acc += (double)a[i] * a[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    const float nb = sqrtf((float)acc);
    for (int i = lane; i < dim; i += 32) a[i] = __fdiv_rn(a[i], nb);
}

// Track.camera_update (strongsort/sort/track.py:129-138): tlbr of the state, both corners through the 3x3 warp, back to
// xyah - the reference's operation order (with the identity warp this is still not an exact no-op in floating point).
__global__ void camera_update_xyah_kernel(int n, double* __restrict__ mean, const double* __restrict__ warp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* m = mean + (size_t)i * 8;
    // to_tlwh: ret[2] *= ret[3]; ret[:2] -= ret[2:] / 2; to_tlbr: ret[2:] = ret[:2] + ret[2:]
    const double w0 = xmul(m[2], m[3]), h0 = m[3];
    double x1 = xsub(m[0], xdiv(w0, 2.0)), y1 = xsub(m[1], xdiv(h0, 2.0));
    double x2 = xadd(x1, w0), y2 = xadd(y1, h0);
    if (warp) {
        // warp_matrix @ [x, y, 1]: three-term dot products, accumulated left to right
        const double a = warp[0], b = warp[1], c = warp[2], d = warp[3], e = warp[4], f = warp[5];
        const double nx1 = xadd(xadd(xmul(a, x1), xmul(b, y1)), c), ny1 = xadd(xadd(xmul(d, x1), xmul(e, y1)), f);
        const double nx2 = xadd(xadd(xmul(a, x2), xmul(b, y2)), c), ny2 = xadd(xadd(xmul(d, x2), xmul(e, y2)), f);
        x1 = nx1; y1 = ny1; x2 = nx2; y2 = ny2;
    }
    const double w = xsub(x2, x1), h = xsub(y2, y1);
    m[0] = xadd(x1, xdiv(w, 2.0)); m[1] = xadd(y1, xdiv(h, 2.0)); m[2] = xdiv(w, h); m[3] = h;
}

template <class F>
int dispatch_kind(int kind, F&& f) {
    switch (kind) {
        case B200TRACK_KF_XYAH: f(std::integral_constant<int, KF_XYAH>{}); return 0;
        case B200TRACK_KF_XYWH: f(std::integral_constant<int, KF_XYWH>{}); return 0;
        case B200TRACK_KF_XYAH_CONF: f(std::integral_constant<int, KF_XYAH_CONF>{}); return 0;
    }
    set_error("unknown kf_kind");
    return B200TRACK_ERR_ARG;
}

}  // namespace
}  // namespace b200

using namespace b200;

#define LAUNCH_CHECK() B200_CU_TRY(cudaGetLastError())

extern "C" int b200track_kf_initiate(int32_t kind, int32_t n, const double* z, double* mean, double* cov, void* st) {
    if (n < 0 || !z || !mean || !cov) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    int rc = dispatch_kind(kind, [&](auto K) { kf_initiate_kernel<decltype(K)::value><<<(n + KF_TPB - 1) / KF_TPB, KF_TPB, 0, (cudaStream_t)st>>>(n, z, mean, cov); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf_predict(int32_t kind, int32_t n, double* mean, double* cov, void* st) {
    if (n < 0 || !mean || !cov) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    int rc = dispatch_kind(kind, [&](auto K) { kf_predict_kernel<decltype(K)::value><<<(n + KF_TPB - 1) / KF_TPB, KF_THREADS, 0, (cudaStream_t)st>>>(n, mean, cov); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf_apply_warp(int32_t n, double* mean, double* cov, const double* warp, const int32_t* warp_index, void* st) {
    if (n < 0 || !mean || !cov || !warp) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    kf_gmc_kernel<<<(n + KF_TPB - 1) / KF_TPB, KF_THREADS, 0, (cudaStream_t)st>>>(n, mean, cov, warp, warp_index);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_aw_max_metric(int32_t batch, int32_t rows, int32_t cols, const double* emb, double w_assoc, double bottom,
                                       double* out, void* st) {
    if (batch < 0 || rows < 0 || cols < 0 || !out || (!emb && batch * rows * cols > 0)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || rows == 0 || cols == 0) return 0;
    if ((size_t)(rows + cols) * 8 > 200 * 1024) { set_error("aw_max_metric: rows + cols too large"); return B200TRACK_ERR_CAPACITY; }
    const size_t smem = (size_t)(rows + cols) * 8;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(aw_max_metric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return B200TRACK_ERR_CUDA; }
    }
    aw_max_metric_kernel<<<batch, 256, smem, (cudaStream_t)st>>>(rows, cols, emb, w_assoc, bottom, out);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf_project(int32_t kind, int32_t n, const double* mean, const double* cov, const double* conf,
                                    double* pmean, double* pcov, void* st) {
    if (n < 0 || !mean || !cov || !pmean || !pcov) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    int rc = dispatch_kind(kind, [&](auto K) { kf_project_kernel<decltype(K)::value><<<(unsigned)(((size_t)n * 4 + 255) / 256), 256, 0, (cudaStream_t)st>>>(n, mean, cov, conf, pmean, pcov); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf_update(int32_t kind, int32_t n, double* mean, double* cov, const double* z, const double* conf, void* st) {
    if (n < 0 || !mean || !cov || !z) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    int rc = dispatch_kind(kind, [&](auto K) { kf_update_kernel<decltype(K)::value><<<(n + KFU_TPB - 1) / KFU_TPB, KFU_TPB * KF_LANES, 0, (cudaStream_t)st>>>(n, mean, cov, z, conf); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
// internal (strongsort_step.cu): Track.update's filter step for the matched slots of all streams, in place
namespace b200 {
cudaError_t launch_kf_update_masked(int kind, int n_streams, int T, int D, double* mean, double* cov, const double* meas,
                                    const double* conf, const int* sel, cudaStream_t st) {
    const int n = n_streams * T;
    if (n == 0) return cudaSuccess;
    if (dispatch_kind(kind, [&](auto K) {
            kf_update_kernel<decltype(K)::value, true><<<(n + KFU_TPB - 1) / KFU_TPB, KFU_TPB * KF_LANES, 0, st>>>(n, mean, cov, meas, conf, sel, T, D); }))
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
}  // namespace b200
extern "C" int b200track_kf_gating_distance(int32_t kind, int32_t T, int32_t D, const double* mean, const double* cov,
                                            const double* meas, int32_t only_position, int32_t metric, const double* conf,
                                            double* out, void* st) {
    if (T < 0 || D < 0 || !mean || !cov || !meas || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (metric != 0 && metric != 1) { set_error("invalid distance metric"); return B200TRACK_ERR_ARG; }
    if (T == 0 || D == 0) return 0;
    if ((size_t)D * 32 > 200 * 1024) { set_error("gating_distance: more than 6400 measurements per problem"); return B200TRACK_ERR_CAPACITY; }
    int rc = dispatch_kind(kind, [&](auto K) {
        auto kern = kf_gating_kernel<decltype(K)::value>;
        if ((size_t)D * 32 > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D * 32);
        kern<<<(T + GD_TRACKS - 1) / GD_TRACKS, 256, (size_t)D * 32, (cudaStream_t)st>>>(T, D, mean, cov, meas, only_position, metric, conf, out, nullptr, 0, 0.0, 0.0); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_gate_cost(int32_t kind, int32_t batch, int32_t T, int32_t D, const double* mean, const double* cov,
                                   const double* meas, int32_t only_position, int32_t fuse, double lambda, const double* conf,
                                   double* cost, void* st) {
    if (batch < 0 || batch > 65535 || T < 0 || D < 0 || !mean || !cov || !meas || !cost) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || T == 0 || D == 0) return 0;
    if ((size_t)D * 32 > 200 * 1024) { set_error("gate_cost: more than 6400 measurements per problem"); return B200TRACK_ERR_CAPACITY; }
    const double thr = only_position ? 5.9915 : 9.4877;            // chi2inv95[2], chi2inv95[4] (matching.py:15-25)
    dim3 grid((T + GD_TRACKS - 1) / GD_TRACKS, batch);
    int rc = dispatch_kind(kind, [&](auto K) {
        auto kern = kf_gating_kernel<decltype(K)::value>;
        if ((size_t)D * 32 > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D * 32);
        kern<<<grid, 256, (size_t)D * 32, (cudaStream_t)st>>>(T, D, mean, cov, meas, only_position, 0, conf, nullptr, cost, fuse ? 1 : 0, thr, lambda); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf_gating_distance_batched(int32_t kind, int32_t batch, int32_t T, int32_t D, const double* mean,
                                                    const double* cov, const double* meas, int32_t only_position, int32_t metric,
                                                    const double* conf, double* out, void* st) {
    if (batch < 0 || batch > 65535 || T < 0 || D < 0 || !mean || !cov || !meas || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (metric != 0 && metric != 1) { set_error("invalid distance metric"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || T == 0 || D == 0) return 0;
    dim3 grid((T + GD_TRACKS - 1) / GD_TRACKS, batch);
    if ((size_t)D * 32 > 200 * 1024) { set_error("gating_distance: more than 6400 measurements per problem"); return B200TRACK_ERR_CAPACITY; }
    int rc = dispatch_kind(kind, [&](auto K) {
        auto kern = kf_gating_kernel<decltype(K)::value>;
        if ((size_t)D * 32 > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, D * 32);
        kern<<<grid, 256, (size_t)D * 32, (cudaStream_t)st>>>(T, D, mean, cov, meas, only_position, metric, conf, out, nullptr, 0, 0.0, 0.0); });
    if (rc) return rc;
    LAUNCH_CHECK();
    return 0;
}
static int launch_sim(int sim, int n, int m, const double* a, const double* b, double W, double H, const double* score,
                      int as_distance, double* out, void* st) {
    if (n < 0 || m < 0 || !a || !b || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (sim < 0 || sim > B200TRACK_SIM_CENTROID) { set_error("Invalid function specified"); return B200TRACK_ERR_ARG; }
    if (n == 0 || m == 0) return 0;
    const size_t total = (size_t)n * m;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    box_similarity_kernel<<<blocks, 256, 0, (cudaStream_t)st>>>(sim, n, m, a, b, W, H, score, as_distance, out);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_box_similarity(int32_t sim, int32_t n, int32_t m, const double* a, const double* b, double W, double H,
                                        double* out, void* st) {
    return launch_sim(sim, n, m, a, b, W, H, nullptr, 0, out, st);
}
extern "C" int b200track_iou_distance(int32_t n, int32_t m, const double* a, const double* b, const double* score, double* out, void* st) {
    return launch_sim(B200TRACK_SIM_IOU, n, m, a, b, 0, 0, score, 1, out, st);
}
extern "C" int b200track_embedding_distance(int32_t n, int32_t m, int32_t dim, const float* a, const float* b, double* out, void* st) {
    if (n < 0 || m < 0 || dim <= 0 || !a || !b || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0 || m == 0) return 0;
    dim3 grid((m + ED_TILE - 1) / ED_TILE, (n + ED_TILE - 1) / ED_TILE);
    embedding_distance_kernel<<<grid, ED_TILE * ED_TILE, 0, (cudaStream_t)st>>>(n, m, dim, a, b, out);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_nn_cosine_distance(int32_t n_tracks, int32_t n_dets, int32_t dim, const float* gallery, const int32_t* seg,
                                            const float* det, double* out, void* st) {
    if (n_tracks < 0 || n_dets < 0 || dim <= 0 || !seg || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n_tracks == 0 || n_dets == 0) return 0;
    if (!gallery || !det) { set_error("NULL argument"); return B200TRACK_ERR_ARG; }
    const int total = n_tracks * n_dets;
    nn_cosine_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)st>>>(n_tracks, n_dets, dim, gallery, seg, det, out);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_lapjv(int32_t batch, int32_t rows, int32_t cols, const double* cost, double limit, int32_t* x, int32_t* y, void* st) {
    if (batch < 0 || rows < 0 || cols < 0 || !x || !y || (!cost && rows * cols > 0)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (rows > 4096 || cols > 4096) { set_error("lapjv: at most 4096 rows / cols"); return B200TRACK_ERR_CAPACITY; }
    if (batch == 0) return 0;
    if (rows == 0 || cols == 0) {
        if (rows) B200_CU_TRY(cudaMemsetAsync(x, 0xff, sizeof(int32_t) * (size_t)batch * rows, (cudaStream_t)st));
        if (cols) B200_CU_TRY(cudaMemsetAsync(y, 0xff, sizeof(int32_t) * (size_t)batch * cols, (cudaStream_t)st));
        return 0;
    }
    if (limit < __builtin_inf()) {
        const int Rpad = (rows + 31) & ~31, Cpad = (cols + 31) & ~31;
        size_t smem = 8 * (size_t)Rpad + 16 * (size_t)Cpad + 4 * (size_t)(Cpad / 32) * Rpad + 4 * (size_t)(Rpad + Cpad) + 4 * (size_t)Rpad +
                      4 * (size_t)Rpad + 14 * (size_t)Cpad + 16 * 18;
        if (smem > 227 * 1024) { set_error("lapjv: problem too large for shared memory"); return B200TRACK_ERR_CAPACITY; }
        B200_CU_TRY(cudaFuncSetAttribute(lapjv_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lapjv_sparse_kernel<<<batch, 256, smem, (cudaStream_t)st>>>(rows, cols, Rpad, Cpad, cost, limit, x, y);
    } else {
        size_t smem = 8 * (size_t)rows + 16 * (size_t)cols + 8 * 36 + 4 * 36 + 8 * (size_t)rows + 8 * (size_t)cols + cols + 16 * 14;
        if (smem > 227 * 1024) { set_error("lapjv: problem too large for shared memory"); return B200TRACK_ERR_CAPACITY; }
        B200_CU_TRY(cudaFuncSetAttribute(lapjv_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lapjv_dense_kernel<<<batch, 256, smem, (cudaStream_t)st>>>(rows, cols, cost, x, y);
    }
    LAUNCH_CHECK();
    return 0;
}

extern "C" int b200track_ema_unit_features(int32_t n, int32_t dim, float* d_trk, const float* d_det, double alpha, void* st) {
    if (n < 0 || dim <= 0 || (n > 0 && (!d_trk || !d_det))) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    // alpha * feature with a Python float alpha and a float32 array is a float32 product with float32(alpha) (numpy)
    ema_unit_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)st>>>(n, dim, d_trk, d_det, (float)alpha, (float)(1.0 - alpha));
    LAUNCH_CHECK();
    return 0;
}

extern "C" int b200track_camera_update_xyah(int32_t n, double* d_mean, const double* d_warp, void* st) {
    if (n < 0 || (n > 0 && !d_mean)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    camera_update_xyah_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)st>>>(n, d_mean, d_warp);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int b200track_unit_features(int32_t n, int32_t dim, float* d_rows, void* st) {
    if (n < 0 || dim <= 0 || (n > 0 && !d_rows)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    unit_rows_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)st>>>(n, dim, d_rows);
    LAUNCH_CHECK();
    return 0;
}
