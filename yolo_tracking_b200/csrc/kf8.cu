// DeepOCSORT's numeric steps as operators (include/b200track.h, "DeepOCSORT operators"): the 8-d [x, y, w, h, vx, vy,
// vw, vh] filter of deep_ocsort.py:103-138 with state-dependent noise, Joseph-form update with an explicit 4x4 inverse
// (deepocsort_kf.py:549-563), the observation-centric re-update of deepocsort_kf.py:433-478 as ONE launch (the whole
// virtual trajectory of a re-found track runs on the device), the velocity-direction + appearance cost of
// association.py:130-172, and the plain embedding similarity of deep_ocsort.py:433.
//
// Covariances are dense here: the camera correction (apply_affine_correction, deepocsort_kf.py:389-405 =
// b200track_kf_apply_warp) breaks the block sparsity.  A CTA stages 64 tracks (mean + covariance, 72 doubles each) in
// shared memory with coalesced loads, one thread then owns one track (row stride 73 doubles: conflict-free), results
// go back with coalesced stores: 2 x 576 B per track, HBM-bound like the other Kalman operators.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200track.h"
#include "api_util.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr int K8_TPB = 64;
constexpr int K8_STRIDE = 73;

__device__ __forceinline__ void k8_in(double* sm, const double* mean, const double* cov, int base, int cnt) {
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) sm[(i >> 3) * K8_STRIDE + (i & 7)] = mean[(size_t)base * 8 + i];
    for (int i = threadIdx.x; i < cnt * 64; i += blockDim.x) sm[(i >> 6) * K8_STRIDE + 8 + (i & 63)] = cov[(size_t)base * 64 + i];
    __syncthreads();
}
__device__ __forceinline__ void k8_out(const double* sm, double* mean, double* cov, int base, int cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) mean[(size_t)base * 8 + i] = sm[(i >> 3) * K8_STRIDE + (i & 7)];
    for (int i = threadIdx.x; i < cnt * 64; i += blockDim.x) cov[(size_t)base * 64 + i] = sm[(i >> 6) * K8_STRIDE + 8 + (i & 63)];
}

// x <- F x, P <- F P F^T + diag(q) with F = [[I, I], [0, I]]; same two-term sums as numpy's products with the 0/1 matrix
__device__ void k8_predict(double* x, double* P, const double* q) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = xadd(x[i], x[i + 4]);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 8; ++j) P[i * 8 + j] = xadd(P[i * 8 + j], P[(i + 4) * 8 + j]);            // F P
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) P[i * 8 + j] = xadd(P[i * 8 + j], P[i * 8 + j + 4]);              // (F P) F^T
    for (int i = 0; i < 8; ++i) P[i * 9] = xadd(P[i * 9], q[i]);
}

// 4x4 inverse by Gauss-Jordan elimination with partial pivoting (np.linalg.inv: LU with partial pivoting)
__device__ void inv4(const double* S, double* out) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = S[i * 4 + j]; a[i][j + 4] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; ++j) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        const double d = xdiv(1.0, a[c][c]);
        for (int j = 0; j < 8; ++j) a[c][j] = xmul(a[c][j], d);
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 8; ++j) a[r][j] = xsub(a[r][j], xmul(f, a[c][j]));
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = a[i][j + 4];
}

// deepocsort_kf.py:549-563 with H = [I 0], R = diag(r): y = z - x[:4]; S = P[:4,:4] + R; K = P[:, :4] S^-1;
// x += K y; P <- (I - K H) P (I - K H)^T + K R K^T
__device__ void k8_update(double* x, double* P, const double* z, const double* r) {
    double S[16], SI[16], K[32], P03[32], M[8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = i == j ? xadd(P[i * 8 + j], r[i]) : P[i * 8 + j];
    inv4(S, SI);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s = xadd(s, xmul(P[i * 8 + k], SI[k * 4 + j]));
            K[i * 4 + j] = s;
        }
    double y[4];
    for (int k = 0; k < 4; ++k) y[k] = xsub(z[k], x[k]);
    for (int i = 0; i < 8; ++i) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s = xadd(s, xmul(K[i * 4 + k], y[k]));
        x[i] = xadd(x[i], s);
    }
    for (int i = 0; i < 32; ++i) P03[i] = P[i];
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < 8; ++j) {                                   // M = row i of (I - K H) P
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s = xadd(s, xmul(K[i * 4 + k], P03[k * 8 + j]));
            M[j] = xsub(P[i * 8 + j], s);
        }
        for (int j = 0; j < 8; ++j) {
            double s = 0.0, t = 0.0;
            for (int k = 0; k < 4; ++k) {
                s = xadd(s, xmul(M[k], K[j * 4 + k]));
                t = xadd(t, xmul(xmul(K[i * 4 + k], r[k]), K[j * 4 + k]));
            }
            P[i * 8 + j] = xadd(xsub(M[j], s), t);
        }
    }
}

// KalmanBoxTracker.predict's filter part (deep_ocsort.py:263-266 + deepocsort_kf.py:340-381): Q from the state's w, h
// BEFORE the motion (new_kf_process_noise :76-80), or Q = I (the filter's default, used inside unfreeze).
// 64 tracks per CTA, 8 lanes per track: lane r owns row r of P (one thread per track was latency-bound at 18 % of the
// warp slots: 106 registers, 64-thread CTAs).
__global__ void __launch_bounds__(K8_TPB * 8) kf8_predict_kernel(int n, double* mean, double* cov, int unit_q) {
    __shared__ double sm[K8_TPB * K8_STRIDE];
    const int base = blockIdx.x * K8_TPB, cnt = min(K8_TPB, n - base);
    const int t = threadIdx.x >> 3, r = threadIdx.x & 7;
    k8_in(sm, mean, cov, base, cnt);
    double row[8], mr = 0.0;
    if (t < cnt) {
        const double* x = sm + t * K8_STRIDE;
        const double* P = x + 8;
        const double ref = x[2 + (r & 1)];                                  // w for the x / w axes, h for y / h
        const double sd = xmul(r < 4 ? 1.0 / 20 : 1.0 / 160, ref);
        const double q = unit_q ? 1.0 : xmul(sd, sd);
        mr = r < 4 ? xadd(x[r], x[r + 4]) : x[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = r < 4 ? xadd(P[r * 8 + j], P[(r + 4) * 8 + j]) : P[r * 8 + j];      // F P
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = xadd(row[j], row[j + 4]);                                            // (F P) F^T
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j == r) row[j] = xadd(row[j], q);
    }
    __syncthreads();
    if (t < cnt) {
        double* x = sm + t * K8_STRIDE;
        x[r] = mr;
#pragma unroll
        for (int c = 0; c < 8; ++c) x[8 + r * 8 + c] = row[c];
    }
    k8_out(sm, mean, cov, base, cnt);
}

// kf.update(z, R) (deepocsort_kf.py:480-569, the observed branch): R = new_kf_measurement_noise(w, h) with the caller's
// w, h (deep_ocsort.py:218: the state BEFORE a possible unfreeze), or R = I when wh is NULL.
// 32 tracks per CTA, 8 lanes per track - lane r owns row r of K, of (I - K H) P and of the result (one thread per track
// needed 210 registers and ran at 0.34 of the HBM roofline).  The first warp factors the 32 innovation matrices, one
// per lane, every lane then solves its own gain row; the gain rows are exchanged through shared memory inside the warp (a track's 8 lanes share a warp).
constexpr int K8U_TPB = 32;
__global__ void __launch_bounds__(K8U_TPB * 8, 5) kf8_update_kernel(int n, double* mean, double* cov, const double* __restrict__ z,
                                                                   const double* __restrict__ wh) {
    __shared__ double sm[K8U_TPB * K8_STRIDE];
    __shared__ double sS[K8U_TPB][19];            // per track: L of S = L L^T (10), 1 / L[i][i] (4), diag R (4); odd stride
    __shared__ double sK[K8U_TPB][33];            // per track: K (8 x 4)
    const int base = blockIdx.x * K8U_TPB, cnt = min(K8U_TPB, n - base);
    const int t = threadIdx.x >> 3, r = threadIdx.x & 7;
    k8_in(sm, mean, cov, base, cnt);
    if ((int)threadIdx.x < cnt) {
        const int q = threadIdx.x;
        const double* P = sm + q * K8_STRIDE + 8;
        double rr[4] = {1.0, 1.0, 1.0, 1.0}, L[4][4];
        if (wh) {
            const double mw = xmul(1.0 / 20, wh[(size_t)(base + q) * 2]), mh = xmul(1.0 / 20, wh[(size_t)(base + q) * 2 + 1]);
            rr[0] = rr[2] = xmul(mw, mw); rr[1] = rr[3] = xmul(mh, mh);
        }
        // S = P[:4, :4] + R is symmetric positive definite: the reference's explicit inverse (np.linalg.inv) times
        // P H^T equals the two triangular solves below to rounding (the 1e-9 bar leaves ~6 digits of room)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double d = xadd(P[j * 8 + j], rr[j]);
#pragma unroll
            for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
            L[j][j] = sqrt(d);
            const double inv = 1.0 / L[j][j];
            sS[q][10 + j] = inv;
#pragma unroll
            for (int i = j + 1; i < 4; ++i) {
                double v = P[i * 8 + j];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
                L[i][j] = v * inv;
            }
        }
        int o = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) sS[q][o++] = L[i][j];
#pragma unroll
        for (int i = 0; i < 4; ++i) sS[q][14 + i] = rr[i];
    }
    __syncthreads();
    double mr = 0.0, K[4], M[8];
    if (t < cnt) {
        const double* m = sm + t * K8_STRIDE;
        const double* P = m + 8;
        {
            double L[4][4], inv[4], y[4];
            int o = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) L[i][j] = sS[t][o++];
#pragma unroll
            for (int i = 0; i < 4; ++i) inv[i] = sS[t][10 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {                 // L y = P[r, :4]
                double v = P[r * 8 + i];
#pragma unroll
                for (int q = 0; q < i; ++q) v -= L[i][q] * y[q];
                y[i] = v * inv[i];
            }
#pragma unroll
            for (int i = 3; i >= 0; --i) {                // L^T K[r, :] = y
                double v = y[i];
#pragma unroll
                for (int q = i + 1; q < 4; ++q) v -= L[q][i] * K[q];
                K[i] = v * inv[i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sK[t][r * 4 + j] = K[j];
        }
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) a = xadd(a, xmul(K[k], xsub(z[(size_t)(base + t) * 4 + k], m[k])));
        mr = xadd(m[r], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) {                                 // row r of (I - K H) P
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) b = xadd(b, xmul(K[k], P[k * 8 + j]));
            M[j] = xsub(P[r * 8 + j], b);
        }
    }
    __syncwarp();                                                     // the track's gain rows are in sK (its 8 lanes share a warp)
    if (t < cnt) {
        // row r of M (I - K H)^T + K R K^T = M[r, j] - sum_k (M[r, k] - K[r, k] R_k) K[j, k]
        double W[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) W[k] = xsub(M[k], xmul(K[k], sS[t][14 + k]));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) b = xadd(b, xmul(W[k], sK[t][j * 4 + k]));
            M[j] = xsub(M[j], b);
        }
    }
    __syncthreads();
    if (t < cnt) {
        double* m = sm + t * K8_STRIDE;
        m[r] = mr;
#pragma unroll
        for (int c = 0; c < 8; ++c) m[8 + r * 8 + c] = M[c];
    }
    k8_out(sm, mean, cov, base, cnt);
}

// unfreeze (deepocsort_kf.py:433-478) on the restored state: box1 = last_measurement, box2 = the new measurement, both
// READ AS [x, y, s, r] although the filter's measurements are [x, y, w, h] (the reference's quirk, kept); `gap` virtual
// boxes on the straight line between them, each applied with R = I, Q = I predicts in between.  last_virtual gets the
// final virtual box: it is the last entry of the filter's observation history afterwards (the next freeze reads it).
__global__ void __launch_bounds__(K8_TPB) kf8_oru_kernel(int n, double* mean, double* cov, const double* __restrict__ box1,
                                                         const double* __restrict__ box2, const int* __restrict__ gap,
                                                         double* __restrict__ last_virtual) {
    __shared__ double sm[K8_TPB * K8_STRIDE];
    const int base = blockIdx.x * K8_TPB, cnt = min(K8_TPB, n - base);
    k8_in(sm, mean, cov, base, cnt);
    if ((int)threadIdx.x < cnt) {
        const int t = base + threadIdx.x;
        double* x = sm + threadIdx.x * K8_STRIDE;
        const double* b1 = box1 + (size_t)t * 4;
        const double* b2 = box2 + (size_t)t * 4;
        const double x1 = b1[0], y1 = b1[1], w1 = sqrt(xmul(b1[2], b1[3])), h1 = sqrt(xdiv(b1[2], b1[3]));
        const double w2 = sqrt(xmul(b2[2], b2[3])), h2 = sqrt(xdiv(b2[2], b2[3]));
        const int g = gap[t];
        const double gd = (double)g;
        const double dx = xdiv(xsub(b2[0], x1), gd), dy = xdiv(xsub(b2[1], y1), gd), dw = xdiv(xsub(w2, w1), gd), dh = xdiv(xsub(h2, h1), gd);
        const double one[8] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        double nb[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < g; ++i) {
            const double k = (double)(i + 1);
            const double w = xadd(w1, xmul(k, dw)), h = xadd(h1, xmul(k, dh));
            nb[0] = xadd(x1, xmul(k, dx)); nb[1] = xadd(y1, xmul(k, dy)); nb[2] = xmul(w, h); nb[3] = xdiv(w, h);
            k8_update(x, x + 8, nb, one);
            if (i != g - 1) k8_predict(x, x + 8, one);
        }
        for (int k = 0; k < 4; ++k) last_virtual[(size_t)t * 4 + k] = nb[k];
    }
    k8_out(sm, mean, cov, base, cnt);
}

// association.py:130-172: cost[d, t] = -(sim[d, t] + angle[d, t] + emb[d, t]) with the velocity-direction term
// angle = valid_t * (pi/2 - |acos(clip(vx_t X + vy_t Y))|) / pi * inertia * score_d, (Y, X) the unit vector from the
// track's previous observation to the detection centre (speed_direction_batch :8-17).  A track without a velocity
// (or without a previous observation) contributes exactly 0, as in numpy (acos(0) = pi/2 there).  dets5 == NULL: no
// angle term (the OCR round's -iou, deep_ocsort.py:478).  The canonical tie-break of the no-limit assignment
// (oracle/lap.py "Ties": + 2^-50 * (d * T + t)) is added here, like csrc/ocsort_step.cu does.
__global__ void __launch_bounds__(256) ocm_cost_kernel(int D, int T, const double* __restrict__ dets5, const double* __restrict__ vel,
                                                       const double* __restrict__ prev5, double inertia, const double* __restrict__ sim,
                                                       const double* __restrict__ emb, double* __restrict__ cost) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= D * T) return;
    const int d = idx / T, t = idx - d * T;
    double s = sim[idx];
    if (dets5) {
        const double* dd = dets5 + (size_t)d * 5;
        const double* pp = prev5 + (size_t)t * 5;
        const double cxd = xdiv(xadd(dd[0], dd[2]), 2.0), cyd = xdiv(xadd(dd[1], dd[3]), 2.0);
        const double cxp = xdiv(xadd(pp[0], pp[2]), 2.0), cyp = xdiv(xadd(pp[1], pp[3]), 2.0);
        const double dx = xsub(cxd, cxp), dy = xsub(cyd, cyp);
        const double norm = xadd(sqrt(xadd(xmul(dx, dx), xmul(dy, dy))), 1e-6);
        const double X = xdiv(dx, norm), Y = xdiv(dy, norm);
        double c = xadd(xmul(vel[t * 2 + 1], X), xmul(vel[t * 2], Y));
        c = fmin(fmax(c, -1.0), 1.0);
        const double PI = 3.141592653589793;
        const double diff = c == 0.0 ? 0.0 : xdiv(xsub(xdiv(PI, 2.0), fabs(acos(c))), PI);
        const double valid = pp[4] < 0.0 ? 0.0 : 1.0;
        s = xadd(s, xmul(xmul(xmul(valid, diff), inertia), dd[4]));
    }
    if (emb) s = xadd(s, emb[idx]);
    cost[idx] = xadd(-s, xmul((double)idx, 0x1p-50));
}

// deep_ocsort.py:433 / :464: dets_embs @ trk_embs.T, fp64 accumulation; a warp per output entry, lanes stride the
// feature axis (coalesced), butterfly sum
__global__ void __launch_bounds__(256) dot_matrix_kernel(int D, int T, int F, const double* __restrict__ a, const double* __restrict__ b,
                                                         double* __restrict__ out) {
    const int idx = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= D * T) return;
    const int d = idx / T, t = idx - d * T;
    const double* ra = a + (size_t)d * F;
    const double* rb = b + (size_t)t * F;
    double s = 0.0;
    for (int k = lane; k < F; k += 32) s = fma(ra[k], rb[k], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[idx] = s;
}

}  // namespace
}  // namespace b200

using namespace b200;
#define LAUNCH_CHECK() B200_CU_TRY(cudaGetLastError())

extern "C" int b200track_kf8_predict(int32_t n, double* mean, double* cov, int32_t unit_q, void* st) {
    if (n < 0 || !mean || !cov) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    kf8_predict_kernel<<<(n + K8_TPB - 1) / K8_TPB, K8_TPB * 8, 0, (cudaStream_t)st>>>(n, mean, cov, unit_q);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf8_update(int32_t n, double* mean, double* cov, const double* z, const double* wh, void* st) {
    if (n < 0 || !mean || !cov || !z) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    kf8_update_kernel<<<(n + K8U_TPB - 1) / K8U_TPB, K8U_TPB * 8, 0, (cudaStream_t)st>>>(n, mean, cov, z, wh);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_kf8_oru(int32_t n, double* mean, double* cov, const double* box1, const double* box2, const int32_t* gap,
                                 double* last_virtual, void* st) {
    if (n < 0 || !mean || !cov || !box1 || !box2 || !gap || !last_virtual) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    kf8_oru_kernel<<<(n + K8_TPB - 1) / K8_TPB, K8_TPB, 0, (cudaStream_t)st>>>(n, mean, cov, box1, box2, gap, last_virtual);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_ocm_cost(int32_t n_dets, int32_t n_tracks, const double* dets5, const double* vel, const double* prev5,
                                  double inertia, const double* sim, const double* emb, double* cost, void* st) {
    if (n_dets < 0 || n_tracks < 0) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n_dets == 0 || n_tracks == 0) return 0;
    if (!sim || !cost || (dets5 && (!vel || !prev5))) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    const long long total = (long long)n_dets * n_tracks;
    ocm_cost_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>(n_dets, n_tracks, dets5, vel, prev5, inertia, sim, emb, cost);
    LAUNCH_CHECK();
    return 0;
}
extern "C" int b200track_dot_matrix(int32_t n, int32_t m, int32_t dim, const double* a, const double* b, double* out, void* st) {
    if (n < 0 || m < 0 || dim < 0) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0 || m == 0) return 0;
    if (!a || !b || !out) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    const long long total = (long long)n * m;
    dot_matrix_kernel<<<(unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)st>>>(n, m, dim, a, b, out);
    LAUNCH_CHECK();
    return 0;
}
