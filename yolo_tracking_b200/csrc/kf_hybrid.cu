// Operator-level form of HybridSORT's 9-d score-carrying filter (include/b200track.h: b200track_kf_xyscr_*): the
// reference's object API KalmanFilter.predict / update / unfreeze (boxmot/motion/kalman_filters/hybridsort_kf.py:339-528,
// configured by KalmanBoxTracker.__init__, hybridsort.py:126-150) on dense [n, 9] / [n, 9, 9] arrays, one thread per track.
// The arithmetic is the block form of kf_hybrid.cuh - the same device functions the fused HybridSORT step runs - so the
// covariance must have the structure every covariance of this filter has (four (position, velocity) 2x2 blocks and P_rr;
// anything else raises *d_err).  HBM bound: 2 x 720 B per track.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200track.h"
#include "api_util.h"
#include "kf_hybrid.cuh"

namespace b200 {
namespace {

constexpr int HS_TPB = 64;
constexpr int HS_STRIDE = 91;        // 9 + 81 + 1 pad (odd stride: conflict-free per-thread rows)

__device__ __forceinline__ void hs_stage_in(double* sm, const double* x, const double* P, int base, int cnt) {
    for (int i = threadIdx.x; i < cnt * 9; i += blockDim.x) sm[(i / 9) * HS_STRIDE + (i % 9)] = x[(size_t)base * 9 + i];
    for (int i = threadIdx.x; i < cnt * 81; i += blockDim.x) sm[(i / 81) * HS_STRIDE + 9 + (i % 81)] = P[(size_t)base * 81 + i];
    __syncthreads();
}
__device__ __forceinline__ void hs_stage_out(const double* sm, double* x, double* P, int base, int cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 9; i += blockDim.x) x[(size_t)base * 9 + i] = sm[(i / 9) * HS_STRIDE + (i % 9)];
    for (int i = threadIdx.x; i < cnt * 81; i += blockDim.x) P[(size_t)base * 81 + i] = sm[(i / 81) * HS_STRIDE + 9 + (i % 81)];
}
// dense row of shared memory -> block form; false if an entry outside the structure is not zero.  State order
// [u, v, s, c, r, du, dv, ds, dc]: position i < 4 pairs with velocity i + 5, r (index 4) stands alone.
__device__ __forceinline__ bool hs_load(const double* m, HyKf& k) {
    const double* P = m + 9;
    for (int c = 0; c < 9; ++c) k.x[c] = m[c];
    bool ok = true;
    for (int a = 0; a < 9; ++a)
        for (int b = 0; b < 9; ++b) {
            const bool on = a == b || (a < 4 && b == a + 5) || (b < 4 && a == b + 5);
            if (!on && P[a * 9 + b] != 0.0) ok = false;
        }
    for (int i = 0; i < 4; ++i) {
        k.pp[i] = P[i * 9 + i]; k.pv[i] = P[i * 9 + i + 5]; k.vv[i] = P[(i + 5) * 9 + i + 5];
        if (P[(i + 5) * 9 + i] != k.pv[i]) ok = false;
    }
    k.prr = P[4 * 9 + 4];
    return ok;
}
__device__ __forceinline__ void hs_store(double* m, const HyKf& k) {
    double* P = m + 9;
    for (int c = 0; c < 9; ++c) m[c] = k.x[c];
    for (int i = 0; i < 81; ++i) P[i] = 0.0;
    for (int i = 0; i < 4; ++i) {
        P[i * 9 + i] = k.pp[i]; P[i * 9 + i + 5] = k.pv[i]; P[(i + 5) * 9 + i] = k.pv[i]; P[(i + 5) * 9 + i + 5] = k.vv[i];
    }
    P[4 * 9 + 4] = k.prr;
}

// mode 0: predict; 1: update(z); 2: unfreeze from the saved state in x / P (virtual trajectory from last_z to z over gap
// frames), then update(z) - the order KalmanFilter.update runs them in (hybridsort_kf.py:480-496)
__global__ void __launch_bounds__(HS_TPB) kf_hybrid_kernel(int mode, int n, double* x, double* P, const double* __restrict__ z,
                                                           const double* __restrict__ last_z, const int* __restrict__ gap,
                                                           double* __restrict__ virtual_last, int* err) {
    extern __shared__ double sm[];
    const int base = blockIdx.x * HS_TPB, cnt = min(HS_TPB, n - base), t = threadIdx.x;
    hs_stage_in(sm, x, P, base, cnt);
    if (t < cnt) {
        HyKf k;
        if (!hs_load(sm + t * HS_STRIDE, k) && err) atomicOr(err, 1);
        const size_t q = (size_t)(base + t);
        if (mode == 0) hy_predict_full(k);
        else {
            double zz[5], vz[5];
            for (int c = 0; c < 5; ++c) vz[c] = zz[c] = z[q * 5 + c];
            if (mode == 2) {
                double lz[5];
                for (int c = 0; c < 5; ++c) lz[c] = last_z[q * 5 + c];
                hy_virtual_trajectory(k, lz, zz, gap[q], vz);
            }
            hy_correct(k, zz);
            if (virtual_last) for (int c = 0; c < 5; ++c) virtual_last[q * 5 + c] = vz[c];
        }
        hs_store(sm + t * HS_STRIDE, k);
    }
    hs_stage_out(sm, x, P, base, cnt);
}

int launch_hy(int mode, int n, double* x, double* P, const double* z, const double* last_z, const int* gap, double* vlast, int* err, void* st) {
    if (n < 0 || (n > 0 && (!x || !P))) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    if (mode >= 1 && !z) { set_error("measurements are NULL"); return B200TRACK_ERR_ARG; }
    if (mode == 2 && (!last_z || !gap)) { set_error("last_z / gap are NULL"); return B200TRACK_ERR_ARG; }
    const size_t smem = (size_t)HS_TPB * HS_STRIDE * sizeof(double);          // 46.6 KB: just under the default limit
    kf_hybrid_kernel<<<(n + HS_TPB - 1) / HS_TPB, HS_TPB, smem, (cudaStream_t)st>>>(mode, n, x, P, z, last_z, gap, vlast, err);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}

}  // namespace
}  // namespace b200

extern "C" int b200track_kf_xyscr_predict(int32_t n, double* d_x, double* d_P, int32_t* d_err, void* st) {
    return b200::launch_hy(0, n, d_x, d_P, nullptr, nullptr, nullptr, nullptr, d_err, st);
}
extern "C" int b200track_kf_xyscr_update(int32_t n, double* d_x, double* d_P, const double* d_z, int32_t* d_err, void* st) {
    return b200::launch_hy(1, n, d_x, d_P, d_z, nullptr, nullptr, nullptr, d_err, st);
}
extern "C" int b200track_kf_xyscr_unfreeze_update(int32_t n, double* d_x, double* d_P, const double* d_last_z, const int32_t* d_gap,
                                                  const double* d_z, double* d_virtual_last, int32_t* d_err, void* st) {
    return b200::launch_hy(2, n, d_x, d_P, d_z, d_last_z, d_gap, d_virtual_last, d_err, st);
}
