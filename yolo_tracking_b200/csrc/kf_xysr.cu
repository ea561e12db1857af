// Operator-level form of OC-SORT's 7-d XYSR filter (include/b200track.h: b200track_kf_xysr_*): the reference's object API
// KalmanFilter.predict / update / unfreeze (boxmot/motion/kalman_filters/ocsort_kf.py:339-526, configured by
// KalmanBoxTracker.__init__, ocsort.py:79-106) on dense [n, 7] / [n, 7, 7] arrays, one thread per track.  The arithmetic is
// the block form of kf_xysr.cuh - the same device functions the fused OC-SORT step runs - so the covariance must have the
// structure every covariance of this filter has (three (position, velocity) 2x2 blocks and P_rr; anything else raises
// *d_err).  HBM bound: 2 x 448 B per track.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200track.h"
#include "api_util.h"
#include "kf_xysr.cuh"

namespace b200 {
namespace {

constexpr int XS_TPB = 64;
constexpr int XS_STRIDE = 57;        // 7 + 49 + 1 pad (odd stride: conflict-free per-thread rows)

__device__ __forceinline__ void xs_stage_in(double* sm, const double* x, const double* P, int base, int cnt) {
    for (int i = threadIdx.x; i < cnt * 7; i += blockDim.x) sm[(i / 7) * XS_STRIDE + (i % 7)] = x[(size_t)base * 7 + i];
    for (int i = threadIdx.x; i < cnt * 49; i += blockDim.x) sm[(i / 49) * XS_STRIDE + 7 + (i % 49)] = P[(size_t)base * 49 + i];
    __syncthreads();
}
__device__ __forceinline__ void xs_stage_out(const double* sm, double* x, double* P, int base, int cnt) {
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 7; i += blockDim.x) x[(size_t)base * 7 + i] = sm[(i / 7) * XS_STRIDE + (i % 7)];
    for (int i = threadIdx.x; i < cnt * 49; i += blockDim.x) P[(size_t)base * 49 + i] = sm[(i / 49) * XS_STRIDE + 7 + (i % 49)];
}
// dense row of shared memory -> block form; false if an entry outside the structure is not zero
__device__ __forceinline__ bool xs_load(const double* m, OcKf& k) {
    const double* P = m + 7;
    for (int c = 0; c < 7; ++c) k.x[c] = m[c];
    bool ok = true;
    for (int a = 0; a < 7; ++a)
        for (int b = 0; b < 7; ++b) {
            const bool on = a == b || (a < 3 && b == a + 4) || (b < 3 && a == b + 4);
            if (!on && P[a * 7 + b] != 0.0) ok = false;
        }
    for (int i = 0; i < 3; ++i) {
        k.pp[i] = P[i * 7 + i]; k.pv[i] = P[i * 7 + i + 4]; k.vv[i] = P[(i + 4) * 7 + i + 4];
        if (P[(i + 4) * 7 + i] != k.pv[i]) ok = false;
    }
    k.prr = P[3 * 7 + 3];
    return ok;
}
__device__ __forceinline__ void xs_store(double* m, const OcKf& k) {
    double* P = m + 7;
    for (int c = 0; c < 7; ++c) m[c] = k.x[c];
    for (int i = 0; i < 49; ++i) P[i] = 0.0;
    for (int i = 0; i < 3; ++i) {
        P[i * 7 + i] = k.pp[i]; P[i * 7 + i + 4] = k.pv[i]; P[(i + 4) * 7 + i] = k.pv[i]; P[(i + 4) * 7 + i + 4] = k.vv[i];
    }
    P[3 * 7 + 3] = k.prr;
}

// mode 0: predict; 1: update(z); 2: unfreeze from the saved state in x / P (virtual trajectory from last_z to z over gap
// frames), then update(z) - the order KalmanFilter.update runs them in (ocsort_kf.py:478-494)
__global__ void __launch_bounds__(XS_TPB) kf_xysr_kernel(int mode, int n, double* x, double* P, const double* __restrict__ z,
                                                         const double* __restrict__ last_z, const int* __restrict__ gap,
                                                         double* __restrict__ virtual_last, int* err) {
    __shared__ double sm[XS_TPB * XS_STRIDE];
    const int base = blockIdx.x * XS_TPB, cnt = min(XS_TPB, n - base), t = threadIdx.x;
    xs_stage_in(sm, x, P, base, cnt);
    if (t < cnt) {
        OcKf k;
        if (!xs_load(sm + t * XS_STRIDE, k) && err) atomicOr(err, 1);
        const size_t q = (size_t)(base + t);
        if (mode == 0) oc_predict_full(k);
        else {
            const double zz[4] = {z[q * 4], z[q * 4 + 1], z[q * 4 + 2], z[q * 4 + 3]};
            double vz[4] = {zz[0], zz[1], zz[2], zz[3]};
            if (mode == 2) {
                const double lz[4] = {last_z[q * 4], last_z[q * 4 + 1], last_z[q * 4 + 2], last_z[q * 4 + 3]};
                oc_virtual_trajectory(k, lz, zz, gap[q], vz);
            }
            oc_correct(k, zz);
            if (virtual_last) for (int c = 0; c < 4; ++c) virtual_last[q * 4 + c] = vz[c];
        }
        xs_store(sm + t * XS_STRIDE, k);
    }
    xs_stage_out(sm, x, P, base, cnt);
}

int launch(int mode, int n, double* x, double* P, const double* z, const double* last_z, const int* gap, double* vlast, int* err, void* st) {
    if (n < 0 || (n > 0 && (!x || !P))) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (n == 0) return 0;
    if (mode >= 1 && !z) { set_error("measurements are NULL"); return B200TRACK_ERR_ARG; }
    if (mode == 2 && (!last_z || !gap)) { set_error("last_z / gap are NULL"); return B200TRACK_ERR_ARG; }
    kf_xysr_kernel<<<(n + XS_TPB - 1) / XS_TPB, XS_TPB, 0, (cudaStream_t)st>>>(mode, n, x, P, z, last_z, gap, vlast, err);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}

}  // namespace
}  // namespace b200

extern "C" int b200track_kf_xysr_predict(int32_t n, double* d_x, double* d_P, int32_t* d_err, void* st) {
    return b200::launch(0, n, d_x, d_P, nullptr, nullptr, nullptr, nullptr, d_err, st);
}
extern "C" int b200track_kf_xysr_update(int32_t n, double* d_x, double* d_P, const double* d_z, int32_t* d_err, void* st) {
    return b200::launch(1, n, d_x, d_P, d_z, nullptr, nullptr, nullptr, d_err, st);
}
extern "C" int b200track_kf_xysr_unfreeze_update(int32_t n, double* d_x, double* d_P, const double* d_last_z, const int32_t* d_gap,
                                                 const double* d_z, double* d_virtual_last, int32_t* d_err, void* st) {
    return b200::launch(2, n, d_x, d_P, d_z, d_last_z, d_gap, d_virtual_last, d_err, st);
}
