// Constant-velocity Kalman filters (XYAH / XYWH) in the exactly block-sparse form.
//
// With F = [[I, I], [0, I]], H = [I 0] and diagonal Q, R, the reference's dense 8x8
// covariance only ever has non-zeros at (i,i), (i,i+4), (i+4,i), (i+4,i+4): four independent
// position/velocity 2x2 filters (verified on the live reference:
// max |off-structure| == 0.0).  The innovation covariance S is diagonal, so the reference's
// cho_factor / cho_solve collapse to sqrt and two divisions per axis.  Each 2x2 block is kept
// as (pp, pv, vv); the operation order below follows the reference's dense products so the
// surviving terms round the same way:
//   predict  bytetrack_kf.py:155-192 / botsort_kf.py:154-191
//   project  bytetrack_kf.py:126-153 / botsort_kf.py:125-152 / strongsort_kf.py:124-155
//   update   bytetrack_kf.py:194-226 (numpy multi_dot evaluates K (S K^T))
//   initiate bytetrack_kf.py:55-86  / botsort_kf.py:55-86
#pragma once
#include "common.cuh"

namespace b200 {

enum KfKind { KF_XYAH = 0, KF_XYWH = 1, KF_XYAH_CONF = 2 };

struct KfState {
    double m[8];
    double pp[4], pv[4], vv[4];
};

constexpr double KF_W_POS = 1.0 / 20;
constexpr double KF_W_VEL = 1.0 / 160;

// reference scale for axis i: XYAH uses h everywhere (axis 2 is a constant); XYWH uses w for 0,2 and h for 1,3
template <int KIND>
__device__ __forceinline__ double kf_ref(const double* m4, int i) {
    if (KIND == KF_XYWH) return (i & 1) ? m4[3] : m4[2];
    return m4[3];
}

template <int KIND>
__device__ __forceinline__ void kf_initiate(const double* z, KfState& s) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s.m[i] = z[i];
        s.m[i + 4] = 0.0;
        double sp, sv;
        if (KIND != KF_XYWH && i == 2) { sp = 1e-2; sv = 1e-5; }
        else {
            const double r = kf_ref<KIND>(z, i);
            sp = xmul(2 * KF_W_POS, r);
            sv = xmul(10 * KF_W_VEL, r);
        }
        s.pp[i] = xmul(sp, sp);
        s.pv[i] = 0.0;
        s.vv[i] = xmul(sv, sv);
    }
}

template <int KIND>
__device__ __forceinline__ void kf_predict(KfState& s) {
    double qp[4], qv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {            // noise from the state BEFORE the motion step
        double sp, sv;
        if (KIND != KF_XYWH && i == 2) { sp = 1e-2; sv = 1e-5; }
        else {
            const double r = kf_ref<KIND>(s.m, i);
            sp = xmul(KF_W_POS, r);
            sv = xmul(KF_W_VEL, r);
        }
        qp[i] = xmul(sp, sp);
        qv[i] = xmul(sv, sv);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s.m[i] = xadd(s.m[i], s.m[i + 4]);
        const double a = xadd(s.pp[i], s.pv[i]);     // (F P)[i,i]
        const double b = xadd(s.pv[i], s.vv[i]);     // (F P)[i,i+4]
        s.pp[i] = xadd(xadd(a, b), qp[i]);
        s.pv[i] = b;
        s.vv[i] = xadd(s.vv[i], qv[i]);
    }
}

// diagonal of the innovation covariance S = H P H^T + R
template <int KIND>
__device__ __forceinline__ void kf_project_diag(const KfState& s, double conf, double* S) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double sd;
        if (KIND != KF_XYWH && i == 2) sd = 1e-1;
        else sd = xmul(KF_W_POS, kf_ref<KIND>(s.m, i));
        if (KIND == KF_XYAH_CONF) sd = xmul(xsub(1.0, conf), sd);
        S[i] = xadd(s.pp[i], xmul(sd, sd));
    }
}

template <int KIND>
__device__ __forceinline__ void kf_update(KfState& s, const double* z, double conf = 0.0) {
    double S[4];
    kf_project_diag<KIND>(s, conf, S);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // cho_factor / cho_solve of a diagonal S reduce to a division by S[i]; one correctly
        // rounded reciprocal per axis (<= 1 ulp from the reference's (b / sqrt(S)) / sqrt(S))
        const double rs = __drcp_rn(S[i]);
        const double kp = xmul(s.pp[i], rs);
        const double kv = xmul(s.pv[i], rs);
        const double y = xsub(z[i], s.m[i]);
        s.m[i] = xadd(s.m[i], xmul(y, kp));
        s.m[i + 4] = xadd(s.m[i + 4], xmul(y, kv));
        const double skp = xmul(S[i], kp), skv = xmul(S[i], kv);   // (S K^T)
        s.pp[i] = xsub(s.pp[i], xmul(kp, skp));
        s.pv[i] = xsub(s.pv[i], xmul(kp, skv));
        s.vv[i] = xsub(s.vv[i], xmul(kv, skv));
    }
}

// squared Mahalanobis distance of one measurement (gating_distance, bytetrack_kf.py:228-270)
template <int KIND>
__device__ __forceinline__ double kf_maha(const KfState& s, const double* S, const double* z, int ndim) {
    double acc = 0.0;
    for (int i = 0; i < ndim; ++i) {
        const double d = xdiv(xsub(z[i], s.m[i]), sqrt(S[i]));
        const double sq = xmul(d, d);
        acc = (i == 0) ? sq : xadd(acc, sq);
    }
    return acc;
}

}  // namespace b200
