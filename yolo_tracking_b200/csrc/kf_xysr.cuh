// The 7-d [x, y, s, r, vx, vy, vs] constant-velocity filter OC-SORT configures (ocsort.py:79-106 on
// boxmot/motion/kalman_filters/ocsort_kf.py): F = I + velocity shifts, H = first four rows, R = diag(1, 1, 10, 10),
// Q = diag(1, 1, 1, 1, .01, .01, 1e-4), P0 = diag(10, 10, 10, 10, 1e4, 1e4, 1e4).  Every covariance reachable from P0 keeps
// three (position, velocity) 2x2 blocks plus P_rr, so the filter runs on 10 numbers instead of 49; one rounding per
// operation, in the reference's operation order.  Shared by the fused frame step (ocsort_step.cu) and the operator
// kernels (kf_xysr.cu).
#pragma once
#include "common.cuh"

namespace b200 {
namespace {

struct OcKf {
    double x[7];
    double pp[3], pv[3], vv[3], prr;
};

__device__ __forceinline__ void oc_predict_cov(OcKf& k) {
    const double qv[3] = {0.01, 0.01, 0.0001};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a = xadd(k.pp[i], k.pv[i]);
        const double b = xadd(k.pv[i], k.vv[i]);
        k.pp[i] = xadd(xadd(a, b), 1.0);
        k.pv[i] = b;
        k.vv[i] = xadd(k.vv[i], qv[i]);
    }
    k.prr = xadd(k.prr, 1.0);
}
__device__ __forceinline__ void oc_predict_full(OcKf& k) {          // kf.predict (no tracker-level guard)
#pragma unroll
    for (int i = 0; i < 3; ++i) k.x[i] = xadd(k.x[i], k.x[i + 4]);
    oc_predict_cov(k);
}
// Joseph-form update, ocsort_kf.py:496-521, on the block-sparse covariance
__device__ __forceinline__ void oc_correct(OcKf& k, const double* z) {
    const double R[4] = {1.0, 1.0, 10.0, 10.0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double S = xadd(k.pp[i], R[i]);
        const double si = xdiv(1.0, S);
        const double kp = xmul(k.pp[i], si), kv = xmul(k.pv[i], si);
        const double y = xsub(z[i], k.x[i]);
        k.x[i] = xadd(k.x[i], xmul(kp, y));
        k.x[i + 4] = xadd(k.x[i + 4], xmul(kv, y));
        const double a = xsub(1.0, kp);
        const double ap00 = xmul(a, k.pp[i]), ap01 = xmul(a, k.pv[i]);
        const double ap10 = xadd(xmul(-kv, k.pp[i]), k.pv[i]), ap11 = xadd(xmul(-kv, k.pv[i]), k.vv[i]);
        const double n00 = xmul(ap00, a);
        const double n01 = xadd(xmul(ap00, -kv), ap01);
        const double n11 = xadd(xmul(ap10, -kv), ap11);
        const double krp = xmul(kp, R[i]), krv = xmul(kv, R[i]);
        k.pp[i] = xadd(n00, xmul(krp, kp));
        k.pv[i] = xadd(n01, xmul(krp, kv));
        k.vv[i] = xadd(n11, xmul(krv, kv));
    }
    {
        const double S = xadd(k.prr, R[3]);
        const double si = xdiv(1.0, S);
        const double kr = xmul(k.prr, si);
        const double y = xsub(z[3], k.x[3]);
        k.x[3] = xadd(k.x[3], xmul(kr, y));
        const double a = xsub(1.0, kr);
        k.prr = xadd(xmul(xmul(a, k.prr), a), xmul(xmul(kr, R[3]), kr));
    }
}

// unfreeze's virtual trajectory (ocsort_kf.py:383-434): from the state saved at the freeze, a straight line of boxes from the
// last measurement `lz` to the new one `z` over `g` frames, each applied as a measurement with a predict in between; the
// caller applies the real measurement on top (KalmanFilter.update does after unfreeze()).  vz receives the last virtual
// box, which the reference leaves at the end of its observation history.
__device__ __forceinline__ bool oc_virtual_trajectory(OcKf& k, const double* lz, const double* z, int g, double* vz) {
    const double x1 = lz[0], y1 = lz[1], s1 = lz[2], r1 = lz[3];
    const double w1 = sqrt(xmul(s1, r1)), h1 = sqrt(xdiv(s1, r1));
    const double w2 = sqrt(xmul(z[2], z[3])), h2 = sqrt(xdiv(z[2], z[3]));
    const double gd = (double)g;
    const double dx = xdiv(xsub(z[0], x1), gd), dy = xdiv(xsub(z[1], y1), gd);
    const double dw = xdiv(xsub(w2, w1), gd), dh = xdiv(xsub(h2, h1), gd);
    for (int i = 0; i < g; ++i) {
        const double f = (double)(i + 1);
        const double w = xadd(w1, xmul(f, dw)), h = xadd(h1, xmul(f, dh));
        vz[0] = xadd(x1, xmul(f, dx)); vz[1] = xadd(y1, xmul(f, dy)); vz[2] = xmul(w, h); vz[3] = xdiv(w, h);
        oc_correct(k, vz);
        if (i != g - 1) oc_predict_full(k);
    }
    return g > 0;
}

}  // namespace
}  // namespace b200
