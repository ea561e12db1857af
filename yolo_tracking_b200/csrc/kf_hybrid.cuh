// The 9-d [u, v, s, c, r, du, dv, ds, dc] constant-velocity filter HybridSORT configures (hybridsort.py:126-150 on
// boxmot/motion/kalman_filters/hybridsort_kf.py): F = I + velocity shifts for u, v, s and the score c; H = first five
// rows; R = diag(1, 1, 10, 10, 10); Q = diag(1, 1, 1, 1, 1, .01, .01, 1e-4, 1e-4); P0 = diag(10 x 5, 1e4 x 4).  Every
// covariance reachable from P0 keeps four (position, velocity) 2x2 blocks plus P_rr, so the filter runs on 13 numbers
// instead of 81; one rounding per operation, in the reference's operation order (the same reduction kf_xysr.cuh makes
// for OC-SORT's 7-d filter).
#pragma once
#include "common.cuh"

namespace b200 {
namespace {

struct HyKf {
    double x[9];
    double pp[4], pv[4], vv[4], prr;
};

__device__ __forceinline__ void hy_predict_cov(HyKf& k) {
    const double qv[4] = {0.01, 0.01, 0.0001, 0.0001};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double a = xadd(k.pp[i], k.pv[i]);
        const double b = xadd(k.pv[i], k.vv[i]);
        k.pp[i] = xadd(xadd(a, b), 1.0);
        k.pv[i] = b;
        k.vv[i] = xadd(k.vv[i], qv[i]);
    }
    k.prr = xadd(k.prr, 1.0);
}
__device__ __forceinline__ void hy_predict_full(HyKf& k) {          // kf.predict (no tracker-level guard)
#pragma unroll
    for (int i = 0; i < 4; ++i) k.x[i] = xadd(k.x[i], k.x[i + 5]);
    hy_predict_cov(k);
}
// Joseph-form update, hybridsort_kf.py:498-523, on the block-sparse covariance; z = [x, y, s, score, r]
__device__ __forceinline__ void hy_correct(HyKf& k, const double* z) {
    const double R[5] = {1.0, 1.0, 10.0, 10.0, 10.0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double S = xadd(k.pp[i], R[i]);
        const double si = xdiv(1.0, S);
        const double kp = xmul(k.pp[i], si), kv = xmul(k.pv[i], si);
        const double y = xsub(z[i], k.x[i]);
        k.x[i] = xadd(k.x[i], xmul(kp, y));
        k.x[i + 5] = xadd(k.x[i + 5], xmul(kv, y));
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
        const double S = xadd(k.prr, R[4]);
        const double si = xdiv(1.0, S);
        const double kr = xmul(k.prr, si);
        const double y = xsub(z[4], k.x[4]);
        k.x[4] = xadd(k.x[4], xmul(kr, y));
        const double a = xsub(1.0, kr);
        k.prr = xadd(xmul(xmul(a, k.prr), a), xmul(xmul(kr, R[4]), kr));
    }
}

// unfreeze's virtual trajectory (hybridsort_kf.py:390-436).  The reference unpacks the five-vector [x, y, s, score, r] as
// `x, y, s, r, c`: the SCORE stands where the aspect ratio is meant (w = sqrt(s * score), h = sqrt(s / score)) and the
// aspect ratio is interpolated linearly like a score - kept.  vz receives the last virtual box, which the reference leaves
// at the end of its observation history.
__device__ __forceinline__ bool hy_virtual_trajectory(HyKf& k, const double* lz, const double* z, int g, double* vz) {
    const double x1 = lz[0], y1 = lz[1], s1 = lz[2], r1 = lz[3], c1 = lz[4];
    const double w1 = sqrt(xmul(s1, r1)), h1 = sqrt(xdiv(s1, r1));
    const double w2 = sqrt(xmul(z[2], z[3])), h2 = sqrt(xdiv(z[2], z[3]));
    const double gd = (double)g;
    const double dx = xdiv(xsub(z[0], x1), gd), dy = xdiv(xsub(z[1], y1), gd);
    const double dw = xdiv(xsub(w2, w1), gd), dh = xdiv(xsub(h2, h1), gd), dc = xdiv(xsub(z[4], c1), gd);
    for (int i = 0; i < g; ++i) {
        const double f = (double)(i + 1);
        const double w = xadd(w1, xmul(f, dw)), h = xadd(h1, xmul(f, dh));
        vz[0] = xadd(x1, xmul(f, dx)); vz[1] = xadd(y1, xmul(f, dy)); vz[2] = xmul(w, h); vz[3] = xdiv(w, h);
        vz[4] = xadd(c1, xmul(f, dc));
        hy_correct(k, vz);
        if (i != g - 1) hy_predict_full(k);
    }
    return g > 0;
}

}  // namespace
}  // namespace b200
