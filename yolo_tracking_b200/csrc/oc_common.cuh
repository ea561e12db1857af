// Pieces shared by the OC-SORT family frame steps (ocsort_step.cu, deepocsort_step.cu): association similarity with the
// exact-zero fast path, the velocity-direction term, block reductions.
#pragma once
#include "boxes.cuh"
#include "common.cuh"

namespace b200 {
namespace {

constexpr int DS_NONE = 0, DS_FREE0 = 1, DS_FREE1 = 2, DS_MATCHED = 3;
constexpr int OCF_ALIVE = 8;
// canonical tie-break of the no-limit assignment (oracle/lap.py "Ties"): cost[r][c] += 2^-50 * (r * C + c)
constexpr double TIE_EPS = 8.8817841970012523e-16;

// a box the pruning rules may reason about: positive, finite extent (everything else is evaluated in full)
__device__ __forceinline__ bool oc_regular_box(double x1, double y1, double x2, double y2) {
    const double w = x2 - x1, h = y2 - y1;
    return w > 0.0 && h > 0.0 && w < 1e100 && h < 1e100 && fabs(x1) < 1e100 && fabs(y1) < 1e100;
}

__device__ __forceinline__ Box oc_x_to_box(double x, double y, double s, double r) {
    const double w = sqrt(xmul(s, r));
    const double h = xdiv(s, w);
    Box b;
    b.x1 = xsub(x, xmul(w, 0.5)); b.y1 = xsub(y, xmul(h, 0.5));
    b.x2 = xadd(x, xmul(w, 0.5)); b.y2 = xadd(y, xmul(h, 0.5));
    return b;
}
__device__ __forceinline__ void oc_box_to_z(double x1, double y1, double x2, double y2, double* z) {
    const double w = xsub(x2, x1), h = xsub(y2, y1);
    z[0] = xadd(x1, xmul(w, 0.5));
    z[1] = xadd(y1, xmul(h, 0.5));
    z[2] = xmul(w, h);
    z[3] = xdiv(w, xadd(h, 1e-6));
}

// run_asso_func (iou.py:191-212).  For iou / giou a pair of disjoint, non-degenerate boxes gives
// exactly +0.0 in the reference's arithmetic (inter = 0, (enc - 0) / enc = 1), so it is returned
// without the divisions; everything else is evaluated in full.
__device__ __forceinline__ double oc_sim(int func, const Box& a, const Box& b, double W, double H) {
    if (func <= 1 && !box_overlap(a, b)) {
        const double un = xadd(xmul(xsub(a.x2, a.x1), xsub(a.y2, a.y1)), xmul(xsub(b.x2, b.x1), xsub(b.y2, b.y1)));
        const double ew = xsub(fmax(a.x2, b.x2), fmin(a.x1, b.x1)), eh = xsub(fmax(a.y2, b.y2), fmin(a.y1, b.y1));
        if (un > 0.0 && un < 1e300 && (func == 0 || (xmul(ew, eh) > 0.0 && xmul(ew, eh) < 1e300))) return 0.0;
    }
    switch (func) {
        case 1: return box_giou(a, b);
        case 2: return box_diou(a, b);
        case 3: return box_ciou(a, b);
        case 4: return box_centroid(a, b, W, H);
        default: return box_iou(a, b);
    }
}

// acos for the velocity-direction term, branch-free.  libm's acos takes different paths for small and large |x| and a warp of
// (track, detection) pairs takes all of them (it was a quarter of the step's instructions at 13 active lanes); here both
// ranges share one polynomial: asin(s) = s + s z g(z) with z = s^2 <= 1/4, where s = |x| for |x| <= 1/2 and
// s = sqrt((1 - |x|) / 2) otherwise (acos(|x|) = 2 asin(s)).  g is the degree-12 Chebyshev interpolant of
// (asin(sqrt z) - sqrt z) / (z sqrt z) on [0, 1/4] (coefficients from a 60-digit fit); the result is within 4.5e-16 of
// numpy's arccos over [-1, 1] (one ulp at pi) - like libm's own distance from glibc, and the term only has to be exact
// at exact ties (see oc_angle).
__device__ __forceinline__ double oc_acos(double x) {
    const double HALF_PI = 1.5707963267948966, PI = 3.141592653589793;
    const double ax = fabs(x);
    const bool big = ax > 0.5;
    const double z = big ? (1.0 - ax) * 0.5 : x * x;
    const double s = big ? sqrt(z) : ax;
    // degree-12 polynomial in z, Estrin's scheme: 5 dependent fma levels instead of Horner's 12 (a cost evaluation is
    // latency bound: the solver and the row reduction wait for single evaluations)
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double p01 = fma(0.07499999999998433, z, 0.16666666666666669);
    const double p23 = fma(0.030381944138531247, z, 0.04464285714635543);
    const double p45 = fma(0.017352392720869973, z, 0.02237217294214989);
    const double p67 = fma(0.011479177415184906, z, 0.013971212973552933);
    const double p89 = fma(0.005457506718640358, z, 0.01032281435018578);
    const double pab = fma(-0.014851887071247204, z, 0.01740087944269402);
    const double q0 = fma(p23, z2, p01), q1 = fma(p67, z2, p45), q2 = fma(pab, z2, p89);
    const double h0 = fma(q1, z4, q0), h1 = fma(0.028757851367421566, z4, q2);
    const double g = fma(h1, z8, h0);
    const double r = fma(s * z, g, s);                     // asin(s)
    return big ? (x > 0.0 ? 2.0 * r : PI - 2.0 * r) : HALF_PI - copysign(r, x);
}

// velocity-direction consistency cost of (track, detection), association.py:134-154
__device__ __forceinline__ double oc_angle(double vy, double vx, double kcx, double kcy, bool valid, double dcx, double dcy,
                                           double inertia, double score) {
    // One reciprocal instead of the reference's two divisions by the norm and a multiplication by 1/pi instead of
    // its division: <= 2 ulp away from numpy's value, like CUDA's acos already is from glibc's; the term only
    // has to be exact at exact ties, and the structural tie (no velocity yet -> exactly 0) is handled by the caller.
    const double HALF_PI = 1.5707963267948966, INV_PI = 0.3183098861837907;
    const double dx = xsub(dcx, kcx), dy = xsub(dcy, kcy);
    const double inv = __drcp_rn(xadd(sqrt(xadd(xmul(dx, dx), xmul(dy, dy))), 1e-6));
    double c = xadd(xmul(vx, xmul(dx, inv)), xmul(vy, xmul(dy, inv)));
    c = fmin(fmax(c, -1.0), 1.0);
    const double diff = xmul(xsub(HALF_PI, oc_acos(c)), INV_PI);           // acos >= 0: the reference's abs() is a no-op
    return xmul(xmul(xmul(valid ? 1.0 : 0.0, diff), inertia), score);
}

// block-wide max of a double and of two ints (all threads get the results)
template <int NT, class SM>
__device__ __forceinline__ void block_max3(SM& sm, double& d, int& a, int& b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        d = fmax(d, __shfl_xor_sync(0xffffffffu, d, s));
        a = max(a, __shfl_xor_sync(0xffffffffu, a, s));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, s));
    }
    __syncthreads();
    if (lane == 0) { sm.red_v[warp] = d; sm.red_i[warp] = a; sm.pred[warp] = b; }
    __syncthreads();
    d = sm.red_v[0]; a = sm.red_i[0]; b = sm.pred[0];
    for (int k = 1; k < NT / 32; ++k) { d = fmax(d, sm.red_v[k]); a = max(a, sm.red_i[k]); b = max(b, sm.pred[k]); }
    __syncthreads();
}

}  // namespace
}  // namespace b200
