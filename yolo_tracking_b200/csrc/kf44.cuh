// Constant-velocity Kalman filters on 8-d box states [x, y, w, h, vx, vy, vw, vh] in the form they keep under an external
// camera-motion warp.
//
// With F = [[I, I], [0, I]], H = [I 0] and diagonal Q, R the covariance of such a filter is four independent
// position / velocity 2x2 blocks (kf.cuh).  A camera warp kron(I4, M) with a 2x2 matrix M (STrack.multi_gmc,
// bot_sort.py:95-111; KalmanFilter.apply_affine_correction, deepocsort_kf.py:389-405) mixes x with y and w with h - and
// nothing else: the covariance becomes two independent 4x4 blocks, (x, y, vx, vy) and (w, h, vw, vh), 2 x 10 numbers per
// track instead of the dense 36 (verified on the live reference's moving-camera goldens: cross-group entries are exactly
// 0.0, tests/test_oracle_golden.py::test_camera_warp_leaves_two_independent_4x4_blocks).  Everything the reference does
// to the filter maps a group onto itself, so a group = (m[4], P[4][4]) is filtered on its own:
//   predict        P <- F P F^T + diag(q)                      bytetrack_kf / botsort_kf multi_predict, deepocsort_kf.py:340-381
//   update (chol)  K = P H^T S^-1, P <- P - K S K^T            botsort_kf.py:193-225 (scipy cho_factor / cho_solve)
//   update (Joseph) P <- (I - K H) P (I - K H)^T + K R K^T     deepocsort_kf.py:549-563 (explicit inverse of S)
//   warp           m <- kron(I2, M) m (+ t on the position), P <- B P B^T
// The sums the reference evaluates with BLAS have two to four terms whose order (and fusing) BLAS does not specify, so the
// results agree with it to a few ulp, not bit for bit; parity is checked at 1e-9 relative.
#pragma once
#include "common.cuh"

namespace b200 {

struct G4 {
    double m[4];        // p0, p1, v0, v1
    double P[4][4];     // symmetric
};

// upper-triangle storage order of a group's covariance: 00 01 02 03 11 12 13 22 23 33
__device__ __forceinline__ void g4_load(G4& g, const double* comp0, size_t stride) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) { g.P[i][j] = g.P[j][i] = comp0[(size_t)k * stride]; ++k; }
}
__device__ __forceinline__ void g4_store(const G4& g, double* comp0, size_t stride) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) { comp0[(size_t)k * stride] = g.P[i][j]; ++k; }
}

__device__ __forceinline__ void g4_predict(G4& g, const double* q) {
    g.m[0] += g.m[2]; g.m[1] += g.m[3];
    double A[2][4];                                       // rows 0, 1 of F P
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) A[i][j] = g.P[i][j] + g.P[i + 2][j];
    // (F P) F^T: columns 0, 1 gain columns 2, 3
    const double p00 = A[0][0] + A[0][2], p01 = A[0][1] + A[0][3], p11 = A[1][1] + A[1][3];
    const double c02 = A[0][2], c03 = A[0][3], c12 = A[1][2], c13 = A[1][3];
    g.P[0][0] = p00 + q[0]; g.P[0][1] = g.P[1][0] = p01; g.P[1][1] = p11 + q[1];
    g.P[0][2] = g.P[2][0] = c02; g.P[0][3] = g.P[3][0] = c03;
    g.P[1][2] = g.P[2][1] = c12; g.P[1][3] = g.P[3][1] = c13;
    g.P[2][2] += q[2]; g.P[3][3] += q[3];
}

// gain K[4][2] = P[:, 0:2] S^-1 with S = P[0:2, 0:2] + diag(r); returns S
__device__ __forceinline__ void g4_gain(const G4& g, const double* r, double K[4][2], double S[2][2]) {
    S[0][0] = g.P[0][0] + r[0]; S[0][1] = g.P[0][1]; S[1][0] = g.P[1][0]; S[1][1] = g.P[1][1] + r[1];
    const double det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
    const double id = 1.0 / det;
    const double i00 = S[1][1] * id, i01 = -S[0][1] * id, i10 = -S[1][0] * id, i11 = S[0][0] * id;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        K[i][0] = g.P[i][0] * i00 + g.P[i][1] * i10;
        K[i][1] = g.P[i][0] * i01 + g.P[i][1] * i11;
    }
}

// deepocsort_kf.py:549-563
__device__ __forceinline__ void g4_update_joseph(G4& g, const double* z, const double* r) {
    double K[4][2], S[2][2];
    g4_gain(g, r, K, S);
    const double y0 = z[0] - g.m[0], y1 = z[1] - g.m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) g.m[i] += K[i][0] * y0 + K[i][1] * y1;
    double A[4][4];                                       // (I - K H) P
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) A[i][j] = g.P[i][j] - (K[i][0] * g.P[0][j] + K[i][1] * g.P[1][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            const double v = (A[i][j] - (A[i][0] * K[j][0] + A[i][1] * K[j][1])) + (K[i][0] * r[0] * K[j][0] + K[i][1] * r[1] * K[j][1]);
            g.P[i][j] = g.P[j][i] = v;
        }
}

// botsort_kf.py:193-225: Cholesky solve of the 2x2 innovation covariance, P <- P - K (S K^T)
__device__ __forceinline__ void g4_update_chol(G4& g, const double* z, const double* r) {
    const double s00 = g.P[0][0] + r[0], s10 = g.P[1][0], s11 = g.P[1][1] + r[1];
    const double l00 = sqrt(s00), l10 = s10 / l00, l11 = sqrt(s11 - l10 * l10);
    double K[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                         // solve S k^T = b^T for every row b of P H^T
        const double y0 = g.P[i][0] / l00, y1 = (g.P[i][1] - l10 * y0) / l11;
        const double k1 = y1 / l11, k0 = (y0 - l10 * k1) / l00;
        K[i][0] = k0; K[i][1] = k1;
    }
    const double y0 = z[0] - g.m[0], y1 = z[1] - g.m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) g.m[i] += y0 * K[i][0] + y1 * K[i][1];
    double SK[2][4];                                      // S K^T
#pragma unroll
    for (int j = 0; j < 4; ++j) { SK[0][j] = s00 * K[j][0] + s10 * K[j][1]; SK[1][j] = s10 * K[j][0] + s11 * K[j][1]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            const double v = g.P[i][j] - (K[i][0] * SK[0][j] + K[i][1] * SK[1][j]);
            g.P[i][j] = g.P[j][i] = v;
        }
}

// (u, v) <- M (u, v) (+ t), every operation rounded on its own
__device__ __forceinline__ void g4_warp_pair(const double* M, const double* t, double& u, double& v) {
    const double nu = xadd(xmul(M[0], u), xmul(M[1], v)), nv = xadd(xmul(M[2], u), xmul(M[3], v));
    u = t ? xadd(nu, t[0]) : nu;
    v = t ? xadd(nv, t[1]) : nv;
}

// m <- kron(I2, M) m (+ t on the position pair), P <- B P B^T with B = kron(I2, M); M row-major 2x2
__device__ __forceinline__ void g4_warp(G4& g, const double* M, const double* t) {
    const double a = M[0], b = M[1], c = M[2], d = M[3];
    // the mean with separately rounded operations: callers that need the warped position elsewhere (association boxes)
    // evaluate g4_warp_pair and must get the same bits
    g4_warp_pair(M, t, g.m[0], g.m[1]);
    g4_warp_pair(M, nullptr, g.m[2], g.m[3]);
    double A[4][4];                                       // B P
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        A[0][j] = a * g.P[0][j] + b * g.P[1][j]; A[1][j] = c * g.P[0][j] + d * g.P[1][j];
        A[2][j] = a * g.P[2][j] + b * g.P[3][j]; A[3][j] = c * g.P[2][j] + d * g.P[3][j];
    }
    double R[4][4];                                       // (B P) B^T
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        R[i][0] = A[i][0] * a + A[i][1] * b; R[i][1] = A[i][0] * c + A[i][1] * d;
        R[i][2] = A[i][2] * a + A[i][3] * b; R[i][3] = A[i][2] * c + A[i][3] * d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) g.P[i][j] = g.P[j][i] = R[i][j];
}

}  // namespace b200
