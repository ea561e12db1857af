// Box-format conversions and pairwise similarities, one rounding per operation in the
// reference's operation order (boxmot/utils/ops.py:7-97, boxmot/utils/iou.py:6-188).
#pragma once
#include "common.cuh"

namespace b200 {

struct Box { double x1, y1, x2, y2; };

// xyxy2xywh (ops.py:17-20)
__device__ __forceinline__ void xyxy_to_xywh(double x1, double y1, double x2, double y2,
                                             double& xc, double& yc, double& w, double& h) {
    xc = xmul(xadd(x1, x2), 0.5);
    yc = xmul(xadd(y1, y2), 0.5);
    w = xsub(x2, x1);
    h = xsub(y2, y1);
}
// xywh2xyxy (ops.py:36-39)
__device__ __forceinline__ Box xywh_to_xyxy(double xc, double yc, double w, double h) {
    Box b;
    const double hw = xmul(w, 0.5), hh = xmul(h, 0.5);   // x / 2 == x * 0.5 exactly
    b.x1 = xsub(xc, hw); b.y1 = xsub(yc, hh);
    b.x2 = xadd(xc, hw); b.y2 = xadd(yc, hh);
    return b;
}
// xywh2tlwh then tlwh2xyah (ops.py:54-57, :93-96): the measurement an STrack feeds the XYAH filter
__device__ __forceinline__ void xywh_to_xyah(double xc, double yc, double w, double h, double* z) {
    const double hw = xmul(w, 0.5), hh = xmul(h, 0.5);
    z[0] = xadd(xsub(xc, hw), hw);
    z[1] = xadd(xsub(yc, hh), hh);
    z[2] = xdiv(w, h);
    z[3] = h;
}

// Intersection area and IoU of iou_batch (iou.py:13-24): no epsilon, 0/0 -> NaN like numpy.
__device__ __forceinline__ double box_inter(const Box& a, const Box& b) {
    const double ix1 = fmax(a.x1, b.x1), iy1 = fmax(a.y1, b.y1);
    const double ix2 = fmin(a.x2, b.x2), iy2 = fmin(a.y2, b.y2);
    const double iw = fmax(0.0, xsub(ix2, ix1)), ih = fmax(0.0, xsub(iy2, iy1));
    return xmul(iw, ih);
}
__device__ __forceinline__ double iou_from_inter(const Box& a, const Box& b, double inter) {
    const double aa = xmul(xsub(a.x2, a.x1), xsub(a.y2, a.y1));
    const double ab = xmul(xsub(b.x2, b.x1), xsub(b.y2, b.y1));
    return xdiv(inter, xsub(xadd(aa, ab), inter));
}
__device__ __forceinline__ double box_iou(const Box& a, const Box& b) {
    return iou_from_inter(a, b, box_inter(a, b));
}
// Strict-overlap test: false  =>  inter == 0 exactly  =>  iou == 0 (for non-degenerate boxes).
__device__ __forceinline__ bool box_overlap(const Box& a, const Box& b) {
    return (b.x1 < a.x2) & (a.x1 < b.x2) & (b.y1 < a.y2) & (a.y1 < b.y2);
}

// giou_batch (iou.py:28-62) - note the reference subtracts the INTERSECTION, not the union.
__device__ __forceinline__ double box_giou(const Box& a, const Box& b) {
    const double inter = box_inter(a, b);
    const double v = iou_from_inter(a, b, inter);
    const double ew = xsub(fmax(a.x2, b.x2), fmin(a.x1, b.x1));
    const double eh = xsub(fmax(a.y2, b.y2), fmin(a.y1, b.y1));
    const double enc = xmul(ew, eh);
    const double g = xsub(v, xdiv(xsub(enc, inter), enc));
    return xmul(xadd(g, 1.0), 0.5);
}
__device__ __forceinline__ void centre_terms(const Box& a, const Box& b, double& inner, double& outer) {
    const double cxa = xmul(xadd(a.x1, a.x2), 0.5), cya = xmul(xadd(a.y1, a.y2), 0.5);
    const double cxb = xmul(xadd(b.x1, b.x2), 0.5), cyb = xmul(xadd(b.y1, b.y2), 0.5);
    const double dx = xsub(cxa, cxb), dy = xsub(cya, cyb);
    inner = xadd(xmul(dx, dx), xmul(dy, dy));
    const double ex = xsub(fmax(a.x2, b.x2), fmin(a.x1, b.x1));
    const double ey = xsub(fmax(a.y2, b.y2), fmin(a.y1, b.y1));
    outer = xadd(xmul(ex, ex), xmul(ey, ey));
}
// diou_batch (iou.py:65-105)
__device__ __forceinline__ double box_diou(const Box& a, const Box& b) {
    const double v = box_iou(a, b);
    double inner, outer;
    centre_terms(a, b, inner, outer);
    return xmul(xadd(xsub(v, xdiv(inner, outer)), 1.0), 0.5);
}
// ciou_batch (iou.py:108-161); atan is CUDA's (<= 1-2 ulp from glibc's - only matters at ties)
__device__ __forceinline__ double box_ciou(const Box& a, const Box& b) {
    const double v = box_iou(a, b);
    double inner, outer;
    centre_terms(a, b, inner, outer);
    const double w1 = xsub(a.x2, a.x1), h1 = xadd(xsub(a.y2, a.y1), 1.0);
    const double w2 = xsub(b.x2, b.x1), h2 = xadd(xsub(b.y2, b.y1), 1.0);
    const double dth = xsub(atan(xdiv(w2, h2)), atan(xdiv(w1, h1)));
    const double PI = 3.141592653589793;
    const double vv = xmul(xdiv(4.0, xmul(PI, PI)), xmul(dth, dth));
    const double S = xsub(1.0, v);
    const double alpha = xdiv(vv, xadd(S, vv));
    const double c = xsub(xsub(v, xdiv(inner, outer)), xmul(alpha, vv));
    return xmul(xadd(c, 1.0), 0.5);
}
// centroid_batch (iou.py:164-188)
__device__ __forceinline__ double box_centroid(const Box& a, const Box& b, double w, double h) {
    const double cxa = xmul(xadd(a.x1, a.x2), 0.5), cya = xmul(xadd(a.y1, a.y2), 0.5);
    const double cxb = xmul(xadd(b.x1, b.x2), 0.5), cyb = xmul(xadd(b.y1, b.y2), 0.5);
    const double dx = xsub(cxa, cxb), dy = xsub(cya, cyb);
    const double dist = sqrt(xadd(xmul(dx, dx), xmul(dy, dy)));
    const double norm = sqrt(xadd(xmul(w, w), xmul(h, h)));
    return xsub(1.0, xdiv(dist, norm));
}

// iou_distance + fuse_score (matching.py:94-119, :213-221): 1 - (1 - (1 - iou)) * score
__device__ __forceinline__ double fused_cost(double iou, double score) {
    const double c = xsub(1.0, iou);
    const double sim = xsub(1.0, c);
    return xsub(1.0, xmul(sim, score));
}

}  // namespace b200
