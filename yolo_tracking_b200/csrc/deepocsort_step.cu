// DeepOCSORT frame step for many independent streams: one kernel launch per frame, one CTA per stream, one thread per
// tracker slot - the OC-SORT step (ocsort_step.cu) with DeepOCSORT's filter, appearance term and camera correction.
//
// Replaces DeepOCSort.update (boxmot/trackers/deepocsort/deep_ocsort.py:357-520) and what it calls:
//   KalmanBoxTracker.apply_affine_correction :222-241 + KalmanFilter.apply_affine_correction
//       (boxmot/motion/kalman_filters/deepocsort_kf.py:389-405)      an externally estimated 2x3 warp per stream
//   KalmanBoxTracker.predict :246-270 (new_kf: 8-d [x, y, w, h, ...] filter, Q from the state's w, h, :76-80)
//   associate (boxmot/utils/association.py:111-201) with emb_cost = dets_embs @ trk_embs.T (:433), zeroed where the
//       similarity is <= 0, weighted by compute_aw_max_metric (:79-108) or w_association_emb
//   the observation-centric recovery round :456-491
//   KalmanBoxTracker.update :183-216: velocity, observation ring, R from the state's w, h (:83-87), KalmanFilter.update
//       (deepocsort_kf.py:480-569: Joseph form) with the observation-centric re-update of unfreeze (:433-478) - including
//       the reference's quirks: the virtual trajectory reads [x, y, w, h] boxes as [x, y, s, r] and runs with R = I, Q = I;
//       afterwards the filter's observation history ends with the last VIRTUAL box, which is the `last_measurement` of the
//       next freeze; last_observation and observations[age] are one array, so a camera correction moves it twice while it
//       is inside the delta_t window
//   update_emb :218-220 (fp64 blend with alpha from the detection confidence, renormalised), new trackers, output scan.
//
// The covariance is kept as the two 4x4 blocks a camera warp leaves (kf44.cuh).  The association is the matrix-free one
// of the OC-SORT step (pair list of overlapping boxes, pruned row reduction, lap_dense.cuh); the appearance term only
// exists where the similarity is positive, i.e. for iou / giou on the pair list: one warp per listed pair takes the dot
// product of the fp32 detection embedding and the fp64 track embedding, the adaptive weights come from per-row /
// per-column top-2 values gathered with shared-memory atomics on order-preserving keys (the implicit zeros of all other
// pairs are merged in when the weights are formed).  For the dense similarities (diou / ciou / centroid) every pair has
// an appearance term: it is evaluated once into a per-stream scratch block in global memory.
#include "boxes.cuh"
#include "kf44.cuh"
#include "lap_dense.cuh"
#include "layout.h"
#include "oc_common.cuh"
#include "step_params.h"

namespace b200 {
namespace {

template <int TMAX, int DMAX>
struct alignas(16) DoSmem {
    double tbox[4][TMAX];           // predicted box, convert_x_to_bbox_new
    double lbox[4][TMAX];           // last_observation box (placeholder -1)
    double kc[2][TMAX];             // centre of k_previous_obs
    double vel[2][TMAX];            // (vy, vx)
    double dbox[4][DMAX];
    double dconf[DMAX];
    double u[DMAX], v[TMAX];
    double red_v[64];
    unsigned long long scratch[40];
    static constexpr int PCAP = 3 * DMAX;
    double pcost[PCAP], psim[PCAP], pemb[PCAP];
    uint32_t ppair[PCAP];           // (row << 16) | column
    unsigned long long rk1[DMAX], rk2[DMAX], ck1[TMAX], ck2[TMAX];   // top-2 appearance keys per row / column; then the weights
    int rn[DMAX], rc1[DMAX], cn[TMAX], cc1[TMAX];
    int segstart[DMAX];
    short segcnt[DMAX];
    int pred[TMAX], xr[DMAX], yc[TMAX], claim[DMAX], partner[TMAX];
    int red_i[64];
    int rowcnt[DMAX], rowmatch[DMAX], colcnt[TMAX];
    int misc[8];
    short hd[DMAX], ht[TMAX], dmatch[DMAX], tmatch[TMAX], ud[DMAX], ut[TMAX], erow[TMAX], freelist[TMAX];
    float4 cboxf[TMAX];
    double dred[8 * (TMAX / 32)];
    int ired[8 * (TMAX / 32)];
    unsigned char kvalid[TMAX], alive[TMAX], dstate[DMAX], tdeg[TMAX], ddeg[DMAX], rowfull[DMAX], rowused[TMAX];
};

__device__ __forceinline__ unsigned long long dkey(double v) {            // order-preserving map double -> uint64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// compute_aw_max_metric's per-row / per-column factor (association.py:84-104) from the two largest LISTED values (first
// F with multiplicity c1, second S2), n_listed of them, and the zeros of the other total - n_listed entries
__device__ __forceinline__ double aw_weight(unsigned long long k1, int c1, unsigned long long k2, int n_listed, int total, double bottom) {
    if (total < 2) return 1.0;
    const double NINF = -__longlong_as_double(0x7ff0000000000000LL);
    double v1 = NINF, v2 = NINF;
    auto push = [&](double v) { if (v > v1) { v2 = v1; v1 = v; } else if (v > v2) v2 = v; };
    if (n_listed > 0) {
        const double F = dunkey(k1);
        push(F);
        if (c1 >= 2) push(F);
        if (n_listed > c1) push(dunkey(k2));
    }
    const int zeros = total - n_listed;
    if (zeros >= 1) push(0.0);
    if (zeros >= 2) push(0.0);
    if (v1 == 0.0) return 0.0;
    return xsub(1.0, xdiv(fmax(xsub(xdiv(v2, v1), bottom), 0.0), xsub(1.0, bottom)));
}

// Cost of (row r = high detection hd[r], column c = live tracker ht[c]) in the first association:
// -((similarity + direction term) + appearance term) plus the canonical tie-break (association.py:130-172).
template <class SM>
struct DoCost1 {
    const SM& sm;
    int func, Cn, ldE;
    double W, H, inertia;
    bool sparse;                    // iou / giou with a non-negative threshold: a disjoint pair has similarity exactly +0.0
    const double* E;                // dense weighted appearance terms [r * ldE + c] (diou / ciou / centroid with embeddings), or null
    __device__ __forceinline__ Box dbox(int j) const { return Box{sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]}; }
    __device__ __forceinline__ Box tbox(int sl) const { return Box{sm.tbox[0][sl], sm.tbox[1][sl], sm.tbox[2][sl], sm.tbox[3][sl]}; }
    __device__ __forceinline__ double bound(int j) const { return xmul(0.5, fabs(xmul(inertia, sm.dconf[j]))) + 1e-6; }
    __device__ __forceinline__ bool free_pair(int j, int sl) const {
        return sparse && !sm.ddeg[j] && !sm.tdeg[sl] && !box_overlap(dbox(j), tbox(sl));
    }
    __device__ __forceinline__ double sim(int r, int c) const { return oc_sim(func, dbox(sm.hd[r]), tbox(sm.ht[c]), W, H); }
    __device__ __forceinline__ double from_sim(int r, int c, double sv, double e) const {
        const int j = sm.hd[r], sl = sm.ht[c];
        double ang = 0.0;
        const double vy = sm.vel[0][sl], vx = sm.vel[1][sl];
        if (sm.kvalid[sl] && !(vx == 0.0 && vy == 0.0))
            ang = oc_angle(vy, vx, sm.kc[0][sl], sm.kc[1][sl], true, xdiv(xadd(sm.dbox[0][j], sm.dbox[2][j]), 2.0),
                           xdiv(xadd(sm.dbox[1][j], sm.dbox[3][j]), 2.0), inertia, sm.dconf[j]);
        return xadd(-xadd(xadd(sv, ang), e), xmul((double)(r * Cn + c), TIE_EPS));
    }
    __device__ __forceinline__ double operator()(int r, int c) const {
        const int j = sm.hd[r], sl = sm.ht[c];
        if (free_pair(j, sl)) return from_sim(r, c, 0.0, 0.0);
        if (sparse) {
            const int n = sm.segcnt[r], b0 = sm.segstart[r];
            for (int k = 0; k < n; ++k)
                if ((int)(sm.ppair[b0 + k] & 0xffff) == c) return sm.pcost[b0 + k];
        }
        return from_sim(r, c, sim(r, c), E ? E[(size_t)r * ldE + c] : 0.0);
    }
    __device__ __forceinline__ double lower(int r, int c) const {
        const int j = sm.hd[r];
        return free_pair(j, sm.ht[c]) ? -bound(j) : -__longlong_as_double(0x7ff0000000000000LL);
    }
};

template <int NT, class SM>
__device__ __forceinline__ DenseLap make_dense(SM& sm) {
    DenseLap w;
    w.u = sm.u; w.v = sm.v; w.pred = sm.pred; w.xr = sm.xr; w.yc = sm.yc; w.claim = sm.claim;
    w.red_v = sm.red_v; w.red_i = sm.red_i; w.freerow = sm.rowmatch; w.dbg = nullptr;
    return w;
}

__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

// new_kf_process_noise (deep_ocsort.py:76-80) split per group: q = ((p ref0)^2, (p ref1)^2, (v ref0)^2, (v ref1)^2)
__device__ __forceinline__ void do_process_noise(double w, double h, double* qA, double* qB) {
    const double pw = xmul(1.0 / 20, w), ph = xmul(1.0 / 20, h), vw = xmul(1.0 / 160, w), vh = xmul(1.0 / 160, h);
    qA[0] = qB[0] = xmul(pw, pw); qA[1] = qB[1] = xmul(ph, ph);
    qA[2] = qB[2] = xmul(vw, vw); qA[3] = qB[3] = xmul(vh, vh);
}

// 2x3 warp applied to the two corner points of a box (deep_ocsort.py:226-241)
__device__ __forceinline__ void warp_box(double* b, const double* wp) {
    const double x1 = wp[0] * b[0] + wp[1] * b[1] + wp[2], y1 = wp[3] * b[0] + wp[4] * b[1] + wp[5];
    const double x2 = wp[0] * b[2] + wp[1] * b[3] + wp[2], y2 = wp[3] * b[2] + wp[4] * b[3] + wp[5];
    b[0] = x1; b[1] = y1; b[2] = x2; b[3] = y2;
}

template <int NT, int TMAX, int DMAX>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : (NT >= 224 ? 2 : (NT == 128 ? 4 : 6))))
deepocsort_step_kernel(const StepParams p) {
    static_assert(NT == TMAX && DMAX <= NT, "one thread per tracker slot; detections fit one pass");
    using SM = DoSmem<TMAX, DMAX>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int s = blockIdx.x, tid = threadIdx.x, t = tid, lane = tid & 31, warp = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int* counts = p.counts + 4 * s;
    int n0 = counts[0];
    const int alive0 = counts[1], id0 = counts[2], frame = counts[3] + 1;
    const bool packed = p.det_off != nullptr;          // packed frames (step_params.h)
    int roff = 0;
    if (packed) roff = p.det_off[s];
    const int nd_in = packed ? p.det_off[s + 1] - roff : p.ndets[s];
    int nd = nd_in;
    int err = 0;
    const int dcap = min(DMAX, p.max_dets), tcap = min(TMAX, p.max_tracks);
    if (nd > dcap) { nd = dcap; err |= B200_ERR_DET_OVERFLOW; }
    if (nd < 0) nd = 0;
    double* gf = p.state_f + (size_t)s * B200_DO_NF * TMAX;
    int* gi = p.state_i + (size_t)s * B200_DO_NI * TMAX;
    const double thr = p.iou_thresh, W = p.img_w, H = p.img_h;
    const int func = p.asso_func, F = p.feat_dim;
    const bool emb_on = !p.embedding_off && F > 0;
    double* pool = emb_on ? p.emb_pool + (size_t)s * TMAX * F : nullptr;
    const float* dfeat = !emb_on ? nullptr : (packed ? p.feats + (size_t)roff * F : p.feats + (size_t)s * p.max_dets * F);
    const double* wp = p.warps ? p.warps + 6 * s : nullptr;


    // ---- HBM -> shared memory: detections ------------------------------------------------------------
    {
        const double* g = packed ? p.dets + (size_t)roff * 6 : p.dets + (size_t)s * p.max_dets * 6;
        const float* g32 = p.dets32 ? p.dets32 + (size_t)roff * 6 : nullptr;
        for (int i = tid; i < nd * 6; i += NT) {
            const double val = g32 ? (double)g32[i] : g[i];
            const int j = i / 6, c = i - 6 * j;
            if (c < 4) sm.dbox[c][j] = val;
            else if (c == 4) sm.dconf[j] = val;
        }
    }
    const double* dets_g = packed ? p.dets + (size_t)roff * 6 : p.dets + (size_t)s * p.max_dets * 6;
    const float* dets32_g = p.dets32 ? p.dets32 + (size_t)roff * 6 : nullptr;
    auto det_cls = [&](int j) -> double { return dets32_g ? (double)dets32_g[j * 6 + 5] : dets_g[j * 6 + 5]; };
    // detections above det_thresh: the only ones that can start a tracker this frame
    __syncthreads();
    const int nhigh = __syncthreads_count(tid < nd && sm.dconf[tid] > p.det_thresh);
    // ---- compaction on demand (see ocsort_step.cu) ------------------------------------------------
    if (n0 > alive0 && n0 + nhigh > tcap) {              // uniform
        bool lv = false;
        if (t < n0) lv = gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE;
        unsigned long long tt;
        const int dst = (int)block_exscan<NT>(lv ? 1ull : 0ull, sm.scratch, tt);
        for (int c0 = 0; c0 < B200_DO_NF; c0 += 8) {
            double tmp[8];
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_DO_NF) tmp[c] = gf[(c0 + c) * TMAX + t];
            }
            __syncthreads();
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_DO_NF) gf[(c0 + c) * TMAX + dst] = tmp[c];
            }
            __syncthreads();
        }
        int itmp[B200_DO_NI];
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_DO_NI; ++c) itmp[c] = gi[c * TMAX + t];
        }
        __syncthreads();
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_DO_NI; ++c) gi[c * TMAX + dst] = itmp[c];
        }
        __syncthreads();
        n0 = (int)tt;
    }

    // ---- tracker side, thread t = slot t: camera correction, predict -----------------------------
    int fl = 0, age = 0, tsu = 0, streak = 0;
    bool live = false;
    double pw = 0.0, ph = 0.0;                       // w, h of the predicted state (R of this frame's update)
    if (t < n0) {
        fl = gi[B200_OCI_FLAGS * TMAX + t];
        live = fl & OCF_ALIVE;
    }
    if (live) {
        age = gi[B200_OCI_AGE * TMAX + t];
        tsu = gi[B200_OCI_TSU * TMAX + t];
        streak = gi[B200_OCI_STREAK * TMAX + t];
        sm.erow[t] = (short)gi[B200_DOI_EROW * TMAX + t];
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = gf[(B200_DO_X + c) * TMAX + t];
        G4 gA, gB;
        g4_load(gA, gf + (size_t)B200_DO_PA * TMAX + t, TMAX);
        g4_load(gB, gf + (size_t)B200_DO_PB * TMAX + t, TMAX);
        gA.m[0] = x[0]; gA.m[1] = x[1]; gA.m[2] = x[4]; gA.m[3] = x[5];
        gB.m[0] = x[2]; gB.m[1] = x[3]; gB.m[2] = x[6]; gB.m[3] = x[7];
        const bool hasobs = fl & B200_OCF_HASOBS;
        double l[4] = {-1.0, -1.0, -1.0, -1.0};
        const double conf0 = gf[B200_DO_CONF * TMAX + t];
        int ra[3] = {-1, -1, -1};
        double rb[3][4];
        if (hasobs) {
#pragma unroll
            for (int c = 0; c < 4; ++c) l[c] = gf[(B200_DO_LAST + c) * TMAX + t];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ra[q] = gi[(B200_OCI_RINGAGE + q) * TMAX + t];
#pragma unroll
                for (int c = 0; c < 4; ++c) rb[q][c] = gf[(B200_DO_RING + 4 * q + c) * TMAX + t];
            }
        }
        if (wp) {
            // KalmanBoxTracker.apply_affine_correction (deep_ocsort.py:222-241).  last_observation IS observations[its age]
            // (one array): it moves once as last_observation (if its sum is positive) and once more while its age is inside
            // the delta_t window - the newest ring entry stands for both.
            if (hasobs) {
                int newest = 0;
                if (ra[1] > ra[newest]) newest = 1;
                if (ra[2] > ra[newest]) newest = 2;
                if (l[0] + l[1] + l[2] + l[3] + conf0 > 0.0) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) if (q == newest) warp_box(rb[q], wp);
                }
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (ra[q] >= 0 && ra[q] >= age - p.delta_t && ra[q] <= age) warp_box(rb[q], wp);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q == newest) { l[0] = rb[q][0]; l[1] = rb[q][1]; l[2] = rb[q][2]; l[3] = rb[q][3]; }
#pragma unroll
                    for (int c = 0; c < 4; ++c) gf[(B200_DO_RING + 4 * q + c) * TMAX + t] = rb[q][c];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) gf[(B200_DO_LAST + c) * TMAX + t] = l[c];
            }
            const double M[4] = {wp[0], wp[1], wp[3], wp[4]}, tv[2] = {wp[2], wp[5]};
            g4_warp(gA, M, tv);
            g4_warp(gB, M, nullptr);
            if (!(fl & B200_OCF_OBSERVED) && (fl & B200_OCF_SAVED)) {    // the frozen state moves too (deepocsort_kf.py:399-404)
                G4 sA, sB;
                double sx[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) sx[c] = gf[(B200_DO_SX + c) * TMAX + t];
                g4_load(sA, gf + (size_t)B200_DO_SPA * TMAX + t, TMAX);
                g4_load(sB, gf + (size_t)B200_DO_SPB * TMAX + t, TMAX);
                sA.m[0] = sx[0]; sA.m[1] = sx[1]; sA.m[2] = sx[4]; sA.m[3] = sx[5];
                sB.m[0] = sx[2]; sB.m[1] = sx[3]; sB.m[2] = sx[6]; sB.m[3] = sx[7];
                g4_warp(sA, M, tv);
                g4_warp(sB, M, nullptr);
                const double so[8] = {sA.m[0], sA.m[1], sB.m[0], sB.m[1], sA.m[2], sA.m[3], sB.m[2], sB.m[3]};
#pragma unroll
                for (int c = 0; c < 8; ++c) gf[(B200_DO_SX + c) * TMAX + t] = so[c];
                g4_store(sA, gf + (size_t)B200_DO_SPA * TMAX + t, TMAX);
                g4_store(sB, gf + (size_t)B200_DO_SPB * TMAX + t, TMAX);
                double lm[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) lm[c] = gf[(B200_DO_LASTZ + c) * TMAX + t];
                const double a0 = M[0] * lm[0] + M[1] * lm[1] + tv[0], a1 = M[2] * lm[0] + M[3] * lm[1] + tv[1];
                const double a2 = M[0] * lm[2] + M[1] * lm[3], a3 = M[2] * lm[2] + M[3] * lm[3];
                gf[(B200_DO_LASTZ + 0) * TMAX + t] = a0; gf[(B200_DO_LASTZ + 1) * TMAX + t] = a1;
                gf[(B200_DO_LASTZ + 2) * TMAX + t] = a2; gf[(B200_DO_LASTZ + 3) * TMAX + t] = a3;
            }
        }
        // KalmanBoxTracker.predict (deep_ocsort.py:246-270)
        if (xadd(gB.m[0], gB.m[2]) <= 0.0) gB.m[2] = 0.0;
        if (xadd(gB.m[1], gB.m[3]) <= 0.0) gB.m[3] = 0.0;
        if (fl & B200_DOF_FROZEN) { gB.m[2] = 0.0; gB.m[3] = 0.0; }
        double qA[4], qB[4];
        do_process_noise(gB.m[0], gB.m[1], qA, qB);
        g4_predict(gA, qA);
        g4_predict(gB, qB);
        age += 1;
        if (tsu > 0) streak = 0;
        tsu += 1;
        pw = gB.m[0]; ph = gB.m[1];
        // the predicted state goes straight back to the slot (the update phase re-reads it: L2)
        const double xo[8] = {gA.m[0], gA.m[1], gB.m[0], gB.m[1], gA.m[2], gA.m[3], gB.m[2], gB.m[3]};
#pragma unroll
        for (int c = 0; c < 8; ++c) gf[(B200_DO_X + c) * TMAX + t] = xo[c];
        g4_store(gA, gf + (size_t)B200_DO_PA * TMAX + t, TMAX);
        g4_store(gB, gf + (size_t)B200_DO_PB * TMAX + t, TMAX);
        Box b;                                                           // convert_x_to_bbox_new
        b.x1 = xsub(xo[0], xdiv(xo[2], 2.0)); b.y1 = xsub(xo[1], xdiv(xo[3], 2.0));
        b.x2 = xadd(xo[0], xdiv(xo[2], 2.0)); b.y2 = xadd(xo[1], xdiv(xo[3], 2.0));
        sm.tbox[0][t] = b.x1; sm.tbox[1][t] = b.y1; sm.tbox[2][t] = b.x2; sm.tbox[3][t] = b.y2;
        sm.tdeg[t] = !oc_regular_box(b.x1, b.y1, b.x2, b.y2);
        if (isnan(b.x1) || isnan(b.y1) || isnan(b.x2) || isnan(b.y2)) { live = false; fl &= ~OCF_ALIVE; }   // :417-421
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.lbox[c][t] = l[c];
        sm.vel[0][t] = gf[(B200_DO_VEL + 0) * TMAX + t];
        sm.vel[1][t] = gf[(B200_DO_VEL + 1) * TMAX + t];
        // k_previous_obs (deep_ocsort.py:15-23): oldest observation among ages age-delta_t .. age-1, else the newest
        double kb[4] = {l[0], l[1], l[2], l[3]};
        if (hasobs) {
            for (int dt = p.delta_t; dt >= 1; --dt) {
                const int a = age - dt;
                if (a < 0) continue;
                const int q = a % 3;
                if (dt <= 3 && ra[q] == a) {
#pragma unroll
                    for (int qq = 0; qq < 3; ++qq) if (qq == q) { kb[0] = rb[qq][0]; kb[1] = rb[qq][1]; kb[2] = rb[qq][2]; kb[3] = rb[qq][3]; }
                    break;
                }
            }
        }
        sm.kc[0][t] = xdiv(xadd(kb[0], kb[2]), 2.0);
        sm.kc[1][t] = xdiv(xadd(kb[1], kb[3]), 2.0);
        sm.kvalid[t] = hasobs;
    }
    if (t < TMAX) { sm.alive[t] = live; sm.tmatch[t] = -1; sm.rowused[t] = 0; }
    if (tid < DMAX) { sm.dmatch[tid] = -1; sm.rowcnt[tid] = 0; sm.rowmatch[tid] = -1; }
    __syncthreads();
    if (tid < DMAX) sm.dstate[tid] = (tid < nd && sm.dconf[tid] > p.det_thresh) ? DS_FREE0 : DS_NONE;    // deep_ocsort.py:377-379

    // compact row (high detections) and column (alive trackers) lists
    int R, Cn;
    {
        const bool isrow = tid < nd && sm.dconf[tid] > p.det_thresh;
        const unsigned long long val = (isrow ? 1ull : 0ull) | (live ? (1ull << 16) : 0ull);
        unsigned long long tot;
        const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);
        if (isrow) sm.hd[ex & 0xffff] = (short)tid;
        if (live) sm.ht[(ex >> 16) & 0xffff] = (short)t;
        R = (int)(tot & 0xffff); Cn = (int)((tot >> 16) & 0xffff);
        __syncthreads();
    }

    // dot product of detection j's fp32 embedding and the fp64 embedding of the tracker in slot sl, by one warp
    auto emb_dot = [&](int j, int sl) -> double {
        const float* a = dfeat + (size_t)j * F;
        const double* b = pool + (size_t)sm.erow[sl] * F;
        double acc = 0.0;
        for (int i = lane * 4; i < F; i += 128) {
            const float4 av = *reinterpret_cast<const float4*>(a + i);
            const double2 b0 = *reinterpret_cast<const double2*>(b + i), b1 = *reinterpret_cast<const double2*>(b + i + 2);
            acc += (double)av.x * b0.x + (double)av.y * b0.y + (double)av.z * b1.x + (double)av.w * b1.y;
        }
        return warp_sum_d(acc);
    };

    // ---- first round: associate(dets, trks, ...) with the appearance term -----------------------------------
    if (tid < DMAX && tid < nd) sm.ddeg[tid] = !oc_regular_box(sm.dbox[0][tid], sm.dbox[1][tid], sm.dbox[2][tid], sm.dbox[3][tid]);
    double* Edense = (emb_on && !(func <= 1 && thr >= 0.0)) ? p.scratch + (size_t)s * TMAX * DMAX : nullptr;
    DoCost1<SM> cost1{sm, func, Cn, TMAX, W, H, p.inertia, func <= 1 && thr >= 0.0, nullptr};
    if (R > 0 && Cn > 0) {
        constexpr int NCH = TMAX / 32;
        double amax = 0.0;                           // bound of the direction term over the rows
        if (tid < Cn) {
            sm.colcnt[tid] = 0;
            const int sl = sm.ht[tid];
            const float FINF = __int_as_float(0x7f800000);
            sm.cboxf[tid] = sm.tdeg[sl] ? make_float4(-FINF, -FINF, FINF, FINF)
                                        : make_float4(__double2float_rd(sm.tbox[0][sl]), __double2float_rd(sm.tbox[1][sl]),
                                                      __double2float_ru(sm.tbox[2][sl]), __double2float_ru(sm.tbox[3][sl]));
            sm.ck1[tid] = 0ull; sm.ck2[tid] = 0ull; sm.cn[tid] = 0; sm.cc1[tid] = 0;
        }
        if (tid < R) { sm.rk1[tid] = 0ull; sm.rk2[tid] = 0ull; sm.rn[tid] = 0; sm.rc1[tid] = 0; }
        if (tid == 0) { sm.misc[0] = 0; sm.misc[1] = SM::PCAP; }
        __syncthreads();
        // A. pair list of the boxes that may overlap (ocsort_step.cu); B. similarity + threshold counts per pair
        int np = 0;
        if (cost1.sparse) {
            for (int r = warp; r < R; r += NT / 32) {
                const int j = sm.hd[r];
                const float FINF = __int_as_float(0x7f800000);
                const bool ddeg = sm.ddeg[j];
                const float dx1 = ddeg ? -FINF : __double2float_rd(sm.dbox[0][j]), dy1 = ddeg ? -FINF : __double2float_rd(sm.dbox[1][j]);
                const float dx2 = ddeg ? FINF : __double2float_ru(sm.dbox[2][j]), dy2 = ddeg ? FINF : __double2float_ru(sm.dbox[3][j]);
                uint32_t bits[NCH];
                int total = 0;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    bits[k] = 0u;
                    if (k * 32 < Cn) {                        // uniform: chunks past the live columns cost nothing
                        const int c = k * 32 + lane;
                        bool cand = false;
                        if (c < Cn) {
                            const float4 tf = sm.cboxf[c];
                            cand = (tf.x < dx2) & (dx1 < tf.z) & (tf.y < dy2) & (dy1 < tf.w);
                        }
                        bits[k] = __ballot_sync(0xffffffffu, cand);
                        total += __popc(bits[k]);
                    }
                }
                int base = 0;
                if (lane == 0 && total) base = atomicAdd(&sm.misc[0], total);
                base = __shfl_sync(0xffffffffu, base, 0);
                const bool fits = base + total <= SM::PCAP;
                if (fits) {
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        if (bits[k]) {
                            if (bits[k] & (1u << lane)) sm.ppair[base + __popc(bits[k] & ((1u << lane) - 1u))] = ((uint32_t)r << 16) | (uint32_t)(k * 32 + lane);
                            base += __popc(bits[k]);
                        }
                    }
                    base -= total;
                }
                if (lane == 0) {
                    sm.segstart[r] = base; sm.segcnt[r] = fits ? (short)total : (short)-1;
                    if (!fits) atomicMin(&sm.misc[1], base);
                }
            }
            __syncthreads();
            np = min(sm.misc[0], sm.misc[1]);
            if (emb_on && sm.misc[0] > SM::PCAP) err |= B200_ERR_BOT_CAPACITY;       // the appearance weights need every overlapping pair
            for (int k = tid; k < np; k += NT) {
                const uint32_t pr = sm.ppair[k];
                const int r = pr >> 16, c = pr & 0xffff;
                const double sv = cost1.sim(r, c);
                if (sv > thr) { atomicAdd(&sm.colcnt[c], 1); atomicAdd(&sm.rowcnt[r], 1); sm.rowmatch[r] = c; }
                sm.psim[k] = sv;
                sm.pemb[k] = 0.0;
            }
        } else {
            // dense similarity: threshold counts over every pair
            for (int idx = tid; idx < R * Cn; idx += NT) {
                const int r = idx / Cn, c = idx - r * Cn;
                if (cost1.sim(r, c) > thr) { atomicAdd(&sm.colcnt[c], 1); atomicAdd(&sm.rowcnt[r], 1); sm.rowmatch[r] = c; }
            }
        }
        __syncthreads();
        int ccnt = tid < Cn ? sm.colcnt[tid] : 0;
        int rc = tid < R ? sm.rowcnt[tid] : 0;
        double dummy = 0.0;
        block_max3<NT>(sm, dummy, ccnt, rc);
        const bool shortcut = (rc == 1 && ccnt == 1);                          // association.py:157-159
        if (shortcut) {
            if (tid < R) sm.xr[tid] = sm.rowcnt[tid] == 1 ? sm.rowmatch[tid] : -1;
            __syncthreads();
        } else {
            // ---- appearance term: raw values where the similarity is positive, adaptive weights, weighted values -------
            double emax = 0.0;
            if (emb_on) {
                const int nent = cost1.sparse ? np : R * Cn;
                for (int k = warp; k < nent; k += NT / 32) {                  // one warp per entry
                    int r, c;
                    double sv;
                    if (cost1.sparse) { const uint32_t pr = sm.ppair[k]; r = pr >> 16; c = pr & 0xffff; sv = sm.psim[k]; }
                    else { r = k / Cn; c = k - r * Cn; sv = cost1.sim(r, c); }
                    double e = 0.0;
                    if (sv > 0.0) {
                        e = emb_dot(sm.hd[r], sm.ht[c]);
                        if (lane == 0 && !p.aw_off) {
                            const unsigned long long key = dkey(e);
                            atomicMax(&sm.rk1[r], key); atomicMax(&sm.ck1[c], key);
                            atomicAdd(&sm.rn[r], 1); atomicAdd(&sm.cn[c], 1);
                        }
                    }
                    if (lane == 0) { if (cost1.sparse) sm.pemb[k] = e; else Edense[(size_t)r * TMAX + c] = e; }
                }
                __syncthreads();
                if (!p.aw_off) {
                    for (int k = tid; k < nent; k += NT) {
                        int r, c;
                        double e;
                        bool pos;
                        if (cost1.sparse) { const uint32_t pr = sm.ppair[k]; r = pr >> 16; c = pr & 0xffff; e = sm.pemb[k]; pos = sm.psim[k] > 0.0; }
                        else { r = k / Cn; c = k - r * Cn; e = Edense[(size_t)r * TMAX + c]; pos = cost1.sim(r, c) > 0.0; }
                        if (!pos) continue;
                        const unsigned long long key = dkey(e);
                        if (key == sm.rk1[r]) atomicAdd(&sm.rc1[r], 1); else atomicMax(&sm.rk2[r], key);
                        if (key == sm.ck1[c]) atomicAdd(&sm.cc1[c], 1); else atomicMax(&sm.ck2[c], key);
                    }
                    __syncthreads();
                    // weights: w_emb = w_assoc * row factor * column factor (association.py:84-107); they replace the keys
                    double wr = 0.0, wc = 0.0;
                    if (tid < R) wr = aw_weight(sm.rk1[tid], sm.rc1[tid], sm.rk2[tid], sm.rn[tid], Cn, p.aw_param);
                    if (tid < Cn) wc = aw_weight(sm.ck1[tid], sm.cc1[tid], sm.ck2[tid], sm.cn[tid], R, p.aw_param);
                    __syncthreads();
                    if (tid < R) sm.rk1[tid] = (unsigned long long)__double_as_longlong(wr);
                    if (tid < Cn) sm.ck1[tid] = (unsigned long long)__double_as_longlong(wc);
                    __syncthreads();
                }
                for (int k = tid; k < nent; k += NT) {
                    int r, c;
                    double e;
                    if (cost1.sparse) { const uint32_t pr = sm.ppair[k]; r = pr >> 16; c = pr & 0xffff; e = sm.pemb[k]; }
                    else { r = k / Cn; c = k - r * Cn; e = Edense[(size_t)r * TMAX + c]; }
                    double wgt = p.w_assoc_emb;
                    if (!p.aw_off) wgt = xmul(xmul(wgt, __longlong_as_double((long long)sm.rk1[r])), __longlong_as_double((long long)sm.ck1[c]));
                    e = p.aw_off ? xmul(e, wgt) : xmul(wgt, e);
                    emax = fmax(emax, fabs(e));
                    if (cost1.sparse) sm.pemb[k] = e; else Edense[(size_t)r * TMAX + c] = e;
                }
                __syncthreads();
                cost1.E = Edense;
            }
            // ---- row reduction: C. minimum over the row's pair segment; D. rows it does not settle, in full ----------
            if (cost1.sparse) {
                for (int k = tid; k < np; k += NT) {
                    const uint32_t pr = sm.ppair[k];
                    sm.pcost[k] = cost1.from_sim(pr >> 16, pr & 0xffff, sm.psim[k], sm.pemb[k]);
                }
                __syncthreads();
                if (tid < R) {
                    const int r = tid, n = sm.segcnt[r], b0 = sm.segstart[r];
                    const double bound = cost1.bound(sm.hd[r]);
                    amax = bound;
                    double m = INF;
                    int a = -1;
                    for (int k = 0; k < n; ++k) {
                        const double cst = sm.pcost[b0 + k];
                        const int c = sm.ppair[b0 + k] & 0xffff;
                        if (cst < m || (cst == m && c < a)) { m = cst; a = c; }
                    }
                    sm.u[r] = m; sm.claim[r] = a;
                    sm.rowfull[r] = n < 0 ? 2 : (m < -bound ? 0 : 1);
                }
            } else if (tid < R) {
                sm.rowfull[tid] = 2;
                amax = cost1.bound(sm.hd[tid]);
            }
            __syncthreads();
            int nfull = 0;
            if (warp == 0) {
                for (int r0 = 0; r0 < R; r0 += 32) {
                    const int r = r0 + lane;
                    const bool fr = r < R && sm.rowfull[r] != 0;
                    const uint32_t mk = __ballot_sync(0xffffffffu, fr);
                    if (fr) sm.partner[nfull + __popc(mk & ((1u << lane) - 1u))] = r;
                    nfull += __popc(mk);
                }
                if (lane == 0) sm.misc[2] = nfull;
            }
            __syncthreads();
            nfull = sm.misc[2];
            const int nch = (Cn + 31) >> 5;
            for (int f0 = 0; f0 < nfull; f0 += 8) {
                const int nf = min(8, nfull - f0);
                for (int task = warp; task < nf * nch; task += NT / 32) {
                    const int fi = task / nch, k = task - fi * nch;
                    const int r = sm.partner[f0 + fi], mode = sm.rowfull[r], j = sm.hd[r];
                    const int c = k * 32 + lane;
                    double m = INF;
                    int a = -1;
                    if (c < Cn) {
                        const int sl = sm.ht[c];
                        const bool fp = cost1.free_pair(j, sl);
                        if (mode == 2 || fp) {
                            m = fp ? cost1.from_sim(r, c, 0.0, 0.0) : cost1.from_sim(r, c, cost1.sim(r, c), Edense ? Edense[(size_t)r * TMAX + c] : 0.0);
                            a = c;
                        }
                    }
#pragma unroll
                    for (int d = 16; d; d >>= 1) {
                        const double om = __shfl_xor_sync(0xffffffffu, m, d);
                        const int oa = __shfl_xor_sync(0xffffffffu, a, d);
                        if (om < m || (om == m && oa >= 0 && (a < 0 || oa < a))) { m = om; a = oa; }
                    }
                    if (lane == 0) { sm.dred[fi * NCH + k] = m; sm.ired[fi * NCH + k] = a; }
                }
                __syncthreads();
                if (tid < nf) {
                    const int r = sm.partner[f0 + tid], mode = sm.rowfull[r];
                    double m = mode == 1 ? sm.u[r] : INF;
                    int a = mode == 1 ? sm.claim[r] : -1;
                    for (int k = 0; k < nch; ++k) {
                        const double om = sm.dred[tid * NCH + k];
                        const int oa = sm.ired[tid * NCH + k];
                        if (oa >= 0 && (om < m || (om == m && (a < 0 || oa < a)))) { m = om; a = oa; }
                    }
                    sm.u[r] = m; sm.claim[r] = a;
                }
                __syncthreads();
            }
            int d1 = 0, d2 = 0;
            block_max3<NT>(sm, amax, d1, d2);
            block_max3<NT>(sm, emax, d1, d2);
            if (tid == 0) atomicAdd(&p.stats[0], 1ull);
            const DenseLap w = make_dense<NT>(sm);
            const double lambda = 2.0 * (amax + emax + 1.0);      // >= lapjv's 2 * (max cost + 1): same assignment (lap_dense.cuh)
            dense_lap_init<NT>(w, R, Cn, lambda);
            dense_lap_augment<NT>(w, cost1, R, Cn, lambda);
        }
        // matched pairs below the similarity threshold fall back to unmatched (association.py:187-193)
        if (tid < R) {
            const int c = sm.xr[tid];
            const int j = sm.hd[tid];
            if (c >= 0) {
                const int sl = sm.ht[c];
                if (cost1.sim(tid, c) < thr) sm.dstate[j] = DS_FREE1;
                else { sm.dstate[j] = DS_MATCHED; sm.dmatch[j] = (short)sl; sm.tmatch[sl] = (short)j; }
            }
        }
        __syncthreads();
    }

    // ---- observation-centric recovery round on the last observations (deep_ocsort.py:456-491) -------------------
    if (t < TMAX) sm.partner[t] = -1;
    __syncthreads();
    if (tid < R && Cn > 0 && sm.xr[tid] >= 0 && sm.dstate[sm.hd[tid]] == DS_FREE1) sm.partner[sm.ht[sm.xr[tid]]] = tid;   // slot -> row of its partner
    __syncthreads();
    bool ocr_ran = false;
    {
        // unmatched lists in associate()'s order (association.py:179-193): never matched first (ascending), then the
        // members of low-similarity matches in match order (= ascending detection index)
        const int ds = tid < DMAX ? sm.dstate[tid] : DS_NONE;
        const bool ufree = live && sm.tmatch[t] < 0;
        const bool ufree1 = ufree && sm.partner[t] >= 0;
        const bool ufree0 = ufree && !ufree1;
        const bool d0 = ds == DS_FREE0, dlate = ds == DS_FREE1;
        const unsigned long long val = (d0 ? 1ull : 0ull) | (dlate ? (1ull << 16) : 0ull) | (ufree0 ? (1ull << 32) : 0ull) |
                                       (ds == DS_FREE1 ? (1ull << 48) : 0ull);
        unsigned long long tot;
        const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);
        const int n0d = (int)(tot & 0xffff), n1d = (int)((tot >> 16) & 0xffff), n0t = (int)((tot >> 32) & 0xffff);
        if (d0) sm.ud[ex & 0xffff] = (short)tid;
        if (dlate) sm.ud[n0d + ((ex >> 16) & 0xffff)] = (short)tid;
        if (ufree0) sm.ut[(ex >> 32) & 0xffff] = (short)t;
        if (ds == DS_FREE1) sm.claim[tid] = (int)((ex >> 48) & 0xffff);          // rank of this detection among FREE1
        __syncthreads();
        if (ufree1) sm.ut[n0t + sm.claim[sm.hd[sm.partner[t]]]] = (short)t;
        const int nr = n0d + n1d, nc = n0t + (int)((tot >> 48) & 0xffff);
        __syncthreads();
        if (nr > 0 && nc > 0) {                                        // uniform
            auto sim2 = [&](int r, int c) -> double {
                const int j = sm.ud[r], sl = sm.ut[c];
                const Box tb = {sm.lbox[0][sl], sm.lbox[1][sl], sm.lbox[2][sl], sm.lbox[3][sl]};
                const Box db = {sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]};
                return oc_sim(func, db, tb, W, H);
            };
            struct Cost2 {
                decltype(sim2)& sim;
                int nc;
                __device__ __forceinline__ double operator()(int r, int c) const { return xadd(-sim(r, c), xmul((double)(r * nc + c), TIE_EPS)); }
                __device__ __forceinline__ double lower(int, int) const { return -__longlong_as_double(0x7ff0000000000000LL); }
            } cost2{sim2, nc};
            double smax = -1e300;
            for (int r = warp; r < nr; r += NT / 32) {
                double m = INF;
                int a = -1;
                for (int c = lane; c < nc; c += 32) {
                    const double sv = sim2(r, c);
                    const double cst = xadd(-sv, xmul((double)(r * nc + c), TIE_EPS));
                    smax = fmax(smax, sv);
                    if (cst < m) { m = cst; a = c; }
                }
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const double om = __shfl_xor_sync(0xffffffffu, m, d);
                    const int oa = __shfl_xor_sync(0xffffffffu, a, d);
                    if (om < m || (om == m && oa >= 0 && (a < 0 || oa < a))) { m = om; a = oa; }
                }
                if (lane == 0) { sm.u[r] = m; sm.claim[r] = a; }
            }
            int d1 = 0, d2 = 0;
            block_max3<NT>(sm, smax, d1, d2);
            if (smax > thr) {
                ocr_ran = true;
                if (tid == 0) atomicAdd(&p.stats[1], 1ull);
                const DenseLap w = make_dense<NT>(sm);
                const double lambda = 2.0 * (1.0 + 1e-6 + 1.0);
                dense_lap_init<NT>(w, nr, nc, lambda);
                dense_lap_augment<NT>(w, cost2, nr, nc, lambda);
                if (tid < nr) {
                    const int c = sm.xr[tid];
                    if (c >= 0) {
                        const int j = sm.ud[tid], sl = sm.ut[c];
                        if (!(sim2(tid, c) < thr)) {
                            if (sm.dstate[j] != DS_NONE) sm.dstate[j] = DS_MATCHED;
                            sm.dmatch[j] = (short)sl; sm.tmatch[sl] = (short)j;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---- Kalman update and bookkeeping, thread t = slot t (deep_ocsort.py:183-216) ---------------------------------
    int hits = 0, det_ind = 0, tid_id = 0;
    double conf = 0.0, cls = 0.0;
    double xo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        G4 gA, gB;
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = gf[(B200_DO_X + c) * TMAX + t];       // the predicted state of phase 1
        g4_load(gA, gf + (size_t)B200_DO_PA * TMAX + t, TMAX);
        g4_load(gB, gf + (size_t)B200_DO_PB * TMAX + t, TMAX);
        gA.m[0] = x[0]; gA.m[1] = x[1]; gA.m[2] = x[4]; gA.m[3] = x[5];
        gB.m[0] = x[2]; gB.m[1] = x[3]; gB.m[2] = x[6]; gB.m[3] = x[7];
        hits = gi[B200_OCI_HITS * TMAX + t];
        det_ind = gi[B200_OCI_DET * TMAX + t];
        tid_id = gi[B200_OCI_ID * TMAX + t];
        conf = gf[B200_DO_CONF * TMAX + t];
        cls = gf[B200_DO_CLS * TMAX + t];
        const int j = sm.tmatch[t];
        bool store_state = false;
        if (j >= 0) {
            const double b0 = sm.dbox[0][j], b1 = sm.dbox[1][j], b2 = sm.dbox[2][j], b3 = sm.dbox[3][j];
            const bool hasobs = fl & B200_OCF_HASOBS;
            const double lsum = hasobs ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
            if (lsum >= 0.0) {                                          // speed_direction(previous_box, bbox)
                const double cx2 = xdiv(xadd(b0, b2), 2.0), cy2 = xdiv(xadd(b1, b3), 2.0);
                const double dy = xsub(cy2, sm.kc[1][t]), dx = xsub(cx2, sm.kc[0][t]);
                const double norm = xadd(sqrt(xadd(xmul(dy, dy), xmul(dx, dx))), 1e-6);
                gf[(B200_DO_VEL + 0) * TMAX + t] = xdiv(dy, norm);
                gf[(B200_DO_VEL + 1) * TMAX + t] = xdiv(dx, norm);
            }
            conf = sm.dconf[j]; cls = det_cls(j); det_ind = j;
            gf[(B200_DO_LAST + 0) * TMAX + t] = b0; gf[(B200_DO_LAST + 1) * TMAX + t] = b1;
            gf[(B200_DO_LAST + 2) * TMAX + t] = b2; gf[(B200_DO_LAST + 3) * TMAX + t] = b3;
            const int rs = age % 3;
            gf[(B200_DO_RING + 4 * rs + 0) * TMAX + t] = b0; gf[(B200_DO_RING + 4 * rs + 1) * TMAX + t] = b1;
            gf[(B200_DO_RING + 4 * rs + 2) * TMAX + t] = b2; gf[(B200_DO_RING + 4 * rs + 3) * TMAX + t] = b3;
            gi[(B200_OCI_RINGAGE + rs) * TMAX + t] = age;
            // convert_bbox_to_z_new; R from the state's w, h BEFORE a possible unfreeze (deep_ocsort.py:211-213)
            const double bw = xsub(b2, b0), bh = xsub(b3, b1);
            const double z[4] = {xadd(b0, xdiv(bw, 2.0)), xadd(b1, xdiv(bh, 2.0)), bw, bh};
            const double mw = xmul(1.0 / 20, pw), mh = xmul(1.0 / 20, ph);
            const double rr[2] = {xmul(mw, mw), xmul(mh, mh)};
            bool virt = false;
            double vz[4] = {0, 0, 0, 0};
            if (!(fl & B200_OCF_OBSERVED) && (fl & B200_OCF_SAVED)) {   // unfreeze (deepocsort_kf.py:433-478)
                double sx[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) sx[c] = gf[(B200_DO_SX + c) * TMAX + t];
                g4_load(gA, gf + (size_t)B200_DO_SPA * TMAX + t, TMAX);
                g4_load(gB, gf + (size_t)B200_DO_SPB * TMAX + t, TMAX);
                gA.m[0] = sx[0]; gA.m[1] = sx[1]; gA.m[2] = sx[4]; gA.m[3] = sx[5];
                gB.m[0] = sx[2]; gB.m[1] = sx[3]; gB.m[2] = sx[6]; gB.m[3] = sx[7];
                const double x1 = gf[(B200_DO_LASTZ + 0) * TMAX + t], y1 = gf[(B200_DO_LASTZ + 1) * TMAX + t];
                const double s1 = gf[(B200_DO_LASTZ + 2) * TMAX + t], r1 = gf[(B200_DO_LASTZ + 3) * TMAX + t];
                // the reference reads both [x, y, w, h] boxes as [x, y, s, r] here - kept
                const double w1 = sqrt(xmul(s1, r1)), h1 = sqrt(xdiv(s1, r1));
                const double w2 = sqrt(xmul(z[2], z[3])), h2 = sqrt(xdiv(z[2], z[3]));
                const int g = tsu;                                      // index2 - index1 of history_obs
                const double gd = (double)g;
                const double dx = xdiv(xsub(z[0], x1), gd), dy = xdiv(xsub(z[1], y1), gd);
                const double dw = xdiv(xsub(w2, w1), gd), dh = xdiv(xsub(h2, h1), gd);
                const double one2[2] = {1.0, 1.0}, one4[4] = {1.0, 1.0, 1.0, 1.0};
                for (int i = 0; i < g; ++i) {
                    const double f = (double)(i + 1);
                    const double w = xadd(w1, xmul(f, dw)), h = xadd(h1, xmul(f, dh));
                    vz[0] = xadd(x1, xmul(f, dx)); vz[1] = xadd(y1, xmul(f, dy)); vz[2] = xmul(w, h); vz[3] = xdiv(w, h);
                    g4_update_joseph(gA, vz, one2);
                    g4_update_joseph(gB, vz + 2, one2);
                    if (i != g - 1) { g4_predict(gA, one4); g4_predict(gB, one4); }
                }
                virt = g > 0;
                fl &= ~B200_OCF_SAVED;
                atomicAdd(&p.stats[2], 1ull);
            }
            fl |= B200_OCF_OBSERVED | B200_OCF_HASOBS;
            fl &= ~B200_DOF_FROZEN;
            g4_update_joseph(gA, z, rr);                                 // the real measurement on top
            g4_update_joseph(gB, z + 2, rr);
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_DO_LASTZ + c) * TMAX + t] = virt ? vz[c] : z[c];
            tsu = 0; hits += 1; streak += 1;
            sm.lbox[0][t] = b0; sm.lbox[1][t] = b1; sm.lbox[2][t] = b2; sm.lbox[3][t] = b3;
            store_state = true;
        } else {                                                        // kf.update(None) (deepocsort_kf.py:506-521), frozen = True
            if (fl & B200_OCF_OBSERVED) {
#pragma unroll
                for (int c = 0; c < 8; ++c) gf[(B200_DO_SX + c) * TMAX + t] = x[c];
                g4_store(gA, gf + (size_t)B200_DO_SPA * TMAX + t, TMAX);
                g4_store(gB, gf + (size_t)B200_DO_SPB * TMAX + t, TMAX);
                fl |= B200_OCF_SAVED;
            }
            fl &= ~B200_OCF_OBSERVED;
            fl |= B200_DOF_FROZEN;
        }
        xo[0] = gA.m[0]; xo[1] = gA.m[1]; xo[2] = gB.m[0]; xo[3] = gB.m[1];
        xo[4] = gA.m[2]; xo[5] = gA.m[3]; xo[6] = gB.m[2]; xo[7] = gB.m[3];
        if (store_state) {
#pragma unroll
            for (int c = 0; c < 8; ++c) gf[(B200_DO_X + c) * TMAX + t] = xo[c];
            g4_store(gA, gf + (size_t)B200_DO_PA * TMAX + t, TMAX);
            g4_store(gB, gf + (size_t)B200_DO_PB * TMAX + t, TMAX);
        }
    }
    __syncthreads();

    // ---- update_emb (deep_ocsort.py:218-220) of every matched tracker, one warp each -----------------------------
    if (emb_on) {
        const double af = p.alpha_fixed_emb;
        for (int q = warp; q < n0; q += NT / 32) {
            if (!sm.alive[q]) continue;
            const int j = sm.tmatch[q];
            if (j < 0) continue;
            const double trust = xdiv(xsub(sm.dconf[j], p.det_thresh), xsub(1.0, p.det_thresh));
            const double alpha = xadd(af, xmul(xsub(1.0, af), xsub(1.0, trust)));
            const double beta = xsub(1.0, alpha);
            double* e = pool + (size_t)sm.erow[q] * F;
            const float* d = dfeat + (size_t)j * F;
            // The renormalisation multiplies by one correctly rounded reciprocal per row instead of dividing every element
            // (a double division is a ~15-instruction sequence; 512 of them per matched tracker were a sixth of the step's
            // instructions): each value is within one ulp of the reference's quotient - the embedding's own tolerance is
            // 1e-6 (float32 products upstream).  Rows of up to 512 values stay in registers between the two passes.
            double acc = 0.0;
            if (F <= 512) {
                double v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int i = lane + 32 * k;
                    v[k] = 0.0;
                    if (i < F) { v[k] = xadd(xmul(alpha, e[i]), xmul(beta, (double)d[i])); acc += v[k] * v[k]; }
                }
                const double rinv = xdiv(1.0, sqrt(warp_sum_d(acc)));
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int i = lane + 32 * k;
                    if (i < F) e[i] = xmul(v[k], rinv);
                }
                continue;
            }
            for (int i = lane; i < F; i += 32) {
                const double v = xadd(xmul(alpha, e[i]), xmul(beta, (double)d[i]));
                e[i] = v;
                acc += v * v;
            }
            const double rinv = xdiv(1.0, sqrt(warp_sum_d(acc)));
            __syncwarp();
            for (int i = lane; i < F; i += 32) e[i] = xmul(e[i], rinv);
        }
    }

    // ---- new trackers (deep_ocsort.py:492-501) and the reversed output scan (:502-520) ------------------------------
    const int ds = tid < DMAX ? sm.dstate[tid] : DS_NONE;
    const bool newborn = ds == DS_FREE0 || ds == DS_FREE1;
    bool emit_old = false, emit_new = false, die = false;
    bool box_from_obs = false;
    if (live) {
        emit_old = tsu < 1 && (streak >= p.min_hits || frame <= p.min_hits);
        die = tsu > p.max_age;
        const double lsum = (fl & B200_OCF_HASOBS) ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
        box_from_obs = !(lsum < 0.0);
        if (!die) sm.rowused[sm.erow[t]] = 1;
    }
    if (newborn) emit_new = (0 >= p.min_hits || frame <= p.min_hits);
    unsigned long long val = (emit_old ? 1ull : 0ull) | (emit_new ? (1ull << 10) : 0ull) | (ds == DS_FREE0 ? (1ull << 20) : 0ull) |
                             (ds == DS_FREE1 ? (1ull << 30) : 0ull) | ((live && !die) ? (1ull << 40) : 0ull);
    unsigned long long tot;
    const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);      // (its barriers publish rowused)
    const int E_old = (int)(tot & 1023), E_new = (int)((tot >> 10) & 1023);
    const int n_free0 = (int)((tot >> 20) & 1023), n_free1 = (int)((tot >> 30) & 1023), n_keep = (int)((tot >> 40) & 1023);
    const int n_new = n_free0 + n_free1;
    // embedding-pool rows of dead trackers are recycled: the k-th new tracker takes the k-th free row
    {
        const bool isfree = t < TMAX && !sm.rowused[t];
        unsigned long long tf;
        const int rank = (int)block_exscan<NT>(isfree ? 1ull : 0ull, sm.scratch, tf);
        if (isfree) sm.freelist[rank] = (short)t;
        __syncthreads();
    }
    double* gout = packed ? nullptr : p.out + (size_t)s * p.max_tracks * 8;
    const int out_cap = packed ? min(p.max_tracks, nd_in) : p.max_tracks;
    if (n0 + n_new > tcap) err |= B200_ERR_TRACK_OVERFLOW;
    auto write_row = [&](int row, const Box& b, int id, double cf, double cl, int di) {
        if (packed) {
            double* o = reinterpret_cast<double*>(p.rows + (size_t)(roff + row) * B200_ROW_BYTE);
            o[0] = b.x1; o[1] = b.y1; o[2] = b.x2; o[3] = b.y2;
            reinterpret_cast<int2*>(o)[4] = make_int2(id, di);
            return;
        }
        double2* o = reinterpret_cast<double2*>(gout + (size_t)row * 8);
        o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
        o[2] = make_double2((double)id, cf); o[3] = make_double2(cl, (double)di);
    };
    auto state_box = [&](const double* x) {
        Box b;
        b.x1 = xsub(x[0], xdiv(x[2], 2.0)); b.y1 = xsub(x[1], xdiv(x[3], 2.0));
        b.x2 = xadd(x[0], xdiv(x[2], 2.0)); b.y2 = xadd(x[1], xdiv(x[3], 2.0));
        return b;
    };
    if (live) {
        gf[B200_DO_CONF * TMAX + t] = conf;
        gf[B200_DO_CLS * TMAX + t] = cls;
        gi[B200_OCI_AGE * TMAX + t] = age;
        gi[B200_OCI_TSU * TMAX + t] = tsu;
        gi[B200_OCI_HITS * TMAX + t] = hits;
        gi[B200_OCI_STREAK * TMAX + t] = streak;
        gi[B200_OCI_DET * TMAX + t] = det_ind;
        gi[B200_OCI_FLAGS * TMAX + t] = die ? (fl & ~OCF_ALIVE) : fl;
        if (emit_old) {
            const int row = E_new + (E_old - 1 - (int)(ex & 1023));
            if (row < out_cap) {
                Box b;
                if (box_from_obs) { b.x1 = sm.lbox[0][t]; b.y1 = sm.lbox[1][t]; b.x2 = sm.lbox[2][t]; b.y2 = sm.lbox[3][t]; }
                else b = state_box(xo);
                write_row(row, b, tid_id + 1, conf, cls, det_ind);
            }
        }
    } else if (t < n0 && (fl & OCF_ALIVE) == 0 && t < TMAX) {
        if (gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE) gi[B200_OCI_FLAGS * TMAX + t] = fl;     // NaN-purged this frame
    }
    int nb_row = -1, nb_det = -1;
    if (newborn) {
        const int j = tid;
        int order;                      // position in the creation order
        if (ocr_ran) order = (int)((ex >> 20) & 1023) + (int)((ex >> 30) & 1023);
        else order = ds == DS_FREE0 ? (int)((ex >> 20) & 1023) : n_free0 + (int)((ex >> 30) & 1023);
        const int dst = n0 + order;
        const int id = id0 + order;
        const double b0 = sm.dbox[0][j], b1 = sm.dbox[1][j], b2 = sm.dbox[2][j], b3 = sm.dbox[3][j];
        const double bw = xsub(b2, b0), bh = xsub(b3, b1);
        const double z[8] = {xadd(b0, xdiv(bw, 2.0)), xadd(b1, xdiv(bh, 2.0)), bw, bh, 0.0, 0.0, 0.0, 0.0};
        if (dst < tcap) {
            double qA[4], qB[4];
            do_process_noise(bw, bh, qA, qB);
#pragma unroll
            for (int c = 0; c < 8; ++c) gf[(B200_DO_X + c) * TMAX + dst] = z[c];
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_DO_LAST + c) * TMAX + dst] = -1.0;
            // P = process noise, position block x 4, velocity block x 100 (deep_ocsort.py:127-129)
            const double dA[4] = {xmul(qA[0], 4.0), xmul(qA[1], 4.0), xmul(qA[2], 100.0), xmul(qA[3], 100.0)};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a; b < 4; ++b) {
                    gf[(B200_DO_PA + k) * TMAX + dst] = a == b ? dA[a] : 0.0;
                    gf[(B200_DO_PB + k) * TMAX + dst] = a == b ? dA[a] : 0.0;
                    ++k;
                }
            gf[B200_DO_CONF * TMAX + dst] = sm.dconf[j];
            gf[B200_DO_CLS * TMAX + dst] = det_cls(j);
            gf[(B200_DO_VEL + 0) * TMAX + dst] = 0.0; gf[(B200_DO_VEL + 1) * TMAX + dst] = 0.0;
            gi[B200_OCI_ID * TMAX + dst] = id;
            gi[B200_OCI_AGE * TMAX + dst] = 0;
            gi[B200_OCI_TSU * TMAX + dst] = 0;
            gi[B200_OCI_HITS * TMAX + dst] = 0;
            gi[B200_OCI_STREAK * TMAX + dst] = 0;
            gi[B200_OCI_DET * TMAX + dst] = j;
            gi[(B200_OCI_RINGAGE + 0) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 1) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 2) * TMAX + dst] = -1;
            gi[B200_OCI_FLAGS * TMAX + dst] = OCF_ALIVE;
            nb_row = sm.freelist[order];
            nb_det = j;
            gi[B200_DOI_EROW * TMAX + dst] = nb_row;
        }
        if (emit_new) {
            const int row = n_new - 1 - order;
            if (row < out_cap) write_row(row, state_box(z), id + 1, sm.dconf[j], det_cls(j), j);
        }
    }
    // a new tracker's embedding is its detection's (fp32 -> fp64 is exact); one warp per new tracker
    if (emb_on) {
        __syncthreads();
        if (tid < DMAX) { sm.hd[tid] = (short)nb_row; sm.ud[tid] = (short)nb_det; }      // hd / ud are free now
        __syncthreads();
        for (int q = warp; q < nd; q += NT / 32) {
            const int row = sm.hd[q], j = sm.ud[q];
            if (row < 0) continue;
            double* e = pool + (size_t)row * F;
            const float* d = dfeat + (size_t)j * F;
            for (int i = lane; i < F; i += 32) e[i] = (double)d[i];
        }
    }
    const int n1 = min(n0 + n_new, tcap);
    const int alive_after = n_keep + min(n_new, tcap - n0 > 0 ? tcap - n0 : 0);
    if (tid == 0) {
        counts[0] = n1;
        counts[1] = alive_after;
        counts[2] = id0 + n_new;
        counts[3] = frame;
        p.nout[s] = min(E_old + E_new, out_cap);
        p.track_updates[s] += (unsigned long long)Cn;
    }
    if (err) { atomicOr(p.err, err); if (p.err_out) atomicOr(p.err_out, err); }
}

template <int TMAX, int DMAX>
cudaError_t launch_do_kernel(const StepParams& p, cudaStream_t stream) {
    auto kern = deepocsort_step_kernel<TMAX, TMAX, DMAX>;
    const size_t smem = sizeof(DoSmem<TMAX, DMAX>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.n_streams, TMAX, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t deepocsort_step_smem(int variant) {
    switch (variant) {
        case 0: return sizeof(DoSmem<64, 64>);
        case 1: return sizeof(DoSmem<128, 128>);
        case 2: return sizeof(DoSmem<224, 224>);
        case 3: return sizeof(DoSmem<256, 256>);
        case 4: return sizeof(DoSmem<512, 512>);
    }
    return 0;
}

cudaError_t launch_deepocsort_step(const StepParams& p, int variant, cudaStream_t stream) {
    switch (variant) {
        case 0: return launch_do_kernel<64, 64>(p, stream);
        case 1: return launch_do_kernel<128, 128>(p, stream);
        case 2: return launch_do_kernel<224, 224>(p, stream);
        case 3: return launch_do_kernel<256, 256>(p, stream);
        case 4: return launch_do_kernel<512, 512>(p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
