// ByteTrack frame step for many independent streams: ONE kernel launch per frame, one CTA per
// stream.  The CTA reads the stream's detections and its whole track state from HBM once,
// keeps everything (Kalman state, boxes, candidate graph, assignment duals) in shared memory
// for the entire step, and writes the state back once, already in the reference's new list
// order, together with the output rows.  Nothing T x D ever touches HBM - or is even computed:
// candidates come from per-frame cell masks (64 x-cells and 64 y-cells, each a bitmask of the
// detections whose box touches the cell), so a track only runs the exact IoU test against the
// handful of detections that share a cell range with it in both axes.
//
// Replaces BYTETracker.update (boxmot/trackers/bytetrack/byte_tracker.py:132-281) and what it
// calls: STrack.multi_predict :35-48 -> KalmanFilter.multi_predict (bytetrack_kf.py:155-192),
// iou_distance / fuse_score / linear_assignment (matching.py:94-119, :213-221, :56-71 ->
// lap.lapjv), STrack.update / re_activate / activate :50-98 -> KalmanFilter.update / initiate,
// joint_stracks / sub_stracks / remove_duplicate_stracks :287-325.
#include "boxes.cuh"
#include "kf.cuh"
#include "lap_sparse.cuh"
#include "layout.h"
#include "step_params.h"

namespace b200 {

namespace {

constexpr int ROLE_TRACKED = 0;   // activated entry of tracked_stracks  -> in strack_pool
constexpr int ROLE_LOST = 1;      // entry of lost_stracks               -> in strack_pool
constexpr int ROLE_UNCONF = 2;    // not yet activated (born last frame) -> "unconfirmed"

constexpr int DF_HIGH = 1, DF_LOW = 2, DF_USED = 4;

constexpr int CAT_NONE = 0, CAT_KEEP = 1, CAT_REFOUND = 2, CAT_LOST_OLD = 3, CAT_LOST_NEW = 4;

constexpr int NCELL = 64;         // cells per axis of the candidate masks

// row types of one association pass: which detection set / limit / cost a row uses
constexpr int RT_NONE = 0, RT_A = 1, RT_B = 2;

struct Sm {
    double *tf, *tbox, *dxywh, *dbox, *dconf, *dcls, *u, *v, *dist;
    unsigned long long* scratch;
    int *ti, *parent, *head, *coldeg, *ncomplex;
    uint32_t *adj, *colbitsA, *colbitsB, *xmask, *ymask;
    float* fext;
    short *rnext, *xr, *yc, *pred, *nextc, *mark, *scn, *lostlist;
    unsigned char *role, *rowtype, *dflag, *cat, *drop;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t carve(Sm* sm, unsigned char* base, int Tmax, int Dmax) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align16(off + bytes); return base ? base + o : (unsigned char*)nullptr; };
    const int DW = Dmax / 32;
    Sm s;
    s.tf = (double*)take(sizeof(double) * B200_NF * Tmax);
    s.tbox = (double*)take(sizeof(double) * 4 * Tmax);
    s.dxywh = (double*)take(sizeof(double) * 4 * Dmax);
    s.dbox = (double*)take(sizeof(double) * 4 * Dmax);
    s.dconf = (double*)take(sizeof(double) * Dmax);
    s.dcls = (double*)take(sizeof(double) * Dmax);
    s.u = (double*)take(sizeof(double) * Tmax);
    s.v = (double*)take(sizeof(double) * Dmax);
    s.dist = (double*)take(sizeof(double) * Dmax);
    s.scratch = (unsigned long long*)take(sizeof(unsigned long long) * 40);
    s.ti = (int*)take(sizeof(int) * B200_NI * Tmax);
    s.parent = (int*)take(sizeof(int) * (Tmax + Dmax));
    s.head = (int*)take(sizeof(int) * Tmax);
    s.coldeg = (int*)take(sizeof(int) * Dmax);
    s.ncomplex = (int*)take(sizeof(int) * 4);
    s.adj = (uint32_t*)take(sizeof(uint32_t) * DW * Tmax);
    s.colbitsA = (uint32_t*)take(sizeof(uint32_t) * DW);
    s.colbitsB = (uint32_t*)take(sizeof(uint32_t) * DW);
    s.xmask = (uint32_t*)take(sizeof(uint32_t) * NCELL * DW);
    s.ymask = (uint32_t*)take(sizeof(uint32_t) * NCELL * DW);
    s.fext = (float*)take(sizeof(float) * 32 * 4);
    s.rnext = (short*)take(sizeof(short) * Tmax);
    s.xr = (short*)take(sizeof(short) * Tmax);
    s.yc = (short*)take(sizeof(short) * Dmax);
    s.pred = (short*)take(sizeof(short) * Dmax);
    s.nextc = (short*)take(sizeof(short) * Dmax);
    s.mark = (short*)take(sizeof(short) * Dmax);
    s.scn = (short*)take(sizeof(short) * Dmax);
    s.lostlist = (short*)take(sizeof(short) * Tmax);
    s.role = take(Tmax);
    s.rowtype = take(Tmax);
    s.dflag = take(Dmax);
    s.cat = take(Tmax);
    s.drop = take(Tmax + Dmax);
    if (sm) *sm = s;
    return off;
}

__device__ __forceinline__ Box load_box(const double* b, int stride, int i) {
    Box r; r.x1 = b[i]; r.y1 = b[stride + i]; r.x2 = b[2 * stride + i]; r.y2 = b[3 * stride + i]; return r;
}

__device__ __forceinline__ void load_kf(const double* tf, int Tmax, int t, KfState& s) {
#pragma unroll
    for (int c = 0; c < 8; ++c) s.m[c] = tf[(B200_TF_MEAN + c) * Tmax + t];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s.pp[i] = tf[(B200_TF_COV + 3 * i + 0) * Tmax + t];
        s.pv[i] = tf[(B200_TF_COV + 3 * i + 1) * Tmax + t];
        s.vv[i] = tf[(B200_TF_COV + 3 * i + 2) * Tmax + t];
    }
}
__device__ __forceinline__ void store_kf(double* tf, int Tmax, int t, const KfState& s) {
#pragma unroll
    for (int c = 0; c < 8; ++c) tf[(B200_TF_MEAN + c) * Tmax + t] = s.m[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        tf[(B200_TF_COV + 3 * i + 0) * Tmax + t] = s.pp[i];
        tf[(B200_TF_COV + 3 * i + 1) * Tmax + t] = s.pv[i];
        tf[(B200_TF_COV + 3 * i + 2) * Tmax + t] = s.vv[i];
    }
}

// STrack.xyxy (byte_tracker.py:100-111): XYAH mean -> (xc, yc, a*h, h) -> corners
template <int KIND>
__device__ __forceinline__ Box mean_to_box(double xc, double yc, double a_or_w, double h) {
    const double w = (KIND == KF_XYWH) ? a_or_w : xmul(a_or_w, h);
    return xywh_to_xyxy(xc, yc, w, h);
}

template <int KIND>
__device__ __forceinline__ void refresh_box(const Sm& sm, int Tmax, int t) {
    const Box b = mean_to_box<KIND>(sm.tf[0 * Tmax + t], sm.tf[1 * Tmax + t], sm.tf[2 * Tmax + t], sm.tf[3 * Tmax + t]);
    sm.tbox[t] = b.x1; sm.tbox[Tmax + t] = b.y1; sm.tbox[2 * Tmax + t] = b.x2; sm.tbox[3 * Tmax + t] = b.y2;
}

// measurement fed to the filter for detection j (STrack.__init__, byte_tracker.py:16-18)
template <int KIND>
__device__ __forceinline__ void det_measurement(const Sm& sm, int Dmax, int j, double* z) {
    const double xc = sm.dxywh[j], yc = sm.dxywh[Dmax + j], w = sm.dxywh[2 * Dmax + j], h = sm.dxywh[3 * Dmax + j];
    if (KIND == KF_XYWH) { z[0] = xc; z[1] = yc; z[2] = w; z[3] = h; }
    else xywh_to_xyah(xc, yc, w, h, z);
}

// cost of (track row, detection): iou_distance, optionally fuse_score, chosen by the row type
struct PassCost {
    const double *tbox, *dbox, *dconf;
    const unsigned char* rowtype;
    int Tmax, Dmax;
    bool fuseA, fuseB;
    __device__ __forceinline__ double eval(const Box& a, int j, bool fuse) const {
        const Box b = load_box(dbox, Dmax, j);
        const double v = box_iou(a, b);
        return fuse ? fused_cost(v, dconf[j]) : xsub(1.0, v);
    }
    __device__ __forceinline__ double operator()(int t, int j) const {
        return eval(load_box(tbox, Tmax, t), j, rowtype[t] == RT_A ? fuseA : fuseB);
    }
};
struct PassLimit {
    const unsigned char* rowtype;
    double limA, limB;
    __device__ __forceinline__ double operator()(int t) const { return rowtype[t] == RT_A ? limA : limB; }
};

struct CellMap {
    float x0, y0, sx, sy;
    __device__ __forceinline__ int cx(double x) const {
        return min(max((int)(((float)x - x0) * sx), 0), NCELL - 1);      // monotone in x
    }
    __device__ __forceinline__ int cy(double y) const {
        return min(max((int)(((float)y - y0) * sy), 0), NCELL - 1);
    }
};

// Candidate graph of one association pass.  Row t (type rowtype[t]) is tested against the
// detections of its column set that share a cell range with it in x AND in y (a superset of
// the overlapping ones because the cell maps are monotone); edge iff the boxes overlap and
// cost <= limit - exact pruning, see lap_sparse.cuh (no overlap => iou = 0 => cost = 1 > limit).
template <int NT>
__device__ void build_graph(const Sm& sm, const LapWork& lw, int Tmax, int Dmax, int n, int words, const CellMap& cm,
                            const PassCost& cost, const PassLimit& lim) {
    const int DW = Dmax / 32;
    for (int t = threadIdx.x; t < n; t += NT) {
        const int rt = sm.rowtype[t];
        if (rt == RT_NONE) {
            for (int wd = 0; wd < words; ++wd) sm.adj[wd * Tmax + t] = 0u;
            continue;
        }
        const Box a = load_box(sm.tbox, Tmax, t);
        const int cx0 = cm.cx(a.x1), cx1 = cm.cx(a.x2), cy0 = cm.cy(a.y1), cy1 = cm.cy(a.y2);
        const uint32_t* colbits = rt == RT_A ? sm.colbitsA : sm.colbitsB;
        const bool fuse = rt == RT_A ? cost.fuseA : cost.fuseB;
        const double limit = rt == RT_A ? lim.limA : lim.limB;
        for (int wd = 0; wd < words; ++wd) {
            uint32_t res = 0u;
            const uint32_t cb = colbits[wd];
            if (cb) {
                uint32_t mx = 0u, my = 0u;
                for (int c = cx0; c <= cx1; ++c) mx |= sm.xmask[c * DW + wd];
                for (int c = cy0; c <= cy1; ++c) my |= sm.ymask[c * DW + wd];
                uint32_t cand = mx & my & cb;
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int j = wd * 32 + b;
                    const Box d = load_box(sm.dbox, Dmax, j);
                    if (box_overlap(a, d)) {
                        const double v = box_iou(a, d);
                        const double c = fuse ? fused_cost(v, sm.dconf[j]) : xsub(1.0, v);
                        if (c <= limit) { res |= 1u << b; atomicAdd(&lw.coldeg[j], 1); }
                    }
                }
            }
            sm.adj[wd * Tmax + t] = res;
        }
    }
    __syncthreads();
}

// STrack.update / re_activate for every matched row (byte_tracker.py:64-98)
template <int NT, int KIND>
__device__ void apply_matches(const Sm& sm, int Tmax, int Dmax, int n, int frame) {
    for (int t = threadIdx.x; t < n; t += NT) {
        if (sm.rowtype[t] == RT_NONE) continue;
        const int j = sm.xr[t];
        if (j < 0) continue;
        KfState s;
        load_kf(sm.tf, Tmax, t, s);
        double z[4];
        det_measurement<KIND>(sm, Dmax, j, z);
        kf_update<KIND>(s, z);
        store_kf(sm.tf, Tmax, t, s);
        int fl = sm.ti[B200_TI_FLAGS * Tmax + t];
        const int st = fl & 3;
        int len = sm.ti[B200_TI_LEN * Tmax + t];
        len = (st == B200_ST_TRACKED) ? len + 1 : 0;
        sm.ti[B200_TI_LEN * Tmax + t] = len;
        sm.ti[B200_TI_FRAME * Tmax + t] = frame;
        sm.ti[B200_TI_DET * Tmax + t] = j;
        sm.ti[B200_TI_FLAGS * Tmax + t] = (fl & ~3) | B200_ST_TRACKED | B200_FLAG_ACTIVATED;
        sm.tf[B200_TF_SCORE * Tmax + t] = sm.dconf[j];
        sm.tf[B200_TF_CLS * Tmax + t] = sm.dcls[j];
        sm.dflag[j] |= DF_USED;
    }
}

template <int NT, int KIND>
__global__ void __launch_bounds__(NT) bytetrack_step_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tmax = p.max_tracks, Dmax = p.max_dets, DW = Dmax / 32;
    Sm sm;
    carve(&sm, smem_raw, Tmax, Dmax);

    int* counts = p.counts + 4 * s;
    const int nT = counts[0], nL = counts[1], id0 = counts[2], frame = counts[3] + 1;
    const int n = nT + nL;
    int nd = p.ndets[s];
    int err = 0;
    if (nd > Dmax) { nd = Dmax; err |= B200_ERR_DET_OVERFLOW; }
    if (nd < 0) nd = 0;
    const int words = (nd + 31) >> 5;

    // ---- HBM -> shared memory, once: detections [nd, 6] (planar) and the track state ----
    {
        const double* g = p.dets + (size_t)s * Dmax * 6;
        for (int i = tid; i < nd * 6; i += NT) {
            const double val = g[i];
            const int j = i / 6, c = i - 6 * j;
            if (c < 4) sm.dbox[c * Dmax + j] = val;
            else if (c == 4) sm.dconf[j] = val;
            else sm.dcls[j] = val;
        }
        const double* gf = p.state_f + (size_t)s * B200_NF * Tmax;
        const int* gi = p.state_i + (size_t)s * B200_NI * Tmax;
        for (int t = tid; t < n; t += NT) {
#pragma unroll
            for (int c = 0; c < B200_NF; ++c) sm.tf[c * Tmax + t] = gf[c * Tmax + t];
#pragma unroll
            for (int c = 0; c < B200_NI; ++c) sm.ti[c * Tmax + t] = gi[c * Tmax + t];
        }
        for (int i = tid; i < NCELL * DW; i += NT) { sm.xmask[i] = 0u; sm.ymask[i] = 0u; }
        for (int i = tid; i < Tmax + Dmax; i += NT) sm.drop[i] = 0;
    }
    __syncthreads();

    // ---- detection side: xyxy -> xywh, the round-trip box used by iou_distance, the two
    // confidence bands (byte_tracker.py:151-158; strict inequalities), frame extents --------
    const float FBIG = 3.0e38f;
    float ex0 = FBIG, ey0 = FBIG, ex1 = -FBIG, ey1 = -FBIG;
    for (int j = tid; j < words * 32; j += NT) {
        int fl = 0;
        if (j < nd) {
            double xc, yc, w, h;
            xyxy_to_xywh(sm.dbox[j], sm.dbox[Dmax + j], sm.dbox[2 * Dmax + j], sm.dbox[3 * Dmax + j], xc, yc, w, h);
            sm.dxywh[j] = xc; sm.dxywh[Dmax + j] = yc; sm.dxywh[2 * Dmax + j] = w; sm.dxywh[3 * Dmax + j] = h;
            const Box b = xywh_to_xyxy(xc, yc, w, h);
            sm.dbox[j] = b.x1; sm.dbox[Dmax + j] = b.y1; sm.dbox[2 * Dmax + j] = b.x2; sm.dbox[3 * Dmax + j] = b.y2;
            const double c = sm.dconf[j];
            if (c > p.track_thresh) fl = DF_HIGH;
            else if (c > p.low_thresh && c < p.track_thresh) fl = DF_LOW;
            if (fl) {
                ex0 = fminf(ex0, (float)b.x1); ey0 = fminf(ey0, (float)b.y1);
                ex1 = fmaxf(ex1, (float)b.x2); ey1 = fmaxf(ey1, (float)b.y2);
            }
        }
        sm.dflag[j] = (unsigned char)fl;
        const uint32_t mh = __ballot_sync(0xffffffffu, fl == DF_HIGH);
        if (lane == 0) sm.colbitsA[j >> 5] = mh;           // first association: all high detections
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        ex0 = fminf(ex0, __shfl_xor_sync(0xffffffffu, ex0, d)); ey0 = fminf(ey0, __shfl_xor_sync(0xffffffffu, ey0, d));
        ex1 = fmaxf(ex1, __shfl_xor_sync(0xffffffffu, ex1, d)); ey1 = fmaxf(ey1, __shfl_xor_sync(0xffffffffu, ey1, d));
    }
    if (lane == 0) { sm.fext[warp * 4] = ex0; sm.fext[warp * 4 + 1] = ey0; sm.fext[warp * 4 + 2] = ex1; sm.fext[warp * 4 + 3] = ey1; }

    // ---- track side: roles, Kalman predict of the pool (unconfirmed tracks are NOT
    // predicted, byte_tracker.py:178-180), boxes ---------------------------------------------
    for (int t = tid; t < n; t += NT) {
        const int fl = sm.ti[B200_TI_FLAGS * Tmax + t];
        const int role = t >= nT ? ROLE_LOST : ((fl & B200_FLAG_ACTIVATED) ? ROLE_TRACKED : ROLE_UNCONF);
        sm.role[t] = (unsigned char)role;
        if (role != ROLE_UNCONF) {
            KfState k;
            load_kf(sm.tf, Tmax, t, k);
            if ((fl & 3) != B200_ST_TRACKED) {          // multi_predict: zero the height (w, h) velocity
                k.m[7] = 0.0;
                if (KIND == KF_XYWH) k.m[6] = 0.0;
            }
            kf_predict<KIND>(k);
            store_kf(sm.tf, Tmax, t, k);
        }
        refresh_box<KIND>(sm, Tmax, t);
        sm.rowtype[t] = role != ROLE_UNCONF ? RT_A : RT_NONE;
        sm.cat[t] = CAT_NONE;
    }

    LapWork lw;
    lw.Tmax = Tmax; lw.Dmax = Dmax; lw.adj = sm.adj; lw.u = sm.u; lw.v = sm.v; lw.dist = sm.dist;
    lw.parent = sm.parent; lw.head = sm.head; lw.rnext = sm.rnext; lw.xr = sm.xr; lw.yc = sm.yc;
    lw.pred = sm.pred; lw.nextc = sm.nextc; lw.mark = sm.mark; lw.scn = sm.scn;
    lw.coldeg = sm.coldeg; lw.ncomplex = sm.ncomplex;
    lap_prepare<NT>(lw, n, words);
    __syncthreads();

    // ---- cell masks over the banded detections ---------------------------------------------
    CellMap cm;
    {
        float x0 = FBIG, y0 = FBIG, x1 = -FBIG, y1 = -FBIG;
        for (int k = 0; k < NT / 32; ++k) {
            x0 = fminf(x0, sm.fext[k * 4]); y0 = fminf(y0, sm.fext[k * 4 + 1]);
            x1 = fmaxf(x1, sm.fext[k * 4 + 2]); y1 = fmaxf(y1, sm.fext[k * 4 + 3]);
        }
        cm.x0 = x0; cm.y0 = y0;
        cm.sx = (x1 > x0) ? (float)NCELL / (x1 - x0) : 0.f;
        cm.sy = (y1 > y0) ? (float)NCELL / (y1 - y0) : 0.f;
    }
    for (int j = tid; j < nd; j += NT) {
        if (!sm.dflag[j]) continue;
        const Box b = load_box(sm.dbox, Dmax, j);
        const uint32_t bit = 1u << (j & 31);
        const int wd = j >> 5;
        const int cx0 = cm.cx(b.x1), cx1 = cm.cx(b.x2), cy0 = cm.cy(b.y1), cy1 = cm.cy(b.y2);
        for (int c = cx0; c <= cx1; ++c) atomicOr(&sm.xmask[c * DW + wd], bit);
        for (int c = cy0; c <= cy1; ++c) atomicOr(&sm.ymask[c * DW + wd], bit);
    }
    __syncthreads();

    PassCost cost;
    cost.tbox = sm.tbox; cost.dbox = sm.dbox; cost.dconf = sm.dconf; cost.rowtype = sm.rowtype; cost.Tmax = Tmax; cost.Dmax = Dmax;
    PassLimit lim;
    lim.rowtype = sm.rowtype;

    // ---- first association: pool x high detections, fused score, limit match_thresh ----
    cost.fuseA = true; cost.fuseB = true;
    lim.limA = p.match_thresh; lim.limB = p.match_thresh;
    build_graph<NT>(sm, lw, Tmax, Dmax, n, words, cm, cost, lim);
    lap_sparse_solve<NT>(lw, n, words, lim, cost);
    apply_matches<NT, KIND>(sm, Tmax, Dmax, n, frame);
    __syncthreads();

    // ---- second pass: two independent problems solved together (disjoint rows AND columns):
    //   A: still-Tracked leftovers x low detections, plain IoU, limit 0.5   (byte_tracker.py:198-226)
    //   B: unconfirmed x remaining high detections, fused score, limit 0.7  (byte_tracker.py:228-240)
    for (int t = tid; t < n; t += NT) {
        const bool matched = sm.rowtype[t] != RT_NONE && sm.xr[t] >= 0;
        const int role = sm.role[t];
        sm.rowtype[t] = (role == ROLE_TRACKED && !matched) ? RT_A : (role == ROLE_UNCONF ? RT_B : RT_NONE);
    }
    for (int j = tid; j < words * 32; j += NT) {
        const int fl = j < nd ? sm.dflag[j] : 0;
        const uint32_t ma = __ballot_sync(0xffffffffu, fl == DF_LOW);
        const uint32_t mb = __ballot_sync(0xffffffffu, fl == DF_HIGH);      // high and not used
        if (lane == 0) { sm.colbitsA[j >> 5] = ma; sm.colbitsB[j >> 5] = mb; }
    }
    __syncthreads();                 // xr of pass 1 fully consumed before lap_prepare resets it
    lap_prepare<NT>(lw, n, words);
    __syncthreads();
    cost.fuseA = false; cost.fuseB = true;
    lim.limA = p.second_thresh; lim.limB = p.unconf_thresh;
    build_graph<NT>(sm, lw, Tmax, Dmax, n, words, cm, cost, lim);
    lap_sparse_solve<NT>(lw, n, words, lim, cost);
    apply_matches<NT, KIND>(sm, Tmax, Dmax, n, frame);
    __syncthreads();

    // ---- lifecycle: lost / removed / aged-out, list categories (byte_tracker.py:222-268) ----
    for (int t = tid; t < n; t += NT) {
        int fl = sm.ti[B200_TI_FLAGS * Tmax + t];
        const int role = sm.role[t];
        const bool unmatched_now = sm.rowtype[t] != RT_NONE && sm.xr[t] < 0;
        if (role == ROLE_TRACKED && unmatched_now) fl = (fl & ~3) | B200_ST_LOST;       // mark_lost; frame_id stays = end_frame
        if (role == ROLE_UNCONF && unmatched_now) fl = (fl & ~3) | B200_ST_REMOVED;     // mark_removed
        int st = fl & 3;
        const bool sticky_old = fl & B200_FLAG_STICKY;      // id already in removed_stracks
        int cat = CAT_NONE;
        if (role == ROLE_LOST) {
            if (st == B200_ST_TRACKED) cat = CAT_REFOUND;
            else {
                if (frame - sm.ti[B200_TI_FRAME * Tmax + t] > p.max_time_lost) {
                    fl = (fl & ~3) | B200_ST_REMOVED;       // stays listed one more frame (removed-lag)
                    st = B200_ST_REMOVED;
                }
                if (!sticky_old) cat = CAT_LOST_OLD;
                if (st == B200_ST_REMOVED) fl |= B200_FLAG_STICKY;
            }
        } else if (st == B200_ST_TRACKED) cat = CAT_KEEP;
        else if (st == B200_ST_LOST && !sticky_old) cat = CAT_LOST_NEW;
        sm.ti[B200_TI_FLAGS * Tmax + t] = fl;
        sm.cat[t] = (unsigned char)cat;
        if (cat != CAT_NONE) refresh_box<KIND>(sm, Tmax, t);
    }
    __syncthreads();

    // compact list of the new lost list (old entries first, then the newly lost) for the
    // duplicate test; packed counters: [0:16) old-lost, [16:32) new-lost
    int nLostOld = 0, nLostNew = 0;
    {
        unsigned long long base = 0;
        for (int c0 = 0; c0 < n; c0 += NT) {
            const int t = c0 + tid;
            const int cat = t < n ? sm.cat[t] : CAT_NONE;
            const unsigned long long val = (cat == CAT_LOST_OLD ? 1ull : 0ull) | (cat == CAT_LOST_NEW ? (1ull << 16) : 0ull);
            unsigned long long tot;
            const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot) + base;
            if (cat == CAT_LOST_OLD) sm.lostlist[ex & 0xffff] = (short)t;
            if (cat == CAT_LOST_NEW) sm.head[(ex >> 16) & 0xffff] = t;      // staged, shifted below
            base += tot;
        }
        nLostOld = (int)(base & 0xffff);
        nLostNew = (int)((base >> 16) & 0xffff);
        __syncthreads();
        for (int k = tid; k < nLostNew; k += NT) sm.lostlist[nLostOld + k] = (short)sm.head[k];
        __syncthreads();
    }
    const int nLostList = nLostOld + nLostNew;

    // ---- remove_duplicate_stracks (byte_tracker.py:312-325): tracked' x lost', 1-iou < 0.15
    // tracked' = kept slots, new tracks (unmatched high detections), re-found slots
    if (nLostList > 0) {
        for (int e = tid; e < n + nd; e += NT) {
            Box a;
            int age;
            if (e < n) {
                const int cat = sm.cat[e];
                if (cat != CAT_KEEP && cat != CAT_REFOUND) continue;
                a = load_box(sm.tbox, Tmax, e);
                age = sm.ti[B200_TI_FRAME * Tmax + e] - sm.ti[B200_TI_START * Tmax + e];
            } else {
                const int j = e - n;
                if ((sm.dflag[j] & (DF_HIGH | DF_USED)) != DF_HIGH) continue;
                if (sm.dconf[j] < p.new_thresh) continue;
                double z[4];
                det_measurement<KIND>(sm, Dmax, j, z);
                a = mean_to_box<KIND>(z[0], z[1], z[2], z[3]);
                age = 0;
            }
            bool dropme = false;
            for (int k = 0; k < nLostList; ++k) {
                const int q = sm.lostlist[k];
                const Box b = load_box(sm.tbox, Tmax, q);
                if (!box_overlap(a, b)) continue;
                if (xsub(1.0, box_iou(a, b)) < p.dup_thresh) {
                    const int ageq = sm.ti[B200_TI_FRAME * Tmax + q] - sm.ti[B200_TI_START * Tmax + q];
                    if (age > ageq) sm.drop[q] = 1; else dropme = true;
                }
            }
            if (dropme) sm.drop[e < n ? e : Tmax + (e - n)] = 1;
        }
        __syncthreads();
    }

    // ---- destinations.  One packed scan (10-bit fields):
    //   slots: [0) keep, [10) refound, [20) lostOld, [30) lostNew ; dets: [40) born kept, [50) born (all)
    // Every CAT_KEEP / CAT_REFOUND entry is activated, so output rows = keep ++ born (frame 1 only) ++ refound.
    const bool born_active = frame == 1;                 // STrack.activate: is_activated only on frame 1
    double* gf = p.state_f + (size_t)s * B200_NF * Tmax;
    int* gi = p.state_i + (size_t)s * B200_NI * Tmax;
    double* gout = p.out + (size_t)s * Tmax * 8;
    const int m = max(n, nd);
    auto elem_val = [&](int i) -> unsigned long long {
        unsigned long long v = 0ull;
        if (i < n && !sm.drop[i]) {
            const int cat = sm.cat[i];
            if (cat != CAT_NONE) v = 1ull << (10 * (cat - 1));
        }
        if (i < nd && (sm.dflag[i] & (DF_HIGH | DF_USED)) == DF_HIGH && !(sm.dconf[i] < p.new_thresh)) {
            v |= 1ull << 50;                              // activate() ran: consumes an id even if dropped below
            if (!sm.drop[Tmax + i]) v |= 1ull << 40;
        }
        return v;
    };
    // block totals first (segment bases depend on them): warp reduce + shared atomics
    if (tid == 0) sm.scratch[36] = 0ull;
    __syncthreads();
    {
        unsigned long long acc = 0ull;
        for (int i = tid; i < m; i += NT) acc += elem_val(i);
#pragma unroll
        for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0 && acc) atomicAdd(&sm.scratch[36], acc);
    }
    __syncthreads();
    const unsigned long long totals = sm.scratch[36];
    const int totKeep = (int)(totals & 1023), totRef = (int)((totals >> 10) & 1023);
    const int totLostOld = (int)((totals >> 20) & 1023), totLostNew = (int)((totals >> 30) & 1023);
    const int totBorn = (int)((totals >> 40) & 1023), totBornAll = (int)((totals >> 50) & 1023);
    const int newT = totKeep + totBorn + totRef;
    const int newL = totLostOld + totLostNew;
    if (newT + newL > Tmax) err |= B200_ERR_TRACK_OVERFLOW;
    const int rowsBorn = born_active ? totBorn : 0;

    unsigned long long base = 0;
    for (int c0 = 0; c0 < m; c0 += NT) {
        const int i = c0 + tid;
        const unsigned long long val = elem_val(i);
        unsigned long long tot;
        const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot) + base;
        base += tot;
        const unsigned long long sv = val & ((1ull << 40) - 1);
        if (sv) {
            const int cat = sm.cat[i];
            int dst, orow = -1;
            if (cat == CAT_KEEP) { dst = (int)(ex & 1023); orow = dst; }
            else if (cat == CAT_REFOUND) { const int k = (int)((ex >> 10) & 1023); dst = totKeep + totBorn + k; orow = totKeep + rowsBorn + k; }
            else if (cat == CAT_LOST_OLD) dst = newT + (int)((ex >> 20) & 1023);
            else dst = newT + totLostOld + (int)((ex >> 30) & 1023);
            if (dst < Tmax) {
#pragma unroll
                for (int c = 0; c < B200_NF; ++c) gf[c * Tmax + dst] = sm.tf[c * Tmax + i];
#pragma unroll
                for (int c = 0; c < B200_NI; ++c) gi[c * Tmax + dst] = sm.ti[c * Tmax + i];
            }
            if (orow >= 0 && orow < Tmax) {
                double* o = gout + (size_t)orow * 8;
                o[0] = sm.tbox[i]; o[1] = sm.tbox[Tmax + i]; o[2] = sm.tbox[2 * Tmax + i]; o[3] = sm.tbox[3 * Tmax + i];
                o[4] = (double)sm.ti[B200_TI_ID * Tmax + i];
                o[5] = sm.tf[B200_TF_SCORE * Tmax + i];
                o[6] = sm.tf[B200_TF_CLS * Tmax + i];
                o[7] = (double)sm.ti[B200_TI_DET * Tmax + i];
            }
        }
        if (val & (1ull << 40)) {                       // STrack.activate (byte_tracker.py:50-62)
            const int j = i;
            const int k = (int)((ex >> 40) & 1023);
            const int dst = totKeep + k;
            const int id = id0 + (int)((ex >> 50) & 1023) + 1;
            double z[4];
            det_measurement<KIND>(sm, Dmax, j, z);
            KfState ks;
            kf_initiate<KIND>(z, ks);
            if (dst < Tmax) {
#pragma unroll
                for (int c = 0; c < 8; ++c) gf[(B200_TF_MEAN + c) * Tmax + dst] = ks.m[c];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    gf[(B200_TF_COV + 3 * a + 0) * Tmax + dst] = ks.pp[a];
                    gf[(B200_TF_COV + 3 * a + 1) * Tmax + dst] = ks.pv[a];
                    gf[(B200_TF_COV + 3 * a + 2) * Tmax + dst] = ks.vv[a];
                }
                gf[B200_TF_SCORE * Tmax + dst] = sm.dconf[j];
                gf[B200_TF_CLS * Tmax + dst] = sm.dcls[j];
                gi[B200_TI_ID * Tmax + dst] = id;
                gi[B200_TI_FRAME * Tmax + dst] = frame;
                gi[B200_TI_START * Tmax + dst] = frame;
                gi[B200_TI_LEN * Tmax + dst] = 0;
                gi[B200_TI_DET * Tmax + dst] = j;
                gi[B200_TI_FLAGS * Tmax + dst] = B200_ST_TRACKED | (born_active ? B200_FLAG_ACTIVATED : 0);
            }
            if (born_active) {
                const int orow = totKeep + k;
                if (orow < Tmax) {
                    const Box b = mean_to_box<KIND>(z[0], z[1], z[2], z[3]);
                    double* o = gout + (size_t)orow * 8;
                    o[0] = b.x1; o[1] = b.y1; o[2] = b.x2; o[3] = b.y2;
                    o[4] = (double)id; o[5] = sm.dconf[j]; o[6] = sm.dcls[j]; o[7] = (double)j;
                }
            }
        }
    }
    if (tid == 0) {
        counts[0] = min(newT, Tmax);
        counts[1] = min(newL, Tmax - min(newT, Tmax));
        counts[2] = id0 + totBornAll;
        counts[3] = frame;
        p.nout[s] = min(totKeep + rowsBorn + totRef, Tmax);
        p.track_updates[s] += (unsigned long long)n;
        if (err) atomicOr(p.err, err);
    }
}

}  // namespace

size_t bytetrack_step_smem(int Tmax, int Dmax) { return carve(nullptr, nullptr, Tmax, Dmax); }

cudaError_t launch_bytetrack_step(const StepParams& p, int kf_kind, cudaStream_t stream) {
    constexpr int NT = 256;
    const size_t smem = bytetrack_step_smem(p.max_tracks, p.max_dets);
    auto kern = kf_kind == KF_XYWH ? bytetrack_step_kernel<NT, KF_XYWH> : bytetrack_step_kernel<NT, KF_XYAH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.n_streams, NT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace b200
