// ByteTrack frame step for many independent streams: ONE kernel launch per frame, one CTA per
// stream.  The CTA reads the stream's detections and its whole track state from HBM once,
// keeps everything (Kalman state, boxes, candidate graph, assignment duals) in shared memory
// for the entire step, and writes the state back once, already in the reference's new list
// order, together with the output rows.  Nothing T x D ever touches HBM.
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

struct Sm {
    double *tf, *tbox, *dxywh, *dbox, *dconf, *dcls, *u, *v, *dist;
    unsigned long long* scratch;
    int *ti, *parent, *head;
    uint32_t *adj, *colbits;
    short *rnext, *xr, *yc, *pred, *nextc, *mark, *scn, *lostlist;
    unsigned char *role, *rowsel, *dflag, *cat, *drop;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t carve(Sm* sm, unsigned char* base, int Tmax, int Dmax) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align16(off + bytes); return base ? base + o : (unsigned char*)nullptr; };
    const int DW = Dmax / 32;
    double* tf = (double*)take(sizeof(double) * B200_NF * Tmax);
    double* tbox = (double*)take(sizeof(double) * 4 * Tmax);
    double* dxywh = (double*)take(sizeof(double) * 4 * Dmax);
    double* dbox = (double*)take(sizeof(double) * 4 * Dmax);
    double* dconf = (double*)take(sizeof(double) * Dmax);
    double* dcls = (double*)take(sizeof(double) * Dmax);
    double* u = (double*)take(sizeof(double) * Tmax);
    double* v = (double*)take(sizeof(double) * Dmax);
    double* dist = (double*)take(sizeof(double) * Dmax);
    unsigned long long* scratch = (unsigned long long*)take(sizeof(unsigned long long) * 40);
    int* ti = (int*)take(sizeof(int) * B200_NI * Tmax);
    int* parent = (int*)take(sizeof(int) * (Tmax + Dmax));
    int* head = (int*)take(sizeof(int) * Tmax);
    uint32_t* adj = (uint32_t*)take(sizeof(uint32_t) * DW * Tmax);
    uint32_t* colbits = (uint32_t*)take(sizeof(uint32_t) * DW);
    short* rnext = (short*)take(sizeof(short) * Tmax);
    short* xr = (short*)take(sizeof(short) * Tmax);
    short* yc = (short*)take(sizeof(short) * Dmax);
    short* pred = (short*)take(sizeof(short) * Dmax);
    short* nextc = (short*)take(sizeof(short) * Dmax);
    short* mark = (short*)take(sizeof(short) * Dmax);
    short* scn = (short*)take(sizeof(short) * Dmax);
    short* lostlist = (short*)take(sizeof(short) * Tmax);
    unsigned char* role = take(Tmax);
    unsigned char* rowsel = take(Tmax);
    unsigned char* dflag = take(Dmax);
    unsigned char* cat = take(Tmax);
    unsigned char* drop = take(Tmax + Dmax);
    if (sm) {
        sm->tf = tf; sm->tbox = tbox; sm->dxywh = dxywh; sm->dbox = dbox; sm->dconf = dconf; sm->dcls = dcls;
        sm->u = u; sm->v = v; sm->dist = dist; sm->scratch = scratch; sm->ti = ti; sm->parent = parent;
        sm->head = head; sm->adj = adj; sm->colbits = colbits; sm->rnext = rnext; sm->xr = xr; sm->yc = yc;
        sm->pred = pred; sm->nextc = nextc; sm->mark = mark; sm->scn = scn; sm->lostlist = lostlist;
        sm->role = role; sm->rowsel = rowsel; sm->dflag = dflag; sm->cat = cat; sm->drop = drop;
    }
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

struct IouCost {
    const double *tbox, *dbox, *dconf;
    int Tmax, Dmax;
    bool fuse;
    __device__ __forceinline__ double operator()(int t, int j) const {
        const Box a = load_box(tbox, Tmax, t), b = load_box(dbox, Dmax, j);
        const double v = box_iou(a, b);
        return fuse ? fused_cost(v, dconf[j]) : xsub(1.0, v);
    }
};

template <int NT>
__device__ void build_colbits(const Sm& sm, int nd, int words, int want, int forbid) {
    for (int j = threadIdx.x; j < words * 32; j += NT) {
        const bool ok = j < nd && (sm.dflag[j] & want) && !(sm.dflag[j] & forbid);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0) sm.colbits[j >> 5] = m;
    }
    __syncthreads();
}

// candidate graph: edge (t, j) iff the boxes overlap and cost <= limit (exact pruning, see
// lap_sparse.cuh).  No overlap => iou == 0 => cost == 1 > limit, so the cheap overlap test
// rejects almost every pair before any division.
template <int NT, class Cost>
__device__ void build_adjacency(const Sm& sm, int Tmax, int Dmax, int n, int words, double limit, const Cost& cost) {
    for (int task = threadIdx.x; task < n * words; task += NT) {
        const int wd = task / n, t = task - wd * n;
        uint32_t bits = sm.rowsel[t] ? sm.colbits[wd] : 0u;
        uint32_t res = 0u;
        if (bits) {
            const Box a = load_box(sm.tbox, Tmax, t);
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const int j = wd * 32 + b;
                const Box d = load_box(sm.dbox, Dmax, j);
                if (box_overlap(a, d)) {
                    if (cost(t, j) <= limit) res |= 1u << b;
                }
            }
        }
        sm.adj[wd * Tmax + t] = res;
    }
    __syncthreads();
}

// STrack.update / re_activate for every matched row (byte_tracker.py:64-98)
template <int NT, int KIND>
__device__ void apply_matches(const Sm& sm, int Tmax, int Dmax, int n, int frame) {
    for (int t = threadIdx.x; t < n; t += NT) {
        if (!sm.rowsel[t]) continue;
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
    __syncthreads();
}

template <int NT, int KIND>
__global__ void __launch_bounds__(NT) bytetrack_step_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    const int tid = threadIdx.x;
    const int Tmax = p.max_tracks, Dmax = p.max_dets;
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

    // ---- detections: [nd, 6] rows -> planar shared memory ------------------------------
    {
        const double* g = p.dets + (size_t)s * Dmax * 6;
        for (int i = tid; i < nd * 6; i += NT) {
            const double val = g[i];
            const int j = i / 6, c = i - 6 * j;
            if (c < 4) sm.dbox[c * Dmax + j] = val;
            else if (c == 4) sm.dconf[j] = val;
            else sm.dcls[j] = val;
        }
    }
    // ---- track state: HBM -> shared memory, once ---------------------------------------
    {
        const double* gf = p.state_f + (size_t)s * B200_NF * Tmax;
        const int* gi = p.state_i + (size_t)s * B200_NI * Tmax;
        for (int t = tid; t < n; t += NT) {
#pragma unroll
            for (int c = 0; c < B200_NF; ++c) sm.tf[c * Tmax + t] = gf[c * Tmax + t];
#pragma unroll
            for (int c = 0; c < B200_NI; ++c) sm.ti[c * Tmax + t] = gi[c * Tmax + t];
        }
    }
    __syncthreads();

    // detection side: xyxy -> xywh, the round-trip box used by iou_distance, and the two
    // confidence bands (byte_tracker.py:151-158; strict inequalities on both sides)
    for (int j = tid; j < Dmax; j += NT) {
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
        }
        sm.dflag[j] = (unsigned char)fl;
    }
    // track side: roles, Kalman predict of the pool (unconfirmed tracks are NOT predicted), boxes
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
        sm.rowsel[t] = role != ROLE_UNCONF;
        sm.cat[t] = CAT_NONE;
    }
    for (int i = tid; i < Tmax + Dmax; i += NT) sm.drop[i] = 0;
    __syncthreads();

    LapWork lw;
    lw.Tmax = Tmax; lw.Dmax = Dmax; lw.adj = sm.adj; lw.u = sm.u; lw.v = sm.v; lw.dist = sm.dist;
    lw.parent = sm.parent; lw.head = sm.head; lw.rnext = sm.rnext; lw.xr = sm.xr; lw.yc = sm.yc;
    lw.pred = sm.pred; lw.nextc = sm.nextc; lw.mark = sm.mark; lw.scn = sm.scn;
    IouCost cost;
    cost.tbox = sm.tbox; cost.dbox = sm.dbox; cost.dconf = sm.dconf; cost.Tmax = Tmax; cost.Dmax = Dmax;

    // ---- first association: pool x high detections, fused score, limit match_thresh ----
    cost.fuse = true;
    build_colbits<NT>(sm, nd, words, DF_HIGH, 0);
    build_adjacency<NT>(sm, Tmax, Dmax, n, words, p.match_thresh, cost);
    lap_sparse_solve<NT>(lw, n, words, p.match_thresh, cost);
    apply_matches<NT, KIND>(sm, Tmax, Dmax, n, frame);

    // ---- second association: still-Tracked leftovers x low detections, plain IoU, 0.5 ---
    for (int t = tid; t < n; t += NT) {
        const bool matched = sm.rowsel[t] && sm.xr[t] >= 0;
        if (matched && sm.role[t] == ROLE_LOST) sm.cat[t] = CAT_REFOUND;
        sm.rowsel[t] = sm.role[t] == ROLE_TRACKED && !matched;
    }
    __syncthreads();
    cost.fuse = false;
    build_colbits<NT>(sm, nd, words, DF_LOW, 0);
    build_adjacency<NT>(sm, Tmax, Dmax, n, words, p.second_thresh, cost);
    lap_sparse_solve<NT>(lw, n, words, p.second_thresh, cost);
    apply_matches<NT, KIND>(sm, Tmax, Dmax, n, frame);
    for (int t = tid; t < n; t += NT) {
        if (sm.rowsel[t] && sm.xr[t] < 0) {             // mark_lost; frame_id stays = end_frame
            const int fl = sm.ti[B200_TI_FLAGS * Tmax + t];
            sm.ti[B200_TI_FLAGS * Tmax + t] = (fl & ~3) | B200_ST_LOST;
        }
        sm.rowsel[t] = sm.role[t] == ROLE_UNCONF;
    }
    __syncthreads();

    // ---- unconfirmed x remaining high detections, fused score, 0.7 ---------------------
    cost.fuse = true;
    build_colbits<NT>(sm, nd, words, DF_HIGH, DF_USED);
    build_adjacency<NT>(sm, Tmax, Dmax, n, words, p.unconf_thresh, cost);
    lap_sparse_solve<NT>(lw, n, words, p.unconf_thresh, cost);
    apply_matches<NT, KIND>(sm, Tmax, Dmax, n, frame);

    // ---- lifecycle: removed / aged-out, list categories (byte_tracker.py:237-268) -------
    for (int t = tid; t < n; t += NT) {
        int fl = sm.ti[B200_TI_FLAGS * Tmax + t];
        const int role = sm.role[t];
        if (role == ROLE_UNCONF && sm.xr[t] < 0) fl = (fl & ~3) | B200_ST_REMOVED;
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

    // ---- destinations.  Packed counters (10 bits each):
    //   slots: keep, keep&activated, refound, lostOld, lostNew ; dets: born, born&activated
    int totKeep, totKeepAct, totRef, totLostOld, totLostNew, totBorn, totBornAct;
    const bool born_active = frame == 1;                 // STrack.activate: is_activated only on frame 1
    double* gf = p.state_f + (size_t)s * B200_NF * Tmax;
    int* gi = p.state_i + (size_t)s * B200_NI * Tmax;
    double* gout = p.out + (size_t)s * Tmax * 8;
    {
        // pass 1: totals
        unsigned long long acc = 0;
        const int m = max(n, nd);
        // per-thread element values are recomputed in pass 2; keep them cheap
        auto slot_val = [&](int t) -> unsigned long long {
            if (t >= n || sm.drop[t]) return 0ull;
            const int cat = sm.cat[t];
            const unsigned long long act = (sm.ti[B200_TI_FLAGS * Tmax + t] & B200_FLAG_ACTIVATED) ? 1ull : 0ull;
            if (cat == CAT_KEEP) return 1ull | (act << 10);
            if (cat == CAT_REFOUND) return 1ull << 20;
            if (cat == CAT_LOST_OLD) return 1ull << 30;
            if (cat == CAT_LOST_NEW) return 1ull << 40;
            return 0ull;
        };
        auto det_val = [&](int j) -> unsigned long long {
            if (j >= nd || sm.drop[Tmax + j]) return 0ull;
            if ((sm.dflag[j] & (DF_HIGH | DF_USED)) != DF_HIGH) return 0ull;
            if (sm.dconf[j] < p.new_thresh) return 0ull;
            return (1ull << 50);
        };
        // block totals first (the segment bases depend on them)
        for (int c0 = 0; c0 < m; c0 += NT) {
            unsigned long long tot;
            block_exscan<NT>(slot_val(c0 + tid) + det_val(c0 + tid), sm.scratch, tot);
            acc += tot;
        }
        totKeep = (int)(acc & 1023); totKeepAct = (int)((acc >> 10) & 1023); totRef = (int)((acc >> 20) & 1023);
        totLostOld = (int)((acc >> 30) & 1023); totLostNew = (int)((acc >> 40) & 1023);
        totBorn = (int)((acc >> 50) & 1023);
        totBornAct = born_active ? totBorn : 0;
        // the born tracks that were dropped as duplicates still consumed an id (activate ran
        // before remove_duplicate_stracks), so ids are numbered over ALL born tracks below.
        int newT = totKeep + totBorn + totRef;
        int newL = totLostOld + totLostNew;
        if (newT + newL > Tmax) err |= B200_ERR_TRACK_OVERFLOW;

        // pass 2: scatter
        unsigned long long base = 0;
        unsigned long long idbase = 0;
        for (int c0 = 0; c0 < m; c0 += NT) {
            const int i = c0 + tid;
            const unsigned long long sv = slot_val(i), dv = det_val(i);
            // id numbering counts every born track, dropped or not
            unsigned long long bornraw = 0ull;
            if (i < nd && (sm.dflag[i] & (DF_HIGH | DF_USED)) == DF_HIGH && !(sm.dconf[i] < p.new_thresh)) bornraw = 1ull;
            unsigned long long tot, idtot;
            const unsigned long long ex = block_exscan<NT>(sv + dv, sm.scratch, tot) + base;
            const unsigned long long idex = block_exscan<NT>(bornraw, sm.scratch, idtot) + idbase;
            base += tot; idbase += idtot;
            if (sv) {
                const int cat = sm.cat[i];
                int dst, orow = -1;
                if (cat == CAT_KEEP) { dst = (int)(ex & 1023); if (sv >> 10) orow = (int)((ex >> 10) & 1023); }
                else if (cat == CAT_REFOUND) { dst = totKeep + totBorn + (int)((ex >> 20) & 1023); orow = totKeepAct + totBornAct + (int)((ex >> 20) & 1023); }
                else if (cat == CAT_LOST_OLD) dst = newT + (int)((ex >> 30) & 1023);
                else dst = newT + totLostOld + (int)((ex >> 40) & 1023);
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
            if (dv) {                                   // STrack.activate (byte_tracker.py:50-62)
                const int j = i;
                const int k = (int)((ex >> 50) & 1023);
                const int dst = totKeep + k;
                double z[4];
                det_measurement<KIND>(sm, Dmax, j, z);
                KfState ks;
                kf_initiate<KIND>(z, ks);
                const int id = id0 + (int)idex + 1;
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
                    const int orow = totKeepAct + k;
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
            counts[2] = id0 + (int)idbase;
            counts[3] = frame;
            p.nout[s] = min(totKeepAct + totBornAct + totRef, Tmax);
            p.track_updates[s] += (unsigned long long)n;
            if (err) atomicOr(p.err, err);
        }
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
