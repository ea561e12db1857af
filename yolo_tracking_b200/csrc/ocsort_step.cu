// OC-SORT frame step for many independent streams: one kernel launch per frame, one CTA per
// stream, one thread per tracker slot.
//
// Replaces OCSort.update (boxmot/trackers/ocsort/ocsort.py:218-379) and what it calls:
// KalmanBoxTracker.predict / update :130-181 -> KalmanFilter.predict / update / freeze /
// unfreeze (boxmot/motion/kalman_filters/ocsort_kf.py:339-526), associate
// (boxmot/utils/association.py:111-201: velocity-direction cost, permutation shortcut, lapjv
// without a limit), the observation-centric recovery round :319-345, new trackers :351-353 and
// the reversed output scan :354-379.
//
// The 7-d filter (F couples x-vx, y-vy, s-vs; H = [I 0]; diagonal Q, R) keeps an exactly
// block-sparse covariance: three position/velocity 2x2 blocks and P_rr, so S is diagonal and the
// reference's inv(S) / Joseph form reduce to a handful of scalar operations per axis.
//
// Unlike ByteTrack the assignment is DENSE: lapjv is called without a cost limit on
// -(similarity + angle cost), every min(D, T) row is matched and pairs are filtered by the
// similarity threshold afterwards, so no pair can be dropped from the problem.  But the matrix
// is never stored (lap_dense.cuh is matrix-free: a cost is re-evaluated from the shared-memory
// resident boxes when the solver asks for it), and the row reduction that starts the solver
// does not evaluate it in full either: for iou / giou a disjoint pair has similarity exactly 0
// and the direction term is bounded by |inertia| * conf / 2, so once the best cost among the
// OVERLAPPING columns of a row (one to three, found by a conservative fp32 test) lies below
// that bound no other column can be the row minimum; only rows without such a column (new
// objects, false positives) are evaluated in full.  Slots are updated in place; a tracker
// that dies leaves a hole and the stream is compacted only when its slot range runs short.
#include <cstdlib>

#include "boxes.cuh"
#include "lap_dense.cuh"
#include "oc_common.cuh"
#include "kf_xysr.cuh"
#include "layout.h"
#include "step_params.h"

namespace b200 {
namespace {

template <int TMAX, int DMAX>
struct alignas(16) OcSmem {
    double tbox[4][TMAX];           // predicted box, convert_x_to_bbox
    double lbox[4][TMAX];           // last_observation box (placeholder -1)
    double kc[2][TMAX];             // centre of k_previous_obs
    double vel[2][TMAX];            // (vy, vx)
    double dbox[4][DMAX];
    double dconf[DMAX];
    double u[DMAX], v[TMAX];
    double red_v[64];
    unsigned long long scratch[40];
    // candidate pairs of the first association (boxes that overlap): row-major segments, exact cost per pair
    static constexpr int PCAP = 3 * DMAX;
    double pcost[PCAP];
    uint32_t ppair[PCAP];           // (row << 16) | column
    int segstart[DMAX];
    short segcnt[DMAX];             // -1: the row's candidates did not fit (evaluated in full instead)
    int pred[TMAX], xr[DMAX], yc[TMAX], claim[DMAX], partner[TMAX];
    int red_i[64];
    int rowcnt[DMAX], rowmatch[DMAX], colcnt[TMAX];
    int misc[8];
    short hd[DMAX], ht[TMAX], dmatch[DMAX], tmatch[TMAX], ud[DMAX], ut[TMAX];
    float4 cboxf[TMAX];             // predicted box of COLUMN c rounded outwards to fp32 (irregular boxes: everything), step A
    double dred[8 * (TMAX / 32)];   // per (row, column chunk) minima of the rows evaluated in full
    int ired[8 * (TMAX / 32)];
    unsigned char kvalid[TMAX], alive[TMAX], dstate[DMAX], tdeg[TMAX], ddeg[DMAX], rowfull[DMAX];
};

// Cost of (row r = high detection hd[r], column c = live tracker ht[c]) in the first association,
// association.py:130-172: -(similarity + direction term) plus the canonical tie-break.  The row reduction and the
// solver evaluate it through this one object, so they see the same bits.
template <class SM>
struct OcCost1 {
    const SM& sm;
    int func, Cn;
    double W, H, inertia;
    bool sparse;                    // iou / giou with a non-negative threshold: a disjoint pair has similarity exactly +0.0
    __device__ __forceinline__ Box dbox(int j) const { return Box{sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]}; }
    __device__ __forceinline__ Box tbox(int sl) const { return Box{sm.tbox[0][sl], sm.tbox[1][sl], sm.tbox[2][sl], sm.tbox[3][sl]}; }
    // |direction term| + ties of a row, rounded up
    __device__ __forceinline__ double bound(int j) const { return xmul(0.5, fabs(xmul(inertia, sm.dconf[j]))) + 1e-6; }
    // provably similarity-free pair: regular boxes that do not overlap
    __device__ __forceinline__ bool free_pair(int j, int sl) const {
        return sparse && !sm.ddeg[j] && !sm.tdeg[sl] && !box_overlap(dbox(j), tbox(sl));
    }
    __device__ __forceinline__ double sim(int r, int c) const { return oc_sim(func, dbox(sm.hd[r]), tbox(sm.ht[c]), W, H); }
    __device__ __forceinline__ double from_sim(int r, int c, double sv) const {
        const int j = sm.hd[r], sl = sm.ht[c];
        double ang = 0.0;
        const double vy = sm.vel[0][sl], vx = sm.vel[1][sl];
        if (sm.kvalid[sl] && !(vx == 0.0 && vy == 0.0))
            ang = oc_angle(vy, vx, sm.kc[0][sl], sm.kc[1][sl], true, xdiv(xadd(sm.dbox[0][j], sm.dbox[2][j]), 2.0),
                           xdiv(xadd(sm.dbox[1][j], sm.dbox[3][j]), 2.0), inertia, sm.dconf[j]);
        return xadd(-xadd(xadd(sv, ang), 0.0), xmul((double)(r * Cn + c), TIE_EPS));
    }
    // A pair that is not similarity-free went through the pair list (same test) unless its row overflowed it: the
    // solver then reads the cost the row reduction computed instead of paying a second ~3000-cycle evaluation - after
    // the first step of a search every evaluation is such a pair.
    __device__ __forceinline__ double operator()(int r, int c) const {
        const int j = sm.hd[r], sl = sm.ht[c];
        if (free_pair(j, sl)) return from_sim(r, c, 0.0);
        if (sparse) {
            const int n = sm.segcnt[r], b0 = sm.segstart[r];
            for (int k = 0; k < n; ++k)
                if ((int)(sm.ppair[b0 + k] & 0xffff) == c) return sm.pcost[b0 + k];
        }
        return from_sim(r, c, sim(r, c));
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

// DENSE = tuned for many streams per SM: three CTAs of the 224 / 256 variants per SM (80 registers, some spills) instead of
// two at 124 registers - 12 % faster once every SM has several streams to overlap, 12 % slower when it has one
// (BASELINE config 2: 64 streams); the launcher picks by stream count.
template <int NT, int TMAX, int DMAX, bool DENSE>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : (NT >= 224 ? (DENSE ? 3 : 2) : (NT == 128 ? 4 : 8))))
ocsort_step_kernel(const StepParams p) {
    static_assert(NT == TMAX && DMAX <= NT, "one thread per tracker slot; detections fit one pass");
    using SM = OcSmem<TMAX, DMAX>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int s = blockIdx.x, tid = threadIdx.x, t = tid;
    // optional per-phase cycle counters (thread 0 of every CTA; b200track_phase_cycles)
    long long ph_last = p.dbg ? clock64() : 0;
#define PHASE(k) do { if (p.dbg && tid == 0) { const long long now_ = clock64(); atomicAdd(&p.dbg[k], (unsigned long long)(now_ - ph_last)); ph_last = now_; } } while (0)
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
    double* gf = p.state_f + (size_t)s * B200_OC_NF * TMAX;
    int* gi = p.state_i + (size_t)s * B200_OC_NI * TMAX;
    const double thr = p.iou_thresh, W = p.img_w, H = p.img_h;
    const int func = p.asso_func;

    // state and detections of a stream a little ahead (one per SM) -> L2: its dependent first loads (flags -> state ->
    // observation ring) then cost L2 hits instead of DRAM round trips
    {
        const int s2 = s + 148;
        if (s2 < p.n_streams) {
            const char* d2 = reinterpret_cast<const char*>(p.dets + (size_t)s2 * p.max_dets * 6);
            int nl_d = (p.max_dets * 48 + 127) >> 7;
            if (packed) {
                const int o2 = p.det_off[s2], n2 = p.det_off[s2 + 1] - o2;
                d2 = p.dets32 ? reinterpret_cast<const char*>(p.dets32 + (size_t)o2 * 6) : reinterpret_cast<const char*>(p.dets + (size_t)o2 * 6);
                nl_d = (n2 * (p.dets32 ? 24 : 48) + 127) >> 7;
            }
            const char* f2 = reinterpret_cast<const char*>(p.state_f + (size_t)s2 * B200_OC_NF * TMAX);
            const char* i2 = reinterpret_cast<const char*>(p.state_i + (size_t)s2 * B200_OC_NI * TMAX);
            // only the slot range in use of every hot component row (a row is TMAX slots long; prefetching whole rows read
            // twice the state from DRAM)
            const int r2 = min(p.counts[4 * s2], TMAX);
            const int lf = (r2 * 8 + 127) >> 7, li = (r2 * 4 + 127) >> 7;            // lines per fp64 / int32 component row
            const int nl_f = B200_OC_SX * lf, nl_i = B200_OC_NI * li;
            for (int l = tid; l < nl_d + nl_f + nl_i; l += NT) {
                const char* a;
                if (l < nl_d) a = d2 + ((size_t)l << 7);
                else if (l < nl_d + nl_f) { const int q = l - nl_d, c = q / lf; a = f2 + (size_t)c * TMAX * 8 + ((size_t)(q - c * lf) << 7); }
                else { const int q = l - nl_d - nl_f, c = q / li; a = i2 + (size_t)c * TMAX * 4 + ((size_t)(q - c * li) << 7); }
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    // ---- HBM -> shared memory: detections, hot part of the tracker state --------------------
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
    // the class column is only read for matched / new trackers, straight from the detection row (L2)
    const double* dets_g = packed ? p.dets + (size_t)roff * 6 : p.dets + (size_t)s * p.max_dets * 6;
    const float* dets32_g = p.dets32 ? p.dets32 + (size_t)roff * 6 : nullptr;
    auto det_cls = [&](int j) -> double { return dets32_g ? (double)dets32_g[j * 6 + 5] : dets_g[j * 6 + 5]; };
    // detections above det_thresh: the only ones that can start a tracker this frame
    __syncthreads();
    const int nhigh = __syncthreads_count(tid < nd && sm.dconf[tid] > p.det_thresh);
    // ---- compaction on demand: dead trackers leave holes in the slot range; the live ones move down (order kept) only
    // when this frame's detections above det_thresh - an upper bound of its new trackers - might not fit behind the range otherwise
    // (every ~10-20 frames at config 2; doing it whenever the range passed max_tracks - max_dets meant every frame)
    if (n0 > alive0 && n0 + nhigh > tcap) {              // uniform
        bool lv = false;
        if (t < n0) lv = gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE;
        unsigned long long tt;
        const int dst = (int)block_exscan<NT>(lv ? 1ull : 0ull, sm.scratch, tt);
        for (int c0 = 0; c0 < B200_OC_NF; c0 += 8) {
            double tmp[8];
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_OC_NF) tmp[c] = gf[(c0 + c) * TMAX + t];
            }
            __syncthreads();
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_OC_NF) gf[(c0 + c) * TMAX + dst] = tmp[c];
            }
            __syncthreads();
        }
        int itmp[B200_OC_NI];
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_OC_NI; ++c) itmp[c] = gi[c * TMAX + t];
        }
        __syncthreads();
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_OC_NI; ++c) gi[c * TMAX + dst] = itmp[c];
        }
        __syncthreads();
        n0 = (int)tt;
    }
    int fl = 0, age = 0, tsu = 0, streak = 0;
    bool live = false;
    if (t < n0) {
        fl = gi[B200_OCI_FLAGS * TMAX + t];
        live = fl & OCF_ALIVE;
    }
    if (live) {
        age = gi[B200_OCI_AGE * TMAX + t];
        tsu = gi[B200_OCI_TSU * TMAX + t];
        streak = gi[B200_OCI_STREAK * TMAX + t];
        double x[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) x[c] = gf[(B200_OC_X + c) * TMAX + t];
        // KalmanBoxTracker.predict (ocsort.py:168-181); the covariance half is deferred
        if (xadd(x[6], x[2]) <= 0.0) x[6] = xmul(x[6], 0.0);
#pragma unroll
        for (int c = 0; c < 3; ++c) x[c] = xadd(x[c], x[c + 4]);
        age += 1;
        if (tsu > 0) streak = 0;
        tsu += 1;
        // the predicted state itself is only needed again by this thread after the association: it is re-derived from
        // the same (L2-resident) loads there instead of occupying 7 x TMAX doubles of shared memory
        const Box b = oc_x_to_box(x[0], x[1], x[2], x[3]);
        sm.tbox[0][t] = b.x1; sm.tbox[1][t] = b.y1; sm.tbox[2][t] = b.x2; sm.tbox[3][t] = b.y2;
        sm.tdeg[t] = !oc_regular_box(b.x1, b.y1, b.x2, b.y2);
        if (isnan(b.x1) || isnan(b.y1) || isnan(b.x2) || isnan(b.y2)) { live = false; fl &= ~OCF_ALIVE; }   // :260-264
        const bool hasobs = fl & B200_OCF_HASOBS;
        double l[4] = {-1.0, -1.0, -1.0, -1.0};
        if (hasobs) {
#pragma unroll
            for (int c = 0; c < 4; ++c) l[c] = gf[(B200_OC_LAST + c) * TMAX + t];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.lbox[c][t] = l[c];
        sm.vel[0][t] = gf[(B200_OC_VEL + 0) * TMAX + t];
        sm.vel[1][t] = gf[(B200_OC_VEL + 1) * TMAX + t];
        // k_previous_obs (ocsort.py:14-22): oldest observation among ages age-3 .. age-1, else the newest
        double kb[4] = {l[0], l[1], l[2], l[3]};
        if (hasobs) {
            const int ra[3] = {gi[(B200_OCI_RINGAGE + 0) * TMAX + t], gi[(B200_OCI_RINGAGE + 1) * TMAX + t], gi[(B200_OCI_RINGAGE + 2) * TMAX + t]};
            for (int dt = p.delta_t; dt >= 1; --dt) {
                const int a = age - dt;
                if (a < 0) continue;
                const int slot = a % 3;
                if (dt <= 3 && ra[slot] == a) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) kb[c] = gf[(B200_OC_RING + 4 * slot + c) * TMAX + t];
                    break;
                }
            }
        }
        sm.kc[0][t] = xdiv(xadd(kb[0], kb[2]), 2.0);
        sm.kc[1][t] = xdiv(xadd(kb[1], kb[3]), 2.0);
        sm.kvalid[t] = hasobs;
    }
    if (t < TMAX) { sm.alive[t] = live; sm.tmatch[t] = -1; }
    if (tid < DMAX) { sm.dmatch[tid] = -1; sm.rowcnt[tid] = 0; sm.rowmatch[tid] = -1; }
    __syncthreads();
    PHASE(1);
    if (tid < DMAX) sm.dstate[tid] = (tid < nd && sm.dconf[tid] > p.det_thresh) ? DS_FREE0 : DS_NONE;    // ocsort.py:250-251

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
    PHASE(2);

    // ---- first round: associate(dets, trks, ...) ------------------------------------------------
    if (tid < DMAX && tid < nd) sm.ddeg[tid] = !oc_regular_box(sm.dbox[0][tid], sm.dbox[1][tid], sm.dbox[2][tid], sm.dbox[3][tid]);
    OcCost1<SM> cost1{sm, func, Cn, W, H, p.inertia, func <= 1 && thr >= 0.0};
    if (R > 0 && Cn > 0) {
        const int lane = tid & 31, warp = tid >> 5;
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        constexpr int NCH = TMAX / 32;
        double amax = 0.0;                           // bound of the direction term over the rows
        if (tid < Cn) {
            sm.colcnt[tid] = 0;
            const int sl = sm.ht[tid];
            const float FINF = __int_as_float(0x7f800000);
            sm.cboxf[tid] = sm.tdeg[sl] ? make_float4(-FINF, -FINF, FINF, FINF)
                                        : make_float4(__double2float_rd(sm.tbox[0][sl]), __double2float_rd(sm.tbox[1][sl]),
                                                      __double2float_ru(sm.tbox[2][sl]), __double2float_ru(sm.tbox[3][sl]));
        }
        if (tid == 0) { sm.misc[0] = 0; sm.misc[1] = SM::PCAP; }
        __syncthreads();
        // Row reduction (lap_dense.cuh step 1) without evaluating the matrix:
        //   A. one warp per row marks the columns whose box may overlap the row's and appends them to the pair list (one
        //      segment per row);
        //   B. one thread per pair: similarity, threshold counts for the permutation shortcut, exact cost;
        //   C. one thread per row: minimum over its segment; if it is below -bound(row) no column outside the segment
        //      (similarity exactly 0, |direction term| <= bound) can be the row minimum;
        //   D. the remaining rows (new objects, false positives; or every row for diou / ciou / centroid, whose similarity
        //      is dense) are evaluated in full, one warp per row.
        if (cost1.sparse) {
            for (int r = warp; r < R; r += NT / 32) {
                const int j = sm.hd[r];
                // conservative fp32 test (boxes rounded outwards; an irregular box overlaps everything): a superset of the
                // pairs the exact test of free_pair() keeps - a pair listed needlessly is evaluated exactly all the same
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
                        if (bits[k]) {                        // uniform
                            if (bits[k] & (1u << lane)) sm.ppair[base + __popc(bits[k] & ((1u << lane) - 1u))] = ((uint32_t)r << 16) | (uint32_t)(k * 32 + lane);
                            base += __popc(bits[k]);
                        }
                    }
                    base -= total;
                }
                if (lane == 0) {
                    sm.segstart[r] = base; sm.segcnt[r] = fits ? (short)total : (short)-1;
                    if (!fits) atomicMin(&sm.misc[1], base);      // bases only grow: every entry below the first overflow is written
                }
            }
            __syncthreads();
            const int np = min(sm.misc[0], sm.misc[1]);
            for (int k = tid; k < np; k += NT) {
                const uint32_t pr = sm.ppair[k];
                const int r = pr >> 16, c = pr & 0xffff;
                const double sv = cost1.sim(r, c);
                if (sv > thr) { atomicAdd(&sm.colcnt[c], 1); atomicAdd(&sm.rowcnt[r], 1); sm.rowmatch[r] = c; }
                sm.pcost[k] = cost1.from_sim(r, c, sv);
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
                sm.rowfull[r] = n < 0 ? 2 : (m < -bound ? 0 : 1);     // 2: nothing of the row was evaluated yet
            }
        } else if (tid < R) {
            sm.rowfull[tid] = 2;
            amax = cost1.bound(sm.hd[tid]);
        }
        __syncthreads();
        // D: tasks = (row to evaluate in full, chunk of 32 columns), dealt over the warps - an evaluation is ~3000 cycles of
        // dependent fp64 work, so a warp walking all chunks of its row would serialise them.  Task results meet in pcost /
        // ppair (the pair list is dead for these rows: mode 1 rows keep theirs, and their tasks skip the listed columns).
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
        for (int f0 = 0; f0 < nfull; f0 += 8) {           // 8 rows per round: their chunk results fit the scratch below
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
                    if (mode == 2 || fp) {                // mode 1: the other columns went through the pair list
                        const double sv = fp ? 0.0 : cost1.sim(r, c);
                        if (sv > thr) { atomicAdd(&sm.colcnt[c], 1); atomicAdd(&sm.rowcnt[r], 1); sm.rowmatch[r] = c; }
                        m = cost1.from_sim(r, c, sv);
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
        __syncthreads();
        PHASE(3);
        int ccnt = tid < Cn ? sm.colcnt[tid] : 0;
        int rc = tid < R ? sm.rowcnt[tid] : 0;
        block_max3<NT>(sm, amax, ccnt, rc);
        const bool shortcut = (rc == 1 && ccnt == 1);                          // association.py:157-159
        if (shortcut) {
            if (tid < R) sm.xr[tid] = sm.rowcnt[tid] == 1 ? sm.rowmatch[tid] : -1;
            __syncthreads();
        } else {
            DenseLap w = make_dense<NT>(sm);
            w.dbg = p.dbg;
            const double lambda = 2.0 * (amax + 1.0);          // >= lapjv's 2 * (max cost + 1): same assignment (lap_dense.cuh)
            dense_lap_init<NT>(w, R, Cn, lambda);
            PHASE(4);
            dense_lap_augment<NT>(w, cost1, R, Cn, lambda);
        }
        PHASE(5);
        // matched pairs below the similarity threshold fall back to unmatched (association.py:187-193)
        if (tid < R) {
            const int c = sm.xr[tid];
            const int j = sm.hd[tid];
            if (c >= 0) {
                const int sl = sm.ht[c];
                const Box tb = {sm.tbox[0][sl], sm.tbox[1][sl], sm.tbox[2][sl], sm.tbox[3][sl]};
                const Box db = {sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]};
                if (oc_sim(func, db, tb, W, H) < thr) sm.dstate[j] = DS_FREE1;
                else { sm.dstate[j] = DS_MATCHED; sm.dmatch[j] = (short)sl; sm.tmatch[sl] = (short)j; }
            }
        }
        __syncthreads();
    }

    // ---- second-round associations on leftovers: BYTE (ocsort.py:293-317, only with use_byte) and the
    // observation-centric recovery (:319-345).  Both are: similarity of a detection list x a tracker list,
    // and if any entry clears the threshold, a no-limit assignment whose pairs are kept when they clear it.
    // rows = sm.ud[0..nr), columns = sm.ut[0..nc) (slot indices); boxes of the columns: predicted or last observed.
    auto second_round = [&](int nr, int nc, bool last_boxes) -> bool {
        if (nr <= 0 || nc <= 0) return false;                        // uniform
        const double (*bx)[TMAX] = last_boxes ? sm.lbox : sm.tbox;
        auto sim2 = [&](int r, int c) -> double {
            const int j = sm.ud[r], sl = sm.ut[c];
            const Box tb = {bx[0][sl], bx[1][sl], bx[2][sl], bx[3][sl]};
            const Box db = {sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]};
            return oc_sim(func, db, tb, W, H);
        };
        struct Cost2 {
            decltype(sim2)& sim;
            int nc;
            __device__ __forceinline__ double operator()(int r, int c) const { return xadd(-sim(r, c), xmul((double)(r * nc + c), TIE_EPS)); }
            __device__ __forceinline__ double lower(int, int) const { return -__longlong_as_double(0x7ff0000000000000LL); }
        } cost2{sim2, nc};
        // row reduction in full (these problems are small and rare), one warp per row
        double smax = -1e300;
        {
            const int lane = tid & 31, warp = tid >> 5;
            const double INF = __longlong_as_double(0x7ff0000000000000LL);
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
        }
        int d1 = 0, d2 = 0;
        block_max3<NT>(sm, smax, d1, d2);
        if (!(smax > thr)) return false;
        const DenseLap w = make_dense<NT>(sm);
        // costs are -similarity (+ ties): at most 1e-6 for every similarity whose range starts at 0, else bounded by 1 + 1e-6
        const double lambda = 2.0 * ((func <= 1 ? 0.0 : 1.0) + 1e-6 + 1.0);
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
        return true;
    };
    // unmatched lists in associate()'s order (association.py:179-193): never matched first (ascending), then the
    // members of low-similarity matches in match order (= ascending detection index).  `sorted_trks`: the tracker
    // list went through np.setdiff1d (ascending) because the BYTE stage ran its assignment.
    if (t < TMAX) sm.partner[t] = -1;
    __syncthreads();
    if (tid < R && Cn > 0 && sm.xr[tid] >= 0 && sm.dstate[sm.hd[tid]] == DS_FREE1) sm.partner[sm.ht[sm.xr[tid]]] = tid;   // slot -> row of its partner
    __syncthreads();
    auto build_lists = [&](bool low_dets, bool sorted_trks, int& nUd, int& nUt) {
        const int ds = tid < DMAX ? sm.dstate[tid] : DS_NONE;
        const bool islow = low_dets && tid < nd && sm.dconf[tid] > 0.1 && sm.dconf[tid] < p.det_thresh;     // ocsort.py:244-249
        const bool ufree = live && sm.tmatch[t] < 0;
        const bool ufree1 = ufree && !sorted_trks && sm.partner[t] >= 0;
        const bool ufree0 = ufree && !ufree1;
        const bool d0 = low_dets ? islow : ds == DS_FREE0, dlate = !low_dets && ds == DS_FREE1;
        const unsigned long long val = (d0 ? 1ull : 0ull) | (dlate ? (1ull << 16) : 0ull) | (ufree0 ? (1ull << 32) : 0ull) |
                                       (ds == DS_FREE1 ? (1ull << 48) : 0ull);
        unsigned long long tot;
        const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);
        const int n0d = (int)(tot & 0xffff), n1d = (int)((tot >> 16) & 0xffff), n0t = (int)((tot >> 32) & 0xffff);
        if (d0) sm.ud[ex & 0xffff] = (short)tid;
        if (dlate) sm.ud[n0d + ((ex >> 16) & 0xffff)] = (short)tid;
        if (ufree0) sm.ut[(ex >> 32) & 0xffff] = (short)t;
        // low-similarity trackers follow in the order of their partner detections
        if (ds == DS_FREE1) sm.claim[tid] = (int)((ex >> 48) & 0xffff);          // rank of this detection among FREE1
        __syncthreads();
        int n1t = 0;
        if (!sorted_trks) {
            if (ufree1) sm.ut[n0t + sm.claim[sm.hd[sm.partner[t]]]] = (short)t;
            n1t = (int)((tot >> 48) & 0xffff);       // every FREE1 detection has exactly one (still unmatched) partner tracker
        }
        nUd = n0d + n1d; nUt = n0t + n1t;
        __syncthreads();
    };
    bool byte_ran = false;
    if (p.use_byte) {
        int nr, nc;
        build_lists(true, false, nr, nc);
        byte_ran = second_round(nr, nc, false);
    }
    bool ocr_ran;
    {
        int nr, nc;
        build_lists(false, byte_ran, nr, nc);
        ocr_ran = second_round(nr, nc, true);
    }
    PHASE(6);

    // ---- deferred Kalman work and bookkeeping, thread t = slot t ---------------------------------
    int hits = 0, det_ind = 0, tid_id = 0;
    double conf = 0.0, cls = 0.0;
    OcKf k;
    if (live) {
#pragma unroll
        for (int c = 0; c < 7; ++c) k.x[c] = gf[(B200_OC_X + c) * TMAX + t];
        if (xadd(k.x[6], k.x[2]) <= 0.0) k.x[6] = xmul(k.x[6], 0.0);          // the motion step of phase 1, same operations
#pragma unroll
        for (int c = 0; c < 3; ++c) k.x[c] = xadd(k.x[c], k.x[c + 4]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            k.pp[i] = gf[(B200_OC_P + 3 * i + 0) * TMAX + t];
            k.pv[i] = gf[(B200_OC_P + 3 * i + 1) * TMAX + t];
            k.vv[i] = gf[(B200_OC_P + 3 * i + 2) * TMAX + t];
        }
        k.prr = gf[(B200_OC_P + 9) * TMAX + t];
        oc_predict_cov(k);
        hits = gi[B200_OCI_HITS * TMAX + t];
        det_ind = gi[B200_OCI_DET * TMAX + t];
        tid_id = gi[B200_OCI_ID * TMAX + t];
        conf = gf[B200_OC_CONF * TMAX + t];
        cls = gf[B200_OC_CLS * TMAX + t];
        const int j = sm.tmatch[t];
        if (j >= 0) {                                                   // KalmanBoxTracker.update(bbox), ocsort.py:130-164
            const double b0 = sm.dbox[0][j], b1 = sm.dbox[1][j], b2 = sm.dbox[2][j], b3 = sm.dbox[3][j];
            const bool hasobs = fl & B200_OCF_HASOBS;
            const double lsum = hasobs ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
            if (lsum >= 0.0) {                                          // speed_direction(previous_box, bbox)
                const double cx2 = xdiv(xadd(b0, b2), 2.0), cy2 = xdiv(xadd(b1, b3), 2.0);
                const double dy = xsub(cy2, sm.kc[1][t]), dx = xsub(cx2, sm.kc[0][t]);
                const double norm = xadd(sqrt(xadd(xmul(dy, dy), xmul(dx, dx))), 1e-6);
                gf[(B200_OC_VEL + 0) * TMAX + t] = xdiv(dy, norm);
                gf[(B200_OC_VEL + 1) * TMAX + t] = xdiv(dx, norm);
            }
            conf = sm.dconf[j]; cls = det_cls(j); det_ind = j;
            gf[(B200_OC_LAST + 0) * TMAX + t] = b0; gf[(B200_OC_LAST + 1) * TMAX + t] = b1;
            gf[(B200_OC_LAST + 2) * TMAX + t] = b2; gf[(B200_OC_LAST + 3) * TMAX + t] = b3;
            const int rs = age % 3;
            gf[(B200_OC_RING + 4 * rs + 0) * TMAX + t] = b0; gf[(B200_OC_RING + 4 * rs + 1) * TMAX + t] = b1;
            gf[(B200_OC_RING + 4 * rs + 2) * TMAX + t] = b2; gf[(B200_OC_RING + 4 * rs + 3) * TMAX + t] = b3;
            gi[(B200_OCI_RINGAGE + rs) * TMAX + t] = age;
            double z[4];
            oc_box_to_z(b0, b1, b2, b3, z);
            bool virt = false;
            double vz[4];
            if (!(fl & B200_OCF_OBSERVED) && (fl & B200_OCF_SAVED)) {   // unfreeze: observation-centric re-update
#pragma unroll
                for (int c = 0; c < 7; ++c) k.x[c] = gf[(B200_OC_SX + c) * TMAX + t];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    k.pp[i] = gf[(B200_OC_SP + 3 * i + 0) * TMAX + t];
                    k.pv[i] = gf[(B200_OC_SP + 3 * i + 1) * TMAX + t];
                    k.vv[i] = gf[(B200_OC_SP + 3 * i + 2) * TMAX + t];
                }
                k.prr = gf[(B200_OC_SP + 9) * TMAX + t];
                const double lz[4] = {gf[(B200_OC_LASTZ + 0) * TMAX + t], gf[(B200_OC_LASTZ + 1) * TMAX + t],
                                      gf[(B200_OC_LASTZ + 2) * TMAX + t], gf[(B200_OC_LASTZ + 3) * TMAX + t]};
                virt = oc_virtual_trajectory(k, lz, z, tsu, vz);       // gap = index2 - index1 of history_obs
                fl &= ~B200_OCF_SAVED;
            }
            fl |= B200_OCF_OBSERVED | B200_OCF_HASOBS;
            oc_correct(k, z);                                            // the real measurement on top
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_OC_LASTZ + c) * TMAX + t] = virt ? vz[c] : z[c];
            tsu = 0; hits += 1; streak += 1;
            sm.lbox[0][t] = b0; sm.lbox[1][t] = b1; sm.lbox[2][t] = b2; sm.lbox[3][t] = b3;
        } else {                                                        // kf.update(None), ocsort_kf.py:465-477
            if (fl & B200_OCF_OBSERVED) {
#pragma unroll
                for (int c = 0; c < 7; ++c) gf[(B200_OC_SX + c) * TMAX + t] = k.x[c];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    gf[(B200_OC_SP + 3 * i + 0) * TMAX + t] = k.pp[i];
                    gf[(B200_OC_SP + 3 * i + 1) * TMAX + t] = k.pv[i];
                    gf[(B200_OC_SP + 3 * i + 2) * TMAX + t] = k.vv[i];
                }
                gf[(B200_OC_SP + 9) * TMAX + t] = k.prr;
                fl |= B200_OCF_SAVED;
            }
            fl &= ~B200_OCF_OBSERVED;
        }
    }

    PHASE(7);
    // ---- new trackers (ocsort.py:351-353) and the reversed output scan (:354-379) -------------------
    // creation order: associate()'s unmatched list order, or ascending when the recovery round ran setdiff1d
    const int ds = tid < DMAX ? sm.dstate[tid] : DS_NONE;
    const bool newborn = ds == DS_FREE0 || ds == DS_FREE1;
    bool emit_old = false, emit_new = false, die = false;
    bool box_from_obs = false;
    if (live) {
        emit_old = tsu < 1 && (streak >= p.min_hits || frame <= p.min_hits);
        die = tsu > p.max_age;
        const double lsum = (fl & B200_OCF_HASOBS) ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
        box_from_obs = !(lsum < 0.0);
    }
    if (newborn) emit_new = (0 >= p.min_hits || frame <= p.min_hits);
    unsigned long long val = (emit_old ? 1ull : 0ull) | (emit_new ? (1ull << 10) : 0ull) | (ds == DS_FREE0 ? (1ull << 20) : 0ull) |
                             (ds == DS_FREE1 ? (1ull << 30) : 0ull) | ((live && !die) ? (1ull << 40) : 0ull);
    unsigned long long tot;
    const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);
    const int E_old = (int)(tot & 1023), E_new = (int)((tot >> 10) & 1023);
    const int n_free0 = (int)((tot >> 20) & 1023), n_free1 = (int)((tot >> 30) & 1023), n_keep = (int)((tot >> 40) & 1023);
    const int n_new = n_free0 + n_free1;
    double* gout = packed ? nullptr : p.out + (size_t)s * p.max_tracks * 8;
    const int out_cap = packed ? min(p.max_tracks, nd_in) : p.max_tracks;
    if (n0 + n_new > tcap) err |= B200_ERR_TRACK_OVERFLOW;
    // result row: the reference's [x1, y1, x2, y2, id, conf, cls, det_ind], or - packed frames - just (id, det_ind): an
    // OC-SORT row's box / conf / cls are the caller's own detection row det_ind (ocsort.py:356-363: last_observation)
    auto write_row = [&](int row, const Box& b, int id, double cf, double cl, int di) {
        if (packed) { reinterpret_cast<int2*>(p.rows)[(size_t)roff + row] = make_int2(id, di); return; }
        double2* o = reinterpret_cast<double2*>(gout + (size_t)row * 8);          // 64-byte rows: four 16-byte stores
        o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
        o[2] = make_double2((double)id, cf); o[3] = make_double2(cl, (double)di);
    };

    if (live) {
        // write the slot back in place
#pragma unroll
        for (int c = 0; c < 7; ++c) gf[(B200_OC_X + c) * TMAX + t] = k.x[c];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gf[(B200_OC_P + 3 * i + 0) * TMAX + t] = k.pp[i];
            gf[(B200_OC_P + 3 * i + 1) * TMAX + t] = k.pv[i];
            gf[(B200_OC_P + 3 * i + 2) * TMAX + t] = k.vv[i];
        }
        gf[(B200_OC_P + 9) * TMAX + t] = k.prr;
        gf[B200_OC_CONF * TMAX + t] = conf;
        gf[B200_OC_CLS * TMAX + t] = cls;
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
                else b = oc_x_to_box(k.x[0], k.x[1], k.x[2], k.x[3]);
                int di = det_ind;
                if (packed && !box_from_obs) {            // the filter's box travels in the exception area
                    const int e = atomicAdd(p.err_out + 1, 1);
                    if (e < p.exc_cap) {
                        unsigned char* x = p.exc + (size_t)e * B200_EXC_OC_BYTES;
                        reinterpret_cast<int2*>(x)[0] = make_int2(roff + row, 0);
                        double* xb = reinterpret_cast<double*>(x + 8);
                        xb[0] = b.x1; xb[1] = b.y1; xb[2] = b.x2; xb[3] = b.y2;
                    } else err |= B200_ERR_PACKED_ROW;
                    di |= B200_ROW_OC_STATE;
                }
                write_row(row, b, tid_id + 1, conf, cls, di);
            }
        }
    } else if (t < n0 && (fl & OCF_ALIVE) == 0 && t < TMAX) {
        // NaN-purged this frame: make the hole permanent
        if (gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE) gi[B200_OCI_FLAGS * TMAX + t] = fl;
    }
    if (newborn) {
        const int j = tid;
        int order;                      // position in the creation order
        if (ocr_ran) order = (int)((ex >> 20) & 1023) + (int)((ex >> 30) & 1023);
        else order = ds == DS_FREE0 ? (int)((ex >> 20) & 1023) : n_free0 + (int)((ex >> 30) & 1023);
        const int dst = n0 + order;
        const int id = id0 + order;
        double z[4];
        oc_box_to_z(sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j], z);
        if (dst < tcap) {
            const double P0[10] = {10.0, 0.0, 1e4, 10.0, 0.0, 1e4, 10.0, 0.0, 1e4, 10.0};
#pragma unroll
            for (int c = 0; c < 4; ++c) { gf[(B200_OC_X + c) * TMAX + dst] = z[c]; gf[(B200_OC_LAST + c) * TMAX + dst] = -1.0; }
#pragma unroll
            for (int c = 4; c < 7; ++c) gf[(B200_OC_X + c) * TMAX + dst] = 0.0;
#pragma unroll
            for (int c = 0; c < 10; ++c) gf[(B200_OC_P + c) * TMAX + dst] = P0[c];
            gf[B200_OC_CONF * TMAX + dst] = sm.dconf[j];
            gf[B200_OC_CLS * TMAX + dst] = det_cls(j);
            gf[(B200_OC_VEL + 0) * TMAX + dst] = 0.0; gf[(B200_OC_VEL + 1) * TMAX + dst] = 0.0;
            gi[B200_OCI_ID * TMAX + dst] = id;
            gi[B200_OCI_AGE * TMAX + dst] = 0;
            gi[B200_OCI_TSU * TMAX + dst] = 0;
            gi[B200_OCI_HITS * TMAX + dst] = 0;
            gi[B200_OCI_STREAK * TMAX + dst] = 0;
            gi[B200_OCI_DET * TMAX + dst] = j;
            gi[(B200_OCI_RINGAGE + 0) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 1) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 2) * TMAX + dst] = -1;
            gi[B200_OCI_FLAGS * TMAX + dst] = OCF_ALIVE;
        }
        if (emit_new) {
            // reversed list order: the newest tracker first; rows of the new trackers precede the old ones
            const int row = n_new - 1 - order;
            // packed: bit 30 marks a new tracker - its box is the detection's round trip through the filter state
            if (row < out_cap) write_row(row, oc_x_to_box(z[0], z[1], z[2], z[3]), id + 1, sm.dconf[j], det_cls(j), packed ? (j | B200_ROW_OC_NEW) : j);
        }
    }
    const int n1 = min(n0 + n_new, tcap);
    const int alive_after = n_keep + min(n_new, tcap - n0 > 0 ? tcap - n0 : 0);
    __syncthreads();
    PHASE(8);

    const int n_final = n1;
    if (tid == 0) {
        counts[0] = n_final;
        counts[1] = alive_after;
        counts[2] = id0 + n_new;
        counts[3] = frame;
        p.nout[s] = min(E_old + E_new, out_cap);
        p.track_updates[s] += (unsigned long long)Cn;
        if (p.dbg) atomicAdd(&p.dbg[0], 1ull);
    }
    if (err) { atomicOr(p.err, err); if (p.err_out) atomicOr(p.err_out, err); }
}
#undef PHASE

template <int TMAX, int DMAX, bool DENSE>
cudaError_t launch_oc_kernel(const StepParams& p, cudaStream_t stream) {
    auto kern = ocsort_step_kernel<TMAX, TMAX, DMAX, DENSE>;
    const size_t smem = sizeof(OcSmem<TMAX, DMAX>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.n_streams, TMAX, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int TMAX, int DMAX>
cudaError_t launch_oc_variant(const StepParams& p, cudaStream_t stream) {
    if constexpr (TMAX >= 224 && TMAX < 512) {
        static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
        static const int force = [] { const char* v = getenv("B200_OC_DENSE"); return v ? atoi(v) : -1; }();     // A/B experiments
        if (force == 1 || (force < 0 && p.n_streams > 2 * sms)) return launch_oc_kernel<TMAX, DMAX, true>(p, stream);
    }
    return launch_oc_kernel<TMAX, DMAX, false>(p, stream);
}

}  // namespace

size_t ocsort_step_smem(int variant) {
    switch (variant) {
        case 0: return sizeof(OcSmem<64, 64>);
        case 1: return sizeof(OcSmem<128, 128>);
        case 2: return sizeof(OcSmem<224, 224>);
        case 3: return sizeof(OcSmem<256, 256>);
        case 4: return sizeof(OcSmem<512, 512>);
    }
    return 0;
}

cudaError_t launch_ocsort_step(const StepParams& p, int variant, cudaStream_t stream) {
    switch (variant) {
        case 0: return launch_oc_variant<64, 64>(p, stream);
        case 1: return launch_oc_variant<128, 128>(p, stream);
        case 2: return launch_oc_variant<224, 224>(p, stream);
        case 3: return launch_oc_variant<256, 256>(p, stream);
        case 4: return launch_oc_variant<512, 512>(p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
