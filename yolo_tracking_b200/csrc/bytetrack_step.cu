// ByteTrack frame step for many independent streams: ONE kernel launch per frame, one CTA per
// stream, one thread per track slot.
//
// Data flow per stream (HBM is touched exactly once in each direction):
//   * thread j fetches detection row j whole, thread t the mean and 3 lifecycle ints of slot t; boxes and the position
//     half of the mean are staged in shared memory; velocities and the 12 covariance terms of a slot are NOT needed for
//     association (ByteTrack costs only use boxes), so each thread loads them at the single deferred predict + update
//     point and parks the updated covariance in shared memory (storage of the dead solver arrays) until the final write;
//   * candidate pairs come from per-frame cell masks (32 x-cells, 16 y-cells; per axis two monotone bitmask families
//     "starts at or before cell c" / "ends at or after cell c", built by warp bit-transposes): a track tests only the
//     detections whose cell range meets its own in both axes - nothing T x D is ever computed, let alone written to HBM;
//   * the assignment is solved on the pruned graph in shared memory (lap_sparse.cuh: greedy tight start, concurrent
//     shortest augmenting paths for the few contested rows); ByteTrack collects the second association's candidates
//     during the first build;
//   * the state is written back already in the reference's new list order (tracked list, then lost list), together
//     with the output rows.
// The step is latency bound (a CTA alone on an SM needs 35 k cycles, four together 53 k): every phase is organised for
// short dependent chains and few barriers - see DESIGN.md 4.1 / 6.
//
// Replaces BYTETracker.update (boxmot/trackers/bytetrack/byte_tracker.py:132-281) and what it
// calls: STrack.multi_predict :35-48 -> KalmanFilter.multi_predict (bytetrack_kf.py:155-192),
// iou_distance / fuse_score / linear_assignment (matching.py:94-119, :213-221, :56-71 ->
// lap.lapjv), STrack.update / re_activate / activate :50-98 -> KalmanFilter.update / initiate,
// joint_stracks / sub_stracks / remove_duplicate_stracks :287-325.
//
// The same kernel, instantiated with BOT = true and the XYWH filter, is the BoT-SORT frame step
// (boxmot/trackers/botsort/bot_sort.py:231-420, camera-motion warp = identity): the association
// costs become min(iou distance, appearance distance) with the appearance term gated by the
// proximity mask (:298-309, :356-368), matched tracks blend the detection embedding into their
// smoothed embedding in fp32 (STrack.update_features :40-48) and vote on their class
// (update_cls :50-67).  The appearance distance is only ever needed for pairs that pass the
// proximity mask (iou distance <= proximity_thresh) - one or two per track - so it is
// evaluated per candidate pair, one warp per pair, in double on the fp32 values exactly like
// matching.py:145-167; nothing T x D x F is computed.  Embeddings live in a per-stream pool of
// rows that never move (a slot carries its row index); rows of dead tracks are recycled.
// This file is compiled twice: as is for the padded interface (b200track_step), and through bytetrack_step_packed.cu with
// B200_STEP_PACKED = 1 for the packed frame interface (b200track_step_packed) - the input / output addressing is a
// compile-time choice, so neither instantiation carries the other's registers.
#include <cstdlib>

#ifndef B200_STEP_PACKED
#define B200_STEP_PACKED 0
#endif

#include "boxes.cuh"
#include "kf.cuh"
#include "kf44.cuh"
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

// Cells per axis of the candidate masks.  A detection covers the cell range [c0, c1] of an axis; instead of one mask
// per cell, two monotone families are kept per axis: lo[c] = detections with c0 <= c, hi[c] = detections with
// c1 >= c.  A track covering [t0, t1] then meets exactly lo[t1] & hi[t0]: two loads per axis whatever the range.
// NCX = 32 = one warp transpose per family; the two 16-cell y families share one transpose.
constexpr int NCX = 32, NCY = 16;

// row types of one association pass: which detection set / limit / cost a row uses
constexpr int RT_NONE = 0, RT_A = 1, RT_B = 2;

// BoT-SORT extras (empty for ByteTrack)
template <int TMAX, int DMAX, bool BOT>
struct BotSmem {};
template <int TMAX, int DMAX>
struct BotSmem<TMAX, DMAX, true> {
    static constexpr int PCAP = 4 * TMAX;
    double pcost[PCAP];              // iou-side cost of a pair waiting for its appearance distance
    uint32_t epairs[PCAP];           // (row << 16) | det
    float dn0[DMAX], dn1[DMAX];      // the two norms a detection's raw embedding row is divided by (curr_feat = (row / dn0) / dn1)
    float dn2[DMAX];                 // norm of a detection's curr_feat (the third in-place normalisation divides by it)
    int nepairs[4];
    short frow[TMAX];                // embedding-pool row of a slot
    short emadet[TMAX];              // detection whose embedding is blended into the slot's, -1 = none
    short freelist[TMAX];
    short nbdet[DMAX], nbrow[DMAX];  // stored newborn k: its detection and its pool row
    unsigned char rowused[TMAX];
};

template <int TMAX, int DMAX, bool BOT = false, bool CAM = false>
struct alignas(16) StepSmem {
    static constexpr int TCAP = TMAX;
    static constexpr int DW = DMAX / 32;
    static constexpr int DWP = (DW + 3) / 4 * 4;      // mask rows padded to 16-byte multiples
    static constexpr int ECAP = 2 * TMAX;             // candidate edge cache (cost computed once, while the graph is built)
    static constexpr int PCAP = 4 * TMAX;             // candidate pairs that passed the conservative fp32 overlap filter
    // second-pass candidates found while the first pass is built ((row << 16) | det and cost); they live in the storage
    // of adj, which is only needed when the edge cache overflows (then the second pass is rebuilt the long way)
    static constexpr int P2CAP = (DW * TMAX * 4) / 12 < TMAX ? (DW * TMAX * 4) / 12 : TMAX;
    double mean[4][TMAX];                             // position half of the mean; velocities stay in HBM / registers
    double dbox[4][DMAX];                             // raw x1, y1, x2, y2
    double dconf[DMAX];
    double u[TMAX];
    unsigned long long scratch[32];
    // Everything the association needs and nothing after it does shares its storage with the parking area of the
    // updated covariances: between the Kalman update and the final write (duplicate removal, scans) the 12 covariance
    // terms of a slot wait here instead of in registers (at 72 registers per thread they would spill, and with
    // 4 x 56 KB of shared memory per SM the L1 is too small to hold the spills).
    union {
        struct {
            double v[DMAX], dist[DMAX];
            double ecost[ECAP];
            int parent[TMAX + DMAX], head[TMAX];
            int ehead[TMAX];
            uint32_t adj[DW][TMAX];
            uint32_t xlo[NCX][DWP], xhi[NCX][DWP], ylo[NCY][DWP], yhi[NCY][DWP];
            uint32_t pairs[PCAP];                     // (row << 16) | det
            short ecol[ECAP], enext[ECAP];
            short rnext[TMAX];
            short yc[DMAX], pred[DMAX], nextc[DMAX], mark[DMAX], scn[DMAX];
        };
        double park[CAM ? 20 : 12][TMAX];       // CAM: the two 4x4 covariance blocks of a camera-corrected filter
    };
    int frame_t[TMAX], start_t[TMAX];
    int coldeg[DMAX], ncomplex[4];
    uint32_t colbitsA[DWP], colbitsB[DWP];
    float fext[16][4];
    short xr[TMAX], match[TMAX], lostlist[TMAX];
    int ecount[4];
    int npairs[4];
    int np2[4];                                       // entries of the second-pass candidate list (see P2CAP)
    float4 dboxf[DMAX];                               // detection boxes rounded outwards to fp32
    unsigned char role[TMAX], rowtype[TMAX], cat[TMAX], drop[TMAX + DMAX], dflag[DMAX];
    BotSmem<TMAX, DMAX, BOT> bot;
    __device__ __forceinline__ double* p2cost() { return reinterpret_cast<double*>(&adj[0][0]); }
    __device__ __forceinline__ uint32_t* p2pair() { return reinterpret_cast<uint32_t*>(&adj[0][0]) + 2 * P2CAP; }
};

// STrack.xyxy (byte_tracker.py:100-111): XYAH mean -> (xc, yc, a*h, h) -> corners
template <int KIND>
__device__ __forceinline__ Box mean_to_box(double xc, double yc, double a_or_w, double h) {
    const double w = (KIND == KF_XYWH) ? a_or_w : xmul(a_or_w, h);
    return xywh_to_xyxy(xc, yc, w, h);
}

template <int KIND, class SM>
__device__ __forceinline__ Box track_box(const SM& sm, int t) {
    return mean_to_box<KIND>(sm.mean[0][t], sm.mean[1][t], sm.mean[2][t], sm.mean[3][t]);
}

// The box iou_distance sees for a detection is the round trip xyxy -> xywh -> xyxy
// (STrack.__init__ + STrack.xyxy with mean None, byte_tracker.py:16, :105-110).
template <class SM>
__device__ __forceinline__ Box det_box(const SM& sm, int j) {
    double xc, yc, w, h;
    xyxy_to_xywh(sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j], xc, yc, w, h);
    return xywh_to_xyxy(xc, yc, w, h);
}

// measurement fed to the filter for detection j (STrack.__init__, byte_tracker.py:16-18)
template <int KIND, class SM>
__device__ __forceinline__ void det_measurement(const SM& sm, int j, double* z) {
    double xc, yc, w, h;
    xyxy_to_xywh(sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j], xc, yc, w, h);
    if (KIND == KF_XYWH) { z[0] = xc; z[1] = yc; z[2] = w; z[3] = h; }
    else xywh_to_xyah(xc, yc, w, h, z);
}

// cost of (track row, detection): iou_distance, optionally fuse_score, chosen by the row type
template <int KIND, class SM>
struct PassCost {
    const SM* sm;
    bool fuseA, fuseB;
    bool embA = false, embB = false;      // BoT-SORT: min(iou cost, gated appearance distance) for this row type
    __device__ __forceinline__ double pair(const Box& a, int j, bool fuse) const {
        const double v = box_iou(a, det_box(*sm, j));
        return fuse ? fused_cost(v, sm->dconf[j]) : xsub(1.0, v);
    }
    __device__ __forceinline__ double operator()(int t, int j) const {
        return pair(track_box<KIND>(*sm, t), j, sm->rowtype[t] == RT_A ? fuseA : fuseB);
    }
};
struct PassLimit {
    const unsigned char* rowtype;
    double limA, limB;
    __device__ __forceinline__ double operator()(int t) const { return rowtype[t] == RT_A ? limA : limB; }
};

// 32 x 32 bit-matrix transpose across the lanes of a warp: lane r holds row r, afterwards lane c holds column c
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
        m ^= m << (j >> 1);
    }
    return x;
}

struct CellMap {
    float x0, y0, sx, sy;
    __device__ __forceinline__ int cx(double x) const {
        return min(max((int)(((float)x - x0) * sx), 0), NCX - 1);      // monotone in x
    }
    __device__ __forceinline__ int cy(double y) const {
        return min(max((int)(((float)y - y0) * sy), 0), NCY - 1);
    }
};

// Candidate graph of one association pass.
//   phase A (thread t = row t): AND the ORs of the cell masks the row's box covers (a superset
//     of the overlapping detections because the cell maps are monotone), run a conservative
//     fp32 overlap test (boxes rounded outwards) on the survivors and append the pairs that
//     pass to a shared list;
//   phase B (thread k = pair k): exact cost; edge iff cost <= limit - exact pruning, see
//     lap_sparse.cuh (no overlap => iou = 0 => cost = 1 > limit).  Every thread evaluates at
//     most a couple of pairs, all lanes busy, one fp64 division each.
template <int KIND, class SM>
__device__ __forceinline__ void graph_phase_a(SM& sm, int t, int n, int words, const CellMap& cm, bool fused = false) {
    constexpr int DW = SM::DW, DWP = SM::DWP;
    const int lane = threadIdx.x & 31;
    const int rt = t < n ? sm.rowtype[t] : RT_NONE;
    // The step is latency bound (few warps, long dependent chains), so the walk is organised for instruction-level
    // parallelism: every iteration tests ONE candidate of EVERY mask word (DW independent chains) instead of
    // draining the words one after the other; hits are collected as bitmasks and emitted the same way.
    uint32_t c[DW], h[DW];
#pragma unroll
    for (int w = 0; w < DW; ++w) { c[w] = 0u; h[w] = 0u; }
    float ax1 = 0.f, ay1 = 0.f, ax2 = 0.f, ay2 = 0.f;
    if (rt != RT_NONE) {
        const Box a = track_box<KIND>(sm, t);
        ax1 = __double2float_rd(a.x1); ay1 = __double2float_rd(a.y1);
        ax2 = __double2float_ru(a.x2); ay2 = __double2float_ru(a.y2);
        const int cx0 = cm.cx(a.x1), cx1 = cm.cx(a.x2), cy0 = cm.cy(a.y1), cy1 = cm.cy(a.y2);
        // classic: row type A walks colbitsA, B walks colbitsB.  fused first pass (colbitsA = high, colbitsB = low
        // detections): pool and unconfirmed rows walk the high detections, still-tracked rows the low ones as well
        const uint32_t* colbits = (fused || rt == RT_A) ? sm.colbitsA : sm.colbitsB;
        const bool also_low = fused && rt == RT_A && sm.role[t] == ROLE_TRACKED;
#pragma unroll
        for (int q = 0; q < DWP / 4; ++q) {
            const uint4 xa = *reinterpret_cast<const uint4*>(&sm.xlo[cx1][q * 4]);
            const uint4 xb = *reinterpret_cast<const uint4*>(&sm.xhi[cx0][q * 4]);
            const uint4 ya = *reinterpret_cast<const uint4*>(&sm.ylo[cy1][q * 4]);
            const uint4 yb = *reinterpret_cast<const uint4*>(&sm.yhi[cy0][q * 4]);
            uint4 cb = *reinterpret_cast<const uint4*>(&colbits[q * 4]);
            if (also_low) {
                const uint4 cl = *reinterpret_cast<const uint4*>(&sm.colbitsB[q * 4]);
                cb.x |= cl.x; cb.y |= cl.y; cb.z |= cl.z; cb.w |= cl.w;
            }
            if (q * 4 + 0 < DW) c[q * 4 + 0] = xa.x & xb.x & ya.x & yb.x & cb.x;
            if (q * 4 + 1 < DW) c[q * 4 + 1] = xa.y & xb.y & ya.y & yb.y & cb.y;
            if (q * 4 + 2 < DW) c[q * 4 + 2] = xa.z & xb.z & ya.z & yb.z & cb.z;
            if (q * 4 + 3 < DW) c[q * 4 + 3] = xa.w & xb.w & ya.w & yb.w & cb.w;
        }
    }
    while (true) {
        uint32_t any = 0u;
#pragma unroll
        for (int w = 0; w < DW; ++w) any |= c[w];
        if (!any) break;
#pragma unroll
        for (int w = 0; w < DW; ++w) {
            const uint32_t cw = c[w];
            const uint32_t low = cw & (0u - cw);              // lowest candidate of this word (0 if none)
            c[w] = cw ^ low;
            const int j = w * 32 + (low ? __ffs(low) - 1 : 0);
            const float4 d = sm.dboxf[j];
            if (d.x < ax2 && ax1 < d.z && d.y < ay2 && ay1 < d.w) h[w] |= low;
        }
    }
    int mine = 0;
#pragma unroll
    for (int w = 0; w < DW; ++w) mine += __popc(h[w]);
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    int base = 0;
    if (lane == 31 && total) base = atomicAdd(&sm.npairs[0], total);      // one list atomic per warp
    base = __shfl_sync(0xffffffffu, base, 31) + inc - mine;
    int off[DW];
#pragma unroll
    for (int w = 0; w < DW; ++w) { off[w] = base; base += __popc(h[w]); }
    while (true) {
        uint32_t any = 0u;
#pragma unroll
        for (int w = 0; w < DW; ++w) any |= h[w];
        if (!any) break;
#pragma unroll
        for (int w = 0; w < DW; ++w) {
            const uint32_t hw = h[w];
            if (hw) {
                const int j = w * 32 + __ffs(hw) - 1;
                h[w] = hw & (hw - 1);
                if (off[w] < SM::PCAP) sm.pairs[off[w]] = ((uint32_t)t << 16) | (uint32_t)j;
                ++off[w];
            }
        }
    }
}

// The bitmask form of the graph (adj) is only read when the edge cache overflowed; it is built on demand
// (graph_build_adj) instead of on every frame.
template <int KIND, class SM>
__device__ __forceinline__ void graph_add_edge(SM& sm, int t, int j, double c) {
    atomicAdd(&sm.coldeg[j], 1);
    const int e = atomicAdd(&sm.ecount[0], 1);
    if (e < SM::ECAP) {
        sm.ecost[e] = c; sm.ecol[e] = (short)j;
        sm.enext[e] = (short)atomicExch(&sm.ehead[t], e);
    }
}

template <int NT, int KIND, bool BOT, class SM>
__device__ __forceinline__ void graph_phase_b(SM& sm, int n, int words, const PassCost<KIND, SM>& cost, const PassLimit& lim,
                                              double proximity, bool fused = false, double lim2A = 0.0, double lim2B = 0.0) {
    const int np = sm.npairs[0];
    if (np <= SM::PCAP) {
        for (int k = threadIdx.x; k < np; k += NT) {
            const uint32_t pr = sm.pairs[k];
            const int t = pr >> 16, j = pr & 0xffff;
            const bool isA = sm.rowtype[t] == RT_A;
            if (!BOT && fused) {
                // one IoU serves both passes: first-pass edge (pool x high, fused score), or a second-pass candidate
                // (tracked x low, plain IoU distance; unconfirmed x high, fused score) kept for after the first solve
                const double v = box_iou(track_box<KIND>(sm, t), det_box(sm, j));
                const bool low = sm.dflag[j] == DF_LOW;
                if (isA && !low) {
                    const double c = fused_cost(v, sm.dconf[j]);
                    if (c <= lim.limA) graph_add_edge<KIND>(sm, t, j, c);
                } else {
                    const double c = isA ? xsub(1.0, v) : fused_cost(v, sm.dconf[j]);
                    if (c <= (isA ? lim2A : lim2B)) {
                        const int e = atomicAdd(&sm.np2[0], 1);
                        if (e < SM::P2CAP) { sm.p2pair()[e] = pr; sm.p2cost()[e] = c; }
                    }
                }
                continue;
            }
            if constexpr (BOT) {
                // bot_sort.py:298-309 / :356-368: the proximity mask is taken on the plain iou distance
                const double v = box_iou(track_box<KIND>(sm, t), det_box(sm, j));
                const double ci = xsub(1.0, v);
                const double c = (isA ? cost.fuseA : cost.fuseB) ? fused_cost(v, sm.dconf[j]) : ci;
                if ((isA ? cost.embA : cost.embB) && !(ci > proximity)) {
                    const int e = atomicAdd(&sm.bot.nepairs[0], 1);
                    sm.bot.epairs[e] = pr; sm.bot.pcost[e] = c;       // e < np <= PCAP
                } else if (c <= (isA ? lim.limA : lim.limB)) graph_add_edge<KIND>(sm, t, j, c);
            } else {
                const double c = cost.pair(track_box<KIND>(sm, t), j, isA ? cost.fuseA : cost.fuseB);
                if (c <= (isA ? lim.limA : lim.limB)) graph_add_edge<KIND>(sm, t, j, c);
            }
        }
    } else {
        // pair list overflowed (pathologically crowded frame): every row re-walks its columns
        // (fused first pass: only the first pass is built here, the second one is rebuilt the long way)
        if (fused && threadIdx.x == 0) sm.np2[0] = SM::P2CAP + 1;
        for (int t = threadIdx.x; t < n; t += NT) {
            const int rt = sm.rowtype[t];
            if (rt == RT_NONE || (fused && rt != RT_A)) continue;
            const Box a = track_box<KIND>(sm, t);
            const uint32_t* colbits = rt == RT_A ? sm.colbitsA : sm.colbitsB;
            for (int wd = 0; wd < words; ++wd) {
                uint32_t cand = colbits[wd];
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int j = wd * 32 + b;
                    if (!box_overlap(a, det_box(sm, j))) continue;
                    const double c = cost.pair(a, j, rt == RT_A ? cost.fuseA : cost.fuseB);
                    if (c <= (rt == RT_A ? lim.limA : lim.limB)) graph_add_edge<KIND>(sm, t, j, c);
                }
            }
        }
    }
}

// Edge cache overflowed (crowded frame): the solver falls back to the bitmask form of the graph and recomputes costs.
// Thread t owns row t of adj, so no atomics: every row walks its columns once.
template <int NT, int KIND, class SM>
__device__ __forceinline__ void graph_build_adj(SM& sm, int n, int words, const PassCost<KIND, SM>& cost, const PassLimit& lim, bool fused = false) {
    if (fused && threadIdx.x == 0) sm.np2[0] = SM::P2CAP + 1;       // adj overwrites the second-pass candidate list
    for (int t = threadIdx.x; t < n; t += NT) {
        int rt = sm.rowtype[t];
        if (fused && rt != RT_A) rt = RT_NONE;
        const Box a = track_box<KIND>(sm, t);
        const uint32_t* colbits = rt == RT_A ? sm.colbitsA : sm.colbitsB;
        for (int wd = 0; wd < words; ++wd) {
            uint32_t bits = 0u;
            uint32_t cand = rt == RT_NONE ? 0u : colbits[wd];
            while (cand) {
                const int b = __ffs(cand) - 1;
                cand &= cand - 1;
                const int j = wd * 32 + b;
                if (!box_overlap(a, det_box(sm, j))) continue;
                const double c = cost.pair(a, j, rt == RT_A ? cost.fuseA : cost.fuseB);
                if (c <= (rt == RT_A ? lim.limA : lim.limB)) bits |= 1u << b;
            }
            sm.adj[wd][t] = bits;
        }
    }
}

// ---- BoT-SORT embedding arithmetic, one warp per row / pair -------------------------------------
// Rows are feat_dim fp32 values, feat_dim a multiple of 128: lane l owns the float4 groups l, l + 32, ...
// Everything the reference does on embeddings is float32 numpy (bot_sort.py:40-48): norms are
// sqrt(dot(x, x)) rounded to fp32 (accumulated in double here; BLAS order is unspecified), divisions and
// the 0.9 / 0.1 blend are separate correctly rounded fp32 operations.
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}
__device__ __forceinline__ float4 f4_div(float4 a, const RowDiv& d) {
    return make_float4(fdiv_row(a.x, d), fdiv_row(a.y, d), fdiv_row(a.z, d), fdiv_row(a.w, d));
}
__device__ __forceinline__ float4 f4_div(float4 a, float n) { return f4_div(a, row_div(n)); }
__device__ __forceinline__ double f4_sq(float4 a) {
    return (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
}
__device__ __forceinline__ float norm_f32(double sumsq) { return sqrtf((float)sumsq); }

// A detection embedding as the reference holds it when costs are computed: get_features row -> STrack.__init__
// (feat /= |feat|, then the aliased smooth_feat /= |smooth_feat|).  The twice-normalised row (curr_feat) is never
// stored: its two divisors (and its own norm, which update_features of a matched track divides by once more) are kept
// per detection, and pair costs, blends and new tracks redo the two row divisions (three operations each with the shared
// reciprocal of common.cuh - the same bits) on the raw input row instead of writing a [max_dets, feat_dim] scratch
// block per stream and reading it back.
// Rows of up to 512 floats are held in registers (four float4 per lane, every load in flight at once): the three
// normalisation passes, the blend and the distance then run without touching memory again.  Written as loops of single
// loads they were serialised on the memory latency - the embedding phases took 90 % of the BoT-SORT step.
struct Row4 { float4 v[4]; };
__device__ __forceinline__ Row4 load_row4(const float4* row, int nv, int lane) {
    Row4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.v[k] = lane + 32 * k < nv ? row[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
    return r;
}
__device__ __forceinline__ double row4_sq(const Row4& r) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) a += f4_sq(r.v[k]);           // zero padding past nv adds exact zeros
    return a;
}

// curr_feat values of a raw row: (x / n0) / n1
__device__ __forceinline__ float4 f4_curr(float4 a, const RowDiv& d0, const RowDiv& d1) { return f4_div(f4_div(a, d0), d1); }
__device__ __forceinline__ Row4 load_curr4(const float4* row, int nv, int lane, const RowDiv& d0, const RowDiv& d1) {
    Row4 r = load_row4(row, nv, lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) r.v[k] = f4_curr(r.v[k], d0, d1);
    return r;
}

// the three norms of a detection's embedding: of the raw row, of row / n0, of (row / n0) / n1
__device__ __forceinline__ float3 det_curr_feat(const float4* row, int nv, int lane) {
    if (nv <= 128) {
        Row4 r = load_row4(row, nv, lane);
        const float n0 = norm_f32(warp_sum(row4_sq(r)));
        const RowDiv d0 = row_div(n0);
#pragma unroll
        for (int k = 0; k < 4; ++k) r.v[k] = f4_div(r.v[k], d0);
        const float n1 = norm_f32(warp_sum(row4_sq(r)));
        const RowDiv d1 = row_div(n1);
#pragma unroll
        for (int k = 0; k < 4; ++k) r.v[k] = f4_div(r.v[k], d1);
        return make_float3(n0, n1, norm_f32(warp_sum(row4_sq(r))));
    }
    double a = 0.0;
    for (int i = lane; i < nv; i += 32) a += f4_sq(row[i]);
    const float n0 = norm_f32(warp_sum(a));
    const RowDiv d0 = row_div(n0);
    a = 0.0;
    for (int i = lane; i < nv; i += 32) a += f4_sq(f4_div(row[i], d0));
    const float n1 = norm_f32(warp_sum(a));
    const RowDiv d1 = row_div(n1);
    a = 0.0;
    for (int i = lane; i < nv; i += 32) a += f4_sq(f4_curr(row[i], d0, d1));
    return make_float3(n0, n1, norm_f32(warp_sum(a)));
}

// embedding_distance (matching.py:145-167) of one (smoothed track embedding, detection curr_feat) pair:
// scipy cdist 'cosine' in double on the fp32 values, clamped at 0
__device__ __forceinline__ double emb_distance(const float4* trk, const float4* det, const RowDiv& d0, const RowDiv& d1, int nv, int lane) {
    double uv = 0.0, uu = 0.0, vv = 0.0;
    if (nv <= 128) {
        const Row4 ra = load_row4(trk, nv, lane), rb = load_curr4(det, nv, lane, d0, d1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 a = ra.v[k], b = rb.v[k];
            uv += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
            uu += f4_sq(a);
            vv += f4_sq(b);
        }
    } else
    for (int i = lane; i < nv; i += 32) {
        const float4 a = trk[i];
        const float4 b = f4_curr(det[i], d0, d1);
        uv += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
        uu += f4_sq(a);
        vv += f4_sq(b);
    }
    uv = warp_sum(uv); uu = warp_sum(uu); vv = warp_sum(vv);
    double c = uv / (sqrt(uu) * sqrt(vv));
    if (fabs(c) > 1.0) c = copysign(1.0, c);
    return fmax(0.0, 1.0 - c);
}

// appearance stage of the candidate graph (BoT-SORT): one warp per pair that passed the proximity mask
template <int NT, int KIND, class SM>
__device__ __forceinline__ void graph_phase_emb(SM& sm, const StepParams& p, int s, const float* dfeat, const PassLimit& lim) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ne = sm.bot.nepairs[0];
    const int nv = p.feat_dim >> 2;
    for (int k = warp; k < ne; k += NT / 32) {
        const uint32_t pr = sm.bot.epairs[k];
        const int t = pr >> 16, j = pr & 0xffff;
        const float4* trk = reinterpret_cast<const float4*>(p.feat_pool + ((size_t)s * SM::TCAP + sm.bot.frow[t]) * p.feat_dim);
        const float4* det = reinterpret_cast<const float4*>(dfeat + (size_t)j * p.feat_dim);
        double e = xmul(emb_distance(trk, det, row_div(sm.bot.dn0[j]), row_div(sm.bot.dn1[j]), nv, lane), 0.5);
        if (e > p.appearance_thresh) e = 1.0;
        const double c = fmin(sm.bot.pcost[k], e);
        if (lane == 0 && c <= (sm.rowtype[t] == RT_A ? lim.limA : lim.limB)) graph_add_edge<KIND>(sm, t, j, c);
    }
}

// update_cls (bot_sort.py:50-67) on a fixed-capacity history {cls[4], score sum[4], n}; returns the new track class
__device__ __forceinline__ double cls_vote(double* h, double cls, double score, int& err) {
    int n = (int)h[8];
    double best = 0.0, out = cls;
    bool found = false;
    for (int k = 0; k < n; ++k) {
        if (cls == h[k]) { h[4 + k] = xadd(h[4 + k], score); found = true; }
        if (h[4 + k] > best) { best = h[4 + k]; out = h[k]; }
    }
    if (!found) {
        if (n < 4) { h[n] = cls; h[4 + n] = score; h[8] = (double)(n + 1); } else err |= B200_ERR_BOT_CAPACITY;
        out = cls;
    }
    return out;
}

template <int NT, int KIND, int TMAX, int DMAX, bool BOT, bool CAM = false>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : (NT == 256 ? (CAM ? 2 : 3) : (NT == 224 ? (BOT ? (CAM ? 2 : 3) : 4) : (NT == 128 ? (CAM ? 4 : 6) : 8)))))
bytetrack_step_kernel(const StepParams p) {
    static_assert(NT == TMAX && DMAX <= NT, "one thread per track slot; detections fit one pass");
    static_assert(!BOT || KIND == KF_XYWH, "BoT-SORT runs on the XYWH filter");
    static_assert(!CAM || BOT, "camera-motion warps belong to BoT-SORT (bot_sort.py:293-295)");
    using SM = StepSmem<TMAX, DMAX, BOT, CAM>;
    // CAM: the covariance is two 4x4 blocks (layout.h: B200_NF_CAM) and the frame's warp is applied after the motion step
    constexpr int NFk = CAM ? B200_NF_CAM : B200_NF, TF_SCORE = CAM ? B200_TFC_SCORE : B200_TF_SCORE, TF_CLS = CAM ? B200_TFC_CLS : B200_TF_CLS;
    const double* wp = (CAM && p.warps) ? p.warps + 6 * blockIdx.x : nullptr;
    constexpr int NI = BOT ? B200_NI_BOT : B200_NI;
    // ByteTrack: the candidates of the second association are collected while the first one is built (one walk, one IoU
    // per pair for both); BoT-SORT keeps two separate builds (its second pass has its own appearance stage)
    constexpr bool FUSE2 = !BOT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    constexpr int DWP = SM::DWP;
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // optional per-phase cycle counters (thread 0 of every CTA; b200track_phase_cycles)
    long long ph_last = p.dbg ? clock64() : 0;
#define PHASE(k) do { if (p.dbg && tid == 0) { const long long now_ = clock64(); atomicAdd(&p.dbg[k], (unsigned long long)(now_ - ph_last)); ph_last = now_; } } while (0)

    // ---- HBM -> shared memory: detections [nd, 6] (planar), means, lifecycle ints ----------
    // Every global load of this phase is issued before the first value is used, and none of them waits for the
    // stream's counters: all TMAX slots and max_dets detection rows are fetched (one memory latency instead of
    // two dependent ones); what lies beyond n / nd is never looked at.
    const int t = tid;                                   // this thread's track slot
    const double* gf = p.state_f + (size_t)s * NFk * TMAX;
    const int* gi = p.state_i + (size_t)s * NI * TMAX;
    // packed frames: the rows of all streams lie back to back, fp32 or fp64 (step_params.h); the row offset of the stream
    // is one more dependent load, which the stream ahead prefetched to L2
    constexpr bool packed = B200_STEP_PACKED != 0;
    int roff = 0, nd_in = 0;
    if (packed) { roff = p.det_off[s]; nd_in = p.det_off[s + 1] - roff; }
    const double* dets_g = packed ? p.dets + (size_t)roff * 6 : p.dets + (size_t)s * p.max_dets * 6;
    const float* dets32_g = p.dets32 ? p.dets32 + (size_t)roff * 6 : nullptr;
    // this stream's raw detection embeddings (BoT-SORT): row j at dfeat + j * feat_dim
    const float* dfeat = (BOT && p.feats) ? p.feats + (packed ? (size_t)roff : (size_t)s * p.max_dets) * p.feat_dim : nullptr;
    auto det_cls = [&](int j) -> double { return dets32_g ? (double)dets32_g[j * 6 + 5] : dets_g[j * 6 + 5]; };
    // thread j fetches detection row j whole (three 16-byte loads: 48-byte rows, 16-byte aligned) - no transposition pass
    double mv[8];
    double2 dr0 = make_double2(0.0, 0.0), dr1 = dr0, dr2 = dr0;
    if (packed) {
        if (tid < min(min(DMAX, p.max_dets), nd_in)) {
            if (dets32_g) {                              // 24-byte rows, 8-byte aligned; widening is exact
                const float2* row = reinterpret_cast<const float2*>(dets32_g + (size_t)tid * 6);
                const float2 a = row[0], b = row[1], c = row[2];
                dr0 = make_double2((double)a.x, (double)a.y); dr1 = make_double2((double)b.x, (double)b.y); dr2 = make_double2((double)c.x, (double)c.y);
            } else {
                const double2* row = reinterpret_cast<const double2*>(dets_g + (size_t)tid * 6);
                dr0 = row[0]; dr1 = row[1]; dr2 = row[2];
            }
        }
    } else if (tid < min(DMAX, p.max_dets)) {
        const double2* row = reinterpret_cast<const double2*>(dets_g + (size_t)tid * 6);
        dr0 = row[0]; dr1 = row[1]; dr2 = row[2];
    }
    // first-touch lines of a later stream -> L2, so that its CTA pays an L2 hit instead of a DRAM round trip.  ~4 CTAs run
    // on each of 148 SMs and CTAs start in index order, so stream s + 148 starts about a quarter of a CTA lifetime from
    // now - long enough for the fetch, short enough not to crowd the L2 (measured: 74 / 148 / 296 / 592 / 1184 streams
    // ahead -> 0.2010 / 0.2009 / 0.2022 / 0.2071 / 0.2094 ms per step)
    {
        constexpr int AHEAD = 148;
        const int s2 = s + AHEAD;
        if (s2 < p.n_streams) {
            const char* d2 = reinterpret_cast<const char*>(p.dets + (size_t)s2 * p.max_dets * 6);
            int nl_d = (p.max_dets * 48 + 127) >> 7;
            if (packed) {
                const int o2 = p.det_off[s2], n2 = p.det_off[s2 + 1] - o2, rb = p.dets32 ? 24 : 48;
                d2 = p.dets32 ? reinterpret_cast<const char*>(p.dets32 + (size_t)o2 * 6) : reinterpret_cast<const char*>(p.dets + (size_t)o2 * 6);
                nl_d = (n2 * rb + 127) >> 7;
                if (tid == 0 && s2 + AHEAD < p.n_streams) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.det_off + s2 + AHEAD));
            }
            const char* f2 = reinterpret_cast<const char*>(p.state_f + (size_t)s2 * NFk * TMAX);
            const char* i2 = reinterpret_cast<const char*>(p.state_i + (size_t)s2 * NI * TMAX);
            const int nl_f = (8 * TMAX * 8 + 127) >> 7, nl_i = (NI * TMAX * 4 + 127) >> 7;   // B200_TF_MEAN == 0
            for (int l = tid; l < nl_d + nl_f + nl_i; l += NT) {
                const char* a = l < nl_d ? d2 + ((size_t)l << 7) : (l < nl_d + nl_f ? f2 + ((size_t)(l - nl_d) << 7) : i2 + ((size_t)(l - nl_d - nl_f) << 7));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) mv[c] = gf[(B200_TF_MEAN + c) * TMAX + t];
    int fl = gi[B200_TI_FLAGS * TMAX + t];
    int frame_t = gi[B200_TI_FRAME * TMAX + t];
    const int start_v = gi[B200_TI_START * TMAX + t];
    int frow_v = 0;
    if constexpr (BOT) frow_v = gi[B200_TI_FROW * TMAX + t];
    int* counts = p.counts + 4 * s;
    const int nT = counts[0], nL = counts[1], id0 = counts[2], frame = counts[3] + 1;
    int nd = packed ? nd_in : p.ndets[s];
    LapWork lw;
    lw.Tmax = TMAX; lw.Dmax = DMAX; lw.adj = &sm.adj[0][0]; lw.u = sm.u; lw.v = sm.v; lw.dist = sm.dist;
    lw.parent = sm.parent; lw.head = sm.head; lw.rnext = sm.rnext; lw.xr = sm.xr; lw.yc = sm.yc;
    lw.pred = sm.pred; lw.nextc = sm.nextc; lw.mark = sm.mark; lw.scn = sm.scn;
    lw.coldeg = sm.coldeg; lw.ncomplex = sm.ncomplex;
    lw.ecost = sm.ecost; lw.ecol = sm.ecol; lw.enext = sm.enext; lw.ehead = sm.ehead; lw.ecount = sm.ecount; lw.ecap = SM::ECAP;
    lw.dbg = p.dbg; lw.dbg_last = &ph_last; lw.dbg_slot = 13;
    if (tid == 0) { sm.npairs[0] = 0; sm.np2[0] = 0; }
    lap_prepare<NT>(lw, TMAX, SM::DW);                   // in the shadow of the loads above (does not wait for n / nd)
    const int n = nT + nL;
    int err = 0;
    if (nd > min(DMAX, p.max_dets)) { nd = min(DMAX, p.max_dets); err |= B200_ERR_DET_OVERFLOW; }
    if (nd < 0) nd = 0;
    const int words = (nd + 31) >> 5;
    double vel[4] = {0.0, 0.0, 0.0, 0.0};
    if (t >= n) { fl = 0; frame_t = 0; }
    {
        if constexpr (BOT) { if (t < n) { sm.bot.frow[t] = (short)frow_v; sm.bot.emadet[t] = -1; } }
        if (tid < nd) {
            sm.dbox[0][tid] = dr0.x; sm.dbox[1][tid] = dr0.y; sm.dbox[2][tid] = dr1.x; sm.dbox[3][tid] = dr1.y;
            sm.dconf[tid] = dr2.x;
            // the class column is only read for matched / new tracks, straight from the detection row (L2)
        }
        if (t < n) {
#pragma unroll
            for (int c = 0; c < 4; ++c) vel[c] = mv[4 + c];
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.mean[c][t] = mv[c];
            sm.start_t[t] = start_v;
        }
        // the covariance / id / score lines are first used after the association: pull them into L2 now
        if (t < n && (t & 15) == 0) {
#pragma unroll
            for (int c = B200_TF_COV; c < NFk; ++c)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gf + c * TMAX + t));
        }
        if (t < n && (t & 31) == 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gi + B200_TI_ID * TMAX + t));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gi + B200_TI_LEN * TMAX + t));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gi + B200_TI_DET * TMAX + t));
        }
        if (tid < DWP && tid >= words) { sm.colbitsA[tid] = 0u; sm.colbitsB[tid] = 0u; }    // words beyond the detections stay empty
        for (int i = tid; i < TMAX + DMAX; i += NT) sm.drop[i] = 0;
        if constexpr (BOT) sm.bot.rowused[t] = 0;
    }
    PHASE(1);            // no barrier: up to the frame extents every thread only touches its own detection / slot

    // ---- detection side (thread j): confidence bands (byte_tracker.py:151-158, strict
    // inequalities), frame extents for the cell maps -------------------------------------------
    const float FBIG = 3.0e38f;
    float ex0 = FBIG, ey0 = FBIG, ex1 = -FBIG, ey1 = -FBIG;
    Box mybox = {0, 0, 0, 0};
    int mydfl = 0;
    if (tid < words * 32) {
        const int j = tid;
        if (j < nd) {
            mybox = det_box(sm, j);
            const double c = sm.dconf[j];
            if (c > p.track_thresh) mydfl = DF_HIGH;
            else if (c > p.low_thresh && c < p.track_thresh) mydfl = DF_LOW;
            if (mydfl) { ex0 = (float)mybox.x1; ey0 = (float)mybox.y1; ex1 = (float)mybox.x2; ey1 = (float)mybox.y2; }
            sm.dboxf[j] = make_float4(__double2float_rd(mybox.x1), __double2float_rd(mybox.y1),
                                      __double2float_ru(mybox.x2), __double2float_ru(mybox.y2));
        }
        sm.dflag[j] = (unsigned char)mydfl;
        const uint32_t mh = __ballot_sync(0xffffffffu, mydfl == DF_HIGH);
        const uint32_t ml = __ballot_sync(0xffffffffu, mydfl == DF_LOW);
        if (lane == 0) { sm.colbitsA[j >> 5] = mh; sm.colbitsB[j >> 5] = ml; }   // first association: all high detections (+ low for the fused build)
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        ex0 = fminf(ex0, __shfl_xor_sync(0xffffffffu, ex0, d)); ey0 = fminf(ey0, __shfl_xor_sync(0xffffffffu, ey0, d));
        ex1 = fmaxf(ex1, __shfl_xor_sync(0xffffffffu, ex1, d)); ey1 = fmaxf(ey1, __shfl_xor_sync(0xffffffffu, ey1, d));
    }
    if (lane == 0) { sm.fext[warp][0] = ex0; sm.fext[warp][1] = ey0; sm.fext[warp][2] = ex1; sm.fext[warp][3] = ey1; }

    // ---- track side (thread t): role; motion step of the mean for the pool (unconfirmed
    // tracks are NOT predicted, byte_tracker.py:178-180).  The covariance half of the predict
    // happens later in registers; it needs the pre-motion w / h, kept in ref_w / ref_h. -------
    int role = ROLE_UNCONF;
    double ref_w = 0.0, ref_h = 0.0;
    if (t < n) {
        role = t >= nT ? ROLE_LOST : ((fl & B200_FLAG_ACTIVATED) ? ROLE_TRACKED : ROLE_UNCONF);
        sm.role[t] = (unsigned char)role;
        if (role != ROLE_UNCONF) {
            ref_w = sm.mean[2][t]; ref_h = sm.mean[3][t];
            if ((fl & 3) != B200_ST_TRACKED) {          // multi_predict: zero the height (w, h) velocity
                vel[3] = 0.0;
                if (KIND == KF_XYWH) vel[2] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) sm.mean[i][t] = xadd(sm.mean[i][t], vel[i]);
        }
        if constexpr (CAM) {
            // STrack.multi_gmc (bot_sort.py:95-111) on the pool AND the unconfirmed tracks: the position half here (the
            // association sees the corrected boxes), velocities and covariance in the Kalman phase with the same operations
            if (wp) {
                const double M[4] = {wp[0], wp[1], wp[3], wp[4]}, tv[2] = {wp[2], wp[5]};
                double x = sm.mean[0][t], y = sm.mean[1][t], w = sm.mean[2][t], h = sm.mean[3][t];
                g4_warp_pair(M, tv, x, y);
                g4_warp_pair(M, nullptr, w, h);
                sm.mean[0][t] = x; sm.mean[1][t] = y; sm.mean[2][t] = w; sm.mean[3][t] = h;
            }
        }
        sm.rowtype[t] = role != ROLE_UNCONF ? RT_A : (FUSE2 ? RT_B : RT_NONE);
        sm.match[t] = -1;
    }

    __syncthreads();
    PHASE(2);

    // ---- cell masks over the banded detections ---------------------------------------------
    CellMap cm;
    {
        float x0 = FBIG, y0 = FBIG, x1 = -FBIG, y1 = -FBIG;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) {
            x0 = fminf(x0, sm.fext[k][0]); y0 = fminf(y0, sm.fext[k][1]);
            x1 = fmaxf(x1, sm.fext[k][2]); y1 = fmaxf(y1, sm.fext[k][3]);
        }
        cm.x0 = x0; cm.y0 = y0;
        cm.sx = (x1 > x0) ? (float)NCX / (x1 - x0) : 0.f;
        cm.sy = (y1 > y0) ? (float)NCY / (y1 - y0) : 0.f;
    }
    if (tid < words * 32) {                              // warp-uniform; lane = detection of word `warp`
        uint32_t tlo = 0u, thi = 0u, ty = 0u;             // thermometer codes over the cells, empty for unbanded detections
        if (mydfl) {
            const int cx0 = cm.cx(mybox.x1), cx1 = cm.cx(mybox.x2), cy0 = cm.cy(mybox.y1), cy1 = cm.cy(mybox.y2);
            tlo = 0xffffffffu << cx0;                     // bit c: cx0 <= c
            thi = 0xffffffffu >> (31 - cx1);              // bit c: cx1 >= c
            ty = ((0xffffu << cy0) & 0xffffu) | ((0xffffu >> (15 - cy1)) << 16);
        }
        tlo = warp_transpose32(tlo, lane); thi = warp_transpose32(thi, lane); ty = warp_transpose32(ty, lane);
        sm.xlo[lane][warp] = tlo; sm.xhi[lane][warp] = thi;
        if (lane < 16) sm.ylo[lane][warp] = ty; else sm.yhi[lane - 16][warp] = ty;
    }
    __syncthreads();
    PHASE(3);
    if constexpr (BOT) {
        // norms of the first-round detection embeddings (features only exist for dets_first, bot_sort.py:266)
        if (p.with_reid) {
            const int nv = p.feat_dim >> 2;
            for (int j = warp; j < nd; j += NT / 32) {
                if (sm.dflag[j] != DF_HIGH) continue;
                const float3 nn = det_curr_feat(reinterpret_cast<const float4*>(dfeat + (size_t)j * p.feat_dim), nv, lane);
                if (lane == 0) { sm.bot.dn0[j] = nn.x; sm.bot.dn1[j] = nn.y; sm.bot.dn2[j] = nn.z; }
            }
        }
        if (tid == 0) sm.bot.nepairs[0] = 0;
    }

    PassCost<KIND, SM> cost;
    cost.sm = &sm;
    PassLimit lim;
    lim.rowtype = sm.rowtype;

    // ---- first association: pool x high detections, fused score, limit match_thresh ----
    cost.fuseA = cost.fuseB = BOT ? (p.fuse_first != 0) : true;          // BoT-SORT: only with fuse_first_associate (bot_sort.py:300-301)
    cost.embA = cost.embB = BOT && p.with_reid;
    lim.limA = p.match_thresh; lim.limB = p.match_thresh;
    graph_phase_a<KIND>(sm, t, n, words, cm, FUSE2);
    __syncthreads();
    PHASE(12);
    graph_phase_b<NT, KIND, BOT>(sm, n, words, cost, lim, p.proximity_thresh, FUSE2, p.second_thresh, p.unconf_thresh);
    __syncthreads();
    if constexpr (BOT) {
        if (sm.npairs[0] > SM::PCAP || sm.ecount[0] > SM::ECAP) err |= B200_ERR_BOT_CAPACITY;
        graph_phase_emb<NT, KIND>(sm, p, s, dfeat, lim);
        __syncthreads();
    }
    if (sm.ecount[0] > SM::ECAP) {                        // uniform
        graph_build_adj<NT, KIND>(sm, n, words, cost, lim, FUSE2);
        __syncthreads();
    }
    PHASE(4);
    lap_sparse_solve<NT>(lw, n, words, lim, cost);
    bool matched1 = false;
    if (t < n && sm.rowtype[t] == RT_A) {
        const int j = sm.xr[t];
        if (j >= 0) { matched1 = true; sm.match[t] = (short)j; sm.dflag[j] |= DF_USED; }
    }
    __syncthreads();
    PHASE(5);

    // ---- second pass: two independent problems solved together (disjoint rows AND columns):
    //   A: still-Tracked leftovers x low detections, plain IoU, limit 0.5   (byte_tracker.py:198-226)
    //   B: unconfirmed x remaining high detections, fused score, limit 0.7  (byte_tracker.py:228-240)
    if (t < n) sm.rowtype[t] = (role == ROLE_TRACKED && !matched1) ? RT_A : (role == ROLE_UNCONF ? RT_B : RT_NONE);
    if (tid < words * 32) {
        const int dfl = tid < nd ? sm.dflag[tid] : 0;
        const uint32_t ma = __ballot_sync(0xffffffffu, dfl == DF_LOW);
        const uint32_t mb = __ballot_sync(0xffffffffu, dfl == DF_HIGH);     // high and not used
        if (lane == 0) { sm.colbitsA[tid >> 5] = ma; sm.colbitsB[tid >> 5] = mb; }
    }
    if (tid == 0) { sm.npairs[0] = 0; if constexpr (BOT) sm.bot.nepairs[0] = 0; }
    lap_prepare<NT>(lw, n, words);
    __syncthreads();
    PHASE(6);
    cost.fuseA = false; cost.fuseB = true;
    cost.embA = false; cost.embB = BOT && p.with_reid;
    lim.limA = p.second_thresh; lim.limB = p.unconf_thresh;
    const int np2 = sm.np2[0];
    if (FUSE2 && np2 <= SM::P2CAP) {
        // the candidates were found during the first build: keep those whose row is still unmatched (tracked rows) /
        // whose detection is still unused (unconfirmed rows); low detections are never used by the first pass
        for (int k = tid; k < np2; k += NT) {
            const uint32_t pr = sm.p2pair()[k];
            const int r = pr >> 16, j = pr & 0xffff;
            const int rt = sm.rowtype[r];
            if (rt == RT_A || (rt == RT_B && !(sm.dflag[j] & DF_USED))) graph_add_edge<KIND>(sm, r, j, sm.p2cost()[k]);
        }
        __syncthreads();
    } else {
        graph_phase_a<KIND>(sm, t, n, words, cm);
        __syncthreads();
        graph_phase_b<NT, KIND, BOT>(sm, n, words, cost, lim, p.proximity_thresh);
        __syncthreads();
        if constexpr (BOT) {
            if (sm.npairs[0] > SM::PCAP || sm.ecount[0] > SM::ECAP) err |= B200_ERR_BOT_CAPACITY;
            graph_phase_emb<NT, KIND>(sm, p, s, dfeat, lim);
            __syncthreads();
        }
    }
    if (sm.ecount[0] > SM::ECAP) {                        // uniform
        graph_build_adj<NT, KIND>(sm, n, words, cost, lim);
        __syncthreads();
    }
    PHASE(7);
    lw.dbg_slot = 14;
    lap_sparse_solve<NT>(lw, n, words, lim, cost);
    PHASE(10);

    // ---- deferred Kalman work + lifecycle, thread t (byte_tracker.py:64-98, :222-253) --------
    // covariance: HBM -> registers (first and only read), predict, update; then parked in shared memory
    // (sm.park, storage of the solver arrays) until the final write; the position half of the mean lives in sm.mean.
    double velo[4] = {0.0, 0.0, 0.0, 0.0};
    int tid_id = 0, len = 0, det_ind = 0, start = 0;
    double score = 0.0, cls = 0.0;
    int cat = CAT_NONE;
    if (t < n) {
        KfState ks;
        bool unmatched2 = false;
        int j = sm.match[t];
        if (sm.rowtype[t] != RT_NONE) {
            j = sm.xr[t];
            if (j >= 0) sm.dflag[j] |= DF_USED; else unmatched2 = true;
        }
        G4 gA, gB;                                      // CAM only
        if constexpr (!CAM) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ks.pp[i] = gf[(B200_TF_COV + 3 * i + 0) * TMAX + t];
            ks.pv[i] = gf[(B200_TF_COV + 3 * i + 1) * TMAX + t];
            ks.vv[i] = gf[(B200_TF_COV + 3 * i + 2) * TMAX + t];
        }
        } else {
            g4_load(gA, gf + (size_t)B200_TFC_COVA * TMAX + t, TMAX);
            g4_load(gB, gf + (size_t)B200_TFC_COVB * TMAX + t, TMAX);
        }
        tid_id = gi[B200_TI_ID * TMAX + t];
        len = gi[B200_TI_LEN * TMAX + t];
        det_ind = gi[B200_TI_DET * TMAX + t];
        score = gf[TF_SCORE * TMAX + t];
        cls = gf[TF_CLS * TMAX + t];
        start = sm.start_t[t];
#pragma unroll
        for (int c = 0; c < 4; ++c) ks.m[c] = sm.mean[c][t];
        // velocities: second (L2) read instead of 8 KB of shared memory; same zeroing rule as the motion step
#pragma unroll
        for (int c = 0; c < 4; ++c) ks.m[4 + c] = gf[(B200_TF_MEAN + 4 + c) * TMAX + t];
        if (role != ROLE_UNCONF && (fl & 3) != B200_ST_TRACKED) {
            ks.m[7] = 0.0;
            if (KIND == KF_XYWH) ks.m[6] = 0.0;
        }
        if constexpr (CAM) {
            // the whole filter step again from the stored state, with the operations of the motion step above: predict
            // (pool only), camera correction, so that the positions equal sm.mean bit for bit
            double m8[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) m8[c] = gf[(B200_TF_MEAN + c) * TMAX + t];
            gA.m[0] = m8[0]; gA.m[1] = m8[1]; gA.m[2] = ks.m[4]; gA.m[3] = ks.m[5];
            gB.m[0] = m8[2]; gB.m[1] = m8[3]; gB.m[2] = ks.m[6]; gB.m[3] = ks.m[7];
            if (role != ROLE_UNCONF) {
                const double spw = xmul(KF_W_POS, ref_w), sph = xmul(KF_W_POS, ref_h), svw = xmul(KF_W_VEL, ref_w), svh = xmul(KF_W_VEL, ref_h);
                const double q[4] = {xmul(spw, spw), xmul(sph, sph), xmul(svw, svw), xmul(svh, svh)};
                g4_predict(gA, q);
                g4_predict(gB, q);
            }
            if (wp) {
                const double M[4] = {wp[0], wp[1], wp[3], wp[4]}, tv[2] = {wp[2], wp[5]};
                g4_warp(gA, M, tv);
                g4_warp(gB, M, nullptr);
            }
        } else if (role != ROLE_UNCONF) {
            // covariance half of multi_predict; the noise uses the pre-motion w / h
            double ref4[4] = {0.0, 0.0, ref_w, ref_h};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double sp, sv;
                if (KIND != KF_XYWH && i == 2) { sp = 1e-2; sv = 1e-5; }
                else { const double r = kf_ref<KIND>(ref4, i); sp = xmul(KF_W_POS, r); sv = xmul(KF_W_VEL, r); }
                const double a = xadd(ks.pp[i], ks.pv[i]);
                const double b = xadd(ks.pv[i], ks.vv[i]);
                ks.pp[i] = xadd(xadd(a, b), xmul(sp, sp));
                ks.pv[i] = b;
                ks.vv[i] = xadd(ks.vv[i], xmul(sv, sv));
            }
        }
        const int st0 = fl & 3;
        if (j >= 0) {                                   // STrack.update / re_activate
            double z[4];
            det_measurement<KIND>(sm, j, z);
            if constexpr (CAM) {
                // botsort_kf.py:193-225 on the two blocks: measurement noise from the (corrected) state's w, h
                const double sw = xmul(KF_W_POS, gB.m[0]), sh = xmul(KF_W_POS, gB.m[1]);
                const double r[2] = {xmul(sw, sw), xmul(sh, sh)};
                g4_update_chol(gA, z, r);
                g4_update_chol(gB, z + 2, r);
                sm.mean[0][t] = gA.m[0]; sm.mean[1][t] = gA.m[1]; sm.mean[2][t] = gB.m[0]; sm.mean[3][t] = gB.m[1];
            } else {
            kf_update<KIND>(ks, z);
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.mean[c][t] = ks.m[c];
            }
            len = (st0 == B200_ST_TRACKED) ? len + 1 : 0;
            frame_t = frame;
            det_ind = j;
            score = sm.dconf[j];
            cls = det_cls(j);
            if constexpr (BOT) {
                double* h = p.cls_hist + ((size_t)s * TMAX + sm.bot.frow[t]) * 9;
                cls = cls_vote(h, cls, score, err);
                if (p.with_reid && (sm.dflag[j] & DF_HIGH)) sm.bot.emadet[t] = (short)j;   // low detections carry no embedding
            }
            fl = (fl & ~3) | B200_ST_TRACKED | B200_FLAG_ACTIVATED;
        } else if (unmatched2) {
            fl = (fl & ~3) | (role == ROLE_TRACKED ? B200_ST_LOST : B200_ST_REMOVED);   // mark_lost / mark_removed
        }
        if constexpr (CAM) {
            g4_store(gA, &sm.park[0][t], TMAX);
            g4_store(gB, &sm.park[10][t], TMAX);
            velo[0] = gA.m[2]; velo[1] = gA.m[3]; velo[2] = gB.m[2]; velo[3] = gB.m[3];
        } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sm.park[3 * i + 0][t] = ks.pp[i]; sm.park[3 * i + 1][t] = ks.pv[i]; sm.park[3 * i + 2][t] = ks.vv[i];
            velo[i] = ks.m[4 + i];
        }
        }
        int st = fl & 3;
        const bool sticky_old = fl & B200_FLAG_STICKY;      // id already in removed_stracks
        if (role == ROLE_LOST) {
            if (st == B200_ST_TRACKED) cat = CAT_REFOUND;
            else {
                if (frame - frame_t > p.max_time_lost) {
                    fl = (fl & ~3) | B200_ST_REMOVED;       // stays listed one more frame (removed-lag)
                    st = B200_ST_REMOVED;
                }
                if (!sticky_old) cat = CAT_LOST_OLD;
                if (st == B200_ST_REMOVED) fl |= B200_FLAG_STICKY;
            }
        } else if (st == B200_ST_TRACKED) cat = CAT_KEEP;
        else if (st == B200_ST_LOST && !sticky_old) cat = CAT_LOST_NEW;
        sm.cat[t] = (unsigned char)cat;
        sm.frame_t[t] = frame_t;
    }
    __syncthreads();
    PHASE(8);
    if constexpr (BOT) {
        // STrack.update_features (bot_sort.py:40-48) of every track matched to a first-round detection
        if (p.with_reid) {
            const int nv = p.feat_dim >> 2;
            const float A = 0.9f, B = 0.1f;             // alpha, float32(1 - alpha)
            for (int q = warp; q < n; q += NT / 32) {
                const int j = sm.bot.emadet[q];
                if (j < 0) continue;
                float4* trk = reinterpret_cast<float4*>(p.feat_pool + ((size_t)s * TMAX + sm.bot.frow[q]) * p.feat_dim);
                const float4* det = reinterpret_cast<const float4*>(dfeat + (size_t)j * p.feat_dim);
                const RowDiv n0 = row_div(sm.bot.dn0[j]), n1 = row_div(sm.bot.dn1[j]), n2 = row_div(sm.bot.dn2[j]);
                auto blend = [&](int i) {
                    const float4 f = f4_div(f4_curr(det[i], n0, n1), n2);
                    const float4 a = trk[i];
                    return make_float4(__fadd_rn(__fmul_rn(A, a.x), __fmul_rn(B, f.x)), __fadd_rn(__fmul_rn(A, a.y), __fmul_rn(B, f.y)),
                                       __fadd_rn(__fmul_rn(A, a.z), __fmul_rn(B, f.z)), __fadd_rn(__fmul_rn(A, a.w), __fmul_rn(B, f.w)));
                };
                if (nv <= 128) {
                    const Row4 ra = load_row4(trk, nv, lane), rd = load_curr4(det, nv, lane, n0, n1);
                    Row4 rb;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 f = f4_div(rd.v[k], n2), a = ra.v[k];
                        rb.v[k] = make_float4(__fadd_rn(__fmul_rn(A, a.x), __fmul_rn(B, f.x)), __fadd_rn(__fmul_rn(A, a.y), __fmul_rn(B, f.y)),
                                              __fadd_rn(__fmul_rn(A, a.z), __fmul_rn(B, f.z)), __fadd_rn(__fmul_rn(A, a.w), __fmul_rn(B, f.w)));
                    }
                    const RowDiv nn = row_div(norm_f32(warp_sum(row4_sq(rb))));
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (lane + 32 * k < nv) trk[lane + 32 * k] = f4_div(rb.v[k], nn);
                    continue;
                }
                double acc = 0.0;
                for (int i = lane; i < nv; i += 32) acc += f4_sq(blend(i));
                const float nn = norm_f32(warp_sum(acc));
                for (int i = lane; i < nv; i += 32) trk[i] = f4_div(blend(i), nn);
            }
        }
    }

    // compact list of the new lost list (old entries first, then the newly lost) for the
    // duplicate test; packed counters: [0:16) old-lost, [16:32) new-lost
    int nLostList;
    {
        const unsigned long long val = (cat == CAT_LOST_OLD ? 1ull : 0ull) | (cat == CAT_LOST_NEW ? (1ull << 16) : 0ull);
        unsigned long long tot;
        const unsigned long long ex = block_exscan1<NT>(val, sm.scratch, tot);
        const int nLostOld = (int)(tot & 0xffff);
        if (cat == CAT_LOST_OLD) sm.lostlist[ex & 0xffff] = (short)t;
        if (cat == CAT_LOST_NEW) sm.lostlist[nLostOld + ((ex >> 16) & 0xffff)] = (short)t;
        nLostList = nLostOld + (int)((tot >> 16) & 0xffff);
        __syncthreads();
        PHASE(9);
    }

    // is detection `tid` the seed of a new track?  (unmatched high detection, byte_tracker.py:242-248)
    const bool born = tid < nd && (sm.dflag[tid] & (DF_HIGH | DF_USED)) == DF_HIGH && !(sm.dconf[tid] < p.new_thresh);

    // ---- remove_duplicate_stracks (byte_tracker.py:312-325): tracked' x lost', 1-iou < 0.15
    // tracked' = kept slots, new tracks, re-found slots
    if (nLostList > 0) {
        // boxes / ages of the lost list once, in scratch (u: 4 planes of TMAX/4; coldeg: ages)
        constexpr int LCAP = TMAX / 4;
        const bool lost_cached = nLostList <= LCAP && nLostList <= DMAX;
        if (lost_cached && tid < nLostList) {
            const int q = sm.lostlist[tid];
            const Box b = track_box<KIND>(sm, q);
            sm.u[tid] = b.x1; sm.u[LCAP + tid] = b.y1; sm.u[2 * LCAP + tid] = b.x2; sm.u[3 * LCAP + tid] = b.y2;
            sm.coldeg[tid] = sm.frame_t[q] - sm.start_t[q];
            // conservative fp32 copy (rounded outwards) for the overlap pre-test; dboxf is free after the last graph build
            sm.dboxf[tid] = make_float4(__double2float_rd(b.x1), __double2float_rd(b.y1), __double2float_ru(b.x2), __double2float_ru(b.y2));
        }
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
            Box a;
            int age;
            if (pass == 0) {
                if (cat != CAT_KEEP && cat != CAT_REFOUND) continue;
                a = track_box<KIND>(sm, t);
                age = frame_t - start;
            } else {
                if (!born) continue;
                double z[4];
                det_measurement<KIND>(sm, tid, z);
                a = mean_to_box<KIND>(z[0], z[1], z[2], z[3]);
                age = 0;
            }
            bool dropme = false;
            const float ax1 = __double2float_rd(a.x1), ay1 = __double2float_rd(a.y1);
            const float ax2 = __double2float_ru(a.x2), ay2 = __double2float_ru(a.y2);
            auto exact = [&](int k, const Box& b) {
                if (!box_overlap(a, b)) return;
                if (xsub(1.0, box_iou(a, b)) < p.dup_thresh) {
                    const int q = sm.lostlist[k];
                    const int ageq = lost_cached ? sm.coldeg[k] : sm.frame_t[q] - sm.start_t[q];
                    if (age > ageq) sm.drop[q] = 1; else dropme = true;
                }
            };
            if (lost_cached) {
                // branch-free sweep first (the loads of consecutive entries overlap), exact test only on the hits;
                // nLostList <= LCAP <= 64 (128 for the 512-slot variant: two sweeps)
                for (int k0 = 0; k0 < nLostList; k0 += 64) {
                    unsigned long long hits = 0ull;
                    const int kn = min(64, nLostList - k0);
#pragma unroll 4
                    for (int k = 0; k < kn; ++k) {
                        const float4 f = sm.dboxf[k0 + k];
                        hits |= (unsigned long long)(f.x < ax2 && ax1 < f.z && f.y < ay2 && ay1 < f.w) << k;   // disjoint even after outward rounding -> 0
                    }
                    while (hits) {
                        const int k = k0 + __ffsll((long long)hits) - 1;
                        hits &= hits - 1;
                        Box b;
                        b.x1 = sm.u[k]; b.y1 = sm.u[LCAP + k]; b.x2 = sm.u[2 * LCAP + k]; b.y2 = sm.u[3 * LCAP + k];
                        exact(k, b);
                    }
                }
            } else {
                for (int k = 0; k < nLostList; ++k) exact(k, track_box<KIND>(sm, sm.lostlist[k]));
            }
            if (dropme) sm.drop[pass == 0 ? t : TMAX + tid] = 1;
        }
        __syncthreads();
        PHASE(11);
    }

    // ---- destinations.  One packed scan (10-bit fields), element i = slot i and detection i:
    //   slots: [0) keep, [10) refound, [20) lostOld, [30) lostNew ; dets: [40) born kept, [50) born (all)
    // Every CAT_KEEP / CAT_REFOUND entry is activated, so output rows = keep ++ born (frame 1 only) ++ refound.
    const bool born_active = frame == 1;                 // STrack.activate: is_activated only on frame 1
    double* wf = p.state_f + (size_t)s * NFk * TMAX;
    int* wi = p.state_i + (size_t)s * NI * TMAX;
    double* gout = packed ? nullptr : p.out + (size_t)s * p.max_tracks * 8;
    const int out_cap = packed ? min(p.max_tracks, nd_in) : p.max_tracks;
    // result row: the reference's [x1, y1, x2, y2, id, conf, cls, det_ind] (64 bytes, four 16-byte stores), or the compact
    // row of the packed interface (layout.h: 40 bytes, BoT-SORT 48) - conf / cls there are the caller's own input columns
    auto write_row = [&](int orow, const Box& b, int id, double conf, double cl, int di) {
        if (!packed) {
            double2* o = reinterpret_cast<double2*>(gout + (size_t)orow * 8);
            o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
            o[2] = make_double2((double)id, conf); o[3] = make_double2(cl, (double)di);
        } else if constexpr (BOT) {
            double2* o = reinterpret_cast<double2*>(p.rows + (size_t)(roff + orow) * B200_ROW_BOT);
            o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
            reinterpret_cast<int4*>(o)[2] = make_int4(id, di, __float_as_int((float)cl), __float_as_int((float)conf));
        } else {
            double* o = reinterpret_cast<double*>(p.rows + (size_t)(roff + orow) * B200_ROW_BYTE);
            o[0] = b.x1; o[1] = b.y1; o[2] = b.x2; o[3] = b.y2;
            reinterpret_cast<int2*>(o)[4] = make_int2(id, di);
        }
    };
    unsigned long long val = 0ull;
    if (cat != CAT_NONE && !sm.drop[t]) val = 1ull << (10 * (cat - 1));
    if (born) {
        val |= 1ull << 50;                                // activate() ran: consumes an id even if dropped
        if (!sm.drop[TMAX + tid]) val |= 1ull << 40;
    }
    if constexpr (BOT) {
        if (val & ((1ull << 40) - 1)) sm.bot.rowused[sm.bot.frow[t]] = 1;      // rows of the tracks that stay listed
    }
    unsigned long long totals;
    const unsigned long long ex = block_exscan1<NT>(val, sm.scratch + 16, totals);
    int nfree = 0;
    if constexpr (BOT) {
        // embedding-pool rows of dead tracks are recycled: the k-th stored newborn takes the k-th free row
        const bool isfree = !sm.bot.rowused[t];
        unsigned long long tf;
        const int rank = (int)block_exscan1<NT>(isfree ? 1ull : 0ull, sm.scratch, tf);
        nfree = (int)tf;
        if (isfree) sm.bot.freelist[rank] = (short)t;
        __syncthreads();
    }
    const int totKeep = (int)(totals & 1023), totRef = (int)((totals >> 10) & 1023);
    const int totLostOld = (int)((totals >> 20) & 1023), totLostNew = (int)((totals >> 30) & 1023);
    const int totBorn = (int)((totals >> 40) & 1023), totBornAll = (int)((totals >> 50) & 1023);
    const int newT = totKeep + totBorn + totRef;
    const int newL = totLostOld + totLostNew;
    if (newT + newL > min(TMAX, p.max_tracks)) err |= B200_ERR_TRACK_OVERFLOW;
    const int cap = min(TMAX, p.max_tracks);
    const int rowsBorn = born_active ? totBorn : 0;

    if (val & ((1ull << 40) - 1)) {
        int dst, orow = -1;
        if (cat == CAT_KEEP) { dst = (int)(ex & 1023); orow = dst; }
        else if (cat == CAT_REFOUND) { const int k = (int)((ex >> 10) & 1023); dst = totKeep + totBorn + k; orow = totKeep + rowsBorn + k; }
        else if (cat == CAT_LOST_OLD) dst = newT + (int)((ex >> 20) & 1023);
        else dst = newT + totLostOld + (int)((ex >> 30) & 1023);
        if (dst < cap) {
#pragma unroll
            for (int c = 0; c < 4; ++c) wf[(B200_TF_MEAN + c) * TMAX + dst] = sm.mean[c][t];
#pragma unroll
            for (int c = 0; c < 4; ++c) wf[(B200_TF_MEAN + 4 + c) * TMAX + dst] = velo[c];
            if constexpr (CAM) {
#pragma unroll
                for (int i = 0; i < 20; ++i) wf[(B200_TFC_COVA + i) * TMAX + dst] = sm.park[i][t];
            } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                wf[(B200_TF_COV + 3 * i + 0) * TMAX + dst] = sm.park[3 * i + 0][t];
                wf[(B200_TF_COV + 3 * i + 1) * TMAX + dst] = sm.park[3 * i + 1][t];
                wf[(B200_TF_COV + 3 * i + 2) * TMAX + dst] = sm.park[3 * i + 2][t];
            }
            }
            wf[TF_SCORE * TMAX + dst] = score;
            wf[TF_CLS * TMAX + dst] = cls;
            wi[B200_TI_ID * TMAX + dst] = tid_id;
            wi[B200_TI_FRAME * TMAX + dst] = frame_t;
            wi[B200_TI_START * TMAX + dst] = start;
            wi[B200_TI_LEN * TMAX + dst] = len;
            wi[B200_TI_DET * TMAX + dst] = det_ind;
            wi[B200_TI_FLAGS * TMAX + dst] = fl;
            if constexpr (BOT) wi[B200_TI_FROW * TMAX + dst] = sm.bot.frow[t];
        }
        if (orow >= 0 && orow < out_cap) write_row(orow, track_box<KIND>(sm, t), tid_id, score, cls, det_ind);
    }
    if (val & (1ull << 40)) {                           // STrack.activate (byte_tracker.py:50-62)
        const int j = tid;
        const int k = (int)((ex >> 40) & 1023);
        const int dst = totKeep + k;
        const int id = id0 + (int)((ex >> 50) & 1023) + 1;
        double z[4];
        det_measurement<KIND>(sm, j, z);
        KfState kn;
        kf_initiate<KIND>(z, kn);
        if (dst < cap) {
#pragma unroll
            for (int c = 0; c < 8; ++c) wf[(B200_TF_MEAN + c) * TMAX + dst] = kn.m[c];
            if constexpr (CAM) {
                // diagonal start: (x, y, vx, vy) and (w, h, vw, vh) blocks, entries 00 11 22 33 at triangle positions 0 4 7 9
#pragma unroll
                for (int i = 0; i < 20; ++i) wf[(B200_TFC_COVA + i) * TMAX + dst] = 0.0;
                wf[(B200_TFC_COVA + 0) * TMAX + dst] = kn.pp[0]; wf[(B200_TFC_COVA + 4) * TMAX + dst] = kn.pp[1];
                wf[(B200_TFC_COVA + 7) * TMAX + dst] = kn.vv[0]; wf[(B200_TFC_COVA + 9) * TMAX + dst] = kn.vv[1];
                wf[(B200_TFC_COVB + 0) * TMAX + dst] = kn.pp[2]; wf[(B200_TFC_COVB + 4) * TMAX + dst] = kn.pp[3];
                wf[(B200_TFC_COVB + 7) * TMAX + dst] = kn.vv[2]; wf[(B200_TFC_COVB + 9) * TMAX + dst] = kn.vv[3];
            } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                wf[(B200_TF_COV + 3 * a + 0) * TMAX + dst] = kn.pp[a];
                wf[(B200_TF_COV + 3 * a + 1) * TMAX + dst] = kn.pv[a];
                wf[(B200_TF_COV + 3 * a + 2) * TMAX + dst] = kn.vv[a];
            }
            }
            wf[TF_SCORE * TMAX + dst] = sm.dconf[j];
            wf[TF_CLS * TMAX + dst] = det_cls(j);
            wi[B200_TI_ID * TMAX + dst] = id;
            wi[B200_TI_FRAME * TMAX + dst] = frame;
            wi[B200_TI_START * TMAX + dst] = frame;
            wi[B200_TI_LEN * TMAX + dst] = 0;
            wi[B200_TI_DET * TMAX + dst] = j;
            wi[B200_TI_FLAGS * TMAX + dst] = B200_ST_TRACKED | (born_active ? B200_FLAG_ACTIVATED : 0);
        }
        if constexpr (BOT) {
            const bool stored = dst < cap && k < nfree;
            const int row = stored ? sm.bot.freelist[k] : -1;
            sm.bot.nbdet[k] = stored ? (short)j : (short)-1;
            sm.bot.nbrow[k] = (short)row;
            if (stored) {
                wi[B200_TI_FROW * TMAX + dst] = row;
                double* h = p.cls_hist + ((size_t)s * TMAX + row) * 9;     // STrack.__init__: cls_hist = [[cls, score]]
                h[0] = det_cls(j); h[4] = sm.dconf[j]; h[8] = 1.0;
            }
        }
        if (born_active) {
            const int orow = totKeep + k;
            if (orow < out_cap) write_row(orow, mean_to_box<KIND>(z[0], z[1], z[2], z[3]), id, sm.dconf[j], det_cls(j), j);
        }
    }
    if constexpr (BOT) {
        // a new track's smoothed embedding is its detection's (twice normalised) embedding (bot_sort.py:40-48)
        if (p.with_reid) {
            __syncthreads();
            const int nv = p.feat_dim >> 2;
            for (int k = warp; k < totBorn; k += NT / 32) {
                const int j = sm.bot.nbdet[k];
                if (j < 0) continue;
                float4* trk = reinterpret_cast<float4*>(p.feat_pool + ((size_t)s * TMAX + sm.bot.nbrow[k]) * p.feat_dim);
                const float4* det = reinterpret_cast<const float4*>(dfeat + (size_t)j * p.feat_dim);
                const RowDiv n0 = row_div(sm.bot.dn0[j]), n1 = row_div(sm.bot.dn1[j]);
                if (nv <= 128) {
                    const Row4 r = load_curr4(det, nv, lane, n0, n1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (lane + 32 * k < nv) trk[lane + 32 * k] = r.v[k];
                } else
                for (int i = lane; i < nv; i += 32) trk[i] = f4_curr(det[i], n0, n1);
            }
        }
    }
    PHASE(15);
    if (err) { atomicOr(p.err, err); if (p.err_out) atomicOr(p.err_out, err); }
    if (tid == 0) {
        if (p.dbg) atomicAdd(&p.dbg[0], 1ull);
        counts[0] = min(newT, cap);
        counts[1] = min(newL, cap - min(newT, cap));
        counts[2] = id0 + totBornAll;
        counts[3] = frame;
        p.nout[s] = min(totKeep + rowsBorn + totRef, out_cap);
        p.track_updates[s] += (unsigned long long)n;
    }
}

#undef PHASE
// four CTAs of the (224, 224) variant share an SM: 227 KB of shared memory, 1 KB reserved per CTA
static_assert(sizeof(StepSmem<224, 224>) + 1024 <= 227 * 1024 / 4, "(224, 224) ByteTrack variant must fit 4 CTAs per SM");
struct Variant { int tmax, dmax; };
constexpr Variant kVariants[] = {{64, 64}, {128, 128}, {224, 224}, {256, 256}, {512, 512}};

template <int KIND, int TMAX, int DMAX, bool BOT, bool CAM = false>
cudaError_t launch_variant(const StepParams& p, cudaStream_t stream) {
    auto kern = bytetrack_step_kernel<TMAX, KIND, TMAX, DMAX, BOT, CAM>;
    // profiling aid: B200_STEP_SMEM_PAD=<bytes> lowers the number of co-resident CTAs (latency vs throughput experiments)
    static const size_t pad = [] { const char* v = getenv("B200_STEP_SMEM_PAD"); return v ? (size_t)atol(v) : (size_t)0; }();
    const size_t smem = sizeof(StepSmem<TMAX, DMAX, BOT, CAM>) + pad;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.n_streams, TMAX, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int KIND, bool BOT, bool CAM = false>
cudaError_t launch_kind(const StepParams& p, int v, cudaStream_t stream) {
    switch (v) {
        case 0: return launch_variant<KIND, 64, 64, BOT, CAM>(p, stream);
        case 1: return launch_variant<KIND, 128, 128, BOT, CAM>(p, stream);
        case 2: return launch_variant<KIND, 224, 224, BOT, CAM>(p, stream);
        case 3: return launch_variant<KIND, 256, 256, BOT, CAM>(p, stream);
        case 4: return launch_variant<KIND, 512, 512, BOT, CAM>(p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

#if !B200_STEP_PACKED
int bytetrack_step_variant(int max_tracks, int max_dets) {
    for (int v = 0; v < (int)(sizeof(kVariants) / sizeof(kVariants[0])); ++v)
        if (max_tracks <= kVariants[v].tmax && max_dets <= kVariants[v].dmax) return v;
    return -1;
}
int bytetrack_step_tmax(int variant) { return kVariants[variant].tmax; }
int step_variant_dmax(int variant) { return kVariants[variant].dmax; }
size_t bytetrack_step_smem(int variant, bool botsort, bool cam) {
    switch (variant) {
        case 0: return cam ? sizeof(StepSmem<64, 64, true, true>) : botsort ? sizeof(StepSmem<64, 64, true>) : sizeof(StepSmem<64, 64>);
        case 1: return cam ? sizeof(StepSmem<128, 128, true, true>) : botsort ? sizeof(StepSmem<128, 128, true>) : sizeof(StepSmem<128, 128>);
        case 2: return cam ? sizeof(StepSmem<224, 224, true, true>) : botsort ? sizeof(StepSmem<224, 224, true>) : sizeof(StepSmem<224, 224>);
        case 3: return cam ? sizeof(StepSmem<256, 256, true, true>) : botsort ? sizeof(StepSmem<256, 256, true>) : sizeof(StepSmem<256, 256>);
        case 4: return cam ? sizeof(StepSmem<512, 512, true, true>) : botsort ? sizeof(StepSmem<512, 512, true>) : sizeof(StepSmem<512, 512>);
    }
    return 0;
}

cudaError_t launch_bytetrack_step(const StepParams& p, int kf_kind, int variant, cudaStream_t stream) {
    return kf_kind == KF_XYWH ? launch_kind<KF_XYWH, false>(p, variant, stream) : launch_kind<KF_XYAH, false>(p, variant, stream);
}

cudaError_t launch_botsort_step(const StepParams& p, int variant, cudaStream_t stream, bool cam) {
    return cam ? launch_kind<KF_XYWH, true, true>(p, variant, stream) : launch_kind<KF_XYWH, true>(p, variant, stream);
}
#else
cudaError_t launch_bytetrack_step_packed(const StepParams& p, int kf_kind, int variant, cudaStream_t stream) {
    return kf_kind == KF_XYWH ? launch_kind<KF_XYWH, false>(p, variant, stream) : launch_kind<KF_XYAH, false>(p, variant, stream);
}

cudaError_t launch_botsort_step_packed(const StepParams& p, int variant, cudaStream_t stream, bool cam) {
    return cam ? launch_kind<KF_XYWH, true, true>(p, variant, stream) : launch_kind<KF_XYWH, true>(p, variant, stream);
}
#endif

}  // namespace b200
