// HybridSORT frame step for many independent streams: one kernel launch per frame, one CTA per stream, one thread per
// tracker slot - the OC-SORT step (ocsort_step.cu) with HybridSORT's score-carrying filter, four corner velocities,
// appearance term and long-term-ReID correction.
//
// Replaces HybridSORT.update (boxmot/trackers/hybridsort/hybridsort.py:373-570) and what it calls:
//   KalmanBoxTracker.predict :299-322 (9-d [u, v, s, c, r, ...] filter of boxmot/motion/kalman_filters/hybridsort_kf.py,
//       the filter's score clipped to [0.6, 1] as the tracker's "kalman score")
//   embedding_distance (boxmot/trackers/hybridsort/association.py:667-684: cdist 'cosine' in float64, clamped at 0)
//   associate_4_points_with_score_with_reid (association.py:495-581): similarity + the velocity-direction term of the four
//       box corners (cost_vel :314-335 on speed_direction_batch_lt/rt/lb/rb :338-383) - 1.3 x embedding distance through
//       lapjv without a limit; a matched pair is dropped again when its embedding distance exceeds 0.4 AND its similarity
//       minus |filter score - detection score| falls below the threshold (:557-566)
//   the observation-centric recovery round :513-545
//   KalmanBoxTracker.update :220-297: corner velocities summed over the delta_t window, observation ring, KalmanFilter.update
//       (hybridsort_kf.py:439-528, Joseph form) with the observation-centric re-update of unfreeze (:390-436, which reads
//       the score as the aspect ratio - kept, kf_hybrid.cuh)
//   update_features :188-205 (float32 blend with alpha 0.8, in-place normalisations), new trackers, the reversed output scan.
//
// What HybridSORT.__init__ fixes (:337-364) is compiled in: TCM_first_step with weight 0, EG_weight_high_score 1.3,
// longterm_reid_weight 0 (the 30-deep feature bank only enters the cost through that zero weight: it is not kept),
// longterm_reid_correction_thresh 0.4, ECC off, use_byte off (tracker_zoo.py:100-115 never forwards it, and the branch
// passes an embedding row where the class goes, :470-474).
// Reference quirk kept: the class and the LAST RESULT COLUMN of a matched / new tracker are read from the unfiltered
// detection array at the FILTERED row index (dets0[row, 5], dets0[row, 6] = that row's score, :396-404 / :451 / :549).
//
// Unlike OC-SORT every (detection, tracker) pair has an appearance term, so nothing can be pruned: the embedding
// distances (one warp per detection row, the row held in registers as doubles, products of fp32 values are exact in
// fp64) and then the full costs go to a per-stream matrix in global memory (L2 resident while the CTA lives), and the
// matrix-free solver of lap_dense.cuh reads its costs from there.
#include "boxes.cuh"
#include "kf_hybrid.cuh"
#include "lap_dense.cuh"
#include "layout.h"
#include "oc_common.cuh"
#include "step_params.h"

namespace b200 {
namespace {

template <int TMAX, int DMAX>
struct alignas(16) HySmem {
    double tbox[4][TMAX];           // predicted box, convert_x_to_bbox
    double lbox[4][TMAX];           // last_observation box (placeholder -1)
    double kbox[4][TMAX];           // k_previous_obs box
    double vel[8][TMAX];            // (dy, dx) of the lt, rt, lb, rb corners
    double kscore[TMAX];            // clip(x[3], 0.6, 1)
    double tnorm[TMAX];             // |smoothed embedding| of a slot (float64 of the fp32 values)
    double dbox[4][DMAX];
    double dconf[DMAX];
    double dnorm[DMAX];
    double u[DMAX], v[TMAX];
    double remb[DMAX];              // embedding distance of the pair a row was assigned
    double red_v[64];
    unsigned long long scratch[40];
    int pred[TMAX], xr[DMAX], yc[TMAX], claim[DMAX], partner[TMAX];
    int red_i[64];
    int rowmatch[DMAX];
    int misc[8];
    short hd[DMAX], ht[TMAX], drow[DMAX], dmatch[DMAX], tmatch[TMAX], ud[DMAX], ut[TMAX], frow[TMAX], freelist[TMAX];
    short nbrow[DMAX], nbdet[DMAX];
    unsigned char kvalid[TMAX], alive[TMAX], dstate[DMAX], rowused[TMAX], ema[TMAX], tdeg[TMAX], ddeg[DMAX];
    // staging of the smoothed track embeddings for the dense cosine matrix: two tiles of eight fp32 rows (up to 512 values)
    float4 etile[2][8][128];
    unsigned long long ebar[2];     // "tile landed" mbarriers of the two buffers
};

// ---- bulk copies (TMA, 1-D) into shared memory, completion on an mbarrier -------------------------------------
__device__ __forceinline__ uint32_t hy_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hy_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hy_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void hy_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hy_smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol error becomes an error flag instead of a hung GPU
__device__ __forceinline__ bool hy_mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = hy_smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void hy_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(hy_smem_u32(dst)), "l"(src), "r"(bytes), "r"(hy_smem_u32(bar)) : "memory");
}

// Cost of (row r = detection hd[r], column c = live tracker ht[c]) in the first association (association.py:508-541):
//   -(similarity + (lt + rt + lb + rb)) + 1.3 * embedding distance, plus the canonical tie-break.
// The embedding distances E are a stored matrix (the dense cosine pass); the geometric part - a similarity and four
// direction terms with a square root, a reciprocal and an arc cosine each - is evaluated ON DEMAND: for iou / giou a pair
// of disjoint regular boxes has similarity exactly +0.0 and the four direction terms are bounded by 2 |inertia| score, so
// lower() = 1.3 E - that bound is a valid lower bound of the cost, and with unrelated embeddings (E around 1) it lies far
// above the cost of a row's real partner: the row reduction and the solver skip nearly every pair without evaluating it.
// Evaluated in full for every pair, this geometry was a quarter of the step's instructions.
template <class SM>
struct HyCost {
    const SM& sm;
    const double* E;
    int ld, Cn, func;
    double W, H, inertia;
    bool sparse;                    // iou / giou: a disjoint pair has similarity exactly +0.0
    __device__ __forceinline__ Box dbox(int j) const { return Box{sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]}; }
    __device__ __forceinline__ Box tbox(int sl) const { return Box{sm.tbox[0][sl], sm.tbox[1][sl], sm.tbox[2][sl], sm.tbox[3][sl]}; }
    __device__ __forceinline__ double bound(int j) const { return xmul(2.0, fabs(xmul(inertia, sm.dconf[j]))) + 1e-6; }
    __device__ __forceinline__ bool free_pair(int j, int sl) const {
        return sparse && !sm.ddeg[j] && !sm.tdeg[sl] && !box_overlap(dbox(j), tbox(sl));
    }
    __device__ __forceinline__ double operator()(int r, int c) const {
        const int j = sm.hd[r], sl = sm.ht[c];
        const Box db = dbox(j);
        const double sv = oc_sim(func, db, tbox(sl), W, H);
        double ang = 0.0;
        if (sm.kvalid[sl]) {
            const double kx1 = sm.kbox[0][sl], ky1 = sm.kbox[1][sl], kx2 = sm.kbox[2][sl], ky2 = sm.kbox[3][sl];
            const double sc = sm.dconf[j];
            // corner order of the reference's sum: lt = (x1, y1), rt = (x1, y2), lb = (x2, y1), rb = (x2, y2)
            double a4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double vy = sm.vel[2 * q][sl], vx = sm.vel[2 * q + 1][sl];
                const double kcx = q < 2 ? kx1 : kx2, kcy = (q & 1) ? ky2 : ky1;
                const double dcx = q < 2 ? db.x1 : db.x2, dcy = (q & 1) ? db.y2 : db.y1;
                a4[q] = (vx == 0.0 && vy == 0.0) ? 0.0 : oc_angle(vy, vx, kcx, kcy, true, dcx, dcy, inertia, sc);
            }
            ang = xadd(xadd(xadd(a4[0], a4[1]), a4[2]), a4[3]);
        }
        const double cst = xadd(-xadd(sv, ang), xmul(1.3, E[(size_t)r * ld + c]));
        return xadd(cst, xmul((double)(r * Cn + c), TIE_EPS));
    }
    __device__ __forceinline__ double lower(int r, int c) const {
        const int j = sm.hd[r];
        return free_pair(j, sm.ht[c]) ? xmul(1.3, E[(size_t)r * ld + c]) - bound(j) : -__longlong_as_double(0x7ff0000000000000LL);
    }
};

template <int NT, class SM>
__device__ __forceinline__ DenseLap make_dense(SM& sm) {
    DenseLap w;
    w.u = sm.u; w.v = sm.v; w.pred = sm.pred; w.xr = sm.xr; w.yc = sm.yc; w.claim = sm.claim;
    w.red_v = sm.red_v; w.red_i = sm.red_i; w.freerow = sm.rowmatch; w.dbg = nullptr;
    return w;
}

__device__ __forceinline__ double hy_warp_sum(double x) {
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}
__device__ __forceinline__ float4 hy_f4_div(float4 a, const RowDiv& d) {       // common.cuh: the quotients of __fdiv_rn
    return make_float4(fdiv_row(a.x, d), fdiv_row(a.y, d), fdiv_row(a.z, d), fdiv_row(a.w, d));
}
__device__ __forceinline__ double hy_f4_sq(float4 a) {
    return (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
}
__device__ __forceinline__ double hy_f4_dot(float4 a, float4 b) {
    return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
}
// np.linalg.norm of a float32 row: sqrt(dot(x, x)) rounded to fp32 (accumulated in double here; BLAS order is unspecified)
__device__ __forceinline__ float hy_norm_f32(double sumsq) { return sqrtf((float)sumsq); }
__device__ __forceinline__ float4 hy_blend(float4 a, float4 f) {      // float32: 0.8 * smooth + float32(1 - 0.8) * feat
    const float A = 0.8f, B = 0.2f;
    return make_float4(__fadd_rn(__fmul_rn(A, a.x), __fmul_rn(B, f.x)), __fadd_rn(__fmul_rn(A, a.y), __fmul_rn(B, f.y)),
                       __fadd_rn(__fmul_rn(A, a.z), __fmul_rn(B, f.z)), __fadd_rn(__fmul_rn(A, a.w), __fmul_rn(B, f.w)));
}
// cdist 'cosine' from the dot product and the two norms, clamped like association.py:683
__device__ __forceinline__ double hy_cosine(double uv, double nu, double nv) {
    double c = uv / (nu * nv);
    if (fabs(c) > 1.0) c = copysign(1.0, c);
    return fmax(0.0, 1.0 - c);
}

// Embedding dot products.  A detection row lives in registers as 16 doubles per lane (value 4 k + e of lane l is element
// 4 (l + 32 k) + e of the row); a lane accumulates its share of a pair with a fixed chain of fused multiply-adds (the
// product of two fp32 values is exact in fp64, so the chain only fixes the ORDER of the additions), and eight such
// per-lane partials - eight pairs - are reduced TOGETHER: three exchange steps that halve the number of values a lane
// still carries (bits 4, 3, 2 of the lane index pick the half it keeps) and two plain steps, 9 shuffles of a double
// instead of 40; afterwards lane 4 q holds the sum of value q and eight lanes finalise their pairs in parallel.  Every
// value is summed over the lanes in the same pairing order, so a pair's bits do not depend on its position in the
// group.
__device__ __forceinline__ double hy_chain(const double (&d)[16], const float4 (&v)[4]) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        acc = fma(d[4 * k], (double)v[k].x, acc); acc = fma(d[4 * k + 1], (double)v[k].y, acc);
        acc = fma(d[4 * k + 2], (double)v[k].z, acc); acc = fma(d[4 * k + 3], (double)v[k].w, acc);
    }
    return acc;
}
__device__ __forceinline__ double hy_reduce8(double (&p)[8], int lane) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double recv = __shfl_xor_sync(0xffffffffu, h4 ? p[i] : p[i + 4], 16);
        p[i] = (h4 ? p[i + 4] : p[i]) + recv;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double recv = __shfl_xor_sync(0xffffffffu, h3 ? p[i] : p[i + 2], 8);
        p[i] = (h3 ? p[i + 2] : p[i]) + recv;
    }
    {
        const double recv = __shfl_xor_sync(0xffffffffu, h2 ? p[0] : p[1], 4);
        p[0] = (h2 ? p[1] : p[0]) + recv;
    }
    p[0] += __shfl_xor_sync(0xffffffffu, p[0], 2);
    p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
    return p[0];                                           // lanes 4 q + {0..3}: value q = 4 * bit4 + 2 * bit3 + bit2
}
// two detection rows x four smoothed rows resident in shared memory (zero padded to 128 float4): value q < 4 is
// (row 0, column q), value 4 + q is (row 1, column q) - a loaded, converted element serves two pairs
__device__ __forceinline__ double hy_dot2x4(const double (&d0)[16], const double (&d1)[16], const float4 (*tile)[128], int lane) {
    double p[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = tile[q][lane + 32 * k];
        p[q] = hy_chain(d0, v);
        p[q + 4] = hy_chain(d1, v);
    }
    return hy_reduce8(p, lane);
}
// the detection row in that register form
__device__ __forceinline__ void hy_load_row16(double (&d)[16], const float4* a, int nv, int lane) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v = lane + 32 * k < nv ? a[lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
        d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
    }
}

template <int NT, int TMAX, int DMAX>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : (NT >= 224 ? 2 : (NT == 128 ? 4 : 6))))
hybridsort_step_kernel(const StepParams p) {
    static_assert(NT == TMAX && DMAX <= NT, "one thread per tracker slot; detections fit one pass");
    using SM = HySmem<TMAX, DMAX>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int s = blockIdx.x, tid = threadIdx.x, t = tid, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int* counts = p.counts + 4 * s;
    int n0 = counts[0];
    const int alive0 = counts[1], id0 = counts[2], frame = counts[3] + 1;
    int nd = p.ndets[s];
    int err = 0;
    const int dcap = min(DMAX, p.max_dets), tcap = min(TMAX, p.max_tracks);
    if (nd > dcap) { nd = dcap; err |= B200_ERR_DET_OVERFLOW; }
    if (nd < 0) nd = 0;
    double* gf = p.state_f + (size_t)s * B200_HY_NF * TMAX;
    int* gi = p.state_i + (size_t)s * B200_HY_NI * TMAX;
    const double thr = p.iou_thresh, W = p.img_w, H = p.img_h;
    const int func = p.asso_func, F = p.feat_dim, nv = F >> 2;
    float* pool = p.feat_pool + (size_t)s * TMAX * F;
    const float* dfeat = p.feats + (size_t)s * p.max_dets * F;          // row j = raw detection j (hybridsort.py:394)
    double* Cm = p.scratch + (size_t)s * DMAX * TMAX;                    // [row][column], leading dimension TMAX
    const double* dets_g = p.dets + (size_t)s * p.max_dets * 6;

    // ---- HBM -> shared memory: detections ------------------------------------------------------------
    for (int i = tid; i < nd * 6; i += NT) {
        const double val = dets_g[i];
        const int j = i / 6, c = i - 6 * j;
        if (c < 4) sm.dbox[c][j] = val;
        else if (c == 4) sm.dconf[j] = val;
    }
    auto det_cls = [&](int j) -> double { return dets_g[j * 6 + 5]; };
    __syncthreads();
    const int nhigh = __syncthreads_count(tid < nd && sm.dconf[tid] > p.det_thresh);
    // ---- compaction on demand (see ocsort_step.cu) ------------------------------------------------
    if (n0 > alive0 && n0 + nhigh > tcap) {              // uniform
        bool lv = false;
        if (t < n0) lv = gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE;
        unsigned long long tt;
        const int dst = (int)block_exscan<NT>(lv ? 1ull : 0ull, sm.scratch, tt);
        for (int c0 = 0; c0 < B200_HY_NF; c0 += 8) {
            double tmp[8];
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_HY_NF) tmp[c] = gf[(c0 + c) * TMAX + t];
            }
            __syncthreads();
            if (lv && dst != t) {
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c0 + c < B200_HY_NF) gf[(c0 + c) * TMAX + dst] = tmp[c];
            }
            __syncthreads();
        }
        int itmp[B200_HY_NI];
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_HY_NI; ++c) itmp[c] = gi[c * TMAX + t];
        }
        __syncthreads();
        if (lv && dst != t) {
#pragma unroll
            for (int c = 0; c < B200_HY_NI; ++c) gi[c * TMAX + dst] = itmp[c];
        }
        __syncthreads();
        n0 = (int)tt;
    }

    // ---- tracker side, thread t = slot t: predict (hybridsort.py:299-322); the covariance half is deferred ----------
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
        sm.frow[t] = (short)gi[B200_HYI_FROW * TMAX + t];
        double x[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) x[c] = gf[(B200_HY_X + c) * TMAX + t];
        if (xadd(x[7], x[2]) <= 0.0) x[7] = xmul(x[7], 0.0);
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = xadd(x[c], x[c + 5]);
        age += 1;
        if (tsu > 0) streak = 0;
        tsu += 1;
        const Box b = oc_x_to_box(x[0], x[1], x[2], x[4]);
        sm.tbox[0][t] = b.x1; sm.tbox[1][t] = b.y1; sm.tbox[2][t] = b.x2; sm.tbox[3][t] = b.y2;
        sm.kscore[t] = fmin(fmax(x[3], 0.6), 1.0);
        if (isnan(b.x1) || isnan(b.y1) || isnan(b.x2) || isnan(b.y2) || isnan(x[3])) { live = false; fl &= ~OCF_ALIVE; }   // :412-418
        const bool hasobs = fl & B200_OCF_HASOBS;
        double l[4] = {-1.0, -1.0, -1.0, -1.0};
        if (hasobs) {
#pragma unroll
            for (int c = 0; c < 4; ++c) l[c] = gf[(B200_HY_LAST + c) * TMAX + t];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.lbox[c][t] = l[c];
#pragma unroll
        for (int c = 0; c < 8; ++c) sm.vel[c][t] = gf[(B200_HY_VEL + c) * TMAX + t];
        // k_previous_obs (hybridsort.py:22-30): oldest observation among ages age-delta_t .. age-1, else the newest
        double kb[4] = {l[0], l[1], l[2], l[3]};
        if (hasobs) {
            const int ra[3] = {gi[(B200_OCI_RINGAGE + 0) * TMAX + t], gi[(B200_OCI_RINGAGE + 1) * TMAX + t], gi[(B200_OCI_RINGAGE + 2) * TMAX + t]};
            for (int dt = p.delta_t; dt >= 1; --dt) {
                const int a = age - dt;
                if (a < 0) continue;
                const int slot = a % 3;
                if (dt <= 3 && ra[slot] == a) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) kb[c] = gf[(B200_HY_RING + 4 * slot + c) * TMAX + t];
                    break;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.kbox[c][t] = kb[c];
        sm.kvalid[t] = hasobs;
    }
    if (t < TMAX) { sm.alive[t] = live; sm.tmatch[t] = -1; sm.rowused[t] = 0; sm.ema[t] = 0; }
    if (tid < DMAX) { sm.dmatch[tid] = -1; sm.drow[tid] = -1; sm.xr[tid] = -1; }
    __syncthreads();
    if (tid < DMAX) sm.dstate[tid] = (tid < nd && sm.dconf[tid] > p.det_thresh) ? DS_FREE0 : DS_NONE;    // hybridsort.py:401-402

    // compact row (detections above det_thresh) and column (alive trackers) lists
    int R, Cn;
    {
        const bool isrow = tid < nd && sm.dconf[tid] > p.det_thresh;
        const unsigned long long val = (isrow ? 1ull : 0ull) | (live ? (1ull << 16) : 0ull);
        unsigned long long tot;
        const unsigned long long ex = block_exscan<NT>(val, sm.scratch, tot);
        if (isrow) { sm.hd[ex & 0xffff] = (short)tid; sm.drow[tid] = (short)(ex & 0xffff); }
        if (live) sm.ht[(ex >> 16) & 0xffff] = (short)t;
        R = (int)(tot & 0xffff); Cn = (int)((tot >> 16) & 0xffff);
        __syncthreads();
    }

    // dot product of raw detection row j and the smoothed embedding of slot sl, by one warp (generic width)
    auto emb_dot = [&](int j, int sl) -> double {
        const float4* a = reinterpret_cast<const float4*>(dfeat + (size_t)j * F);
        const float4* b = reinterpret_cast<const float4*>(pool + (size_t)sm.frow[sl] * F);
        double acc = 0.0;
        for (int i = lane; i < nv; i += 32) acc += hy_f4_dot(a[i], b[i]);
        return hy_warp_sum(acc);
    };

    // ---- first round: associate_4_points_with_score_with_reid ------------------------------------------
    if (R > 0 && Cn > 0) {
        // norms of the detection rows and of the live trackers' smoothed embeddings, one warp per row
        for (int q = warp; q < R + Cn; q += NW) {
            const bool isdet = q < R;
            const int idx = isdet ? sm.hd[q] : sm.ht[q - R];
            const float4* a = isdet ? reinterpret_cast<const float4*>(dfeat + (size_t)idx * F)
                                    : reinterpret_cast<const float4*>(pool + (size_t)sm.frow[idx] * F);
            double acc = 0.0;
            for (int i = lane; i < nv; i += 32) acc += hy_f4_sq(a[i]);
            acc = sqrt(hy_warp_sum(acc));
            if (lane == 0) { if (isdet) sm.dnorm[idx] = acc; else sm.tnorm[idx] = acc; }
        }
        __syncthreads();
        // embedding distances -> Cm: one warp per detection row, the row in registers as doubles (rows of up to 512 values)
        if (nv <= 128) {
            // Every warp holds TWO detection rows in registers; the smoothed rows of the trackers pass through shared memory
            // in tiles of eight (1-D TMA bulk copies, double buffered), so a tile fetched once from L2 serves all 2 NW detection rows of
            // the round - read per pair from L2 instead, the 2 KB rows made the step L2-bandwidth bound (83 MB per stream
            // and frame at config 4).
            for (int i = tid; i < 2 * 8 * 128; i += NT) (&sm.etile[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // padding stays zero
            if (tid == 0) {
                hy_mbar_init(&sm.ebar[0], 1); hy_mbar_init(&sm.ebar[1], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill is ordered before the bulk copies
            __syncthreads();
            const int ntile = (Cn + 7) >> 3;
            const uint32_t rowbytes = (uint32_t)nv * 16u;
            // one elected thread moves a tile: eight 1-D bulk copies (a tracker's row is contiguous, the rows are not), their
            // bytes counted on the buffer's mbarrier
            auto issue_tile = [&](int ti, int buf) {
                if (tid == 0) {
                    const int c0 = ti * 8, ncols = min(8, Cn - c0);
                    hy_mbar_expect_tx(&sm.ebar[buf], rowbytes * ncols);
                    for (int qq = 0; qq < ncols; ++qq)
                        hy_bulk_g2s(&sm.etile[buf][qq][0], pool + (size_t)sm.frow[sm.ht[c0 + qq]] * F, rowbytes, &sm.ebar[buf]);
                }
            };
            const int qi = lane >> 2;                                 // the value of a group this lane finalises
            int g = 0;                                                // tiles consumed so far: buffer g & 1, its use (g >> 1)
            for (int r0 = 0; r0 < R; r0 += 2 * NW) {
                const int rA = r0 + 2 * warp, rB = rA + 1;
                const bool actA = rA < R, actB = rB < R;              // warp-uniform; idle warps still synchronise
                const int jA = actA ? sm.hd[rA] : 0, jB = actB ? sm.hd[rB] : jA;
                double d0[16], d1[16];
                hy_load_row16(d0, reinterpret_cast<const float4*>(dfeat + (size_t)jA * F), nv, lane);
                hy_load_row16(d1, reinterpret_cast<const float4*>(dfeat + (size_t)jB * F), nv, lane);
                const int myr = (qi & 4) ? rB : rA;
                const bool myact = (lane & 3) == 0 && ((qi & 4) ? actB : actA);
                const double nj = sm.dnorm[(qi & 4) ? jB : jA];
                issue_tile(0, g & 1);
                for (int ti = 0; ti < ntile; ++ti) {
                    const int cur = g + ti, buf = cur & 1;
                    if (ti + 1 < ntile) issue_tile(ti + 1, buf ^ 1);  // its last readers passed the barrier below
                    if (!hy_mbar_wait(&sm.ebar[buf], (uint32_t)(cur >> 1) & 1u)) err |= B200_ERR_PIPELINE;
                    if (actA) {
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            const int c0 = ti * 8 + 4 * gq;
                            if (c0 < Cn) {                            // uniform
                                const double uv = hy_dot2x4(d0, d1, &sm.etile[buf][4 * gq], lane);
                                const int c = c0 + (qi & 3);
                                if (myact && c < Cn) Cm[(size_t)myr * TMAX + c] = hy_cosine(uv, sm.tnorm[sm.ht[c]], nj);
                            }
                        }
                    }
                    __syncthreads();                                  // the buffer may be refilled
                }
                g += ntile;
            }
        } else {
            for (int k = warp; k < R * Cn; k += NW) {
                const int r = k / Cn, c = k - r * Cn;
                const int j = sm.hd[r], sl = sm.ht[c];
                const double uv = emb_dot(j, sl);
                if (lane == 0) Cm[(size_t)r * TMAX + c] = hy_cosine(uv, sm.tnorm[sl], sm.dnorm[j]);
            }
        }
        __syncthreads();
        // ---- row reduction (lap_dense.cuh step 1) with on-demand costs, one warp per row -------------------------------
        if (tid < Cn) sm.tdeg[sm.ht[tid]] = !oc_regular_box(sm.tbox[0][sm.ht[tid]], sm.tbox[1][sm.ht[tid]], sm.tbox[2][sm.ht[tid]], sm.tbox[3][sm.ht[tid]]);
        if (tid < R) sm.ddeg[sm.hd[tid]] = !oc_regular_box(sm.dbox[0][sm.hd[tid]], sm.dbox[1][sm.hd[tid]], sm.dbox[2][sm.hd[tid]], sm.dbox[3][sm.hd[tid]]);
        __syncthreads();
        const HyCost<SM> cost{sm, Cm, TMAX, Cn, func, W, H, p.inertia, func <= 1};
        double bmax = 0.0;
        for (int r = warp; r < R; r += NW) {
            // pass 1: pairs without a usable bound (overlapping or irregular boxes; every pair for diou / ciou / centroid)
            // are evaluated; of the others the smallest bound is remembered
            double m = INF, lbmin = INF;
            int a = -1, lbc = -1;
            for (int c = lane; c < Cn; c += 32) {
                const double lb = cost.lower(r, c);
                if (lb == -INF) {
                    const double cst = cost(r, c);
                    if (cst < m) { m = cst; a = c; }
                } else if (lb < lbmin) { lbmin = lb; lbc = c; }
            }
            double mall = m;
#pragma unroll
            for (int d = 16; d; d >>= 1) mall = fmin(mall, __shfl_xor_sync(0xffffffffu, mall, d));
            if (mall == INF) {
                // a row without such a pair (a new object, a false positive): the pair with the smallest bound sets the bar
                int lbc_all = lbc;
                double lball = lbmin;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const double ol = __shfl_xor_sync(0xffffffffu, lball, d);
                    const int oc = __shfl_xor_sync(0xffffffffu, lbc_all, d);
                    if (ol < lball || (ol == lball && oc >= 0 && (lbc_all < 0 || oc < lbc_all))) { lball = ol; lbc_all = oc; }
                }
                if (lbc_all >= 0) mall = cost(r, lbc_all);                  // every lane: the same value
            }
            // pass 2: a bounded pair can only be the row minimum (or tie with it) if its bound does not exceed the bar
            for (int c = lane; c < Cn; c += 32) {
                const double lb = cost.lower(r, c);
                if (lb != -INF && lb <= mall) {
                    const double cst = cost(r, c);
                    if (cst < m || (cst == m && c < a)) { m = cst; a = c; }
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, m, d);
                const int oa = __shfl_xor_sync(0xffffffffu, a, d);
                if (om < m || (om == m && oa >= 0 && (a < 0 || oa < a))) { m = om; a = oa; }
            }
            if (lane == 0) { sm.u[r] = m; sm.claim[r] = a; }
            bmax = fmax(bmax, cost.bound(sm.hd[r]));
        }
        int d1 = 0, d2 = 0;
        block_max3<NT>(sm, bmax, d1, d2);
        if (tid == 0) atomicAdd(&p.stats[0], 1ull);
        {
            const DenseLap w = make_dense<NT>(sm);
            // every cost is below 1 (similarity) + the direction bound + 1.3 * 2 (cosine distance): any upper bound of
            // lapjv's 2 * (max cost + 1) gives the same assignment (lap_dense.cuh)
            const double lambda = 2.0 * (1.0 + bmax + 2.6 + 1e-3 + 1.0);
            dense_lap_init<NT>(w, R, Cn, lambda);
            dense_lap_augment<NT>(w, cost, R, Cn, lambda);
        }
        if (tid < R && sm.xr[tid] >= 0) sm.remb[tid] = Cm[(size_t)tid * TMAX + sm.xr[tid]];     // the matrix still holds the distances
        __syncthreads();
        // long-term-ReID correction (association.py:557-566): far in appearance AND below the score-penalised threshold
        if (tid < R) {
            const int c = sm.xr[tid];
            const int j = sm.hd[tid];
            if (c >= 0) {
                const int sl = sm.ht[c];
                const Box db = {sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]};
                const Box tb = {sm.tbox[0][sl], sm.tbox[1][sl], sm.tbox[2][sl], sm.tbox[3][sl]};
                const double sdif = fabs(xsub(sm.kscore[sl], sm.dconf[j]));
                if (sm.remb[tid] > 0.4 && xsub(oc_sim(func, db, tb, W, H), sdif) < thr) { sm.dstate[j] = DS_FREE1; atomicAdd(&p.stats[3], 1ull); }
                else { sm.dstate[j] = DS_MATCHED; sm.dmatch[j] = (short)sl; sm.tmatch[sl] = (short)j; sm.ema[sl] = 1; }
            }
        }
        __syncthreads();
    }

    // ---- observation-centric recovery round on the last observations (hybridsort.py:513-545) -------------------
    if (t < TMAX) sm.partner[t] = -1;
    __syncthreads();
    if (tid < R && Cn > 0 && sm.xr[tid] >= 0 && sm.dstate[sm.hd[tid]] == DS_FREE1) sm.partner[sm.ht[sm.xr[tid]]] = tid;   // slot -> row of its partner
    __syncthreads();
    bool ocr_ran = false;
    {
        // unmatched lists in the association's order (association.py:543-566): never matched first (ascending), then the
        // members of corrected matches in match order (= ascending detection row)
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
            for (int r = warp; r < nr; r += NW) {
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
                            sm.dstate[j] = DS_MATCHED;
                            sm.dmatch[j] = (short)sl; sm.tmatch[sl] = (short)j;       // update_feature=False: sm.ema stays 0
                        }
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---- Kalman update and bookkeeping, thread t = slot t (hybridsort.py:220-297) ---------------------------------
    int hits = 0, det_ind = 0, tid_id = 0;
    double conf = 0.0, cls = 0.0;
    HyKf k;
    if (live) {
#pragma unroll
        for (int c = 0; c < 9; ++c) k.x[c] = gf[(B200_HY_X + c) * TMAX + t];
        if (xadd(k.x[7], k.x[2]) <= 0.0) k.x[7] = xmul(k.x[7], 0.0);          // the motion step of the first phase, same operations
#pragma unroll
        for (int c = 0; c < 4; ++c) k.x[c] = xadd(k.x[c], k.x[c + 5]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            k.pp[i] = gf[(B200_HY_P + 3 * i + 0) * TMAX + t];
            k.pv[i] = gf[(B200_HY_P + 3 * i + 1) * TMAX + t];
            k.vv[i] = gf[(B200_HY_P + 3 * i + 2) * TMAX + t];
        }
        k.prr = gf[(B200_HY_P + 12) * TMAX + t];
        hy_predict_cov(k);
        hits = gi[B200_OCI_HITS * TMAX + t];
        det_ind = gi[B200_OCI_DET * TMAX + t];
        tid_id = gi[B200_OCI_ID * TMAX + t];
        conf = gf[B200_HY_CONF * TMAX + t];
        cls = gf[B200_HY_CLS * TMAX + t];
        const int j = sm.tmatch[t];
        if (j >= 0) {
            const int row = sm.drow[j];                                  // the FILTERED row index the reference indexes dets0 with
            const double bb[4] = {sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j]};
            const bool hasobs = fl & B200_OCF_HASOBS;
            const double lsum = hasobs ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
            if (lsum >= 0.0) {
                // every observation of ages age-1 .. age-delta_t adds the unit directions of its four corners (:232-247);
                // without one the last observation stands in
                double vsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                bool found = false;
                auto add_dir = [&](const double* pb) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ix = q < 2 ? 0 : 2, iy = (q & 1) ? 3 : 1;
                        const double dy = xsub(bb[iy], pb[iy]), dx = xsub(bb[ix], pb[ix]);
                        const double norm = xadd(sqrt(xadd(xmul(dy, dy), xmul(dx, dx))), 1e-6);
                        const double uy = xdiv(dy, norm), ux = xdiv(dx, norm);
                        vsum[2 * q] = found ? xadd(vsum[2 * q], uy) : uy;
                        vsum[2 * q + 1] = found ? xadd(vsum[2 * q + 1], ux) : ux;
                    }
                    found = true;
                };
                for (int i = 0; i < p.delta_t; ++i) {
                    const int a = age - i - 1;
                    if (a < 0) continue;
                    const int slot = a % 3;
                    if (gi[(B200_OCI_RINGAGE + slot) * TMAX + t] == a) {
                        double pb[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) pb[c] = gf[(B200_HY_RING + 4 * slot + c) * TMAX + t];
                        add_dir(pb);
                    }
                }
                if (!found) {
                    const double pb[4] = {sm.lbox[0][t], sm.lbox[1][t], sm.lbox[2][t], sm.lbox[3][t]};
                    add_dir(pb);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) gf[(B200_HY_VEL + c) * TMAX + t] = vsum[c];
            }
            conf = sm.dconf[j]; cls = det_cls(row); det_ind = row;
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_HY_LAST + c) * TMAX + t] = bb[c];
            const int rs = age % 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_HY_RING + 4 * rs + c) * TMAX + t] = bb[c];
            gi[(B200_OCI_RINGAGE + rs) * TMAX + t] = age;
            double z4[4];
            oc_box_to_z(bb[0], bb[1], bb[2], bb[3], z4);
            const double z[5] = {z4[0], z4[1], z4[2], conf, z4[3]};      // convert_bbox_to_z: [x, y, s, score, r]
            bool virt = false;
            double vz[5];
            if (!(fl & B200_OCF_OBSERVED) && (fl & B200_OCF_SAVED)) {   // unfreeze: observation-centric re-update
#pragma unroll
                for (int c = 0; c < 9; ++c) k.x[c] = gf[(B200_HY_SX + c) * TMAX + t];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    k.pp[i] = gf[(B200_HY_SP + 3 * i + 0) * TMAX + t];
                    k.pv[i] = gf[(B200_HY_SP + 3 * i + 1) * TMAX + t];
                    k.vv[i] = gf[(B200_HY_SP + 3 * i + 2) * TMAX + t];
                }
                k.prr = gf[(B200_HY_SP + 12) * TMAX + t];
                double lz[5];
#pragma unroll
                for (int c = 0; c < 5; ++c) lz[c] = gf[(B200_HY_LASTZ + c) * TMAX + t];
                virt = hy_virtual_trajectory(k, lz, z, tsu, vz);       // gap = index2 - index1 of history_obs
                fl &= ~B200_OCF_SAVED;
                atomicAdd(&p.stats[2], 1ull);
            }
            fl |= B200_OCF_OBSERVED | B200_OCF_HASOBS;
            hy_correct(k, z);                                            // the real measurement on top
#pragma unroll
            for (int c = 0; c < 5; ++c) gf[(B200_HY_LASTZ + c) * TMAX + t] = virt ? vz[c] : z[c];
            tsu = 0; hits += 1; streak += 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.lbox[c][t] = bb[c];
        } else {                                                        // kf.update(None), hybridsort_kf.py:467-479
            if (fl & B200_OCF_OBSERVED) {
#pragma unroll
                for (int c = 0; c < 9; ++c) gf[(B200_HY_SX + c) * TMAX + t] = k.x[c];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    gf[(B200_HY_SP + 3 * i + 0) * TMAX + t] = k.pp[i];
                    gf[(B200_HY_SP + 3 * i + 1) * TMAX + t] = k.pv[i];
                    gf[(B200_HY_SP + 3 * i + 2) * TMAX + t] = k.vv[i];
                }
                gf[(B200_HY_SP + 12) * TMAX + t] = k.prr;
                fl |= B200_OCF_SAVED;
            }
            fl &= ~B200_OCF_OBSERVED;
        }
    }
    __syncthreads();

    // ---- update_features (hybridsort.py:188-205) of every tracker matched in the first round, one warp each ---------
    for (int q = warp; q < n0; q += NW) {
        if (!sm.alive[q] || !sm.ema[q]) continue;
        const int j = sm.tmatch[q];
        float4* trk = reinterpret_cast<float4*>(pool + (size_t)sm.frow[q] * F);
        const float4* det = reinterpret_cast<const float4*>(dfeat + (size_t)j * F);
        // feat /= |feat| (float32 norm), blend, renormalise
        double acc = 0.0;
        for (int i = lane; i < nv; i += 32) acc += hy_f4_sq(det[i]);
        const RowDiv n1 = row_div(hy_norm_f32(hy_warp_sum(acc)));
        acc = 0.0;
        for (int i = lane; i < nv; i += 32) acc += hy_f4_sq(hy_blend(trk[i], hy_f4_div(det[i], n1)));
        const RowDiv n2 = row_div(hy_norm_f32(hy_warp_sum(acc)));
        for (int i = lane; i < nv; i += 32) trk[i] = hy_f4_div(hy_blend(trk[i], hy_f4_div(det[i], n1)), n2);
    }

    // ---- new trackers (hybridsort.py:550-553) and the reversed output scan (:554-570) ------------------------------
    const int ds = tid < DMAX ? sm.dstate[tid] : DS_NONE;
    const bool newborn = ds == DS_FREE0 || ds == DS_FREE1;
    bool emit_old = false, emit_new = false, die = false;
    bool box_from_obs = false;
    if (live) {
        emit_old = tsu < 1 && (streak >= p.min_hits || frame <= p.min_hits);
        die = tsu > p.max_age;
        const double lsum = (fl & B200_OCF_HASOBS) ? xadd(xadd(xadd(xadd(sm.lbox[0][t], sm.lbox[1][t]), sm.lbox[2][t]), sm.lbox[3][t]), conf) : -5.0;
        box_from_obs = !(lsum < 0.0);
        if (!die) sm.rowused[sm.frow[t]] = 1;
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
    double* gout = p.out + (size_t)s * p.max_tracks * 8;
    const int out_cap = p.max_tracks;
    if (n0 + n_new > tcap) err |= B200_ERR_TRACK_OVERFLOW;
    // result row [x1, y1, x2, y2, id, conf, cls, score of input row `row`] (the reference's det_ind column, see the header)
    auto write_row = [&](int orow, const Box& b, int id, double cf, double cl, double last) {
        double2* o = reinterpret_cast<double2*>(gout + (size_t)orow * 8);
        o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
        o[2] = make_double2((double)id, cf); o[3] = make_double2(cl, last);
    };
    if (live) {
#pragma unroll
        for (int c = 0; c < 9; ++c) gf[(B200_HY_X + c) * TMAX + t] = k.x[c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            gf[(B200_HY_P + 3 * i + 0) * TMAX + t] = k.pp[i];
            gf[(B200_HY_P + 3 * i + 1) * TMAX + t] = k.pv[i];
            gf[(B200_HY_P + 3 * i + 2) * TMAX + t] = k.vv[i];
        }
        gf[(B200_HY_P + 12) * TMAX + t] = k.prr;
        gf[B200_HY_CONF * TMAX + t] = conf;
        gf[B200_HY_CLS * TMAX + t] = cls;
        gi[B200_OCI_AGE * TMAX + t] = age;
        gi[B200_OCI_TSU * TMAX + t] = tsu;
        gi[B200_OCI_HITS * TMAX + t] = hits;
        gi[B200_OCI_STREAK * TMAX + t] = streak;
        gi[B200_OCI_DET * TMAX + t] = det_ind;
        gi[B200_OCI_FLAGS * TMAX + t] = die ? (fl & ~OCF_ALIVE) : fl;
        if (emit_old) {
            const int orow = E_new + (E_old - 1 - (int)(ex & 1023));
            if (orow < out_cap) {
                Box b;
                if (box_from_obs) { b.x1 = sm.lbox[0][t]; b.y1 = sm.lbox[1][t]; b.x2 = sm.lbox[2][t]; b.y2 = sm.lbox[3][t]; }
                else b = oc_x_to_box(k.x[0], k.x[1], k.x[2], k.x[4]);
                write_row(orow, b, tid_id + 1, conf, cls, (det_ind >= 0 && det_ind < nd) ? sm.dconf[det_ind] : 0.0);
            }
        }
    } else if (t < n0 && (fl & OCF_ALIVE) == 0 && t < TMAX) {
        if (gi[B200_OCI_FLAGS * TMAX + t] & OCF_ALIVE) gi[B200_OCI_FLAGS * TMAX + t] = fl;     // NaN-purged this frame
    }
    int nb_row = -1, nb_det = -1;
    if (newborn) {
        const int j = tid, row = sm.drow[j];
        int order;                      // position in the creation order
        if (ocr_ran) order = (int)((ex >> 20) & 1023) + (int)((ex >> 30) & 1023);
        else order = ds == DS_FREE0 ? (int)((ex >> 20) & 1023) : n_free0 + (int)((ex >> 30) & 1023);
        const int dst = n0 + order;
        const int id = id0 + order;
        double z4[4];
        oc_box_to_z(sm.dbox[0][j], sm.dbox[1][j], sm.dbox[2][j], sm.dbox[3][j], z4);
        const double cl = det_cls(row);
        if (dst < tcap) {
            const double x0[9] = {z4[0], z4[1], z4[2], sm.dconf[j], z4[3], 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int c = 0; c < 9; ++c) gf[(B200_HY_X + c) * TMAX + dst] = x0[c];
#pragma unroll
            for (int c = 0; c < 4; ++c) gf[(B200_HY_LAST + c) * TMAX + dst] = -1.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gf[(B200_HY_P + 3 * i + 0) * TMAX + dst] = 10.0;
                gf[(B200_HY_P + 3 * i + 1) * TMAX + dst] = 0.0;
                gf[(B200_HY_P + 3 * i + 2) * TMAX + dst] = 1e4;
            }
            gf[(B200_HY_P + 12) * TMAX + dst] = 10.0;
            gf[B200_HY_CONF * TMAX + dst] = sm.dconf[j];
            gf[B200_HY_CLS * TMAX + dst] = cl;
#pragma unroll
            for (int c = 0; c < 8; ++c) gf[(B200_HY_VEL + c) * TMAX + dst] = 0.0;
            gi[B200_OCI_ID * TMAX + dst] = id;
            gi[B200_OCI_AGE * TMAX + dst] = 0;
            gi[B200_OCI_TSU * TMAX + dst] = 0;
            gi[B200_OCI_HITS * TMAX + dst] = 0;
            gi[B200_OCI_STREAK * TMAX + dst] = 0;
            gi[B200_OCI_DET * TMAX + dst] = row;
            gi[(B200_OCI_RINGAGE + 0) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 1) * TMAX + dst] = -1; gi[(B200_OCI_RINGAGE + 2) * TMAX + dst] = -1;
            gi[B200_OCI_FLAGS * TMAX + dst] = OCF_ALIVE;
            nb_row = sm.freelist[order];
            nb_det = j;
            gi[B200_HYI_FROW * TMAX + dst] = nb_row;
        }
        if (emit_new) {
            const int orow = n_new - 1 - order;
            if (orow < out_cap) write_row(orow, oc_x_to_box(z4[0], z4[1], z4[2], z4[3]), id + 1, sm.dconf[j], cl, sm.dconf[row]);
        }
    }
    // a new tracker's smoothed embedding is its detection's row normalised twice (hybridsort.py:189, :205); one warp each
    __syncthreads();
    if (tid < DMAX) { sm.nbrow[tid] = (short)nb_row; sm.nbdet[tid] = (short)nb_det; }
    __syncthreads();
    for (int q = warp; q < nd; q += NW) {
        const int prow = sm.nbrow[q], j = sm.nbdet[q];
        if (prow < 0) continue;
        float4* e = reinterpret_cast<float4*>(pool + (size_t)prow * F);
        const float4* d = reinterpret_cast<const float4*>(dfeat + (size_t)j * F);
        double acc = 0.0;
        for (int i = lane; i < nv; i += 32) acc += hy_f4_sq(d[i]);
        const RowDiv n1 = row_div(hy_norm_f32(hy_warp_sum(acc)));
        acc = 0.0;
        for (int i = lane; i < nv; i += 32) acc += hy_f4_sq(hy_f4_div(d[i], n1));
        const RowDiv n2 = row_div(hy_norm_f32(hy_warp_sum(acc)));
        for (int i = lane; i < nv; i += 32) e[i] = hy_f4_div(hy_f4_div(d[i], n1), n2);
    }
    const int n1c = min(n0 + n_new, tcap);
    const int alive_after = n_keep + min(n_new, tcap - n0 > 0 ? tcap - n0 : 0);
    if (tid == 0) {
        counts[0] = n1c;
        counts[1] = alive_after;
        counts[2] = id0 + n_new;
        counts[3] = frame;
        p.nout[s] = min(E_old + E_new, out_cap);
        p.track_updates[s] += (unsigned long long)Cn;
    }
    if (err) { atomicOr(p.err, err); if (p.err_out) atomicOr(p.err_out, err); }
}

template <int TMAX, int DMAX>
cudaError_t launch_hy_kernel(const StepParams& p, cudaStream_t stream) {
    auto kern = hybridsort_step_kernel<TMAX, TMAX, DMAX>;
    const size_t smem = sizeof(HySmem<TMAX, DMAX>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.n_streams, TMAX, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t hybridsort_step_smem(int variant) {
    switch (variant) {
        case 0: return sizeof(HySmem<64, 64>);
        case 1: return sizeof(HySmem<128, 128>);
        case 2: return sizeof(HySmem<224, 224>);
        case 3: return sizeof(HySmem<256, 256>);
        case 4: return sizeof(HySmem<512, 512>);
    }
    return 0;
}

cudaError_t launch_hybridsort_step(const StepParams& p, int variant, cudaStream_t stream) {
    switch (variant) {
        case 0: return launch_hy_kernel<64, 64>(p, stream);
        case 1: return launch_hy_kernel<128, 128>(p, stream);
        case 2: return launch_hy_kernel<224, 224>(p, stream);
        case 3: return launch_hy_kernel<256, 256>(p, stream);
        case 4: return launch_hy_kernel<512, 512>(p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200
