// Batched StrongSORT frame step: every stream of a context advances one frame in a fixed sequence of launches, nothing
// returns to the host in between.
//
//   ss_pre_kernel      one CTA per stream: Track.camera_update on every listed track, age / time_since_update counters,
//                      gallery row counts of the confirmed tracks, detections -> tlwh / xyah
//   kf_predict         (ops.cu operator kernel, all slots of all streams)                        strongsort_kf.py:88-122
//   gallery_cost       (emb_gemm.cu: bf16 tcgen05 pre-filter + exact float32 values, one CTA per (slot, stream)):
//                      min cosine distance to the stored features of every confirmed track        matching.py:247-378
//   gate_cost          (ops.cu: Mahalanobis gate + motion fusion, in place)                       linear_assignment.py:144-200
//   ss_match_kernel    one CTA per stream: min_cost_matching on the confirmed tracks (scipy's linear_sum_assignment on the
//                      clipped matrix, bit-faithful on ties: lsa_scipy.cuh), the unmatched set in CPython's set order
//                      (pyset.cuh), the IoU round on unconfirmed + just-missed tracks               tracker.py:104-155
//   kf_update (masked) (ops.cu: the matched slots of all streams, confidence-scaled noise)        strongsort_kf.py:157-189
//   ss_post_kernel     one CTA per stream: Track.update (feature smoothing, confirmation), mark_missed, new tracks,
//                      the new track list, gallery ring append (partial_fit), result rows          tracker.py:73-102, track.py:152-185
//
// Slots never move: `order` holds the reference's self.tracks list (new tracks are appended, deleted ones drop out), the
// gallery ring and the smoothed feature of a track live at its slot.  Replaces StrongSORT.update
// (boxmot/trackers/strongsort/strong_sort.py:43-99) for many streams at once; ids, confirmation / deletion, gallery sizes
// equal the live reference's goldens frame by frame (tests/test_strongsort_gpu.py).
#include "strongsort_step.h"

#include <cuda_bf16.h>

#include "../../include/b200track.h"
#include "api_util.h"
#include "boxes.cuh"
#include "kf.cuh"
#include "layout.h"
#include "lsa_scipy.cuh"
#include "pyset.cuh"

namespace b200 {
namespace {

constexpr int NT = LSA_NT;               // 256 threads: one per list position / detection
constexpr int CAP = 256;                 // slots / detections per stream
constexpr double INFTY_COST = 1e5;       // linear_assignment.py:10

struct MatchSmem {
    unsigned long long scratch[40];
    short slot_of[CAP], tsu_of[CAP];
    short conf_k[CAP], unconf_k[CAP], matchedB[CAP], ut_a[CAP], cand[CAP];
    short ud1[CAP], ud2[CAP], rmatch[CAP];
    unsigned char colused[CAP], inb[CAP];
    short pybufs[6 * 1024];
    int nUT, nK;
};
constexpr size_t MATCH_LSA_OFF = (sizeof(MatchSmem) + 15) & ~size_t(15);      // lsa_work_bytes(CAP, CAP) follow

// stable compaction of one value per thread; the result is visible when the call returns
__device__ __forceinline__ int compact(bool flag, short value, short* out, unsigned long long* scratch) {
    unsigned long long tot;
    const unsigned long long ex = block_exscan<NT>(flag ? 1ull : 0ull, scratch, tot);
    if (flag) out[ex] = value;
    __syncthreads();
    return (int)tot;
}

// Track.to_tlbr (track.py:101-127): xyah mean -> corners, the reference's operation order (x / 2 == x * 0.5 exactly)
__device__ __forceinline__ Box mean_tlbr(const double* m) {
    Box b;
    const double w = xmul(m[2], m[3]), h = m[3];
    b.x1 = xsub(m[0], xmul(w, 0.5)); b.y1 = xsub(m[1], xmul(h, 0.5));
    b.x2 = xadd(b.x1, w); b.y2 = xadd(b.y1, h);
    return b;
}

__global__ void __launch_bounds__(NT) ss_pre_kernel(const SSParams p, const double* __restrict__ dets, const int* __restrict__ ndets,
                                                    const double* __restrict__ warps, int* err_step) {
    const int s = blockIdx.x, tid = threadIdx.x;
    const int T = p.T, D = p.D;
    int* ti = p.ti + (size_t)s * SS_NI * T;
    const int n = p.counts[4 * s];
    const double* wp = warps ? warps + 6 * s : nullptr;
    for (int k = tid; k < n; k += NT) {
        const int slot = p.order[(size_t)s * T + k];
        double* m = p.mean + ((size_t)s * T + slot) * 8;
        // Track.camera_update (track.py:129-138): with the identity warp still not an exact no-op in floating point
        Box b = mean_tlbr(m);
        if (wp) {
            const double a = wp[0], bb = wp[1], c = wp[2], d = wp[3], e = wp[4], f = wp[5];
            const double nx1 = xadd(xadd(xmul(a, b.x1), xmul(bb, b.y1)), c), ny1 = xadd(xadd(xmul(d, b.x1), xmul(e, b.y1)), f);
            const double nx2 = xadd(xadd(xmul(a, b.x2), xmul(bb, b.y2)), c), ny2 = xadd(xadd(xmul(d, b.x2), xmul(e, b.y2)), f);
            b.x1 = nx1; b.y1 = ny1; b.x2 = nx2; b.y2 = ny2;
        }
        const double w = xsub(b.x2, b.x1), h = xsub(b.y2, b.y1);
        m[0] = xadd(b.x1, xmul(w, 0.5)); m[1] = xadd(b.y1, xmul(h, 0.5)); m[2] = xdiv(w, h); m[3] = h;
        ti[SSI_AGE * T + slot] += 1;                   // Track.predict (track.py:144-150)
        ti[SSI_TSU * T + slot] += 1;
    }
    int grows = 0;
    for (int t = tid; t < T; t += NT) {
        const int st = ti[SSI_STATE * T + t];
        const int g = st == SS_CONFIRMED ? min(ti[SSI_APPENDED * T + t], p.budget) : 0;
        p.gcount[(size_t)s * T + t] = g;
        p.match[(size_t)s * T + t] = -1;
        grows += g;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) grows += __shfl_xor_sync(0xffffffffu, grows, d);
    if ((tid & 31) == 0 && grows && p.stats) atomicAdd(&p.stats[3], (unsigned long long)grows);
    int nd = ndets[s];
    if (nd > D) { nd = D; if (tid == 0) { atomicOr(p.err, B200_ERR_DET_OVERFLOW); if (err_step) atomicOr(err_step, B200_ERR_DET_OVERFLOW); } }
    for (int j = tid; j < D; j += NT) {
        double x = 0.0, y = 0.0, w = 0.0, h = 0.0, c = 0.0;
        if (j < nd) {
            const double* r = dets + ((size_t)s * D + j) * 6;
            x = r[0]; y = r[1]; w = xsub(r[2], r[0]); h = xsub(r[3], r[1]); c = r[4];      // strong_sort.py:67-75 xyxy -> tlwh
        }
        double* tl = p.tlwh + ((size_t)s * D + j) * 4;
        tl[0] = x; tl[1] = y; tl[2] = w; tl[3] = h;
        double* z = p.meas + ((size_t)s * D + j) * 4;                                       // Detection.to_xyah (detection.py:34-41)
        z[0] = xadd(x, xmul(w, 0.5)); z[1] = xadd(y, xmul(h, 0.5)); z[2] = j < nd ? xdiv(w, h) : 0.0; z[3] = h;
        p.dconf[(size_t)s * D + j] = c;
    }
}

__global__ void __launch_bounds__(NT) ss_match_kernel(const SSParams p, const int* __restrict__ ndets) {
    extern __shared__ __align__(16) unsigned char raw[];
    MatchSmem& sm = *reinterpret_cast<MatchSmem*>(raw);
    const int s = blockIdx.x, tid = threadIdx.x;
    const int T = p.T, D = p.D;
    const int* ti = p.ti + (size_t)s * SS_NI * T;
    const int n = min(p.counts[4 * s], T);
    const int nd = max(0, min(ndets[s], D));
    int* match = p.match + (size_t)s * T;
    const double* costmat = p.cost + (size_t)s * T * D;
    // ---- the track list: confirmed / unconfirmed positions in list order (tracker.py:121-123)
    int state = SS_FREE;
    if (tid < n) {
        const int slot = p.order[(size_t)s * T + tid];
        sm.slot_of[tid] = (short)slot;
        sm.tsu_of[tid] = (short)min(ti[SSI_TSU * T + slot], 32000);
        state = ti[SSI_STATE * T + slot];
    }
    const int nC = compact(tid < n && state == SS_CONFIRMED, (short)tid, sm.conf_k, sm.scratch);
    const int nU = compact(tid < n && state != SS_CONFIRMED, (short)tid, sm.unconf_k, sm.scratch);
    // ---- min_cost_matching(gated_metric, max_dist, confirmed, all detections) (linear_assignment.py:14-79)
    int nB = 0, nUD1 = 0;
    if (nC > 0 && nd > 0) {
        const double lim = p.max_dist, clipped = p.max_dist + 1e-5;
        auto orig = [&](int r, int c) {
            const double v = costmat[(size_t)sm.slot_of[sm.conf_k[r]] * D + c];
            return v > lim ? clipped : v;
        };
        const bool tr = nd < nC;                       // scipy solves the transposed problem when there are more rows
        const int nr = tr ? nd : nC, nc = tr ? nC : nd;
        const LsaWork w = lsa_carve(raw + MATCH_LSA_OFF, nr, nc);
        const bool ok = tr ? lsa_scipy_solve(nr, nc, [&](int i, int j) { return orig(j, i); }, w)
                           : lsa_scipy_solve(nr, nc, [&](int i, int j) { return orig(i, j); }, w);
        if (!ok && tid == 0) atomicOr(p.err, B200_ERR_LSA);
        if (tid < nC) sm.rmatch[tid] = -1;
        if (tid < nd) sm.colused[tid] = 0;
        __syncthreads();
        if (ok && tid < nr) {
            const int o = w.col4row[tid];
            if (tr) { sm.rmatch[o] = (short)tid; sm.colused[tid] = 1; }
            else { sm.rmatch[tid] = (short)o; sm.colused[o] = 1; }
        }
        __syncthreads();
        bool acc = false, rej = false;
        int j = -1;
        if (tid < nC) {
            j = sm.rmatch[tid];
            if (j >= 0) { rej = orig(tid, j) > lim; acc = !rej; }
            if (acc) match[sm.slot_of[sm.conf_k[tid]]] = j;
        }
        nB = compact(acc, tid < nC ? sm.conf_k[tid] : (short)0, sm.matchedB, sm.scratch);
        const int n1 = compact(tid < nd && !sm.colused[tid], (short)tid, sm.ud1, sm.scratch);
        const int n2 = compact(rej, (short)j, sm.ud1 + n1, sm.scratch);
        nUD1 = n1 + n2;
    } else {
        if (tid < nd) sm.ud1[tid] = (short)tid;
        nUD1 = nd;
        __syncthreads();
    }
    // ---- unmatched_tracks_a = list(set(track_indices) - set(k for k, _ in matches)) (linear_assignment.py:141), then the
    // candidates of the IoU round: unconfirmed + [k in unmatched_a if time_since_update == 1] (tracker.py:136-142)
    if (tid == 0) {
        const int nUT = pyset_difference_order(sm.conf_k, nC, sm.matchedB, nB, n, sm.pybufs, 1024, sm.inb, sm.ut_a);
        int nK = 0;
        for (int k = 0; k < nU; ++k) sm.cand[nK++] = sm.unconf_k[k];
        for (int k = 0; k < nUT; ++k)
            if (sm.tsu_of[sm.ut_a[k]] == 1) sm.cand[nK++] = sm.ut_a[k];
        sm.nUT = nUT; sm.nK = nK;
    }
    __syncthreads();
    const int nK = sm.nK;
    // ---- min_cost_matching(iou_cost, max_iou_dist, candidates, unmatched detections) (iou_matching.py:50-87)
    int nUD2 = nUD1;
    short* udf = sm.ud1;
    if (nK > 0 && nUD1 > 0) {
        double* M = p.iou + (size_t)s * T * D;
        const double lim = p.max_iou_dist, clipped = p.max_iou_dist + 1e-5;
        for (int idx = tid; idx < nK * nUD1; idx += NT) {
            const int r = idx / nUD1, c = idx - r * nUD1;
            const int k = sm.cand[r];
            double v = INFTY_COST;
            if (!(sm.tsu_of[k] > 1)) {
                const Box a = mean_tlbr(p.mean + ((size_t)s * T + sm.slot_of[k]) * 8);
                const double* tl = p.tlwh + ((size_t)s * D + sm.ud1[c]) * 4;
                Box b;
                b.x1 = tl[0]; b.y1 = tl[1]; b.x2 = xadd(tl[0], tl[2]); b.y2 = xadd(tl[1], tl[3]);
                v = xsub(1.0, box_iou(a, b));
            }
            M[idx] = v > lim ? clipped : v;
        }
        __syncthreads();
        const int ldm = nUD1;
        const bool tr = nUD1 < nK;
        const int nr = tr ? nUD1 : nK, nc = tr ? nK : nUD1;
        const LsaWork w = lsa_carve(raw + MATCH_LSA_OFF, nr, nc);
        const bool ok = tr ? lsa_scipy_solve(nr, nc, [&](int i, int j) { return M[(size_t)j * ldm + i]; }, w)
                           : lsa_scipy_solve(nr, nc, [&](int i, int j) { return M[(size_t)i * ldm + j]; }, w);
        if (!ok && tid == 0) atomicOr(p.err, B200_ERR_LSA);
        if (tid < nK) sm.rmatch[tid] = -1;
        if (tid < nUD1) sm.colused[tid] = 0;
        __syncthreads();
        if (ok && tid < nr) {
            const int o = w.col4row[tid];
            if (tr) { sm.rmatch[o] = (short)tid; sm.colused[tid] = 1; }
            else { sm.rmatch[tid] = (short)o; sm.colused[o] = 1; }
        }
        __syncthreads();
        bool rej = false;
        int j = -1;
        if (tid < nK) {
            j = sm.rmatch[tid];
            if (j >= 0) {
                rej = M[(size_t)tid * ldm + j] > lim;
                if (!rej) match[sm.slot_of[sm.cand[tid]]] = sm.ud1[j];
            }
        }
        const int n1 = compact(tid < nUD1 && !sm.colused[tid], tid < nUD1 ? sm.ud1[tid] : (short)0, sm.ud2, sm.scratch);
        const int n2 = compact(rej, j >= 0 ? sm.ud1[j] : (short)0, sm.ud2 + n1, sm.scratch);
        nUD2 = n1 + n2;
        udf = sm.ud2;
    }
    if (tid < nUD2) p.ud[(size_t)s * D + tid] = udf[tid];
    if (tid == 0) p.nud[s] = nUD2;
}

// ---- float32 feature arithmetic, one warp per row (norms accumulated in double: the reference's come from BLAS) ----
__device__ __forceinline__ double warp_sum_d(double x) {
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

struct PostSmem {
    unsigned long long scratch[40];
    short slot_of[CAP], mdet[CAP], keep_slot[CAP], freelist[CAP], born_slot[CAP];
    unsigned char st_after[CAP], slot_state[CAP];
};

__global__ void __launch_bounds__(NT) ss_post_kernel(const SSParams p, const double* __restrict__ dets, const int* __restrict__ ndets,
                                                     const float* __restrict__ feats, double* __restrict__ out, int* __restrict__ nout,
                                                     int* err_step) {
    __shared__ PostSmem sm;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.T, D = p.D, F = p.F;
    int* ti = p.ti + (size_t)s * SS_NI * T;
    const int n = min(p.counts[4 * s], T);
    const int next_id = p.counts[4 * s + 1];
    const double* drow = dets + (size_t)s * D * 6;
    // ---- Track.update / mark_missed (track.py:152-185)
    if (tid < T) sm.slot_state[tid] = (unsigned char)ti[SSI_STATE * T + tid];
    __syncthreads();
    int slot = 0, st = SS_FREE, tsu = 0, j = -1;
    if (tid < n) {
        slot = p.order[(size_t)s * T + tid];
        sm.slot_of[tid] = (short)slot;
        st = sm.slot_state[slot];
        tsu = ti[SSI_TSU * T + slot];
        j = p.match[(size_t)s * T + slot];
        if (j >= 0) {
            p.conf[(size_t)s * T + slot] = drow[j * 6 + 4];
            p.cls[(size_t)s * T + slot] = drow[j * 6 + 5];
            ti[SSI_DET * T + slot] = j;
            const int hits = ti[SSI_HITS * T + slot] + 1;
            ti[SSI_HITS * T + slot] = hits;
            tsu = 0;
            ti[SSI_TSU * T + slot] = 0;
            if (st == SS_TENTATIVE && hits >= p.n_init) st = SS_CONFIRMED;
        } else if (st == SS_TENTATIVE || tsu > p.max_age) st = SS_FREE;       // deleted
        ti[SSI_STATE * T + slot] = st;
        if (st == SS_FREE) ti[SSI_APPENDED * T + slot] = 0;
        sm.mdet[tid] = (short)j;
        sm.st_after[tid] = (unsigned char)st;
    }
    __syncthreads();
    if (tid < n) sm.slot_state[slot] = (unsigned char)st;
    __syncthreads();
    // ---- the new list: survivors in list order, then the new tracks (tracker.py:96-100)
    const int nKeep = compact(tid < n && st != SS_FREE, (short)slot, sm.keep_slot, sm.scratch);
    const int nFree = compact(tid < T && sm.slot_state[tid] == SS_FREE, (short)tid, sm.freelist, sm.scratch);
    const int nud = p.nud[s];
    const int nBorn = min(nud, nFree);
    if (nud > nFree && tid == 0) { atomicOr(p.err, B200_ERR_TRACK_OVERFLOW); if (err_step) atomicOr(err_step, B200_ERR_TRACK_OVERFLOW); }
    if (tid < nBorn) {
        // Tracker._initiate_track -> Track.__init__ (track.py:72-99), KalmanFilter.initiate (strongsort_kf.py:55-86)
        const int d = p.ud[(size_t)s * D + tid];
        const int ns = sm.freelist[tid];
        sm.born_slot[tid] = (short)ns;
        const double* z = p.meas + ((size_t)s * D + d) * 4;
        double* m = p.mean + ((size_t)s * T + ns) * 8;
        double* P = p.cov + ((size_t)s * T + ns) * 64;
        const double h = z[3];
        const double sp = xmul(2 * KF_W_POS, h), sv = xmul(10 * KF_W_VEL, h);
        const double sd[8] = {sp, sp, 1e-2, sp, sv, sv, 1e-5, sv};
        for (int i = 0; i < 64; ++i) P[i] = 0.0;
        for (int i = 0; i < 4; ++i) { m[i] = z[i]; m[i + 4] = 0.0; }
        for (int i = 0; i < 8; ++i) P[i * 9] = xmul(sd[i], sd[i]);
        p.conf[(size_t)s * T + ns] = drow[d * 6 + 4];
        p.cls[(size_t)s * T + ns] = drow[d * 6 + 5];
        ti[SSI_ID * T + ns] = next_id + tid;
        ti[SSI_STATE * T + ns] = SS_TENTATIVE;
        ti[SSI_HITS * T + ns] = 1; ti[SSI_AGE * T + ns] = 1; ti[SSI_TSU * T + ns] = 0;
        ti[SSI_DET * T + ns] = d; ti[SSI_APPENDED * T + ns] = 0;
    }
    __syncthreads();
    // ---- features: smoothing of the matched tracks (track.py:166-172), gallery append of the confirmed ones
    // (NearestNeighborDistanceMetric.partial_fit, matching.py:343-358), first feature of the new tracks (track.py:88-90)
    const int nv = F >> 2;
    for (int k = warp; k < n; k += NT / 32) {
        if (sm.st_after[k] == SS_FREE) continue;
        const int sl = sm.slot_of[k];
        float* a = p.feat + ((size_t)s * T + sl) * F;
        const int dj = sm.mdet[k];
        if (dj >= 0) {
            // feature /= |feature|; smooth = alpha * smooth + beta * feature; smooth /= |smooth| - float32 like numpy.  The row
            // (F <= 512: four float4 per lane) stays in registers across the passes, and the divisions of a row share one
            // refined reciprocal (common.cuh: the same quotients as __fdiv_rn).
            const float4* b4 = reinterpret_cast<const float4*>(feats + ((size_t)s * D + dj) * F);
            float4* a4 = reinterpret_cast<float4*>(a);
            if (nv <= 128) {
                float4 bv[4], av[4];
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = lane + 32 * q < nv;
                    bv[q] = in ? b4[lane + 32 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    av[q] = in ? a4[lane + 32 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc += (double)bv[q].x * bv[q].x + (double)bv[q].y * bv[q].y + (double)bv[q].z * bv[q].z + (double)bv[q].w * bv[q].w;
                }
                const RowDiv nb = row_div(sqrtf((float)warp_sum_d(acc)));
                acc = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4& v = av[q];
                    v.x = __fadd_rn(__fmul_rn(p.ema_alpha, v.x), __fmul_rn(p.ema_beta, fdiv_row(bv[q].x, nb)));
                    v.y = __fadd_rn(__fmul_rn(p.ema_alpha, v.y), __fmul_rn(p.ema_beta, fdiv_row(bv[q].y, nb)));
                    v.z = __fadd_rn(__fmul_rn(p.ema_alpha, v.z), __fmul_rn(p.ema_beta, fdiv_row(bv[q].z, nb)));
                    v.w = __fadd_rn(__fmul_rn(p.ema_alpha, v.w), __fmul_rn(p.ema_beta, fdiv_row(bv[q].w, nb)));
                    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
                }
                const RowDiv ns = row_div(sqrtf((float)warp_sum_d(acc)));
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (lane + 32 * q < nv)
                        a4[lane + 32 * q] = make_float4(fdiv_row(av[q].x, ns), fdiv_row(av[q].y, ns), fdiv_row(av[q].z, ns), fdiv_row(av[q].w, ns));
            } else {
                const float* b = feats + ((size_t)s * D + dj) * F;
                double acc = 0.0;
                for (int i = lane; i < F; i += 32) acc += (double)b[i] * b[i];
                const RowDiv nb = row_div(sqrtf((float)warp_sum_d(acc)));
                acc = 0.0;
                for (int i = lane; i < F; i += 32) {
                    const float v = __fadd_rn(__fmul_rn(p.ema_alpha, a[i]), __fmul_rn(p.ema_beta, fdiv_row(b[i], nb)));
                    a[i] = v;
                    acc += (double)v * v;
                }
                const RowDiv ns = row_div(sqrtf((float)warp_sum_d(acc)));
                for (int i = lane; i < F; i += 32) a[i] = fdiv_row(a[i], ns);
            }
            __syncwarp();
        }
        if (sm.st_after[k] == SS_CONFIRMED) {
            const int appended = ti[SSI_APPENDED * T + sl];
            const size_t dst = (((size_t)s * T + sl) * p.budget + (appended % p.budget)) * F;
            const float4* src = reinterpret_cast<const float4*>(a);
            float4* o32 = reinterpret_cast<float4*>(p.gal32 + dst);
            __nv_bfloat162* o16 = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.gal16) + dst);
            float acc = 0.f;
            for (int i = lane; i < nv; i += 32) { const float4 v = src[i]; o32[i] = v; acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
#pragma unroll
            for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            const float inv = acc > 0.f ? rsqrtf(acc) : 0.f;
            for (int i = lane; i < nv; i += 32) {
                const float4 v = src[i];
                o16[2 * i] = __floats2bfloat162_rn(v.x * inv, v.y * inv);
                o16[2 * i + 1] = __floats2bfloat162_rn(v.z * inv, v.w * inv);
            }
            __syncwarp();
            if (lane == 0) ti[SSI_APPENDED * T + sl] = appended + 1;
        }
    }
    for (int k = warp; k < nBorn; k += NT / 32) {
        const int d = p.ud[(size_t)s * D + k];
        const float* b = feats + ((size_t)s * D + d) * F;
        float* a = p.feat + ((size_t)s * T + sm.born_slot[k]) * F;
        double acc = 0.0;
        for (int i = lane; i < F; i += 32) acc += (double)b[i] * b[i];
        const RowDiv nb = row_div(sqrtf((float)warp_sum_d(acc)));
        for (int i = lane; i < F; i += 32) a[i] = fdiv_row(b[i], nb);
    }
    // ---- result rows: confirmed tracks updated this frame, list order (strong_sort.py:84-99)
    const bool listed = tid < n && st == SS_CONFIRMED && tsu < 1;
    unsigned long long tot;
    const int orow = (int)block_exscan<NT>(listed ? 1ull : 0ull, sm.scratch, tot);
    if (listed) {
        const Box b = mean_tlbr(p.mean + ((size_t)s * T + slot) * 8);
        double2* o = reinterpret_cast<double2*>(out + ((size_t)s * T + orow) * 8);
        o[0] = make_double2(b.x1, b.y1); o[1] = make_double2(b.x2, b.y2);
        o[2] = make_double2((double)ti[SSI_ID * T + slot], p.conf[(size_t)s * T + slot]);
        o[3] = make_double2(p.cls[(size_t)s * T + slot], (double)ti[SSI_DET * T + slot]);
    }
    if (tid < nKeep) p.order[(size_t)s * T + tid] = sm.keep_slot[tid];
    if (tid < nBorn) p.order[(size_t)s * T + nKeep + tid] = sm.born_slot[tid];
    if (tid == 0) {
        p.counts[4 * s] = nKeep + nBorn;
        p.counts[4 * s + 1] = next_id + nud;           // _next_id advances for every unmatched detection
        p.counts[4 * s + 2] += 1;
        nout[s] = (int)tot;
        p.track_updates[s] += (unsigned long long)n;
    }
}

}  // namespace

size_t strongsort_match_smem() { return MATCH_LSA_OFF + lsa_work_bytes(CAP, CAP) + 64; }

#define SS_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return B200TRACK_ERR_CUDA; } } while (0)

int launch_strongsort_step(const SSParams& p, const double* dets, const int* ndets, const float* feats, const double* warps,
                           double* out, int* nout, int* err_step, cudaStream_t st) {
    const int S = p.n_streams;
    ss_pre_kernel<<<S, NT, 0, st>>>(p, dets, ndets, warps, err_step);
    SS_TRY(cudaGetLastError());
    if (int rc = b200track_kf_predict(B200TRACK_KF_XYAH_CONF, S * p.T, p.mean, p.cov, st)) return rc;
    // a cosine distance above max_dist / mc_lambda cannot survive the fused cost's clip at max_dist
    const double thr = p.max_dist / p.mc_lambda * (1.0 + 1e-12);
    if (int rc = b200track_gallery_cost(S, p.T, p.budget, p.D, p.F, p.gal32, p.gal16, p.gcount, feats, thr, thr + 1e-5, p.cost,
                                        p.ws, p.ws_bytes, reinterpret_cast<uint64_t*>(p.gstats), st)) return rc;
    if (int rc = b200track_gate_cost(B200TRACK_KF_XYAH_CONF, S, p.T, p.D, p.mean, p.cov, p.meas, 0, 1, p.mc_lambda, nullptr, p.cost, st)) return rc;
    const size_t smem = strongsort_match_smem();
    SS_TRY(cudaFuncSetAttribute(ss_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ss_match_kernel<<<S, NT, smem, st>>>(p, ndets);
    SS_TRY(cudaGetLastError());
    SS_TRY(launch_kf_update_masked(B200TRACK_KF_XYAH_CONF, S, p.T, p.D, p.mean, p.cov, p.meas, p.dconf, p.match, st));
    ss_post_kernel<<<S, NT, 0, st>>>(p, dets, ndets, feats, out, nout, err_step);
    SS_TRY(cudaGetLastError());
    return 0;
}

}  // namespace b200
