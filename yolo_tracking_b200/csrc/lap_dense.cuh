// Dense linear assignment with lap.lapjv(cost, extend_cost=True) semantics (no cost_limit), one
// CTA per problem, MATRIX-FREE: the cost of a pair is a functor evaluated from shared-memory
// resident boxes whenever the solver needs it, the R x C matrix is never stored (at 100 x 130
// fp64 it would be 104 KB per stream - more than a CTA's share of shared memory, and as a global
// scratch block it was 40 % of the OC-SORT step's DRAM traffic).
// Reference call site: boxmot/utils/association.py:20-24 (OC-SORT family).
//
// Without a limit lapjv pads with max(cost)+1, so an unmatched (row, column) pair costs
// lambda = 2 * (max + 1): every min(R, C) row is matched - nothing can be pruned from the
// problem.  Which pairs are matched does not depend on the value of lambda once it exceeds
// every cost by more than 2 (leaving a matchable pair unmatched can then never pay), so callers
// may pass any upper bound of 2 * (max + 1).  The solver keeps a private "stay unmatched" column
// of cost lambda per row (the extended matrix without its dummy block) and runs
//   1. row reduction: u[r] = min_c cost(r, c); the row takes its arg-min column unless a lower
//      row claimed it (dual feasible, complementary slack, free columns keep v = 0) - on
//      tracking matrices this assigns almost every real pair at once.  The caller computes the
//      row minima (it can prune: see ocsort_step.cu);
//   2. shortest augmenting paths for the rows still free (under one per frame on tracking
//      data): thread j owns column j (distance, scanned flag and v[j] in registers), one
//      block-wide arg-min with ONE barrier per Dijkstra step - the per-warp minima are
//      double-buffered and every thread reduces them itself, so there is no serial section.
//      A column whose cheap lower bound cannot improve its distance is not evaluated: after the
//      first step of a search that is nearly every column (a matched row's dual sits ~0.9 below
//      the bound of its non-overlapping columns).
// On tie-free inputs the optimum is unique, hence identical to lapjv's x, y.
#pragma once
#include "common.cuh"

namespace b200 {

struct DenseLap {
    double* u;              // [R]
    double* v;              // [C]
    int* pred;              // [C]
    int* xr;                // [R]  column of row, -1 = unmatched
    int* yc;                // [C]  row of column, -1 = free
    int* claim;             // [R]
    double* red_v;          // [2][32]
    int* red_i;             // [2][32]
    int* freerow;           // [R]  rows left free by the row reduction, ascending
    unsigned long long* dbg;   // optional counters (profiling aid): [9] += searches, [10] += Dijkstra steps, [11] += cost evaluations
};

// Step 1.  The caller reduced every row: w.u[r] = row minimum, w.claim[r] = its column (-1: none).
// Rows are the side that must be matched or pay lambda; columns may stay free, so a free column must
// keep v = 0 for the optimality proof - which is why the reduction runs over rows.
template <int NT>
__device__ void dense_lap_init(const DenseLap& w, int R, int Cn, double lambda) {
    const int tid = threadIdx.x;
    for (int c = tid; c < Cn; c += NT) { w.v[c] = 0.0; w.yc[c] = -1; w.pred[c] = 0x7fffffff; }
    for (int r = tid; r < R; r += NT) w.xr[r] = -1;
    __syncthreads();
    for (int r = tid; r < R; r += NT) {
        const double m = w.u[r];
        const int a = w.claim[r];
        const bool take = a >= 0 && m <= lambda;
        w.u[r] = take ? m : lambda;
        w.claim[r] = take ? a : -1;
        if (take) atomicMin(&w.pred[a], r);
    }
    __syncthreads();
    for (int r = tid; r < R; r += NT) {
        const int a = w.claim[r];
        if (a >= 0 && w.pred[a] == r) { w.xr[r] = a; w.yc[a] = r; }
    }
    __syncthreads();
}

// Step 2 for every row that is still free.  cost(r, c) must return the same bits the row reduction saw;
// cost.lower(r, c) <= cost(r, c).
// Needs Cn <= NT (one column per thread).
template <int NT, class CostFn>
__device__ void dense_lap_augment(const DenseLap& w, const CostFn& cost, int R, int Cn, double lambda) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = tid;
    const bool mine = j < Cn;
    const int nwa = (Cn + 31) >> 5;                         // warps that own columns: only they reduce, only their minima are read
    int par = 0;
    // the rows the row reduction left free, ascending (warp 0 compacts them; usually none or one)
    int nfree = 0;
    if (warp == 0) {
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            const bool fr = r < R && w.xr[r] < 0;
            const uint32_t m = __ballot_sync(0xffffffffu, fr);
            if (fr) w.freerow[nfree + __popc(m & ((1u << lane) - 1u))] = r;
            nfree += __popc(m);
        }
        if (lane == 0) w.red_i[63] = nfree;
    }
    __syncthreads();
    nfree = w.red_i[63];
    for (int f = 0; f < nfree; ++f) {
        const int i0 = w.freerow[f];
        if (w.dbg && tid == 0) atomicAdd(&w.dbg[9], 1ull);
        double dist = INF;
        bool scn = false;
        const double vj = mine ? w.v[j] : 0.0;
        if (mine) w.pred[j] = -1;
        int i = i0, bestRow = -1, sink = -2;
        double minVal = 0.0, bestDummy = INF;
        while (true) {
            const double ui = w.u[i];
            double cand = INF;
            int cj = -1;
            if (mine && !scn) {
                // cost.lower(i, j) <= cost(i, j) is cheap (a bound for pairs that provably do not overlap); fp addition is
                // monotone, so the bound below never exceeds the value it stands for and skipping is exact
                if (minVal + cost.lower(i, j) - ui - vj < dist) {
                    if (w.dbg) atomicAdd(&w.dbg[11], 1ull);
                    const double r = minVal + cost(i, j) - ui - vj;
                    if (r < dist) { dist = r; w.pred[j] = i; }
                }
                cand = dist; cj = j;
            }
            if (warp < nwa) {                               // warps without columns have nothing to offer
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, cand, d);
                    const int oj = __shfl_xor_sync(0xffffffffu, cj, d);
                    if (ob < cand || (ob == cand && oj >= 0 && (cj < 0 || oj < cj))) { cand = ob; cj = oj; }
                }
                if (lane == 0) { w.red_v[par * 32 + warp] = cand; w.red_i[par * 32 + warp] = cj; }
            }
            if (w.dbg && tid == 0) atomicAdd(&w.dbg[10], 1ull);
            __syncthreads();
            double b = w.red_v[par * 32];
            int bj = w.red_i[par * 32];
            for (int k = 1; k < nwa; ++k) {
                const double ob = w.red_v[par * 32 + k];
                const int oj = w.red_i[par * 32 + k];
                if (ob < b || (ob == b && oj >= 0 && (bj < 0 || oj < bj))) { b = ob; bj = oj; }
            }
            par ^= 1;
            const double dd = minVal + lambda - ui;         // row i may stay unmatched
            if (dd < bestDummy) { bestDummy = dd; bestRow = i; }
            if (bj < 0 || bestDummy <= b) { sink = -1; minVal = bestDummy; break; }
            minVal = b;
            if (j == bj) scn = true;
            const int owner = w.yc[bj];                     // not modified during a search
            if (owner < 0) { sink = bj; break; }
            i = owner;
        }
        if (mine && scn) {
            const double delta = minVal - dist;
            const int r = w.yc[j];
            if (r >= 0) w.u[r] += delta;
            w.v[j] = vj - delta;
        }
        __syncthreads();
        if (tid == 0) {
            w.u[i0] += minVal;
            int c = -1;
            bool go = true;
            if (sink >= 0) c = sink;
            else if (bestRow == i0) go = false;
            else { c = w.xr[bestRow]; w.xr[bestRow] = -1; }
            while (go) {
                const int r = w.pred[c];
                w.yc[c] = r;
                const int t = w.xr[r];
                w.xr[r] = c;
                c = t;
                if (r == i0) break;
            }
        }
        __syncthreads();
    }
}

}  // namespace b200
