// Dense linear assignment with lap.lapjv(cost, extend_cost=True) semantics (no cost_limit), one
// CTA per problem, on a cost MATRIX in global memory of any size up to 4096 x 4096: the form behind the
// operator b200track_lapjv(cost_limit = inf).  The frame steps use the matrix-free solver of lap_dense.cuh.
// Reference call site: boxmot/utils/association.py:20-24 (OC-SORT family).
//
// Without a limit lapjv pads with max(cost)+1, so an unmatched (row, column) pair costs
// lambda = 2 * (max + 1): every min(R, C) row is matched - nothing can be pruned.  The solver
// keeps a private "stay unmatched" column of cost lambda per row (the extended matrix without
// its dummy block) and runs
//   1. row reduction: u[r] = min_c cost[r][c]; the row takes its arg-min column unless a lower
//      row claimed it (dual feasible, complementary slack, free columns keep v = 0) - on
//      tracking matrices this assigns almost every real pair at once;
//   2. shortest augmenting paths for the rows still free: threads own columns, one block-wide
//      arg-min per Dijkstra step.
// On tie-free inputs the optimum is unique, hence identical to lapjv's x, y.
#pragma once
#include "common.cuh"

namespace b200 {

struct DenseLapM {
    double* u;              // [R]
    double* v;              // [C]
    double* dist;           // [C]
    int* pred;              // [C]
    int* xr;                // [R]  column of row, -1 = unmatched
    int* yc;                // [C]  row of column, -1 = free
    int* claim;             // [R]
    unsigned char* scn;     // [C]
    double* red_v;          // [32]
    int* red_i;             // [32]
    double* sh_d;           // [4]  s_min, s_bestDummy
    int* sh_i;              // [4]  s_cur, s_sink, s_bestRow
};

// Step 1, row reduction: u[r] = min_c cost[r][c] (capped by lambda), v = 0; a row takes its
// arg-min column unless a lower row claimed it.  Rows are the side that must be matched or pay
// lambda; columns may stay free, so a free column must keep v = 0 for the optimality proof -
// which is why the reduction runs over rows (a column reduction would leave free columns with
// v = colmin != 0).  One warp per row, lanes sweep the columns (coalesced reads of the matrix).
template <int NT>
__device__ void dense_lapm_init(const DenseLapM& w, const double* C, int ld, int R, int Cn, double lambda, bool have_rowmin = false) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < Cn; c += NT) { w.v[c] = 0.0; w.yc[c] = -1; w.pred[c] = 0x7fffffff; }
    for (int r = tid; r < R; r += NT) w.xr[r] = -1;
    __syncthreads();
    if (have_rowmin) {
        // the caller reduced every row while it filled the matrix: w.u[r] = row minimum, w.claim[r] = its column
        for (int r = tid; r < R; r += NT) {
            const double m = w.u[r];
            const int a = w.claim[r];
            const bool take = a >= 0 && m <= lambda;
            w.u[r] = take ? m : lambda;
            w.claim[r] = take ? a : -1;
            if (take) atomicMin(&w.pred[a], r);
        }
    } else {
        for (int r = warp; r < R; r += NT / 32) {
            double m = INF; int a = -1;
            const double* row = C + (size_t)r * ld;
            for (int j = lane; j < Cn; j += 32) { const double x = row[j]; if (x < m) { m = x; a = j; } }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, m, d);
                const int oa = __shfl_xor_sync(0xffffffffu, a, d);
                if (om < m || (om == m && oa >= 0 && (a < 0 || oa < a))) { m = om; a = oa; }
            }
            if (lane == 0) {
                const bool take = a >= 0 && m <= lambda;
                w.u[r] = take ? m : lambda;
                w.claim[r] = take ? a : -1;
                if (take) atomicMin(&w.pred[a], r);
            }
        }
    }
    __syncthreads();
    for (int r = tid; r < R; r += NT) {
        const int a = w.claim[r];
        if (a >= 0 && w.pred[a] == r) { w.xr[r] = a; w.yc[a] = r; }
    }
    __syncthreads();
}

// Step 2 for every row that is still free.
template <int NT>
__device__ void dense_lapm_augment(const DenseLapM& w, const double* C, int ld, int R, int Cn, double lambda) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i0 = 0; i0 < R; ++i0) {
        if (w.xr[i0] >= 0) continue;            // uniform: shared memory, read after a barrier
        for (int j = tid; j < Cn; j += NT) { w.dist[j] = INF; w.scn[j] = 0; w.pred[j] = -1; }
        if (tid == 0) { w.sh_d[0] = 0.0; w.sh_i[0] = i0; w.sh_d[1] = INF; w.sh_i[2] = -1; w.sh_i[1] = -2; }
        __syncthreads();
        while (true) {
            const int i = w.sh_i[0];
            const double minVal = w.sh_d[0], ui = w.u[i];
            double best = INF; int bj = -1;
            for (int j = tid; j < Cn; j += NT) {
                if (w.scn[j]) continue;
                const double r = minVal + C[(size_t)i * ld + j] - ui - w.v[j];
                double dj = w.dist[j];
                if (r < dj) { dj = r; w.dist[j] = r; w.pred[j] = i; }
                if (dj < best) { best = dj; bj = j; }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, d);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
                if (ob < best || (ob == best && oj >= 0 && (bj < 0 || oj < bj))) { best = ob; bj = oj; }
            }
            if (lane == 0) { w.red_v[warp] = best; w.red_i[warp] = bj; }
            __syncthreads();
            if (tid == 0) {
                double b = w.red_v[0]; int j = w.red_i[0];
                for (int k = 1; k < NT / 32; ++k)
                    if (w.red_v[k] < b || (w.red_v[k] == b && w.red_i[k] >= 0 && (j < 0 || w.red_i[k] < j))) { b = w.red_v[k]; j = w.red_i[k]; }
                const double dd = minVal + lambda - ui;         // row i may stay unmatched
                if (dd < w.sh_d[1]) { w.sh_d[1] = dd; w.sh_i[2] = i; }
                if (j < 0 || w.sh_d[1] <= b) { w.sh_i[1] = -1; w.sh_d[0] = w.sh_d[1]; }
                else {
                    w.sh_d[0] = b; w.scn[j] = 1;
                    if (w.yc[j] < 0) w.sh_i[1] = j; else w.sh_i[0] = w.yc[j];
                }
            }
            __syncthreads();
            if (w.sh_i[1] != -2) break;
        }
        const double minVal = w.sh_d[0];
        const int sink = w.sh_i[1];
        for (int j = tid; j < Cn; j += NT) {
            if (!w.scn[j]) continue;
            const double delta = minVal - w.dist[j];
            const int r = w.yc[j];
            if (r >= 0) w.u[r] += delta;
            w.v[j] -= delta;
        }
        __syncthreads();
        if (tid == 0) {
            w.u[i0] += minVal;
            int j = -1;
            bool go = true;
            if (sink >= 0) j = sink;
            else if (w.sh_i[2] == i0) go = false;
            else { j = w.xr[w.sh_i[2]]; w.xr[w.sh_i[2]] = -1; }
            while (go) {
                const int r = w.pred[j];
                w.yc[j] = r;
                const int t = w.xr[r];
                w.xr[r] = j;
                j = t;
                if (r == i0) break;
            }
        }
        __syncthreads();
    }
}

}  // namespace b200
