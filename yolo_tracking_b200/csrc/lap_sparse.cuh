// Exact linear assignment with lap.lapjv(cost, extend_cost=True, cost_limit=L) semantics on a
// pruned (sparse) candidate graph, one CTA per stream.
//
// lapjv's (R+C)x(R+C) extended matrix (reference call site
// boxmot/utils/matching.py:56-71) has the optimum of
//     sum_{matched} c_ij + L/2 * (#unmatched rows + #unmatched cols)
//  =  const + sum_{matched} (c_ij - L),
// so a pair with c_ij > L is never used and can be pruned EXACTLY.  What is left is a sparse
// bipartite graph that falls apart into many small connected components; each component is
// an independent assignment problem.  Here:
//   1. all threads build the candidate graph (caller): per-row edge lists with cached costs, or bitmask rows,
//   2. rows are placed by a greedy tight start (edge lists) or components are found with a lock-free union-find
//      (bitmask rows),
//   3. what is left is solved with the shortest-augmenting-path (Hungarian / JV augmentation) algorithm over the
//      sparse rows.  Every row owns a
//      private "stay unmatched" column of cost L, real edges keep their cost c_ij (columns
//      stay unmatched for free): the same objective up to a constant, with the dummy block
//      of the extended matrix never materialised and no cost ever rescaled.
// The dual variables, predecessor links and work lists are shared-memory arrays indexed by
// the stream-global row / column id, so concurrent components never touch the same entry.
// On tie-free inputs the optimum is unique, hence identical to lapjv's x, y.
#pragma once
#include "common.cuh"

namespace b200 {

struct LapWork {
    int Tmax, Dmax;
    uint32_t* adj;      // [Dmax/32][Tmax]  only read when there is no (valid) edge cache
    double* u;          // [Tmax]
    double* v;          // [Dmax]
    double* dist;       // [Dmax]
    int* parent;        // [Tmax + Dmax]
    int* head;          // [Tmax]  compact list of the rows sent to the general solver
    short* rnext;       // [Tmax]
    short* xr;          // [Tmax]  column of row, -1 = unmatched
    short* yc;          // [Dmax]  row of column, -1 = free
    short* pred;        // [Dmax]
    short* nextc;       // [Dmax]
    short* mark;        // [Dmax]  stamp when reached
    short* scn;         // [Dmax]  stamp when scanned
    int* coldeg;        // [Dmax]  candidate rows per column (filled while adj is built)
    int* ncomplex;      // [2]     rows that need an augmentation; rows whose concurrent search gave up
    // optional edge cache (costs computed while the graph was built): per-row linked lists in a
    // fixed pool.  ecount[0] > ecap means the pool overflowed and costs are recomputed instead.
    double* ecost = nullptr;   // [ecap]
    short* ecol = nullptr;     // [ecap]
    short* enext = nullptr;    // [ecap]
    int* ehead = nullptr;      // [Tmax]
    int* ecount = nullptr;     // [1]
    int ecap = 0;
    // optional profiling hook: cycles up to the end of the classification stage are added to dbg[dbg_slot] (thread 0)
    unsigned long long* dbg = nullptr;
    int dbg_slot = 0;
    long long* dbg_last = nullptr;
};

__device__ __forceinline__ int uf_find(volatile int* parent, int x) {
    while (true) {
        const int p = parent[x];
        if (p == x) return x;
        const int g = parent[p];
        parent[x] = g;          // path halving: g is always an ancestor of x, races are benign
        x = g;
    }
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        if (atomicCAS(&parent[b], b, a) == b) return;   // link the larger root under the smaller
    }
}

// Insert one row into the matching of its component (one shortest augmenting path).
// CONCURRENT: several rows are inserted at the same time by different threads.  A search claims every column it
// touches (compare-and-swap on claim[j], "free" = Tmax + j as left by lap_prepare in parent[Tmax + j]); the rows of
// its tree are the mates of claimed columns, so two searches that never meet on a column work on edge-disjoint
// parts of the graph and their dual updates and augmentations commute.  A search that meets a foreign claim gives
// up WITHOUT having modified the matching or the duals (returns false; the row is inserted afterwards, alone).
template <bool CONCURRENT = false, class Cost, class Lambda>
__device__ bool lap_insert_row(const LapWork& w, int words, const Lambda& lambda, const Cost& cost, int i0) {
    const short stamp = (short)(i0 + 1 + (CONCURRENT ? w.Tmax : 0));       // a retry never sees its own stale marks
    int* claim = w.parent + w.Tmax;
    bool conflict = false;
    const bool cached = w.ecost != nullptr && *w.ecount <= w.ecap;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double minVal = 0.0;
    int i = i0;
    double bestDummy = INF;
    int bestDummyRow = -1;
    int rhead = -1, rtail = -1;
    int sink = -1;
    while (true) {
        const double ui = w.u[i];
        const double dd = minVal + lambda(i) - ui;     // row i may stay unmatched at cost lambda(i)
        if (dd < bestDummy) { bestDummy = dd; bestDummyRow = i; }
        auto relax = [&](int j, double c) {
            if (CONCURRENT) {
                const int o = claim[j];
                if (o != i0 && (o != w.Tmax + j || atomicCAS(&claim[j], w.Tmax + j, i0) != w.Tmax + j)) { conflict = true; return; }
            }
            if (w.scn[j] == stamp) return;
            const double r = minVal + c - ui - w.v[j];
            if (w.mark[j] != stamp) {
                w.mark[j] = stamp; w.dist[j] = r; w.pred[j] = (short)i; w.nextc[j] = -1;
                if (rtail < 0) rhead = j; else w.nextc[rtail] = (short)j;
                rtail = j;
            } else if (r < w.dist[j]) { w.dist[j] = r; w.pred[j] = (short)i; }
        };
        if (cached) {
            for (int e = w.ehead[i]; e >= 0; e = w.enext[e]) relax(w.ecol[e], w.ecost[e]);
            if (CONCURRENT && conflict) return false;
        } else {
            for (int wd = 0; wd < words; ++wd) {
                uint32_t bits = w.adj[wd * w.Tmax + i];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int j = wd * 32 + b;
                    relax(j, cost(i, j));
                }
            }
        }
        int jmin = -1;
        double dmin = INF;
        for (int j = rhead; j >= 0; j = w.nextc[j])
            if (w.scn[j] != stamp && w.dist[j] < dmin) { dmin = w.dist[j]; jmin = j; }
        if (jmin < 0 || bestDummy <= dmin) { sink = -1; minVal = bestDummy; break; }
        minVal = dmin;
        w.scn[jmin] = stamp;
        if (w.yc[jmin] < 0) { sink = jmin; break; }
        i = w.yc[jmin];
    }
    // dual update (rows of the tree are the mates of the scanned columns, plus i0)
    w.u[i0] += minVal;
    for (int j = rhead; j >= 0; j = w.nextc[j]) {
        if (w.scn[j] != stamp) continue;
        const double delta = minVal - w.dist[j];
        const int r = w.yc[j];
        if (r >= 0) w.u[r] += delta;
        w.v[j] -= delta;
    }
    // augment
    int j;
    if (sink >= 0) j = sink;
    else {
        if (bestDummyRow == i0) return true;          // i0 itself stays unmatched
        j = w.xr[bestDummyRow];
        w.xr[bestDummyRow] = -1;
    }
    while (true) {
        const int r = w.pred[j];
        w.yc[j] = (short)r;
        const int t = w.xr[r];
        w.xr[r] = (short)j;
        j = t;
        if (r == i0) break;
    }
    return true;
}

// Zero the per-problem work arrays.  Call (all threads) BEFORE the candidate graph is built:
// the builder accumulates coldeg[j] with atomicAdd for every edge it writes into adj.
template <int NT>
__device__ __forceinline__ void lap_prepare(const LapWork& w, int nrows, int words) {
    const int tid = threadIdx.x;
    const int ncols = words * 32;
    for (int t = tid; t < nrows; t += NT) {
        w.xr[t] = -1; w.u[t] = 0.0; w.parent[t] = t; w.head[t] = -1; w.rnext[t] = -1;
        if (w.ehead) w.ehead[t] = -1;
    }
    for (int j = tid; j < ncols; j += NT) {
        w.yc[j] = -1; w.v[j] = 0.0; w.parent[w.Tmax + j] = w.Tmax + j; w.mark[j] = 0; w.scn[j] = 0; w.coldeg[j] = 0;
    }
    if (tid == 0) { w.ncomplex[0] = 0; w.ncomplex[1] = 0; if (w.ecount) *w.ecount = 0; }
}

// Whole-CTA solve.  coldeg[] and either the edge cache or adj[word][row] must be complete (zero for rows / columns
// not taking part) and visible (__syncthreads() after the build); rows are 0..nrows-1, columns 0..32*words-1.
// lambda(row) is the cost limit of that row's problem.  Results in xr / yc.
//
// With a valid edge cache (the frame steps):
//   * a row whose only edge leads to a column nobody else wants is matched directly;
//   * every other row takes part in a greedy TIGHT start: u[row] = its cheapest edge (never above lambda: dearer edges
//     were pruned), v = 0, and the row claims that column with a compare-and-swap.  Duals are feasible and every
//     claimed edge is tight, so rows that got their column are optimally placed unless an augmentation re-routes them;
//   * the rows that lost the race (two tracks preferring one detection - a handful per frame) are inserted by
//     shortest augmenting paths, concurrently (one per warp): a search never leaves the component of its row, and
//     searches that do meet are detected by column claims and redone one after the other.
// Without an edge cache (operator kernel, overflowed cache): components by union-find over the bitmask rows, each
// solved by the thread of its smallest row.
template <int NT, class Cost, class Lambda>
__device__ void lap_sparse_solve(const LapWork& w, int nrows, int words, const Lambda& lambda, const Cost& cost) {
    const int tid = threadIdx.x;
    const bool small_ok = w.ecost != nullptr && *w.ecount <= w.ecap;
    if (small_ok) {
        for (int t = tid; t < nrows; t += NT) {
            const int e0 = w.ehead[t];
            if (e0 < 0) continue;
            int bj = w.ecol[e0];
            int e = w.enext[e0];
            if (e < 0 && w.coldeg[bj] == 1) { w.xr[t] = (short)bj; w.yc[bj] = (short)t; continue; }
            double best = w.ecost[e0];
            for (; e >= 0; e = w.enext[e]) {
                const double c = w.ecost[e];
                if (c < best) { best = c; bj = w.ecol[e]; }
            }
            w.u[t] = best;
            if (atomicCAS(reinterpret_cast<unsigned short*>(&w.yc[bj]), (unsigned short)0xffffu, (unsigned short)t) == 0xffffu) w.xr[t] = (short)bj;
            else w.head[atomicAdd(w.ncomplex, 1)] = t;
        }
        __syncthreads();
        if (w.dbg && tid == 0) { const long long now_ = clock64(); atomicAdd(&w.dbg[w.dbg_slot], (unsigned long long)(now_ - *w.dbg_last)); *w.dbg_last = now_; }
        const int nc = *w.ncomplex;
        if (nc == 0) return;                              // uniform: every thread reads the same value
        // concurrent insertions, one per warp at a time (lane 0): searches in different components never meet
        if ((tid & 31) == 0) {
            for (int k = tid >> 5; k < nc; k += NT / 32) {
                const int r = w.head[k];
                if (!lap_insert_row<true>(w, words, lambda, cost, r)) w.rnext[atomicAdd(w.ncomplex + 1, 1)] = (short)r;
            }
        }
        __syncthreads();
        const int nretry = w.ncomplex[1];
        if (nretry == 0) return;                          // uniform
        if (tid == 0)
            for (int k = 0; k < nretry; ++k) lap_insert_row(w, words, lambda, cost, w.rnext[k]);
        __syncthreads();
        return;
    }
    for (int t = tid; t < nrows; t += NT) {
        int deg = 0, first = -1;
        for (int wd = 0; wd < words; ++wd) {
            const uint32_t bits = w.adj[wd * w.Tmax + t];
            if (bits) { if (first < 0) first = wd * 32 + __ffs(bits) - 1; deg += __popc(bits); }
        }
        if (deg == 0) continue;
        if (deg == 1 && w.coldeg[first] == 1) { w.xr[t] = (short)first; w.yc[first] = (short)t; continue; }
        atomicAdd(w.ncomplex, 1);
        for (int wd = 0; wd < words; ++wd) {
            uint32_t bits = w.adj[wd * w.Tmax + t];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                uf_union(w.parent, t, w.Tmax + wd * 32 + b);
            }
        }
        w.rnext[t] = -2;                                  // marks "complex row" for the next phase
    }
    __syncthreads();
    if (*w.ncomplex == 0) return;                         // uniform
    for (int t = tid; t < nrows; t += NT) {
        if (w.rnext[t] != -2) continue;
        w.rnext[t] = (short)atomicExch(&w.head[uf_find(w.parent, t)], t);   // the root is the component's smallest row
    }
    __syncthreads();
    for (int t = tid; t < nrows; t += NT)
        for (int r = w.head[t]; r >= 0; r = w.rnext[r]) lap_insert_row(w, words, lambda, cost, r);
    __syncthreads();
}

}  // namespace b200
