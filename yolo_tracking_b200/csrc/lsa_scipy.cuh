// scipy.optimize.linear_sum_assignment as a CTA-wide device routine (see lsa_scipy.cu for the restated tie rules): used by
// the operator kernel and by the batched StrongSORT step (strongsort_step.cu), which solves its two assignment problems
// per stream on matrices of per-stream size.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace b200 {

struct LsaCand {
    double val;
    int it;          // position in `remaining`, -1 = none
    int un;          // column is unassigned
};
__device__ __forceinline__ bool lsa_better(const LsaCand& a, const LsaCand& b) {      // a wins over b
    if (b.it < 0) return a.it >= 0;
    if (a.it < 0) return false;
    if (a.val < b.val) return true;
    if (a.val > b.val) return false;
    if (a.un != b.un) return a.un > b.un;
    return a.un ? a.it > b.it : a.it < b.it;
}

constexpr int LSA_NT = 256;

// work arrays of one problem with nr <= nc (shared memory)
struct LsaWork {
    double *u, *v, *sp;                  // [nr], [nc], [nc]
    int *path, *row4col, *col4row, *remaining;   // [nc], [nc], [nr], [nc]
    unsigned char *SR, *SC;              // [nr], [nc]
};
__host__ __device__ inline size_t lsa_work_bytes(size_t nr, size_t nc) {
    return 8 * nr + 16 + 2 * (8 * nc + 16) + 3 * (4 * nc + 16) + 4 * nr + 16 + nr + 16 + nc + 16;
}
__device__ __forceinline__ LsaWork lsa_carve(unsigned char* raw, size_t nr, size_t nc) {
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = raw + off; off = (off + bytes + 15) & ~size_t(15); return p; };
    LsaWork w;
    w.u = (double*)take(8 * nr); w.v = (double*)take(8 * nc); w.sp = (double*)take(8 * nc);
    w.path = (int*)take(4 * nc); w.row4col = (int*)take(4 * nc); w.col4row = (int*)take(4 * nr); w.remaining = (int*)take(4 * nc);
    w.SR = take(nr); w.SC = take(nc);
    return w;
}

// Solves min sum cost(i, col4row[i]) over the nr x nc problem (nr <= nc: scipy transposes tall problems before it
// solves them, the caller passes the accessor of the orientation scipy would solve).  All LSA_NT threads of the CTA
// call it; w.col4row holds the assignment afterwards (visible after the final barrier).  Returns false when the problem
// is infeasible (inf / nan costs).
template <class CostFn>
__device__ __forceinline__ bool lsa_scipy_solve(int nr, int nc, CostFn cost, const LsaWork& w) {
    __shared__ LsaCand red[LSA_NT / 32];
    __shared__ int s_i, s_sink, s_nrem;
    __shared__ double s_min;
    double* u = w.u; double* v = w.v; double* sp = w.sp;
    int* path = w.path; int* row4col = w.row4col; int* col4row = w.col4row; int* remaining = w.remaining;
    unsigned char* SR = w.SR; unsigned char* SC = w.SC;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool feasible = true;
    __syncthreads();                     // the previous user of the shared words / work arrays is done
    for (int i = tid; i < nr; i += LSA_NT) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = tid; j < nc; j += LSA_NT) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    __syncthreads();
    for (int cur = 0; cur < nr; ++cur) {
        for (int j = tid; j < nc; j += LSA_NT) { remaining[j] = nc - j - 1; SC[j] = 0; sp[j] = INF; }
        for (int i = tid; i < nr; i += LSA_NT) SR[i] = 0;
        if (tid == 0) { s_i = cur; s_sink = -1; s_nrem = nc; s_min = 0.0; }
        __syncthreads();
        while (true) {
            const int i = s_i, nrem = s_nrem;
            const double minVal = s_min, ui = u[i];
            LsaCand best = {INF, -1, 0};
            for (int it = tid; it < nrem; it += LSA_NT) {
                const int j = remaining[it];
                const double r = xsub(xsub(xadd(minVal, cost(i, j)), ui), v[j]);
                double s = sp[j];
                if (r < s) { path[j] = i; sp[j] = r; s = r; }
                const LsaCand c = {s, it, row4col[j] == -1 ? 1 : 0};
                // sequential rule: take on strictly lower, or on equal when the column is unassigned
                if (best.it < 0 ? (s < INF || c.un) : lsa_better(c, best)) best = c;
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                LsaCand o;
                o.val = __shfl_xor_sync(0xffffffffu, best.val, d);
                o.it = __shfl_xor_sync(0xffffffffu, best.it, d);
                o.un = __shfl_xor_sync(0xffffffffu, best.un, d);
                if (lsa_better(o, best)) best = o;
            }
            if (lane == 0) red[warp] = best;
            __syncthreads();
            if (tid == 0) {
                LsaCand b = red[0];
                for (int k = 1; k < LSA_NT / 32; ++k) if (lsa_better(red[k], b)) b = red[k];
                SR[i] = 1;
                if (b.it < 0 || !(b.val < INF)) s_sink = -2;                             // infeasible (inf / nan costs)
                else {
                    s_min = b.val;
                    const int j = remaining[b.it];
                    if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
                    SC[j] = 1;
                    remaining[b.it] = remaining[nrem - 1];
                    s_nrem = nrem - 1;
                }
            }
            __syncthreads();
            if (s_sink != -1) break;
        }
        if (s_sink == -2) { feasible = false; break; }
        const double minVal = s_min;
        for (int i = tid; i < nr; i += LSA_NT)
            if (i == cur) u[i] = xadd(u[i], minVal);
            else if (SR[i]) u[i] = xadd(u[i], xsub(minVal, sp[col4row[i]]));
        for (int j = tid; j < nc; j += LSA_NT)
            if (SC[j]) v[j] = xsub(v[j], xsub(minVal, sp[j]));
        __syncthreads();
        if (tid == 0) {
            int j = s_sink;
            while (true) {
                const int i = path[j];
                row4col[j] = i;
                const int t = col4row[i];
                col4row[i] = j;
                j = t;
                if (i == cur) break;
            }
        }
        __syncthreads();
    }
    return feasible;
}

}  // namespace b200
