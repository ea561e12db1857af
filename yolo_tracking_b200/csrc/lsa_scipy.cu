// scipy.optimize.linear_sum_assignment, bit-faithful including its behaviour on exactly tied costs, one CTA per
// problem.  Reference call site: boxmot/trackers/strongsort/sort/linear_assignment.py:59-61 (min_cost_matching), where
// every entry above max_distance is clipped to max_distance + 1e-5, so most of the matrix is one repeated value and
// WHICH of the tied junk pairs the solver forms decides the order of `unmatched_detections`, hence the ids of new
// tracks.  scipy's solver (rectangular_lsap.cpp: Crouse's shortest augmenting paths, rows processed in order, the
// `remaining` column list filled in reverse and swap-removed, ties resolved towards a column that gives a new sink)
// is restated here with the column scan of every Dijkstra step spread over the threads; the reduction reproduces the
// sequential tie rule exactly (lowest value; among equal values the LAST unassigned column in scan order, else the
// FIRST column in scan order).  Arithmetic is kept in scipy's order (minVal + c - u[i] - v[j], one rounding each).
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200track.h"
#include "api_util.h"
#include "common.cuh"

namespace b200 {
namespace {

struct Cand {
    double val;
    int it;          // position in `remaining`, -1 = none
    int un;          // column is unassigned
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {      // a wins over b
    if (b.it < 0) return a.it >= 0;
    if (a.it < 0) return false;
    if (a.val < b.val) return true;
    if (a.val > b.val) return false;
    if (a.un != b.un) return a.un > b.un;
    return a.un ? a.it > b.it : a.it < b.it;
}

constexpr int LSA_NT = 256;

__global__ void __launch_bounds__(LSA_NT) lsa_scipy_kernel(int rows, int cols, const double* __restrict__ cost_all,
                                                           int* __restrict__ col4row_all, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char raw[];
    // scipy transposes tall problems so that rows <= columns
    const bool tr = cols < rows;
    const int nr = tr ? cols : rows, nc = tr ? rows : cols;
    const double* C = cost_all + (size_t)blockIdx.x * rows * cols;
    auto cost = [&](int i, int j) { return tr ? C[(size_t)j * cols + i] : C[(size_t)i * cols + j]; };
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = raw + off; off = (off + bytes + 15) & ~size_t(15); return p; };
    double* u = (double*)take(8 * (size_t)nr);
    double* v = (double*)take(8 * (size_t)nc);
    double* sp = (double*)take(8 * (size_t)nc);
    int* path = (int*)take(4 * (size_t)nc);
    int* row4col = (int*)take(4 * (size_t)nc);
    int* col4row = (int*)take(4 * (size_t)nr);
    int* remaining = (int*)take(4 * (size_t)nc);
    unsigned char* SR = take(nr);
    unsigned char* SC = take(nc);
    __shared__ Cand red[LSA_NT / 32];
    __shared__ int s_i, s_sink, s_nrem;
    __shared__ double s_min;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
            Cand best = {INF, -1, 0};
            for (int it = tid; it < nrem; it += LSA_NT) {
                const int j = remaining[it];
                const double r = xsub(xsub(xadd(minVal, cost(i, j)), ui), v[j]);
                double s = sp[j];
                if (r < s) { path[j] = i; sp[j] = r; s = r; }
                const Cand c = {s, it, row4col[j] == -1 ? 1 : 0};
                // sequential rule: take on strictly lower, or on equal when the column is unassigned
                if (best.it < 0 ? (s < INF || c.un) : better(c, best)) best = c;
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                Cand o;
                o.val = __shfl_xor_sync(0xffffffffu, best.val, d);
                o.it = __shfl_xor_sync(0xffffffffu, best.it, d);
                o.un = __shfl_xor_sync(0xffffffffu, best.un, d);
                if (better(o, best)) best = o;
            }
            if (lane == 0) red[warp] = best;
            __syncthreads();
            if (tid == 0) {
                Cand b = red[0];
                for (int k = 1; k < LSA_NT / 32; ++k) if (better(red[k], b)) b = red[k];
                SR[i] = 1;
                if (b.it < 0 || !(b.val < INF)) { s_sink = -2; atomicOr(err, 1); }       // infeasible (inf / nan costs)
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
        if (s_sink == -2) break;
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
    // internal (possibly transposed) assignment: col4row[nr]; the host side restores scipy's (row_ind, col_ind)
    int* out = col4row_all + (size_t)blockIdx.x * (rows < cols ? rows : cols);
    for (int i = tid; i < nr; i += LSA_NT) out[i] = col4row[i];
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200track_linear_sum_assignment(int32_t batch, int32_t rows, int32_t cols, const double* d_cost,
                                               int32_t* d_col4row, int32_t* d_err, void* st) {
    if (batch < 0 || rows < 0 || cols < 0 || !d_col4row || !d_err || (!d_cost && rows * cols > 0)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || rows == 0 || cols == 0) return 0;
    const size_t nr = rows < cols ? rows : cols, nc = rows < cols ? cols : rows;
    const size_t smem = 8 * nr + 16 + 2 * (8 * nc + 16) + 3 * (4 * nc + 16) + 4 * nr + 16 + nr + 16 + nc + 16;
    if (smem > 200 * 1024) { set_error("linear_sum_assignment: problem too large for shared memory"); return B200TRACK_ERR_CAPACITY; }
    B200_CU_TRY(cudaFuncSetAttribute(lsa_scipy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsa_scipy_kernel<<<batch, LSA_NT, smem, (cudaStream_t)st>>>(rows, cols, d_cost, d_col4row, d_err);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}
