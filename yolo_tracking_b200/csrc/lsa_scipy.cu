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
#include "lsa_scipy.cuh"

namespace b200 {
namespace {

__global__ void __launch_bounds__(LSA_NT) lsa_scipy_kernel(int rows, int cols, const double* __restrict__ cost_all,
                                                           int* __restrict__ col4row_all, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char raw[];
    // scipy transposes tall problems so that rows <= columns
    const bool tr = cols < rows;
    const int nr = tr ? cols : rows, nc = tr ? rows : cols;
    const double* C = cost_all + (size_t)blockIdx.x * rows * cols;
    auto cost = [&](int i, int j) { return tr ? C[(size_t)j * cols + i] : C[(size_t)i * cols + j]; };
    const LsaWork w = lsa_carve(raw, nr, nc);
    if (!lsa_scipy_solve(nr, nc, cost, w) && threadIdx.x == 0) atomicOr(err, 1);
    // internal (possibly transposed) assignment: col4row[nr]; the host side restores scipy's (row_ind, col_ind)
    int* out = col4row_all + (size_t)blockIdx.x * (rows < cols ? rows : cols);
    for (int i = threadIdx.x; i < nr; i += LSA_NT) out[i] = w.col4row[i];
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200track_linear_sum_assignment(int32_t batch, int32_t rows, int32_t cols, const double* d_cost,
                                               int32_t* d_col4row, int32_t* d_err, void* st) {
    if (batch < 0 || rows < 0 || cols < 0 || !d_col4row || !d_err || (!d_cost && rows * cols > 0)) { set_error("bad argument"); return B200TRACK_ERR_ARG; }
    if (batch == 0 || rows == 0 || cols == 0) return 0;
    const size_t nr = rows < cols ? rows : cols, nc = rows < cols ? cols : rows;
    const size_t smem = lsa_work_bytes(nr, nc);
    if (smem > 200 * 1024) { set_error("linear_sum_assignment: problem too large for shared memory"); return B200TRACK_ERR_CAPACITY; }
    B200_CU_TRY(cudaFuncSetAttribute(lsa_scipy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsa_scipy_kernel<<<batch, LSA_NT, smem, (cudaStream_t)st>>>(rows, cols, d_cost, d_col4row, d_err);
    B200_CU_TRY(cudaGetLastError());
    return 0;
}
