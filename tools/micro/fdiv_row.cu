// Bit-for-bit check of the row division used by the BoT-SORT embedding arithmetic (csrc/bytetrack_step.cu: row_div /
// fdiv_row) against __fdiv_rn: for every divisor n of a sweep and 2^20 numerators each, the quotient formed from the shared
// refined reciprocal must equal the correctly rounded one.  Numerators cover the magnitudes embedding components take
// (1e-12 .. 1e3, both signs); divisors the norms (1e-3 .. 1e3).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fdiv_row tools/micro/fdiv_row.cu && ./fdiv_row
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct RowDiv { float n, y; };
__device__ __forceinline__ RowDiv row_div(float n) {
    float y0;
    asm("rcp.approx.f32 %0, %1;" : "=f"(y0) : "f"(n));
    return RowDiv{n, __fmaf_rn(y0, __fmaf_rn(-n, y0, 1.0f), y0)};
}
__device__ __forceinline__ float fdiv_row(float a, const RowDiv& d) {
    const float q0 = __fmul_rn(a, d.y);
    return __fmaf_rn(d.y, __fmaf_rn(-d.n, q0, a), q0);
}
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void check(int ndiv, unsigned long long* bad, unsigned long long* total) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mism = 0, cnt = 0;
    for (int k = 0; k < ndiv; ++k) {
        // divisor: random mantissa, exponent in [-10, 10]
        const uint32_t hb = hash(0x9e3779b9u * (k + 1));
        const float n = __uint_as_float(((127u - 10u + hb % 21u) << 23) | (hash(hb) & 0x7fffffu));
        const RowDiv d = row_div(n);
        for (int r = 0; r < 16; ++r) {
            const uint32_t ha = hash(gid * 16u + r + 0x85ebca6bu * (k + 1));
            const float a = __uint_as_float(((ha & 1u) << 31) | ((127u - 40u + (ha >> 1) % 51u) << 23) | (hash(ha) & 0x7fffffu));
            const float q = fdiv_row(a, d), ref = __fdiv_rn(a, n);
            mism += __float_as_uint(q) != __float_as_uint(ref);
            ++cnt;
        }
    }
    atomicAdd(bad, mism);
    atomicAdd(total, cnt);
}

int main() {
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    check<<<256, 256>>>(4096, d, d + 1);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("{\"check\": \"fdiv_row vs __fdiv_rn\", \"quotients\": %llu, \"mismatches\": %llu}\n", h[1], h[0]);
    return h[0] != 0 || h[1] == 0;
}
