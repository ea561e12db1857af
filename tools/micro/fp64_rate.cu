// Micro-benchmark: sustained fp64 FMA rate and dependent-FMA latency of one B200 (context for the roofline discussion in
// the tracking kernels compute in fp64).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(const char* name, int blocks, int threads, int iters) {
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * threads * iters * ILP;
    printf("%-44s %8.3f ms  %8.2f TFLOP/s fp64  (%.2f cycles per warp-FMA per SM sub-partition at 1.965 GHz)\n", name, ms, 2.0 * fmas / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (fmas / 32.0 / (148.0 * 4.0)));
    cudaFree(out);
}

int main() {
    run<8>("148x8 CTAs x 256 thr, 8 independent chains", 148 * 8, 256, 1 << 14);
    run<1>("148x8 CTAs x 256 thr, 1 dependent chain", 148 * 8, 256, 1 << 15);
    run<1>("148 CTAs x 32 thr (1 warp/SM), dependent chain", 148, 32, 1 << 16);
    run<4>("148 CTAs x 224 thr (7 warps/SM), 4 chains", 148, 224, 1 << 15);
    return 0;
}
