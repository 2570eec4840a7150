// FP64 pipe micro-benchmark: DFMA throughput per SM versus warps per SM and
// independent chains per thread (ILP).  nvcc -arch=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = (threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 1234.5678) out[0] = s;
}
template <int ILP>
void run(int warps, double *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<ILP><<<148, warps * 32>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<148, warps * 32>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)iters * ILP * warps * 32 * 148;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("warps/SM %2d ILP %2d : %.2f TFLOP/s  %.2f DFMA lanes/clk/SM (at %d MHz), %.1f clk per warp-DFMA per warp\n", warps, ILP,
           2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000,
           (ms * 1e-3) * (clk * 1e3) / ((double)iters * ILP));
}
int main() {
    double *d; cudaMalloc(&d, 64);
    for (int w : {4, 8, 12, 16, 32}) {
        run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d); run<32>(w, d);
    }
    return 0;
}
