// Is the FP64 tensor path (mma.sync m8n8k4 f64) a pipe of its own on B200, and how
// fast is it next to DFMA?  nvcc -arch=sm_100a -O3 -o dmma_ubench dmma_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NM, int NF>
__global__ void k(double *out, int iters, double a, double b) {
    double c[NM > 0 ? 2 * NM : 2];
    double x[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < 2 * NM; ++i) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NM; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += x[i];
    if (s == 1234.5678) out[0] = s;
}
template <int NM, int NF>
void run(int warps, double *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<NM, NF><<<148, warps * 32>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<NM, NF><<<148, warps * 32>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl_m = 2.0 * 256 * NM * (double)iters * warps * 148;      // 8x8x4 FMA per warp mma
    double fl_f = 2.0 * 32 * NF * (double)iters * warps * 148;
    printf("warps/SM %2d  mma/iter %2d dfma/iter %2d : DMMA %.2f TF/s  DFMA %.2f TF/s  total %.2f\n", warps, NM, NF,
           fl_m / ms / 1e9, fl_f / ms / 1e9, (fl_m + fl_f) / ms / 1e9);
}
int main() {
    double *d; cudaMalloc(&d, 64);
    for (int w : {4, 8, 16, 32}) {
        run<8, 0>(w, d); run<0, 16>(w, d); run<8, 8>(w, d); run<8, 16>(w, d); run<4, 16>(w, d); run<8, 32>(w, d);
    }
    return 0;
}
