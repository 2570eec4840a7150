// DMMA (mma.sync.m8n8k4.f64) dependent-issue latency and per-warp throughput versus
// the number of independent accumulators, 1 and 2 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC, bool VARA>
__global__ void k(double *out, int iters, double a, double b, long long *clk) {
    double c[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) c[i] = threadIdx.x * 1e-3 + i;
    double aa[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) aa[i] = a + i * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            dmma(c[2 * i], c[2 * i + 1], aa[i], b);
            if (VARA) aa[i] = fma(aa[i], 0.999999, 1e-9);   // an FMA feeding the next A operand
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += c[i];
    if (s == 1234.5678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int NACC, bool VARA>
void run(int warps, double *d, long long *dc) {
    int iters = 4000;
    k<NACC, VARA><<<148, warps * 32>>>(d, iters, 0.999999, 1e-9, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  accumulators %d  feedA %d : %.1f clk per DMMA per warp, %.1f clk per DMMA per sub-partition\n", warps, NACC, (int)VARA,
           (double)c / iters / NACC, (double)c / iters / NACC / (warps / 4.0));
}
int main() {
    double *d; long long *dc; cudaMalloc(&d, 64); cudaMalloc(&dc, 64);
    for (int w : {4, 8, 16}) {
        run<1, false>(w, d, dc); run<2, false>(w, d, dc); run<3, false>(w, d, dc); run<4, false>(w, d, dc);
        run<6, false>(w, d, dc); run<8, false>(w, d, dc); run<6, true>(w, d, dc); run<12, false>(w, d, dc);
    }
    return 0;
}
