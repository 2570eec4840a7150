// The harmonic pass's consumer loop in isolation: 8 warps per SM run k-steps of
// 4 LDS.64 + 7 FP64 ops + 6 DMMA from a static shared-memory tile; optionally 8 more
// warps spin on an mbarrier (what idle producers do).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TR = 256, TRP = TR + 4, MT = 6;
struct Tile { double e[8][TRP]; double v[8][TRP]; };
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>   // 0: full loop, 1: no recurrence (A from smem), 2: DMMA only (operands loaded once)
__global__ void __launch_bounds__(512, 1) k(double *out, int tiles, int spin, long long *clk) {
    extern __shared__ __align__(16) unsigned char sm[];
    Tile *T = reinterpret_cast<Tile *>(sm);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 3 * sizeof(Tile));
    for (int i = threadIdx.x; i < 3 * (int)sizeof(Tile) / 8; i += blockDim.x)
        reinterpret_cast<double *>(sm)[i] = 1e-3 * (i % 97);
    if (threadIdx.x == 0) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 8) {   // spinning "producers"
        if (!spin) return;
        uint32_t a = (uint32_t)__cvta_generic_to_shared(bar), ok = 0;
        while (!ok) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(a), "r"(0u) : "memory");
        }
        return;
    }
    const int m8 = lane >> 2, r4 = lane & 3;
    double c[MT][2];
    for (int j = 0; j < MT; ++j) c[j][0] = c[j][1] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
        const Tile &TT = T[it % 3];
#pragma unroll 4
        for (int ks = 0; ks < 8; ++ks) {
            const int row = warp * 32 + ks * 4 + r4;
            if (MODE == 0) {
                const double a0 = TT.e[m8][row], c4 = TT.e[6][row], pv = TT.e[(m8 + 3) & 7][row], bv = TT.v[m8][row];
                const double tc = c4 + c4;
                double am = fma(0.5, pv, 0.25), ak = a0;
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    dmma(c[j][0], c[j][1], ak, bv);
                    if (j + 1 < MT) { const double an = fma(tc, ak, -am); am = ak; ak = an; }
                }
            } else if (MODE == 1) {
                const double bv = TT.v[m8][row];
                double av[MT];
#pragma unroll
                for (int j = 0; j < MT; ++j) av[j] = TT.e[(m8 + j) & 7][row];
#pragma unroll
                for (int j = 0; j < MT; ++j) dmma(c[j][0], c[j][1], av[j], bv);
            } else {
                const double bv = 0.5 + row, av = 0.25;
#pragma unroll
                for (int j = 0; j < MT; ++j) dmma(c[j][0], c[j][1], av, bv);
            }
        }
    }
    long long t1 = clock64();
    if (spin && threadIdx.x == 0) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
    }
    double s = 0;
    for (int j = 0; j < MT; ++j) s += c[j][0] + c[j][1];
    if (s == 1234.5678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE>
void run(int spin, double *d, long long *dc) {
    int tiles = 400, smem = 3 * sizeof(Tile) + 64;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE><<<148, 512, smem>>>(d, tiles, spin, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d spin %d: %s  %.1f clk per k-step per warp = %.1f clk per DMMA per sub-partition\n", MODE, spin,
           cudaGetErrorString(e), (double)c / tiles / 8, (double)c / tiles / 8 / 6 / 2);
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *b, unsigned ph) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
    } while (!ok);
}
// EPOW: consumers compute the k=2..4 powers per tile; HANDOFF: full/empty mbarriers with 8 producer warps;
// PWORK: producers write 10 doubles per row into the tile (STS) per tile
template <bool EPOW, bool HANDOFF, bool PWORK>
__global__ void __launch_bounds__(512, 1) k2(double *out, int tiles, long long *clk) {
    extern __shared__ __align__(16) unsigned char sm[];
    Tile *T = reinterpret_cast<Tile *>(sm);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + 3 * sizeof(Tile)), *empty = full + 3;
    for (int i = threadIdx.x; i < 3 * (int)sizeof(Tile) / 8; i += blockDim.x)
        reinterpret_cast<double *>(sm)[i] = 1e-3 * (i % 97);
    if (threadIdx.x == 0) for (int b = 0; b < 3; ++b) { mb_init(&full[b], 8); mb_init(&empty[b], 8); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 8) {
        if (!HANDOFF) return;
        const int rr = threadIdx.x - 256;
        for (int it = 0; it < tiles; ++it) {
            const int b = it % 3;
            if (it >= 3) mb_wait(&empty[b], (unsigned)(it / 3 - 1) & 1u);
            if (PWORK) {
                Tile &TT = T[b];
                double x = 1e-3 * (rr + it);
#pragma unroll
                for (int q = 0; q < 2; ++q) TT.e[q][rr] = x * (q + 1);
#pragma unroll
                for (int q = 0; q < 8; ++q) TT.v[q][rr] = x + q;
            }
            __syncwarp();
            if (lane == 0) mb_arrive(&full[b]);
        }
        return;
    }
    const int m8 = lane >> 2, r4 = lane & 3;
    double c[MT][2];
    for (int j = 0; j < MT; ++j) c[j][0] = c[j][1] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
        const int b = it % 3;
        Tile &TT = T[b];
        if (HANDOFF) mb_wait(&full[b], (unsigned)(it / 3) & 1u);
        if (EPOW) {
            const int row = warp * 32 + lane;
            const double cx = TT.e[0][row], sx = TT.e[1][row];
            const double c2 = fma(cx, cx, -(sx * sx)), s2 = 2.0 * (cx * sx);
            const double c3 = fma(c2, cx, -(s2 * sx)), s3 = fma(c2, sx, s2 * cx);
            const double c4 = fma(c2, c2, -(s2 * s2)), s4 = 2.0 * (c2 * s2);
            TT.e[2][row] = c2; TT.e[3][row] = s2; TT.e[4][row] = c3; TT.e[5][row] = s3; TT.e[6][row] = c4; TT.e[7][row] = s4;
            __syncwarp();
        }
#pragma unroll 4
        for (int ks = 0; ks < 8; ++ks) {
            const int row = warp * 32 + ks * 4 + r4;
            const double a0 = TT.e[m8][row], c4 = TT.e[6][row], pv = TT.e[(m8 + 3) & 7][row], bv = TT.v[m8][row];
            const double tc = c4 + c4;
            double am = fma(0.5, pv, 0.25), ak = a0;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                dmma(c[j][0], c[j][1], ak, bv);
                if (j + 1 < MT) { const double an = fma(tc, ak, -am); am = ak; ak = an; }
            }
        }
        if (HANDOFF) { __syncwarp(); if (lane == 0) mb_arrive(&empty[b]); }
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < MT; ++j) s += c[j][0] + c[j][1];
    if (s == 1234.5678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <bool EPOW, bool HANDOFF, bool PWORK>
void run2(double *d, long long *dc) {
    int tiles = 400, smem = 3 * sizeof(Tile) + 64;
    cudaFuncSetAttribute(k2<EPOW, HANDOFF, PWORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k2<EPOW, HANDOFF, PWORK><<<148, 512, smem>>>(d, tiles, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("epow %d handoff %d pwork %d: %s  %.0f clk per tile = %.1f clk per DMMA per sub-partition\n", (int)EPOW, (int)HANDOFF, (int)PWORK,
           cudaGetErrorString(e), (double)c / tiles, (double)c / tiles / 8 / 6 / 2);
}
int main() {
    double *d; long long *dc; cudaMalloc(&d, 64); cudaMalloc(&dc, 64);
    run<0>(0, d, dc);
    run2<false,false,false>(d, dc); run2<true,false,false>(d, dc); run2<true,true,false>(d, dc); run2<true,true,true>(d, dc); run2<false,true,true>(d, dc);
    return 0;
}
