// i8mma_test.cu -- known-answer test + issue-rate measurement of the int8 tcgen05 MMA
// (kind::i8, M = 64, N = 48, K = 32, both operands MN-major, no swizzle) in the shared
// memory layout the harmonic-sum kernel uses.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8mma_test i8mma_test.cu
//   ./i8mma_test [swap_lbo_sbo]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 64, N = 48, KB = 32;          // one MMA
constexpr int A_SBO = 160, A_LBO = 4 * A_SBO;   // bytes: atom stride along M, 8-row block stride along K
constexpr int B_SBO = 128, B_LBO = 3 * B_SBO;
constexpr int A_TILE = 4 * A_LBO, B_TILE = 4 * B_LBO;

__host__ __device__ inline int a_off(int m, int k) { return (m & 15) + 16 * (k & 7) + A_SBO * (m >> 4) + A_LBO * (k >> 3); }
__host__ __device__ inline int b_off(int n, int k) { return (n & 15) + 16 * (k & 7) + B_SBO * (n >> 4) + B_LBO * (k >> 3); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .b32 rx;\n .reg .pred px;\n elect.sync rx|px, %1;\n selp.u32 %0, 1, 0, px;\n}\n"
                 : "=r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // base offset 0, SWIZZLE_NONE
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128) k_test(const int8_t *A, const int8_t *B, int nk, int swap, int reps,
                                              int32_t *D, long long *cycles, int dcol = 0) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    unsigned char *sa = smem, *sb = smem + nk * A_TILE;
    for (int i = threadIdx.x; i < nk * A_TILE; i += blockDim.x) sa[i] = 0;
    for (int i = threadIdx.x; i < nk * B_TILE; i += blockDim.x) sb[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nk * KB * M; i += blockDim.x) {
        int k = i / M, m = i % M;
        sa[(k / KB) * A_TILE + a_off(m, k % KB)] = (unsigned char)A[i];     // A[k][m]
    }
    for (int i = threadIdx.x; i < nk * KB * N; i += blockDim.x) {
        int k = i / N, n = i % N;
        sb[(k / KB) * B_TILE + b_off(n, k % KB)] = (unsigned char)B[i];     // B[k][n]
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    // instruction descriptor: D = s32, A/B = signed int8, both MN-major, N = 48, M = 64
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int kb = 0; kb < nk; ++kb) {
                uint64_t da = make_desc(smem_u32(sa + kb * A_TILE), swap ? A_SBO : A_LBO, swap ? A_LBO : A_SBO);
                uint64_t db = make_desc(smem_u32(sb + kb * B_TILE), swap ? B_SBO : B_LBO, swap ? B_LBO : B_SBO);
                mma_i8(tmem + dcol, da, db, idesc, (r > 0 || kb > 0) ? 1u : 0u);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
        } while (!ok);
    }
    if (threadIdx.x == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // TMEM -> registers: lane 32*warp + l, 16 columns at a time
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0 + dcol;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];   // [tmem lane][col]
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(64));
}

// issue-rate sweep: shapes / majors / accumulator rotation (timing only; operands are zero)
__global__ void __launch_bounds__(128) k_rate(int m, int n, int mn_major, int nacc, int reps, int kind_f8,
                                              long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    for (int i = threadIdx.x; i < 65536; i += blockDim.x) smem[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (kind_f8 ? (1u << 4) : (2u << 4) | (1u << 7) | (1u << 10)) | ((uint32_t)mn_major << 15) |
                           ((uint32_t)mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    if (warp == 0 && elect_one()) {
        // MN-major: atoms of 16 along MN (SBO 128), 8-row K blocks (LBO = atoms * 128)
        // K-major: 8-row MN groups (SBO 256), two 16-byte K halves (LBO 128)
        const uint32_t a_lbo = mn_major ? (m / 16) * 128 : 128, a_sbo = mn_major ? 128 : 256;
        const uint32_t b_lbo = mn_major ? (n / 16) * 128 : 128, b_sbo = mn_major ? 128 : 256;
        const uint64_t da = make_desc(smem_u32(smem), a_lbo, a_sbo);
        const uint64_t db = make_desc(smem_u32(smem + 16384), b_lbo, b_sbo);
        uint32_t dcol[8];
        for (int j = 0; j < 8; ++j) dcol[j] = tmem + (uint32_t)((j % nacc) * n);
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (kind_f8)
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                                 " tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol[j]), "l"(da),
                                 "l"(db), "r"(idesc), "r"(1u) : "memory");
                else
                    mma_i8(dcol[j], da, db, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
        } while (!ok);
        cycles[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

int main(int argc, char **argv) {
    const int nk = 4;
    std::vector<int8_t> A(nk * KB * M), B(nk * KB * N);
    srand(7);
    for (auto &x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto &x : B) x = (int8_t)(rand() % 256 - 128);
    std::vector<long long> ref(M * N, 0);
    for (int k = 0; k < nk * KB; ++k)
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) ref[m * N + n] += (long long)A[k * M + m] * B[k * N + n];
    int8_t *dA, *dB;
    int32_t *dD;
    long long *dC;
    CK(cudaMalloc(&dA, A.size()));
    CK(cudaMalloc(&dB, B.size()));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const int smem = nk * (A_TILE + B_TILE);
    for (int swap = 0; swap < 2; ++swap) {
        CK(cudaMemset(dD, 0xff, 128 * N * 4));
        k_test<<<1, 128, smem>>>(dA, dB, nk, swap, 1, dD, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
        std::vector<int32_t> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        // M = 64 accumulator: row m -> TMEM lane (m % 16) + 32 (m / 16)
        long long bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n)
                if (D[((m & 15) + 32 * (m >> 4)) * N + n] != ref[m * N + n]) ++bad;
        long long bad_lin = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n)
                if (D[m * N + n] != ref[m * N + n]) ++bad_lin;
        printf("swap=%d (LBO/SBO %s): mismatches 16x4 layout %lld, linear layout %lld of %d; D[0][0..3] = %d %d %d %d ref %lld %lld %lld %lld\n",
               swap, swap ? "swapped" : "as documented", bad, bad_lin, M * N, D[0], D[1], D[2], D[3], ref[0], ref[1],
               ref[2], ref[3]);
    }
    // issue rate (results overflow; only the time matters)
    for (int reps : {64, 256}) {
        k_test<<<1, 128, smem>>>(dA, dB, nk, 0, reps, dD, dC);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        printf("%d MMAs (64x48x32 i8): %lld cycles, %.1f per MMA\n", reps * nk, cyc, (double)cyc / (reps * nk));
    }
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    struct Cfg { int m, n, mn, nacc, f8; };
    const Cfg cfgs[] = {{64, 48, 1, 1, 0}, {64, 48, 1, 6, 0}, {64, 48, 0, 1, 0}, {64, 48, 0, 6, 0}, {128, 48, 1, 1, 0},
                        {128, 48, 0, 1, 0}, {128, 48, 0, 4, 0}, {64, 96, 1, 1, 0}, {64, 96, 0, 1, 0}, {128, 96, 0, 1, 0}, {128, 128, 0, 1, 0},
                        {128, 256, 0, 1, 0}, {128, 256, 1, 1, 0}, {64, 256, 0, 1, 0}, {64, 48, 0, 1, 1}, {128, 48, 0, 1, 1}, {128, 256, 0, 1, 1}};
    for (const Cfg &c : cfgs) {
        k_rate<<<1, 128, 65536>>>(c.m, c.n, c.mn, c.nacc, 512, c.f8, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate M=%d N=%d: %s\n", c.m, c.n, cudaGetErrorString(e)); return 1; }
        long long cyc;
        CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        printf("%s M=%3d N=%3d K=32 %s-major, %d accumulators: %.1f cycles per MMA (floor max(M,128)*N/256 = %d)\n",
               c.f8 ? "f8" : "i8", c.m, c.n, c.mn ? "MN" : "K", c.nacc, (double)cyc / 512, (c.m > 128 ? c.m : 128) * c.n / 256);
    }
    // accumulator at a column offset that is not a multiple of 4 / 8 / 16
    for (int dcol : {16, 8, 4, 2, 1}) {
        CK(cudaMemset(dD, 0xff, 128 * N * 4));
        k_test<<<1, 128, smem>>>(dA, dB, nk, 0, 1, dD, dC, dcol);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("D column offset %d: CUDA error %s (the accumulator base must be aligned)\n", dcol, cudaGetErrorString(e)); return 0; }
        std::vector<int32_t> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n)
                if (D[((m & 15) + 32 * (m >> 4)) * N + n] != ref[m * N + n]) ++bad;
        printf("D column offset %d: mismatches %lld of %d\n", dcol, bad, M * N);
    }
    return 0;
}
