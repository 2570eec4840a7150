// sqrt4 (csrc/gppd_device.cuh) against sqrt(): bit-for-bit over random and edge-case inputs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gppupildemodulation.jl_b200/csrc tools/sqrt4_test.cu -o tools/sqrt4_test
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "gppd_device.cuh"

__global__ void k_test(unsigned long long seed, long long n, unsigned long long *bad, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    double h[4], r[4];
    for (int k = 0; k < 4; ++k) {
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        if (mode == 0) {            // random bit patterns of non-negative doubles (all exponents, incl. denormal / inf / nan)
            h[k] = __longlong_as_double((long long)(x & 0x7fffffffffffffffull));
        } else if (mode == 1) {     // the range of the kernel: sums of squares of volts
            h[k] = (double)((x >> 11) * (1.0 / 9007199254740992.0)) * 100.0;
        } else {                    // around the fast-path limits and perfect squares
            const unsigned hi = (mode == 2 ? 0x03500000u : 0x7fefffffu - 0x100u) + (unsigned)(x & 0x1ffu) - 0x100u;
            h[k] = __hiloint2double((int)hi, (int)(x >> 32));
        }
        x += 0x9E3779B97F4A7C15ull;
    }
    if (mode == 1 && (i & 7) == 0) h[1] = 0.0;
    gppd::sqrt4(h, r);
    for (int k = 0; k < 4; ++k) {
        const double ref = sqrt(h[k]);
        const bool same = __double_as_longlong(ref) == __double_as_longlong(r[k]) || (ref != ref && r[k] != r[k]);
        if (!same) atomicAdd(bad, 1ull);
    }
}

int main() {
    unsigned long long *bad;
    cudaMallocManaged(&bad, 8);
    *bad = 0;
    const long long n = 1ll << 26;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 4; ++rep) k_test<<<(unsigned)(n / 256), 256>>>(1234567ull + 977ull * rep + 31ull * mode, n, bad, mode);
        cudaDeviceSynchronize();
        printf("mode %d: %lld x 4 values x 4 seeds, mismatches so far %llu\n", mode, n, *bad);
    }
    printf("%s\n", *bad == 0 ? "sqrt4 == sqrt: OK" : "MISMATCH");
    return *bad != 0;
}
