// harm_tc32_kernels.cu -- the OPTIONAL reduced-precision form of the harmonic sums
// (GPPD_FP32: "an optional FP32 path within 1e-5, stated separately").
//
// Same sums, same pipeline and same partial-sum layout as harm_tc_kernels.cu, with both
// factors of the contraction  C[64 x 48] = V^T E  carried at float32-class precision:
//   * V and E are computed in float32 (the stream values z = w conj(p)(d - mu) from the
//     float32 VOLT values, the harmonics (cos, sin)(k theta) by float complex products from
//     the row's double-precision basis) and rounded to 23-bit fixed point by one FFMA,
//         v 2^F + (1.5 * 2^23 + 0x8080)  ->  mantissa = X + 0x408080,
//     whose bytes are the balanced digits a_0 + 128, a_1 + 128 and (7 bits) a_2 + 64:
//     three signed-byte digit planes per operand instead of six;
//   * two MMAs per 32 rows instead of four:  [V d2 ; V d1] x [E d0 | E d1 | E d2]  (N = 144:
//     TMEM column block j holds 256^(2+j) in lanes 0..63 and 256^(1+j) in lanes 64..127) and
//     [V d0 ; 0] x [E d1 | E d2]  (N = 96, blocks 3 and 4: 256^1, 256^2).  Only the pair (0, 0)
//     -- 2^-25 of a full-scale product -- is not formed.
// The int32 accumulation stays exact; what is lost is the rounding of the inputs: 2^-18 of
// the largest sampled |V| and 2^-21 absolute for E per element, i.e. ~1e-7 .. 1e-6 on a sum.
// The constant sums (sum w |d|^2, sum z, ...) are accumulated per thread in float32 over the
// 96 rows a thread sees and added up in double.  Everything downstream (harmonic table, the
// FP64 NEWUOA fit, the demodulation) is unchanged, so the fitted parameters carry the error of
// the sums only.  Measured: the harmonic pass of the 100-table night 1.72 -> see DESIGN.md.
#include <cstdlib>

#include "fit_math.cuh"
#include "gppd_device.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "tma.cuh"

namespace gppd {

#ifndef TC_SAMPLE_UNROLL_N
#define TC_SAMPLE_UNROLL_N 1
#endif
constexpr int TC_SAMPLE_UNROLL = TC_SAMPLE_UNROLL_N;   // iterations of the scale sampling in flight

constexpr int T3_SEG_ROWS = 6144;             // = TC_SEG_ROWS: the partial buffers are shared
constexpr int T3_KB = 32;                     // rows per K-block = K of one int8 MMA
#ifndef T3_RS_N
#define T3_RS_N 8
#endif
#ifndef T3_OS_N
#define T3_OS_N 4
#endif
#ifndef T3_EW_N
#define T3_EW_N 3
#endif
#ifndef T3_VSETS_N
#define T3_VSETS_N 2
#endif
constexpr int T3_RS = T3_RS_N;                // raw ring stages
constexpr int T3_OS = T3_OS_N;                // operand ring stages
constexpr int T3_VSETS = T3_VSETS_N;
constexpr int T3_VW = 8 * T3_VSETS, T3_EW = T3_EW_N;
constexpr int T3_MMA_WARP = T3_VW + T3_EW;
constexpr int T3_WARPS = T3_MMA_WARP + 1;
constexpr int T3_THREADS = T3_WARPS * 32;
constexpr int T3_RAW_VOLT = T3_KB * 320;
constexpr int T3_RAW_BYTES = T3_RAW_VOLT + T3_KB * 16;
constexpr int T3_ND = 3;
// operand tiles (MN-major, no swizzle): byte (mn, k) at (mn % 16) + 16 (k % 8) + SBO (mn / 16) + LBO (k / 8)
// V: per 8-row block 16 atoms of 16 values: [digit 2: 4 atoms][digit 1: 4][digit 0: 4][zero: 4]
constexpr int V3_SBO = 160;
constexpr int V3_LBO = 16 * V3_SBO, V3_TILE = 4 * V3_LBO;
// E: per 8-row block 9 atoms: digit j, harmonics 8 a + 1 .. 8 a + 8 at atom 3 j + a
constexpr int E3_SBO = 128;
constexpr int E3_LBO = 3 * T3_ND * E3_SBO, E3_TILE = 4 * E3_LBO;
constexpr int T3_OP_BYTES = V3_TILE + E3_TILE;
constexpr int T3_SMEM = T3_RS * T3_RAW_BYTES + T3_OS * T3_OP_BYTES + 128;
constexpr int T3_TMEM_COLS = 256;             // 5 blocks x 48 columns used
constexpr int T3_EBITS = 21;                  // E = X 2^-21
constexpr int T3_VBITS = 18;                  // sampled max |V| -> below 2^18 (16x headroom in 2^22)
static_assert(T3_SMEM <= 227 * 1024, "shared memory");
static_assert(T3_SEG_ROWS * 16384ll < (1ll << 31), "int32 accumulators: one digit pair per column block");

#define T3_MAGIC 12615808.0f                  // 1.5 * 2^23 + 0x8080
constexpr uint32_t T3_MAGIC_EXP = 0x4B000000u; // sign / exponent bits of a value in [2^23, 2^24)

// digits of four values: byte j of m[i] -> byte i of out[j] as signed digits
__device__ __forceinline__ void t3_digits4(const uint32_t (&m)[4], uint32_t (&out)[T3_ND]) {
    const uint32_t t0 = __byte_perm(m[0], m[1], 0x5140), t1 = __byte_perm(m[2], m[3], 0x5140);
    const uint32_t t2 = __byte_perm(m[0], m[1], 0x0062), t3 = __byte_perm(m[2], m[3], 0x0062);
    out[0] = __byte_perm(t0, t1, 0x5410) ^ 0x80808080u;
    out[1] = __byte_perm(t0, t1, 0x7632) ^ 0x80808080u;
    // top digit: 7 bits, a_2 + 64 (bit 7 of the byte is the exponent's lowest bit)
    out[2] = (__byte_perm(t2, t3, 0x5410) + 0x40404040u) ^ 0x80808080u;
}

__device__ __forceinline__ float2 t3_cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 t3_csqr(float2 a) {
    return make_float2(fmaf(a.x, a.x, -(a.y * a.y)), 2.0f * (a.x * a.y));
}
// unit phasor of the FC sample in float32 (exp(1im*angle(fc)), :388); angle(0) = 0 -> 1
__device__ __forceinline__ float2 t3_fc_unit(float x, float y) {
    const float h2 = fmaf(x, x, y * y);
    if (h2 > 1.0e-30f && h2 < 1.0e30f) {
        float inv = rsqrtf(h2);
        inv = inv * fmaf(-0.5f * h2, inv * inv, 1.5f);      // one Newton step: full float accuracy
        return make_float2(x * inv, y * inv);
    }
    const double2 u = fc_phasor(make_double2((double)x, (double)y));
    return make_float2((float)u.x, (float)u.y);
}

struct T3Shared {
    uint64_t raw_full[T3_RS], raw_empty[T3_RS], op_full[T3_OS], op_empty[T3_OS], acc_full;
    float2 stats[16][NGROUP];      // (mean, weight) [diode * 4 + state][group]
    double2 statsd[16][NGROUP];    // the same in double (constant sums of the epilogue)
    float2 voff[5][NGROUP];        // centres (+ the first sample when the offsets are fitted) of a group's
    float2 vlo[5][NGROUP];         //   4 diodes + FC as float32 pairs hi + lo, [channel][group]
    float vscale[4][NGROUP];       // fixed-point scale, [diode][group]
    double inv[NDIODE];
    double2 mud[NDIODE];           // the job's first sample (fitted offsets), else 0
    unsigned int vmax[NDIODE];     // float bits of the sampled max
    int ovf[NGROUP];
    uint32_t tmem;
};

// stream values of the 4 diodes of a group for one row (float32 form of tc_values)
template <int KIND, bool OFFS, bool ACC, bool FAINT>
__device__ __forceinline__ void t3_values(int st, const float2 (&dd)[4], float2 fcs, const float2 *stats16,
                                          float2 (&vv)[4], float *cst) {
    constexpr int NACC = KIND == 0 ? (OFFS ? 5 : 3) : 2;
    const float2 fc = t3_fc_unit(fcs.x, fcs.y);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        float wpr = fc.x, wpi = fc.y, w = 1.0f;
        if (FAINT) {
            const float2 mw = stats16[(d * 4 + (st & 3)) * NGROUP];
            w = mw.y;
            const float wm = mw.y * mw.x;
            wpr = wm * fc.x;
            wpi = wm * fc.y;
        }
        if (KIND == 0) {
            const float dr = dd[d].x, di = dd[d].y;      // (the first sample is already subtracted)
            vv[d].x = fmaf(wpr, dr, wpi * di);
            vv[d].y = fmaf(wpr, di, -(wpi * dr));
            if (ACC) {
                cst[d * NACC + 0] = fmaf(w, fmaf(dr, dr, di * di), cst[d * NACC + 0]);
                cst[d * NACC + 1] += vv[d].x;
                cst[d * NACC + 2] += vv[d].y;
                if (OFFS) {
                    cst[d * NACC + 3] = fmaf(w, dr, cst[d * NACC + 3]);
                    cst[d * NACC + 4] = fmaf(w, di, cst[d * NACC + 4]);
                }
            }
        } else {
            vv[d].x = wpr;
            vv[d].y = wpi;
            if (ACC) {
                cst[d * NACC + 0] += wpr;
                cst[d * NACC + 1] += wpi;
            }
        }
    }
}

// sample of channel ch at row i minus mu, rounded to float32 (the scale sampling only)
__device__ __forceinline__ float2 t3_sample(const TableView &tv, long long i, int ch, double2 mu) {
    const double2 s = row_sample(tv, i, ch);
    return make_float2((float)(s.x - mu.x), (float)(s.y - mu.y));
}

// V producer: thread = (row 4 wv + r4 of the K-block, group g), K-blocks vset, vset + 2, ...
template <int KIND, bool OFFS, bool FAINT>
__device__ __forceinline__ void t3_v_producer(T3Shared &S, unsigned char *raw_ring, unsigned char *op_ring,
                                              const TableDesc &tb, unsigned flags, long long rbase, int nseg,
                                              int nkb, int warp, int lane, float *cst, unsigned long long &cnt) {
    const int r4 = lane >> 3, g = lane & 7;
    const int vset = warp >> 3, krow = 4 * (warp & 7) + r4;
    const uint32_t b_raw_full = smem_u32(&S.raw_full[0]), b_raw_empty = smem_u32(&S.raw_empty[0]);
    const uint32_t b_op_full = smem_u32(&S.op_full[0]), b_op_empty = smem_u32(&S.op_empty[0]);
    const int8_t *stp = FAINT ? tb.state + rbase + krow : nullptr;
    int st_next = ST_NORMAL;
    if (FAINT && vset * T3_KB + krow < nseg) st_next = stp[vset * T3_KB];
    uint32_t ovf = 0;
    const bool be = tb.tv.big_endian != 0;
    const float2 *off = &S.voff[0][g], *olo = &S.vlo[0][g];     // [d * NGROUP]
    const float *sc = &S.vscale[0][g];
    const unsigned char *rw0 = raw_ring + krow * 320 + 32 * g;
    // digit j of this thread's 8 values: atom (2 - j) * 4 + g / 2, k-row krow, bytes 8 (g % 2) ..
    unsigned char *vt0 = op_ring + (g >> 1) * V3_SBO + (krow >> 3) * V3_LBO + (krow & 7) * 16 + 8 * (g & 1);
#pragma unroll 1
    for (int kb = vset; kb < nkb; kb += T3_VSETS) {
        const int rs = kb % T3_RS, os = kb % T3_OS;
        const int i = kb * T3_KB + krow;
        const int st = st_next;
        if (FAINT && i + T3_VSETS * T3_KB < nseg) st_next = stp[(long long)(kb + T3_VSETS) * T3_KB];
        tc_wait(b_raw_full + 8 * rs, (kb / T3_RS) & 1);
        const unsigned char *rw = rw0 + rs * T3_RAW_BYTES;
        const uint4 wa = *reinterpret_cast<const uint4 *>(rw);
        const uint4 wb = *reinterpret_cast<const uint4 *>(rw + 16);
        uint2 wf = *reinterpret_cast<const uint2 *>(rw + 256 - 24 * g);     // row + 256 + 8 g
        uint32_t w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        if (be) {
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = bswap32(w[k]);
            wf.x = bswap32(wf.x);
            wf.y = bswap32(wf.y);
        }
        float2 vv[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) vv[d] = make_float2(0.0f, 0.0f);
#ifdef T3_SKIP_V
        const bool valid = false;
#else
        const bool valid = i < nseg && (!FAINT || row_valid(st, flags));
#endif
        if (valid) {
            float2 dd[4];
#pragma unroll
            for (int d = 0; d < 4; ++d)
                dd[d] = make_float2((__uint_as_float(w[2 * d]) - off[d * NGROUP].x) - olo[d * NGROUP].x,
                                    (__uint_as_float(w[2 * d + 1]) - off[d * NGROUP].y) - olo[d * NGROUP].y);
            const float2 fcs = make_float2((__uint_as_float(wf.x) - off[4 * NGROUP].x) - olo[4 * NGROUP].x,
                                           (__uint_as_float(wf.y) - off[4 * NGROUP].y) - olo[4 * NGROUP].y);
            if (FAINT) cnt += 1ull << (16 * (st & 3));     // (bright: every row of the segment, set by the caller)
            t3_values<KIND, OFFS, true, FAINT>(st, dd, fcs, &S.stats[0][g], vv, cst);
        }
        uint32_t m[8];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            m[2 * d] = __float_as_uint(fmaf(vv[d].x, sc[d * NGROUP], T3_MAGIC));
            m[2 * d + 1] = __float_as_uint(fmaf(vv[d].y, sc[d * NGROUP], T3_MAGIC));
            ovf |= (m[2 * d] ^ T3_MAGIC_EXP) | (m[2 * d + 1] ^ T3_MAGIC_EXP);
        }
        uint32_t dlo[T3_ND], dhi[T3_ND];
        {
            const uint32_t m0[4] = {m[0], m[1], m[2], m[3]}, m1[4] = {m[4], m[5], m[6], m[7]};
            t3_digits4(m0, dlo);
            t3_digits4(m1, dhi);
        }
        __syncwarp();                                   // every lane has consumed its raw bytes
        if (lane == 0) tc_arrive(b_raw_empty + 8 * rs);
        tc_wait(b_op_empty + 8 * os, ((kb / T3_OS) & 1) ^ 1);
        unsigned char *vt = vt0 + os * T3_OP_BYTES;
#pragma unroll
        for (int j = 0; j < T3_ND; ++j)
            *reinterpret_cast<uint2 *>(vt + ((2 - j) * 4) * V3_SBO) = make_uint2(dlo[j], dhi[j]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) tc_arrive(b_op_full + 8 * os);
    }
    if (ovf & 0xff800000u) S.ovf[g] = 1;             // a value left [2^23, 2^24): it did not fit 2^22
}

template <int KIND, bool OFFS>
__global__ void __launch_bounds__(T3_THREADS, 1)
k_harm_tc32(const TableDesc *tabs, const JobInfo *jobs, unsigned flags, int P, const double *stats,
            double *partial) {
    extern __shared__ __align__(128) unsigned char t3_smem[];
    __shared__ __align__(16) T3Shared S;

    constexpr int NCONST = KIND == 0 ? 7 : 2;
    constexpr int NACC = KIND == 0 ? (OFFS ? 5 : 3) : 2;
    constexpr int HP = KIND == 0 ? HP_Z : HP_Y;
    const int job = blockIdx.x, p = blockIdx.y;
    const JobInfo ji = jobs[job];
    const long long seg0 = (long long)p * T3_SEG_ROWS;
    if (seg0 >= ji.nrows) return;
    const TableDesc tb = tabs[ji.table];
    const TableView &tv = tb.tv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nseg = (int)((ji.nrows - seg0) < T3_SEG_ROWS ? (ji.nrows - seg0) : T3_SEG_ROWS);
    const int nkb = (nseg + T3_KB - 1) / T3_KB;
    const bool faint = tb.state != nullptr;
    const long long rbase = ji.row0 + seg0;           // first table row of the segment

    unsigned char *raw_ring = t3_smem;
    unsigned char *op_ring = t3_smem + T3_RS * T3_RAW_BYTES;

    // ---- set-up: barriers, TMEM, tables, the zero atoms of the V tiles ----------------------
    if (threadIdx.x == 0) {
        for (int i = 0; i < T3_RS; ++i) {
            mbar_init(&S.raw_full[i], 1);
            mbar_init(&S.raw_empty[i], 8 + 1);        // 8 V warps + 1 E warp per K-block
        }
        for (int i = 0; i < T3_OS; ++i) {
            mbar_init(&S.op_full[i], 8 + 1);
            mbar_init(&S.op_empty[i], 1);
        }
        mbar_init(&S.acc_full, 1);
    }
    if (warp == T3_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&S.tmem)),
                     "n"(T3_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (threadIdx.x < NGROUP * 16) {
        const int g = threadIdx.x >> 4, k = threadIdx.x & 15;
        const double2 mw = faint ? stats_mean_weight(stats, job * NGROUP + g, k >> 2, k & 3) : make_double2(1.0, 1.0);
        S.statsd[k][g] = mw;
        S.stats[k][g] = make_float2((float)mw.x, (float)mw.y);
    } else if (threadIdx.x < NGROUP * 16 + NCHAN) {
        const int ch = threadIdx.x - NGROUP * 16;
        // what a sample is measured from: the centre, plus the job's first sample when the offsets
        // are fitted, as float32 hi + lo: (raw - hi) - lo rounds at the size of the difference
        double2 o = tv.offsets ? __ldg(tv.offsets + ch) : make_double2(0.0, 0.0);
        if (ch < 32) {
            const double2 m = OFFS ? row_sample(tv, ji.row0, ch) : make_double2(0.0, 0.0);
            S.mud[ch] = m;
            o.x += m.x;
            o.y += m.y;
        }
        const float2 hi = make_float2((float)o.x, (float)o.y);
        const float2 lo = make_float2((float)(o.x - (double)hi.x), (float)(o.y - (double)hi.y));
        if (ch < 32) { S.voff[ch & 3][ch >> 2] = hi; S.vlo[ch & 3][ch >> 2] = lo; }
        else { S.voff[4][ch - 32] = hi; S.vlo[4][ch - 32] = lo; }
    } else if (threadIdx.x < NGROUP * 16 + NCHAN + NDIODE) {
        S.vmax[threadIdx.x - NGROUP * 16 - NCHAN] = 0u;
    } else if (threadIdx.x < NGROUP * 16 + NCHAN + NDIODE + NGROUP) {
        S.ovf[threadIdx.x - NGROUP * 16 - NCHAN - NDIODE] = 0;
    }
    // the zero half of the second MMA's A operand: atoms 12..15 of every 8-row block of every stage
    for (int q = threadIdx.x; q < T3_OS * 4 * 4 * (128 / 16); q += T3_THREADS) {
        const int os = q / 128, kb8 = (q / 32) & 3, atom = (q / 8) & 3, w16 = q & 7;
        *reinterpret_cast<uint4 *>(op_ring + os * T3_OP_BYTES + kb8 * V3_LBO + (12 + atom) * V3_SBO + 16 * w16) =
            make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem;

    const int r4 = lane >> 3, g = lane & 7;
    if (warp < 8) {
        // 128 rows spread over the segment: the largest |V| component of each diode
        float mx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const int r = threadIdx.x >> 3;
#pragma unroll TC_SAMPLE_UNROLL
        for (int it = 0; it < 4; ++it) {
            const int i = (int)(((long long)(it * 32 + r) * nseg) >> 7);
            const long long row = rbase + i;
            const int st = faint ? tb.state[row] : ST_NORMAL;
            if (faint && !row_valid(st, flags)) continue;
            float2 dd[4], vv[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) dd[d] = t3_sample(tv, row, g * 4 + d, S.mud[g * 4 + d]);
            const float2 fcs = t3_sample(tv, row, fc_channel(g), make_double2(0.0, 0.0));
            if (faint) t3_values<KIND, OFFS, false, true>(st, dd, fcs, &S.stats[0][g], vv, nullptr);
            else t3_values<KIND, OFFS, false, false>(st, dd, fcs, &S.stats[0][g], vv, nullptr);
#pragma unroll
            for (int d = 0; d < 4; ++d) mx[d] = fmaxf(mx[d], fmaxf(fabsf(vv[d].x), fabsf(vv[d].y)));
        }
        if (faint && r == 0) {
            // FAINT: rows of a state that the sample missed: bound |V| from the per-state table,
            // |d| <= mean + 8 sigma
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float amu = OFFS ? (float)hypot(S.mud[g * 4 + d].x, S.mud[g * 4 + d].y) : 0.0f;
                for (int st = 0; st < 4; ++st) {
                    const float2 mw = S.stats[d * 4 + st][g];
                    if (!(mw.y > 0.0f && mw.y < 1.0e30f && mw.x >= 0.0f && mw.x < 1.0e30f)) continue;
                    const float wm = mw.y * mw.x;
                    const float bound = KIND == 0 ? wm * (mw.x + amu + 8.0f * rsqrtf(mw.y)) : wm;
                    mx[d] = fmaxf(mx[d], bound);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (mx[d] > 0.0f && mx[d] < 1.0e30f) atomicMax(&S.vmax[g * 4 + d], __float_as_uint(mx[d]));
    }
    __syncthreads();
    if (threadIdx.x < NDIODE) {
        const float m = __uint_as_float(S.vmax[threadIdx.x]);
        int e = 0;
        if (m > 0.0f) frexpf(m, &e);                      // m < 2^e
        int F = T3_VBITS - e;
        F = F > 100 ? 100 : (F < -100 ? -100 : F);        // (float exponent range)
        S.vscale[threadIdx.x & 3][threadIdx.x >> 2] = scalbnf(1.0f, F);
        S.inv[threadIdx.x] = scalbn(1.0, -(F + T3_EBITS));
    }
    __syncthreads();

    float cst[NACC * 4];
#pragma unroll
    for (int q = 0; q < NACC * 4; ++q) cst[q] = 0.0f;
    unsigned long long cnt = 0;

    const char *volt = reinterpret_cast<const char *>(tv.volt);
    auto load = [&](int kb) {
        const int rs = kb % T3_RS;
        mbar_wait(&S.raw_empty[rs], ((kb / T3_RS) & 1) ^ 1);
        const int rows = min(T3_KB, nseg - kb * T3_KB);
        const long long row = rbase + (long long)kb * T3_KB;
        unsigned char *dst = raw_ring + rs * T3_RAW_BYTES;
        if (lane == 0) {
            mbar_expect_tx(&S.raw_full[rs], (unsigned)rows * 336u);
            bulk_g2s(dst, volt + row * 320, (unsigned)rows * 320u, &S.raw_full[rs]);
        } else if (lane == 1) {
            bulk_g2s(dst + T3_RAW_VOLT, tb.basis + row, (unsigned)rows * 16u, &S.raw_full[rs]);
        }
        __syncwarp();
    };
    constexpr int T3_LEAD = T3_RS - T3_EW;

    if (warp == T3_MMA_WARP) {
        // ---- control warp: MMA issuer (and loader of the first T3_RS K-blocks) ------------------
        constexpr uint32_t ID144 = tc_idesc(144), ID96 = tc_idesc(96);
        const uint32_t b_op_full = smem_u32(&S.op_full[0]);
        const uint64_t dv0 = tc_desc(smem_u32(op_ring), V3_LBO, V3_SBO);
        const uint64_t de0 = tc_desc(smem_u32(op_ring) + V3_TILE, E3_LBO, E3_SBO);
        for (int kb = 0; kb < min(T3_RS, nkb); ++kb) load(kb);
        for (int kb = 0; kb < nkb; ++kb) {
            const int os = kb % T3_OS;
            tc_wait(b_op_full + 8 * os, (kb / T3_OS) & 1);
            tc_fence_after();
            if (tc_elect()) {
                const uint64_t so = (uint64_t)((os * T3_OP_BYTES) >> 4);
                const uint64_t av = dv0 + so, be = de0 + so;
                // [d2 ; d1] x E digits 0..2 -> blocks 0..2;  [d0 ; 0] x E digits 1, 2 -> blocks 3, 4
                const uint32_t acc = kb > 0 ? 1u : 0u;
                tc_mma(tmem, av, be, ID144, acc);
                tc_mma(tmem + 144, av + ((8 * V3_SBO) >> 4), be + ((3 * E3_SBO) >> 4), ID96, acc);
                tc_commit(&S.op_empty[os]);
                if (kb == nkb - 1) tc_commit(&S.acc_full);
            }
            __syncwarp();
        }
    } else if (warp >= T3_VW) {
        // ---- E producers: lane = row, K-blocks e, e + T3_EW, ...; they also issue the TMA loads --
        const int e = warp - T3_VW;
        const uint32_t b_raw_full = smem_u32(&S.raw_full[0]), b_raw_empty = smem_u32(&S.raw_empty[0]);
        const uint32_t b_op_full = smem_u32(&S.op_full[0]), b_op_empty = smem_u32(&S.op_empty[0]);
#pragma unroll 1
        for (int kb = e; kb < nkb; kb += T3_EW) {
            const int rs = kb % T3_RS, os = kb % T3_OS;
            if (kb + T3_LEAD >= T3_RS && kb + T3_LEAD < nkb) load(kb + T3_LEAD);
            tc_wait(b_raw_full + 8 * rs, (kb / T3_RS) & 1);
            const uint4 bw = *reinterpret_cast<const uint4 *>(raw_ring + rs * T3_RAW_BYTES + T3_RAW_VOLT + lane * 16);
            // basis = (sin theta, cos theta)
            float2 e1 = make_float2((float)__hiloint2double(bw.w, bw.z), (float)__hiloint2double(bw.y, bw.x));
            if (kb * T3_KB + lane >= nseg) e1 = make_float2(1.0f, 0.0f);
            float2 eh[8];
            eh[0] = e1;
            eh[1] = t3_csqr(e1);
            eh[2] = t3_cmul(eh[1], e1);
            eh[3] = t3_csqr(eh[1]);
            eh[4] = t3_cmul(eh[3], e1);
            eh[5] = t3_csqr(eh[2]);
            eh[6] = t3_cmul(eh[5], e1);
            eh[7] = t3_csqr(eh[3]);
            const float2 e8 = eh[7];
            __syncwarp();
            if (lane == 0) tc_arrive(b_raw_empty + 8 * rs);
            tc_wait(b_op_empty + 8 * os, ((kb / T3_OS) & 1) ^ 1);
            unsigned char *et = op_ring + os * T3_OP_BYTES + V3_TILE + (lane >> 3) * E3_LBO + (lane & 7) * 16;
#ifdef T3_SKIP_E
            for (int a = 0; a < 0; ++a) {
#else
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#endif
                // harmonics 8 a + 1 .. 8 a + 8: 16 values = one 16-byte atom row per digit
                uint32_t dg[4][T3_ND];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t m[4];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const float2 v = eh[2 * q + s];
                        m[2 * s] = __float_as_uint(fmaf(v.x, 2097152.0f, T3_MAGIC));       // 2^21
                        m[2 * s + 1] = __float_as_uint(fmaf(v.y, 2097152.0f, T3_MAGIC));
                    }
                    t3_digits4(m, dg[q]);
                }
#pragma unroll
                for (int j = 0; j < T3_ND; ++j)
                    *reinterpret_cast<uint4 *>(et + (3 * j + a) * E3_SBO) = make_uint4(dg[0][j], dg[1][j], dg[2][j], dg[3][j]);
                if (a < 2) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) eh[q] = t3_cmul(eh[q], e8);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tc_arrive(b_op_full + 8 * os);
        }
    } else {
        if (faint) t3_v_producer<KIND, OFFS, true>(S, raw_ring, op_ring, tb, flags, rbase, nseg, nkb, warp, lane, cst, cnt);
        else t3_v_producer<KIND, OFFS, false>(S, raw_ring, op_ring, tb, flags, rbase, nseg, nkb, warp, lane, cst, cnt);
        // bright tables: all nseg rows are valid and NORMAL; one lane per group carries the count
        if (!faint) cnt = (warp == 0 && r4 == 0) ? (unsigned long long)nseg << (16 * (ST_NORMAL & 3)) : 0ull;
    }
    __syncthreads();      // every producer is done with the raw ring: it now holds the reductions

    double *s_red = reinterpret_cast<double *>(raw_ring);                          // [T3_VW][NGROUP][20]
    unsigned long long *s_cnt = reinterpret_cast<unsigned long long *>(s_red + T3_VW * NGROUP * 20);
    if (warp < T3_VW) {
        // constant sums: the 4 row lanes of a group (in double from here on), then the V warps in order
#pragma unroll
        for (int q = 0; q < NACC * 4; ++q) {
            double sv = (double)cst[q];
            sv += __shfl_xor_sync(0xffffffffu, sv, 8);
            sv += __shfl_xor_sync(0xffffffffu, sv, 16);
            if (r4 == 0) s_red[(warp * NGROUP + g) * 20 + q] = sv;
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
        if (r4 == 0) s_cnt[warp * NGROUP + g] = cnt;
    }
    __syncthreads();

    // ---- epilogue: TMEM -> FP64 sums -> the partial layout of k_harm_ws ---------------------
    double *stage = reinterpret_cast<double *>(op_ring);         // [64][49] second-half sums
    if (warp < 4) {
        mbar_wait(&S.acc_full, 0);
        tc_fence_after();
        const int tl = warp * 32 + lane;                          // TMEM lane
        const int h = tl >> 6, c = tl & 63;                       // digit half (0: d2 / d0, 1: d1), V column
        const int cg = c >> 3, d = (c >> 1) & 3, im = c & 1;
        const double inv = S.inv[cg * 4 + d];
        double *out = partial + ((long long)(job * NGROUP + cg) * P + p) * 4 * HP + d * HP + NCONST;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {                             // 16 of the 48 E columns at a time
            double acc[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = 0.0;
#pragma unroll
            for (int b = 0; b < (h ? 3 : 5); ++b) {
                // blocks 0..2: V digit 2 (lanes 0..63) / 1 (64..127) x E digit b; 3, 4: V digit 0 x E digit b - 2
                const double wgt = scalbn(1.0, 8 * (b < 3 ? b + (h ? 1 : 2) : b - 2));
                uint32_t v[16];
                tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * 48 + q * 16), v);
#pragma unroll
                for (int t = 0; t < 16; ++t) acc[t] = fma((double)(int)v[t], wgt, acc[t]);
            }
            if (h) {
#pragma unroll
                for (int t = 0; t < 16; ++t) stage[c * 49 + q * 16 + t] = acc[t];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (!h) {
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int m = q * 16 + t, k = m >> 1, sn = m & 1;
                    const int slot = sn ? (im ? 1 : 3) : (im ? 2 : 0);
                    out[k * 4 + slot] = (acc[t] + stage[c * 49 + m]) * inv;
                }
            }
        }
        tc_fence_before();
    }
    // constants: thread = (group, diode, constant)
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NGROUP * 4 * NCONST) {
        const int t = threadIdx.x - 128;
        const int cg = t / (4 * NCONST), d = (t / NCONST) & 3, cc = t % NCONST;
        unsigned long long cn = 0;
        for (int w = 0; w < T3_VW; ++w) cn += s_cnt[w * NGROUP + cg];
        double sv = 0.0;
        int src = -1;
        if (KIND == 0) {
            if (cc == 1) src = 0;
            else if (cc == 5) src = 1;
            else if (cc == 6) src = 2;
            else if (OFFS && cc == 3) src = 3;
            else if (OFFS && cc == 4) src = 4;
        } else {
            src = cc;
        }
        if (src >= 0) {
            for (int w = 0; w < T3_VW; ++w) sv += s_red[(w * NGROUP + cg) * 20 + d * NACC + src];
        } else if (cc == 0 || cc == 2) {
            for (int st = 0; st < 4; ++st) {
                const double n_s = (double)((cn >> (16 * st)) & 0xffffull);
                const double2 mw = S.statsd[d * 4 + st][cg];
                if (n_s > 0.0) sv += cc == 0 ? n_s * mw.y : n_s * (mw.y * (mw.x * mw.x));
            }
        }
        if (S.ovf[cg] && cc == (KIND == 0 ? 1 : 0)) sv = __longlong_as_double(0x7ff8000000000000ll);
        partial[((long long)(job * NGROUP + cg) * P + p) * 4 * HP + d * HP + cc] = sv;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T3_MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(T3_TMEM_COLS));
}

void launch_harmonics_tc32(const Launcher &L, const TableDesc *d_tabs, const JobInfo *d_jobs, int njobs,
                           unsigned flags, int P, const double *d_spart2, double *d_partZ, double *d_partY) {
    cudaFuncSetAttribute(k_harm_tc32<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    cudaFuncSetAttribute(k_harm_tc32<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    cudaFuncSetAttribute(k_harm_tc32<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM);
    dim3 grid(njobs, P);
    if (flags & 2u) {
        k_harm_tc32<0, true><<<grid, T3_THREADS, T3_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partZ);
        k_harm_tc32<1, true><<<grid, T3_THREADS, T3_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partY);
        *L.counter += 2;
    } else {
        k_harm_tc32<0, false><<<grid, T3_THREADS, T3_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partZ);
        *L.counter += 1;
    }
}

}  // namespace gppd
