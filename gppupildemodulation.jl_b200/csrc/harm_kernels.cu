// harm_kernels.cu -- the Jacobi-Anger ("harmonic") form of the objective.
//
// The reference evaluates, ~35 times per fit, sums over all rows of
//     conj(e_n) z_n,   e_n = exp(j b sin(theta_n + q))
// (reference src/Modulation.jl:137-145; z_n = w_n conj(p_n) d_n).  With
//     exp(j b sin x) = sum_k J_k(b) exp(j k x)
// the row sums factor:  sum_n conj(e_n) z_n = sum_k J_k(b) e^{-jkq} Z_k,
//     Z_k = sum_n z_n exp(-j k theta_n),   k = -HK..HK
// so ONE pass over the rows (this file) yields 2*HK+1 complex harmonics per fit
// and every later objective call costs O(HK) instead of O(N) (fit_kernels.cu,
// k_fit_harmonic).  Truncation at HK = 24: |error| <= 2 sum_{k>24} |J_k(b)| < 1.3e-15
// for |b| <= 5; larger |b| (or a job without a uniform phase quantum) falls back to
// the direct evaluator.
//
// Z_k and Z_{-k} share four real sums (c = cos k theta, s = sin k theta, z = x + j y):
//     A = sum c x, B = sum s y, C = sum c y, D = sum s x
//     Z_k = (A + B) + j (C - D),   Z_{-k} = (A - B) + j (C + D)
// i.e. 4 FMAs per (row, diode, |k|) for two harmonics.
//
// Per group (4 diodes) and row tile the sums are the product
//     C[48 x 8] += E^T[48 x rows] V[rows x 8],
//     E[row][2(k-1) + {0,1}] = (cos, sin)(k theta_row), k = 1..24,
//     V[row][2d + {0,1}]     = (x, y) of diode d's stream value,
// whose entries are exactly the four sums: (cos,x) = A, (sin,y) = B, (cos,y) = C,
// (sin,x) = D.  The FP64 units take this contraction as warp-level
// mma.sync.m8n8k4.f64 (DMMA: one instruction = 256 FMAs on the same FP64 units as
// DFMA, measured 37.1 against 33.9 TFLOP/s), which removes the instruction-issue and
// register pressure that kept the plain-FMA form near half of the pipe's peak.  E is
// never stored: the lane that owns element (m, row) of an A fragment generates the six
// harmonics k0, k0 + 4, ..., k0 + 20 it needs with the three-term recurrence
//     trig((k + 4) t) = 2 cos(4 t) trig(k t) - trig((k - 4) t)      (one FMA each)
// from the row's (cos, sin)(t .. 4t), which the producers tabulate.
//
// Kernel structure (k_harm_ws): one block of 8 warps per (job, group, 6 144-row
// segment), two blocks per SM.  The kernel is WARP-SYNCHRONOUS: every warp owns the
// 32-row chunks warp, warp + 8, ... of the segment and does everything for them --
//   1. stages the raw table bytes of its next chunk with cp.async (16-byte copies: the
//      group's 8 VOLT floats, its FC pair, the row's basis, the state byte; double
//      buffered, no register dependency, so the global-memory latency is off the
//      critical path),
//   2. "produces" a private 32-row compute tile, one row per lane: (cos, sin)(k theta),
//      k = 1..4, and the four diodes' z (or y) values, plus the constant sums,
//   3. "consumes" it: 8 k-steps of 4 rows, per k-step 5 recurrence FMAs and 6 DMMAs
//      into the 12 accumulator registers that hold the warp's 48 x 8 partial C --
// with only __syncwarp in the main loop.  While some warps of a sub-partition are in
// their latency-bound produce phase the others keep the FP64 units busy with DMMAs
// (ncu: the FP64 "shared" pipe 75 % busy, DMMA 56 %).  An earlier version with
// dedicated producer and consumer warps handing tiles over through mbarriers was
// slower (3.1 ms against 2.96 ms per 100-table night): the producers' in-order FP64
// chains queued behind the consumers' 16-cycle DMMAs and stalled the hand-over.
// A segment is a FIXED run of rows of the job (independent of the batch and of the
// launch shape), the warps' partial C are added in warp order, and k_harm_reduce adds
// the segments in index order, so a fit's sums -- hence its whole NEWUOA trajectory --
// do not depend on what else is in the batch.
#include <cstdlib>
#include <type_traits>

#include "fit_math.cuh"
#include "gppd_device.cuh"
#include "kernels.h"

namespace gppd {

constexpr int HARM_SEG_ROWS = 6144;           // rows per segment
constexpr int MTILES = 2 * HK / 8;            // 8-row tiles of the 48 (harmonic, cos|sin) rows of C
static_assert(2 * HK == 8 * MTILES && MTILES == 6, "lane k0 + 4j must cover k = 1..HK");

__host__ __device__ inline int harm_segments(long long nrows) {
    return (int)((nrows + HARM_SEG_ROWS - 1) / HARM_SEG_ROWS);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 csqr(double2 a) {
    return make_double2(fma(a.x, a.x, -(a.y * a.y)), 2.0 * (a.x * a.y));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// unit phasor of the FC sample: (x, y) / |(x, y)|  (= exp(1im*angle(fc)), :388)
__device__ __forceinline__ double2 fc_unit(double x, double y) {
    const double h2 = fma(x, x, y * y);
    if (h2 > 1.0e-280 && h2 < 1.0e280) {
        const double inv = rsqrt(h2);
        return make_double2(x * inv, y * inv);
    }
    return fc_phasor(make_double2(x, y));
}

// C[8x8] += A[8x4] B[4x8] on the FP64 units: lane t holds A[t/4][t%4], B[t%4][t/4]
// and C[t/4][2(t%4)], C[t/4][2(t%4) + 1]
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// The constant (b, phi independent) sums of a fit, per diode:
//   [0] sum w, [1] sum w |d - mu|^2, [2] sum w |p|^2, [3..4] sum w (d - mu), [5..6] Z_0
// sum w and sum w |p|^2 (|FCphasor| = 1) follow from the per-state row counts of the
// segment: sum_s n_s w_s and sum_s n_s w_s m_s^2.
constexpr int WS_WARPS = 8;                    // warps per block, two blocks per SM
constexpr int WS_THREADS = WS_WARPS * 32;
constexpr int WS_CH = 32;                      // rows per chunk (one per lane)
constexpr int WS_CHP = WS_CH + 4;              // padded component stride (bank-conflict free)

struct WsRaw {                 // raw bytes of one chunk, as copied by cp.async
    uint4 dio[4][WS_CH];
    uint4 fc[WS_CH];
    uint4 basis[WS_CH];
    uint32_t state[WS_CH];
};
struct WsWarp {
    double e[8][WS_CHP];
    double v[8][WS_CHP];
    WsRaw raw[2];
};
static_assert(sizeof(WsWarp) >= 48 * 8 * 8, "a warp's partial C is parked in its own area");
constexpr int WS_SMEM = WS_WARPS * (int)sizeof(WsWarp);

template <int KIND, bool OFFS>
__global__ void __launch_bounds__(WS_THREADS, 2)
k_harm_ws(const TableDesc *tabs, const JobInfo *jobs, unsigned flags, int P, const double *stats,
          double *partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double2 s_stats[16];
    __shared__ double2 s_off[5];
    __shared__ double s_red[WS_WARPS][24];
    __shared__ unsigned long long s_cnt[WS_WARPS];

    constexpr int NCONST = KIND == 0 ? 7 : 2;
    constexpr int NACC = KIND == 0 ? (OFFS ? 5 : 3) : 2;
    constexpr int HP = KIND == 0 ? HP_Z : HP_Y;
    const int jg = blockIdx.x, p = blockIdx.y;
    const int job = jg >> 3, group = jg & 7;
    if (!group_on(flags, group)) return;       // gppd_options.group_mask
    const JobInfo ji = jobs[job];
    const TableDesc tb = tabs[ji.table];
    const TableView &tv = tb.tv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long seg0 = (long long)p * HARM_SEG_ROWS;
    if (seg0 >= ji.nrows) return;
    const int nseg = (int)((ji.nrows - seg0) < HARM_SEG_ROWS ? (ji.nrows - seg0) : HARM_SEG_ROWS);
    const int nchunks = (nseg + WS_CH - 1) / WS_CH;
    WsWarp &W = reinterpret_cast<WsWarp *>(smem_raw)[warp];

    if (threadIdx.x < 16) {
        s_stats[threadIdx.x] = tb.state ? stats_mean_weight(stats, jg, threadIdx.x >> 2, threadIdx.x & 3)
                                        : make_double2(1.0, 1.0);
    } else if (threadIdx.x < 21) {
        const int k = threadIdx.x - 16;
        const int ch = k < 4 ? group * 4 + k : fc_channel(group);
        s_off[k] = (tv.kind == 0 && tv.offsets) ? __ldg(tv.offsets + ch) : make_double2(0.0, 0.0);
    }
    __syncthreads();

    const bool async_ok = tv.kind == 1 ||
        ((reinterpret_cast<unsigned long long>(tv.volt) & 15ull) == 0 && (tv.volt_stride & 15) == 0);
    const bool faint = tb.state != nullptr;

    // consumer side: lane (m8, r) owns A[m8][r] = trig(k theta_row) of harmonic
    // k = k0 + 4j in m-tile j (k0 = m8/2 + 1, trig = cos for even m8, sin for odd) and
    // B[r][m8] = V[row][m8], row = (k-step base) + r;
    // trig((k0 - 4) t) = +cos((4 - k0) t) or -sin((4 - k0) t); k0 = 4: cos 0 = 1, sin 0 = 0
    const int m8 = lane >> 2, r4 = lane & 3;
    const int k0 = (m8 >> 1) + 1, trig = m8 & 1;
    const int pidx = k0 < 4 ? 2 * (3 - k0) + trig : 0;
    const double psgn = k0 < 4 ? (trig ? -1.0 : 1.0) : 0.0;
    const double padd = (k0 == 4 && !trig) ? 1.0 : 0.0;
    double c[MTILES][2];
#pragma unroll
    for (int j = 0; j < MTILES; ++j) c[j][0] = c[j][1] = 0.0;
    double cst[NACC * 4];
#pragma unroll
    for (int q = 0; q < NACC * 4; ++q) cst[q] = 0.0;
    unsigned long long cnt = 0;
    double2 mu[4];
#pragma unroll
    for (int d = 0; d < 4; ++d)
        mu[d] = OFFS ? row_sample(tv, ji.row0, group * 4 + d) : make_double2(0.0, 0.0);

    auto issue = [&](int chunk, WsRaw &S) {     // lane -> row of the chunk
        const int i = chunk * WS_CH + lane;
        if (i >= nseg) return;
        const long long r = ji.row0 + seg0 + i;
        cp_async16(&S.basis[lane], tb.basis + r);
        if (tv.kind == 0) {
            const char *row = reinterpret_cast<const char *>(tv.volt) + r * tv.volt_stride;
            cp_async16(&S.dio[0][lane], row + 32 * group);
            cp_async16(&S.dio[1][lane], row + 32 * group + 16);
            cp_async16(&S.fc[lane], row + 256 + 16 * (group >> 1));
        } else {
#pragma unroll
            for (int d = 0; d < 4; ++d)
                cp_async16(&S.dio[d][lane], tv.data + (long long)(group * 4 + d) * tv.n + r);
            cp_async16(&S.fc[lane], tv.data + (long long)fc_channel(group) * tv.n + r);
        }
        if (faint) {
            const int8_t *sp = tb.state + r;
            const unsigned long long aw = reinterpret_cast<unsigned long long>(sp) & ~3ull;
            if (aw >= reinterpret_cast<unsigned long long>(tb.state) &&
                aw + 4 <= reinterpret_cast<unsigned long long>(tb.state + tv.n)) {
                cp_async4(&S.state[lane], reinterpret_cast<const void *>(aw));
            } else {
                const unsigned sh = 8u * (unsigned)(reinterpret_cast<unsigned long long>(sp) & 3ull);
                S.state[lane] = ((unsigned)(unsigned char)*sp) << sh;
            }
        }
    };

    auto produce = [&](auto faint_tag, int chunk, const WsRaw &S) {
        constexpr bool FAINT = decltype(faint_tag)::value;
        const int i = chunk * WS_CH + lane;
        double2 e1 = make_double2(1.0, 0.0);
        double2 vv[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) vv[d] = make_double2(0.0, 0.0);
        if (i < nseg) {
            const long long r = ji.row0 + seg0 + i;
            double2 sc, fcs, dd[4];
            int st = ST_NORMAL;
            if (async_ok) {
                const uint4 bw = S.basis[lane];
                sc = make_double2(__hiloint2double(bw.y, bw.x), __hiloint2double(bw.w, bw.z));
                if (tv.kind == 0) {
                    uint4 a = S.dio[0][lane], b = S.dio[1][lane], f = S.fc[lane];
                    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    uint32_t fx = (group & 1) ? f.z : f.x, fy = (group & 1) ? f.w : f.y;
                    if (tv.big_endian) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) w[k] = bswap32(w[k]);
                        fx = bswap32(fx);
                        fy = bswap32(fy);
                    }
#pragma unroll
                    for (int d = 0; d < 4; ++d)
                        dd[d] = make_double2((double)__uint_as_float(w[2 * d]) - s_off[d].x,
                                             (double)__uint_as_float(w[2 * d + 1]) - s_off[d].y);
                    fcs = make_double2((double)__uint_as_float(fx) - s_off[4].x,
                                       (double)__uint_as_float(fy) - s_off[4].y);
                } else {
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        const uint4 q = S.dio[d][lane];
                        dd[d] = make_double2(__hiloint2double(q.y, q.x), __hiloint2double(q.w, q.z));
                    }
                    const uint4 q = S.fc[lane];
                    fcs = make_double2(__hiloint2double(q.y, q.x), __hiloint2double(q.w, q.z));
                }
                if (FAINT) {
                    const unsigned sh = 8u * (unsigned)(reinterpret_cast<unsigned long long>(tb.state + r) & 3ull);
                    st = (int)(signed char)((S.state[lane] >> sh) & 0xffu);
                }
            } else {
                sc = tb.basis[r];
#pragma unroll
                for (int d = 0; d < 4; ++d) dd[d] = row_sample(tv, r, group * 4 + d);
                fcs = row_sample(tv, r, fc_channel(group));
                if (FAINT) st = tb.state[r];
            }
            e1 = make_double2(sc.y, sc.x);
            const bool valid = FAINT ? row_valid(st, flags) : true;
            if (valid) {
                const double2 fc = fc_unit(fcs.x, fcs.y);
                cnt += 1ull << (16 * (st & 3));
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    double wpr = fc.x, wpi = fc.y, w = 1.0;   // bright: w = 1, p = FCphasor
                    if (FAINT) {
                        const double2 mw = s_stats[d * 4 + (st & 3)];
                        w = mw.y;
                        const double wm = mw.y * mw.x;        // p = power .* FCphasor
                        wpr = wm * fc.x;
                        wpi = wm * fc.y;
                    }
                    if (KIND == 0) {
                        double dr = dd[d].x, di = dd[d].y;
                        if (OFFS) { dr -= mu[d].x; di -= mu[d].y; }
                        vv[d].x = fma(wpr, dr, wpi * di);
                        vv[d].y = fma(wpr, di, -(wpi * dr));
                        cst[d * NACC + 0] = fma(w, fma(dr, dr, di * di), cst[d * NACC + 0]);
                        cst[d * NACC + 1] += vv[d].x;
                        cst[d * NACC + 2] += vv[d].y;
                        if (OFFS) {
                            cst[d * NACC + 3] = fma(w, dr, cst[d * NACC + 3]);
                            cst[d * NACC + 4] = fma(w, di, cst[d * NACC + 4]);
                        }
                    } else {
                        vv[d].x = wpr;
                        vv[d].y = wpi;
                        cst[d * NACC + 0] += wpr;
                        cst[d * NACC + 1] += wpi;
                    }
                }
            }
        }
        const double2 e2 = csqr(e1), e3 = cmul(e2, e1), e4 = csqr(e2);
        W.e[0][lane] = e1.x; W.e[1][lane] = e1.y;
        W.e[2][lane] = e2.x; W.e[3][lane] = e2.y;
        W.e[4][lane] = e3.x; W.e[5][lane] = e3.y;
        W.e[6][lane] = e4.x; W.e[7][lane] = e4.y;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            W.v[2 * d][lane] = vv[d].x;
            W.v[2 * d + 1][lane] = vv[d].y;
        }
    };

    // chunks warp, warp + WS_WARPS, ... of the segment
    int ck = warp;
    if (async_ok && ck < nchunks) issue(ck, W.raw[0]);
    cp_async_commit();
    for (int n = 0; ck < nchunks; ck += WS_WARPS, ++n) {
        if (async_ok) {
            if (ck + WS_WARPS < nchunks) issue(ck + WS_WARPS, W.raw[(n + 1) & 1]);
            cp_async_commit();
            cp_async_wait<1>();          // this chunk's bytes have landed (each lane reads its own)
        }
        if (faint) produce(std::true_type{}, ck, W.raw[n & 1]);
        else produce(std::false_type{}, ck, W.raw[n & 1]);
        __syncwarp();
#pragma unroll 4
        for (int ks = 0; ks < WS_CH / 4; ++ks) {
            const int row = ks * 4 + r4;
            const double a0 = W.e[m8][row];
            const double c4 = W.e[6][row];
            const double pv = W.e[pidx][row];
            const double bv = W.v[m8][row];
            const double tc = c4 + c4;
            double am = fma(psgn, pv, padd);
            double ak = a0;
#pragma unroll
            for (int j = 0; j < MTILES; ++j) {
                dmma_m8n8k4(c[j][0], c[j][1], ak, bv);
                if (j + 1 < MTILES) {
                    const double an = fma(tc, ak, -am);
                    am = ak;
                    ak = an;
                }
            }
        }
        __syncwarp();                    // the tile is rewritten by the next chunk
    }
    if (async_ok) cp_async_wait<0>();
    __syncwarp();

    // park the warp's partial C in its own area, reduce the constant sums
    double *cpart = reinterpret_cast<double *>(&W);
#pragma unroll
    for (int j = 0; j < MTILES; ++j) {
        double *dst = cpart + ((8 * j + m8) * 8 + 2 * r4);
        dst[0] = c[j][0];
        dst[1] = c[j][1];
    }
#pragma unroll
    for (int q = 0; q < NACC * 4; ++q) {
        double sv = cst[q];
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
        if (lane == 0) s_red[warp][q] = sv;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();

    double *out = partial + ((long long)jg * P + p) * 4 * HP;
    for (int idx = threadIdx.x; idx < 48 * 8; idx += WS_THREADS) {
        const int m = idx >> 3, n = idx & 7;
        double sv = 0.0;
#pragma unroll
        for (int w = 0; w < WS_WARPS; ++w)
            sv += reinterpret_cast<const double *>(&reinterpret_cast<WsWarp *>(smem_raw)[w])[m * 8 + n];
        const int k = m >> 1, sn = m & 1, d = n >> 1, im = n & 1;
        const int slot = sn ? (im ? 1 : 3) : (im ? 2 : 0);
        out[d * HP + NCONST + k * 4 + slot] = sv;
    }
    if (threadIdx.x < NCONST * 4) {
        const int d = threadIdx.x / NCONST, cc = threadIdx.x % NCONST;
        unsigned long long cn = 0;
        for (int w = 0; w < WS_WARPS; ++w) cn += s_cnt[w];
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
            for (int w = 0; w < WS_WARPS; ++w) sv += s_red[w][d * NACC + src];
        } else if (cc == 0 || cc == 2) {
            for (int st = 0; st < 4; ++st) {
                const double n_s = (double)((cn >> (16 * st)) & 0xffffull);
                const double2 mw = s_stats[d * 4 + st];
                if (n_s > 0.0) sv += cc == 0 ? n_s * mw.y : n_s * (mw.y * (mw.x * mw.x));
            }
        }
        out[d * HP + cc] = sv;
    }
}

// partial sums -> per-fit harmonic table, the job's segments added in index order
__global__ void k_harm_reduce(const JobInfo *jobs, const double *partZ, const double *partY, int P,
                              int nfits, double *htab) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int nvals = partY ? HV_COUNT : HV_Y0R;
    if (idx >= (long long)nfits * nvals) return;
    const int v = (int)(idx / nfits), fit = (int)(idx % nfits);
    const int job = fit / NDIODE, ch = fit % NDIODE;
    const int jg = job * NGROUP + ch / 4, d = ch & 3;
    const int nseg = harm_segments(jobs[job].nrows);
    double s = 0.0;
    if (v < HV_Y0R) {
        for (int p = 0; p < nseg; ++p) s += partZ[(((long long)jg * P + p) * 4 + d) * HP_Z + v];
    } else {
        const int vy = v - HV_Y0R;
        for (int p = 0; p < nseg; ++p) s += partY[(((long long)jg * P + p) * 4 + d) * HP_Y + vy];
    }
    htab[(long long)v * nfits + fit] = s;
}

// The same for very long jobs (more than HARM_LONG segments = 393 216 rows: the 1e8-row
// exposure of BASELINE config 4 has 16 277): one WARP per (value, fit) instead of one thread,
// which would walk the segments one dependent load after the other (measured: 2.5 ms of the
// 1e8-row global fit).  Lane l adds the chunks l, l + 32, ... of HARM_LONG consecutive segments
// each in index order, the lanes' sums are then added in lane order: a fixed order that depends
// on the job's length only, so the result is deterministic and independent of the batch.
constexpr int HARM_LONG = 64;
__global__ void k_harm_reduce_long(const JobInfo *jobs, const double *partZ, const double *partY, int P,
                                   int nfits, double *htab) {
    const long long idx = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nvals = partY ? HV_COUNT : HV_Y0R;
    if (idx >= (long long)nfits * nvals) return;
    const int v = (int)(idx / nfits), fit = (int)(idx % nfits);
    const int job = fit / NDIODE, ch = fit % NDIODE;
    const int jg = job * NGROUP + ch / 4, d = ch & 3;
    const int nseg = harm_segments(jobs[job].nrows);
    const bool zs = v < HV_Y0R;
    const double *base = zs ? partZ + (((long long)jg * P) * 4 + d) * HP_Z + v
                            : partY + (((long long)jg * P) * 4 + d) * HP_Y + (v - HV_Y0R);
    const long long stride = zs ? 4ll * HP_Z : 4ll * HP_Y;
    double s = 0.0;
    for (int c0 = lane * HARM_LONG; c0 < nseg; c0 += 32 * HARM_LONG) {
        const int c1 = c0 + HARM_LONG < nseg ? c0 + HARM_LONG : nseg;
        double cs = 0.0;
        for (int p = c0; p < c1; ++p) cs += base[(long long)p * stride];
        s += cs;
    }
    double tot = 0.0;
    for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, s, l);     // lane order
    if (lane == 0) htab[(long long)v * nfits + fit] = tot;
}

int harm_max_segments(long long max_rows_per_job) { return harm_segments(max_rows_per_job); }

// P = segments of the longest job of the batch.  A batch that holds a job of more than
// HARM_LONG segments uses the warp-per-value kernel for all its jobs; for a job of up to
// HARM_LONG segments (every real table) the two kernels add the same numbers in the same
// order (one chunk on lane 0, the other lanes contribute +0.0), so a job's sums still do not
// depend on the batch it is in.
static void launch_harm_reduce(const Launcher &L, const JobInfo *d_jobs, const double *d_partZ,
                               const double *d_partY, int P, int nfits, long long tot, double *d_htab) {
    if (P > HARM_LONG)
        k_harm_reduce_long<<<(unsigned)((tot * 32 + 255) / 256), 256, 0, L.stream>>>(d_jobs, d_partZ, d_partY, P,
                                                                                    nfits, d_htab);
    else
        k_harm_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, L.stream>>>(d_jobs, d_partZ, d_partY, P, nfits,
                                                                           d_htab);
    *L.counter += 1;
}

void launch_harmonics(const Launcher &L, const TableDesc *d_tabs, const JobInfo *d_jobs, int njobs,
                      unsigned flags, int P, int SP, const double *d_spart1,
                      const double *d_spart2, double *d_partZ, double *d_partY, double *d_htab,
                      int tensor) {
    const bool offs = (flags & 2u) != 0;
    const int nfits = njobs * NDIODE;
    const long long tot = (long long)nfits * (offs ? HV_COUNT : HV_Y0R);
    if (tensor) {
        // GPPD_FP32 (bit 6): the float32-class form of the same sums
        if ((flags & 64u) && tensor == 1)
            launch_harmonics_tc32(L, d_tabs, d_jobs, njobs, flags, P, d_spart2, d_partZ, d_partY);
        else
            launch_harmonics_tc(L, d_tabs, d_jobs, njobs, flags, P, d_spart2, d_partZ, d_partY, tensor == 2);
        launch_harm_reduce(L, d_jobs, d_partZ, offs ? d_partY : nullptr, P, nfits, tot, d_htab);
        return;
    }
    cudaFuncSetAttribute(k_harm_ws<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    cudaFuncSetAttribute(k_harm_ws<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    cudaFuncSetAttribute(k_harm_ws<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    dim3 grid(njobs * NGROUP, P);
    if (offs) {
        k_harm_ws<0, true><<<grid, WS_THREADS, WS_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partZ);
        k_harm_ws<1, true><<<grid, WS_THREADS, WS_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partY);
        *L.counter += 2;
    } else {
        k_harm_ws<0, false><<<grid, WS_THREADS, WS_SMEM, L.stream>>>(d_tabs, d_jobs, flags, P, d_spart2, d_partZ);
        *L.counter += 1;
    }
    launch_harm_reduce(L, d_jobs, d_partZ, offs ? d_partY : nullptr, P, nfits, tot, d_htab);
}

}  // namespace gppd
