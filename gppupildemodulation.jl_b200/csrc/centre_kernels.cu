// centre_kernels.cu -- `--center empirical`: one circle per channel.
//
// Reference: compute_offsets, src/GPPupilDemodulation.jl:105-125 -- for each of the 40
// channels `fit(Circle, re, im)` over all samples of the table, or over its HIGH samples
// when the table has FAINT states, and processmetrology subtracts the circle centres
// (:153-154).  `Circle` is not defined in the reference's environment, so the reference
// throws there; this is the algebraic (Kasa) least-squares circle the call was written
// for: minimise sum (x^2 + y^2 - 2 x0 x - 2 y0 y - c)^2, a linear problem in
// (x0, y0, c), i.e. ten moment sums per channel and a 2x2 solve.
//
// The sums are taken relative to a pivot on the circle (the channel's sample of the
// table's first row), which keeps every moment at the scale of the radius whatever the
// distance of the circle from the origin, and over FIXED row segments added in index
// order, so a table's centres do not depend on the batch it is in.
#include "kernels.h"

namespace gppd {

constexpr int CIRC_LANES = 8;                       // rows in flight per block iteration
constexpr int CIRC_THREADS = CIRC_LANES * NCHAN;    // thread = (row lane, channel)
constexpr int CIRC_UNROLL = 8;

int circle_max_segments(long long max_rows) {
    return (int)((max_rows + CIRC_SEG_ROWS - 1) / CIRC_SEG_ROWS);
}

// centred sample (u, v) -> the ten moments
struct Moments {
    double n, u, v, uu, uv, vv, uuu, uuv, uvv, vvv;
};

__device__ __forceinline__ void add_sample(Moments &m, double u, double v) {
    const double uu = u * u, vv = v * v;
    m.n += 1.0;
    m.u += u;
    m.v += v;
    m.uu += uu;
    m.uv = fma(u, v, m.uv);
    m.vv += vv;
    m.uuu = fma(uu, u, m.uuu);
    m.uuv = fma(uu, v, m.uuv);
    m.uvv = fma(u, vv, m.uvv);
    m.vvv = fma(vv, v, m.vvv);
}

// grid (segment, table); a row is 40 consecutive float2 = one coalesced 320-byte read
// of the 40 threads of a row lane
__global__ void __launch_bounds__(CIRC_THREADS) k_circle_seg(const TableDesc *tabs, int P,
                                                             double *part) {
    __shared__ double s_m[CIRC_LANES][CIRC_VALS][NCHAN];
    const TableDesc &tb = tabs[blockIdx.y];
    if (tb.tv.kind != 0 || !tb.tv.offsets) return;
    const long long n = tb.tv.n;
    const long long r0 = (long long)blockIdx.x * CIRC_SEG_ROWS;
    if (r0 >= n) return;
    const long long r1 = min(n, r0 + (long long)CIRC_SEG_ROWS);
    const int lane = threadIdx.x / NCHAN, ch = threadIdx.x - lane * NCHAN;
    const int8_t *state = tb.state;
    TableView tv = tb.tv;
    tv.offsets = nullptr;                      // raw volts
    const double2 piv = row_sample(tv, 0, ch);

    Moments m = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long r = r0 + lane; r < r1; r += CIRC_LANES * CIRC_UNROLL) {
        double2 d[CIRC_UNROLL];
        int use[CIRC_UNROLL];
#pragma unroll
        for (int j = 0; j < CIRC_UNROLL; ++j) {
            const long long rr = r + (long long)j * CIRC_LANES;
            use[j] = rr < r1;
            if (use[j]) {
                d[j] = row_sample(tv, rr, ch);
                if (state) use[j] = state[rr] == ST_HIGH;   // cmplxV[state .== HIGH, ch], :108
            }
        }
#pragma unroll
        for (int j = 0; j < CIRC_UNROLL; ++j)
            if (use[j]) add_sample(m, d[j].x - piv.x, d[j].y - piv.y);
    }
    double *sm = &s_m[lane][0][ch];
    sm[0 * NCHAN] = m.n;   sm[1 * NCHAN] = m.u;   sm[2 * NCHAN] = m.v;   sm[3 * NCHAN] = m.uu;
    sm[4 * NCHAN] = m.uv;  sm[5 * NCHAN] = m.vv;  sm[6 * NCHAN] = m.uuu; sm[7 * NCHAN] = m.uuv;
    sm[8 * NCHAN] = m.uvv; sm[9 * NCHAN] = m.vvv;
    __syncthreads();
    // lanes added in lane order: [table][segment][value][channel]
    double *dst = part + ((size_t)blockIdx.y * P + blockIdx.x) * (CIRC_VALS * NCHAN);
    for (int i = threadIdx.x; i < CIRC_VALS * NCHAN; i += CIRC_THREADS) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < CIRC_LANES; ++l) acc += (&s_m[l][0][0])[i];
        dst[i] = acc;
    }
}

// one thread per (table, channel): segments in index order, then the 2x2 solve.  A
// channel with fewer than 3 samples, or whose samples lie on one line, has no circle:
// its centre is 0 (the channel is left as it is).
__global__ void k_circle_solve(const TableDesc *tabs, int ntables, int P, const double *part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntables * NCHAN) return;
    const int t = i / NCHAN, ch = i - t * NCHAN;
    const TableDesc &tb = tabs[t];
    if (tb.tv.kind != 0 || !tb.tv.offsets) return;
    const int nseg = (int)((tb.tv.n + CIRC_SEG_ROWS - 1) / CIRC_SEG_ROWS);
    double s[CIRC_VALS];
#pragma unroll
    for (int k = 0; k < CIRC_VALS; ++k) s[k] = 0.0;
    for (int p = 0; p < nseg; ++p) {
        const double *src = part + ((size_t)t * P + p) * (CIRC_VALS * NCHAN) + ch;
#pragma unroll
        for (int k = 0; k < CIRC_VALS; ++k) s[k] += src[k * NCHAN];
    }
    double2 centre = make_double2(0.0, 0.0);
    const double N = s[0];
    if (N >= 3.0) {
        const double mu = s[1] / N, mv = s[2] / N;
        // central moments from the pivot-relative ones
        const double cuu = s[3] - N * mu * mu;
        const double cuv = s[4] - N * mu * mv;
        const double cvv = s[5] - N * mv * mv;
        const double cuuu = s[6] - 3.0 * mu * s[3] + 2.0 * N * mu * mu * mu;
        const double cvvv = s[9] - 3.0 * mv * s[5] + 2.0 * N * mv * mv * mv;
        const double cuuv = s[7] - 2.0 * mu * s[4] - mv * s[3] + 2.0 * N * mu * mu * mv;
        const double cuvv = s[8] - 2.0 * mv * s[4] - mu * s[5] + 2.0 * N * mu * mv * mv;
        const double det = cuu * cvv - cuv * cuv;
        if (det > 1.0e-12 * cuu * cvv && det > 0.0) {
            const double bu = 0.5 * (cuuu + cuvv), bv = 0.5 * (cvvv + cuuv);
            const double uc = (bu * cvv - bv * cuv) / det;
            const double vc = (bv * cuu - bu * cuv) / det;
            TableView tv = tb.tv;
            tv.offsets = nullptr;
            const double2 piv = row_sample(tv, 0, ch);
            centre = make_double2(piv.x + (mu + uc), piv.y + (mv + vc));
        }
    }
    const_cast<double2 *>(tb.tv.offsets)[ch] = centre;
}

void launch_circle(const Launcher &L, const TableDesc *d_tabs, int ntables, long long max_rows,
                   double *d_part) {
    const int P = circle_max_segments(max_rows);
    k_circle_seg<<<dim3(P, ntables), CIRC_THREADS, 0, L.stream>>>(d_tabs, P, d_part);
    k_circle_solve<<<(ntables * NCHAN + 127) / 128, 128, 0, L.stream>>>(d_tabs, ntables, P, d_part);
    *L.counter += 2;
}

}  // namespace gppd
