// fit_math.cuh -- arithmetic shared by the fit evaluators: closed-form linear
// parameters and chi2 from the reduced sums (reference src/Modulation.jl:140-145,
// :174-215, :325), Bessel coefficients for the Jacobi-Anger evaluator, and the
// per-state (mean, weight) lookup (reference src/Faint.jl:89-100).
#pragma once
#include "gppd_device.cuh"

namespace gppd {

// Constant (b, phi independent) sums of one fit.
struct FitConsts {
    double sw, sdd, sgg;   // sum w, sum w|d-mu|^2, sum w |p|^2
    double sdr, sdi;       // sum w (d - mu)
    double mur, mui;       // reference point mu (fitoffsets only), else 0
    double nvalid;
};

// Linear parameters and chi2 from the sums.  With g = p e (p = power * FCphasor,
// e = exp(j b sin(.))):  S_gd = sum w conj(g) (d - mu),  S_g = sum w g.
//   no offsets:  a = S_gd / S_gg,                         chi2 N = S_dd - |S_gd|^2 / S_gg
//   offsets:     [S_w  S_g; conj(S_g)  S_gg] [c'; a] = [S_d'; S_gd]  (Cramer, like the
//                reference's StaticArrays 2x2 solve), c = c' + mu,
//                chi2 N = S_dd - Re(conj(c') S_d' + conj(a) S_gd)
// which is sum w |c + a g - d|^2 at the least-squares (c, a) (reference :325).
__device__ __forceinline__ double solve_linear(const FitConsts &k, bool offs, double sgdr,
                                               double sgdi, double sgr, double sgi, double &cre,
                                               double &cim, double &are, double &aim) {
    if (!offs) {
        are = sgdr / k.sgg;
        aim = sgdi / k.sgg;
        cre = 0.0;
        cim = 0.0;
        double num = fma(sgdr, sgdr, sgdi * sgdi);
        return (k.sdd - num / k.sgg) / k.nvalid;
    }
    double det = fma(k.sw, k.sgg, -fma(sgr, sgr, sgi * sgi));
    double t1r = fma(sgr, sgdr, -(sgi * sgdi)), t1i = fma(sgr, sgdi, sgi * sgdr);
    double cpr = (k.sgg * k.sdr - t1r) / det, cpi = (k.sgg * k.sdi - t1i) / det;
    double t2r = fma(sgr, k.sdr, sgi * k.sdi), t2i = fma(sgr, k.sdi, -(sgi * k.sdr));
    are = (k.sw * sgdr - t2r) / det;
    aim = (k.sw * sgdi - t2i) / det;
    cre = cpr + k.mur;
    cim = cpi + k.mui;
    double proj = fma(cpr, k.sdr, cpi * k.sdi) + fma(are, sgdr, aim * sgdi);
    return (k.sdd - proj) / k.nvalid;
}

// J_0(b) .. J_HK(b) by Miller's backward recurrence J_{k-1} = (2k/b) J_k - J_{k+1}
// from k = 56, normalised with J_0 + 2 sum J_{2k} = 1 (|b| <= HARM_BMAX keeps the
// start order far in the decaying region).  Valid for either sign of b.
__device__ __forceinline__ void bessel_j(double b, double *J) {
    if (b == 0.0) {
        J[0] = 1.0;
#pragma unroll 1
        for (int k = 1; k <= HK; ++k) J[k] = 0.0;
        return;
    }
    const int M = 56;
    const double tb = 2.0 / b;
    double jp = 0.0, jc = 1.0e-250, sum = 0.0;
#pragma unroll 1
    for (int k = M; k >= 1; --k) {
        double jm = fma((double)k * tb, jc, -jp);  // J_{k-1}
        jp = jc;
        jc = jm;
        if (k - 1 <= HK) J[k - 1] = jc;
        if (((k - 1) & 1) == 0) sum += (k - 1 == 0) ? jc : 2.0 * jc;
        if (fabs(jc) > 1.0e200) {  // rescale (tiny |b|: the recurrence grows like (2k/b)^k)
            jc *= 1.0e-200;
            jp *= 1.0e-200;
            sum *= 1.0e-200;
#pragma unroll 1
            for (int i = k - 1; i <= HK; ++i) J[i] *= 1.0e-200;
        }
    }
    double inv = 1.0 / sum;
#pragma unroll 1
    for (int k = 0; k <= HK; ++k) J[k] *= inv;
}

// Per-state statistics of one (job, group) from the two partial-sum passes:
//   part1[(jg*P + p)*STATS_VALS + dio*4 + st] = sum |d|,  [.. + 16 + st] = count
//   part2[(jg*P + p)*16 + dio*4 + st]         = sum (|d| - mean)^2
// mean = sum/n, weight = 1/var = (n-1)/M2 (reference src/Faint.jl:93-98).  P is the
// stride (max segments per job in the batch), nseg the job's own segment count;
// segments are added in index order so every kernel gets bit-identical values.
constexpr int STATS_SEG_ROWS = 4096;
__host__ __device__ inline int stats_segments(long long nrows) {
    return (int)((nrows + STATS_SEG_ROWS - 1) / STATS_SEG_ROWS);
}
__device__ __forceinline__ double stats_mean(const double *part1, int jg, int P, int nseg, int dio,
                                             int st) {
    double s = 0.0, n = 0.0;
    for (int p = 0; p < nseg; ++p) {
        const double *q = part1 + ((long long)jg * P + p) * STATS_VALS;
        s += q[dio * 4 + st];
        n += q[16 + st];
    }
    return s / n;
}
__device__ __forceinline__ double2 stats_mean_weight(const double *part1, const double *part2,
                                                     int jg, int P, int nseg, int dio, int st) {
    double s = 0.0, n = 0.0, m2 = 0.0;
    for (int p = 0; p < nseg; ++p) {
        const double *q = part1 + ((long long)jg * P + p) * STATS_VALS;
        s += q[dio * 4 + st];
        n += q[16 + st];
        m2 += part2[((long long)jg * P + p) * 16 + dio * 4 + st];
    }
    double var = m2 / (n - 1.0);   // n == 1 -> 0/0 = NaN as in Julia
    return make_double2(s / n, 1.0 / var);
}

}  // namespace gppd
