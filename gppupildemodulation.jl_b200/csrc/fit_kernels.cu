// fit_kernels.cu -- the per-diode modulation fit (reference
// src/Modulation.jl:387-416: objective :323-326, model update :122-148,
// closed-form linear parameters :174-215, NEWUOA :332-336).
//
// Compiled with -fmad=false: the solver state machine must take exactly the
// oracle's floating-point steps (see newuoa2.cuh).  Hot loops therefore spell
// out their fused multiply-adds with fma().
//
// Direct evaluator (this file, k_fit_direct): one thread block per fit.
//   prologue  z_n = w_n conj(p_n) (d_n - mu),  y_n = w_n p_n  -> HBM scratch,
//             constant sums S_w, S_d, S_dd, S_gg         (p = power * FCphasor)
//   per objective call, one pass over the rows (16 B basis + 16 B z per row):
//             e_n = exp(j b sin(theta_n + q));  S_gd = sum conj(e_n) z_n
//             [offsets: S_g = sum e_n y_n]
//             (c, a) in closed form, chi2 = (S_dd - Re(conj(c) S_d + conj(a) S_gd)) / N
//   which is the reference's  sum w |c + a g - d|^2 / N  at its own least-squares
//   (c, a), with g = p e.  Reductions run in a fixed order (deterministic).
#include "fit_driver.cuh"
#include "gppd_device.cuh"
#include "kernels.h"

namespace gppd {

constexpr int FIT_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_sum_vec(double (&v)[NV], double *red /* [NV][8] */) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 8 + w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < FIT_THREADS / 32; ++j) s += red[k * 8 + j];
        v[k] = s;
    }
}

// Constant (b, phi independent) sums of one fit.
struct FitConsts {
    double sw, sdd, sgg;   // sum w, sum w|d-mu|^2, sum w |p|^2
    double sdr, sdi;       // sum w (d - mu)
    double mur, mui;       // mu = weighted mean of d (offsets mode), else 0
    double nvalid;
};

// Linear parameters and chi2 from the sums (reference :140-145, :174-215, :325).
// Offsets mode solves the centred system, c = c' + mu.
__device__ __forceinline__ double solve_linear(const FitConsts &k, bool offs, double sgdr,
                                               double sgdi, double sgr, double sgi, double &cre,
                                               double &cim, double &are, double &aim) {
    if (!offs) {
        // a = (mw . d) / (mw . model) with mw . model = sum w |g|^2 (real)
        are = sgdr / k.sgg;
        aim = sgdi / k.sgg;
        cre = 0.0;
        cim = 0.0;
        double num = fma(sgdr, sgdr, sgdi * sgdi);
        return (k.sdd - num / k.sgg) / k.nvalid;
    }
    // [sw  S_g; conj(S_g)  sgg] [c'; a] = [S_d'; S_gd]   (Cramer, StaticArrays 2x2)
    double det = fma(k.sw, k.sgg, -fma(sgr, sgr, sgi * sgi));
    // c' = (sgg S_d' - S_g S_gd) / det
    double t1r = fma(sgr, sgdr, -(sgi * sgdi)), t1i = fma(sgr, sgdi, sgi * sgdr);
    double cpr = (k.sgg * k.sdr - t1r) / det, cpi = (k.sgg * k.sdi - t1i) / det;
    // a = (sw S_gd - conj(S_g) S_d') / det
    double t2r = fma(sgr, k.sdr, sgi * k.sdi), t2i = fma(sgr, k.sdi, -(sgi * k.sdr));
    are = (k.sw * sgdr - t2r) / det;
    aim = (k.sw * sgdi - t2i) / det;
    cre = cpr + k.mur;
    cim = cpi + k.mui;
    // chi2 N = S_dd - Re(conj(c') S_d' + conj(a) S_gd)
    double proj = fma(cpr, k.sdr, cpi * k.sdi) + fma(are, sgdr, aim * sgdi);
    return (k.sdd - proj) / k.nvalid;
}

template <bool OFFS>
__global__ void __launch_bounds__(FIT_THREADS)
k_fit_direct(TableView tv, const JobInfo *jobs, const int8_t *state, const double2 *stats,
             const double2 *basis, double2 *zbuf, double2 *ybuf, FitOptions opt,
             const int *fit_list, FitResult *results, double *trace) {
    __shared__ double red[5 * 8];
    const int fit = fit_list ? fit_list[blockIdx.x] : blockIdx.x;
    const int job = fit / NDIODE, ch = fit % NDIODE;
    const int fcch = fc_channel(ch / 4);
    const JobInfo ji = jobs[job];
    const unsigned flags = opt.flags;
    double2 *z = zbuf + (long long)ch * tv.n + ji.row0;
    double2 *y = OFFS ? ybuf + (long long)ch * tv.n + ji.row0 : nullptr;
    const double2 *bas = basis + ji.row0;

    // per-state (mean |d|, 1/var |d|), reference compute_mean_var_power
    double2 st4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
        st4[s] = state ? stats[(long long)fit * 4 + s] : make_double2(1.0, 1.0);

    FitConsts kc;
    kc.mur = kc.mui = 0.0;
    kc.nvalid = (double)ji.nvalid;
    if (OFFS) {  // weighted mean of d, to centre the 2x2 system
        double acc[3] = {0, 0, 0};
        for (int i = threadIdx.x; i < ji.nrows; i += FIT_THREADS) {
            long long r = ji.row0 + i;
            double w = 1.0;
            if (state) {
                int st = state[r];
                if (!row_valid(st, flags)) continue;
                w = st4[st & 3].y;
            }
            double2 d = row_sample(tv, r, ch);
            acc[0] += w;
            acc[1] = fma(w, d.x, acc[1]);
            acc[2] = fma(w, d.y, acc[2]);
        }
        block_sum_vec<3>(acc, red);
        kc.mur = acc[1] / acc[0];
        kc.mui = acc[2] / acc[0];
    }
    {
        double acc[5] = {0, 0, 0, 0, 0};  // sw, sdd, sgg, sdr, sdi
        for (int i = threadIdx.x; i < ji.nrows; i += FIT_THREADS) {
            long long r = ji.row0 + i;
            double w = 1.0, m = 1.0;
            bool valid = true;
            if (state) {
                int st = state[r];
                valid = row_valid(st, flags);
                w = st4[st & 3].y;
                m = st4[st & 3].x;
            }
            double2 zz = make_double2(0.0, 0.0), yy = zz;
            if (valid) {
                double2 d = row_sample(tv, r, ch);
                double2 fc = fc_phasor(row_sample(tv, r, fcch));
                double dr = d.x - kc.mur, di = d.y - kc.mui;
                double pr = m * fc.x, pi = m * fc.y;  // p = power .* FCphasor, :396
                double wpr = w * pr, wpi = w * pi;
                // z = w conj(p) (d - mu)
                zz.x = fma(wpr, dr, wpi * di);
                zz.y = fma(wpr, di, -(wpi * dr));
                yy.x = wpr;
                yy.y = wpi;
                acc[0] += w;
                acc[1] = fma(w, fma(dr, dr, di * di), acc[1]);
                acc[2] = fma(w, fma(pr, pr, pi * pi), acc[2]);
                acc[3] = fma(w, dr, acc[3]);
                acc[4] = fma(w, di, acc[4]);
            }
            z[i] = zz;
            if (OFFS) y[i] = yy;
        }
        block_sum_vec<5>(acc, red);
        kc.sw = acc[0];
        kc.sdd = acc[1];
        kc.sgg = acc[2];
        kc.sdr = acc[3];
        kc.sdi = acc[4];
    }
    __syncthreads();  // z/y visible to the whole block

    FitDriver drv;
    drv.start(opt);
    double cre = 0, cim = 0, are = 0, aim = 0, f = 0;
    double *tr = trace ? trace + (long long)fit * (3 * 160) : nullptr;
    for (;;) {
        const double b = drv.b, phi = drv.phi;
        const PhaseQ pq = make_phaseq(phi, ji.thmin, ji.thmax);
        double acc[4] = {0, 0, 0, 0};
        for (int i = threadIdx.x; i < ji.nrows; i += FIT_THREADS) {
            double2 sc = bas[i];
            double2 zz = z[i];
            double sn;
            if (pq.uniform) {
                sn = fma(sc.x, pq.cq, sc.y * pq.sq);
            } else {
                sn = sin_arg(pq, phi, row_theta(tv, ji.row0 + i), sc);
            }
            double su, cu;
            sincos(b * sn, &su, &cu);
            // conj(e) z
            acc[0] = fma(cu, zz.x, fma(su, zz.y, acc[0]));
            acc[1] = fma(cu, zz.y, fma(-su, zz.x, acc[1]));
            if (OFFS) {  // e y
                double2 yy = y[i];
                acc[2] = fma(cu, yy.x, fma(-su, yy.y, acc[2]));
                acc[3] = fma(cu, yy.y, fma(su, yy.x, acc[3]));
            }
        }
        if (OFFS) {
            block_sum_vec<4>(acc, red);
        } else {
            double a2[2] = {acc[0], acc[1]};
            block_sum_vec<2>(a2, red);
            acc[0] = a2[0];
            acc[1] = a2[1];
        }
        f = solve_linear(kc, OFFS, acc[0], acc[1], acc[2], acc[3], cre, cim, are, aim);
        if (tr && threadIdx.x == 0 && drv.nfev < 160) {
            tr[3 * drv.nfev] = b;
            tr[3 * drv.nfev + 1] = phi;
            tr[3 * drv.nfev + 2] = f;
        }
        if (!drv.step(opt, f)) break;
    }
    if (threadIdx.x == 0) {
        FitResult r;
        r.cre = cre; r.cim = cim; r.are = are; r.aim = aim;
        r.b = drv.b; r.phi = drv.phi;
        r.alpha = atan2(aim, are);
        r.chi2 = drv.chi2;
        PhaseQ pq = make_phaseq(drv.phi, ji.thmin, ji.thmax);
        r.q = pq.q; r.cq = pq.cq; r.sq = pq.sq; r.uniform = pq.uniform;
        r.nfev = drv.nfev; r.status = drv.status; r.method = 1; r.second = drv.second;
        results[fit] = r;
    }
}

void launch_fit_direct(const Launcher &L, const TableView &tv, int nfits, const JobInfo *d_jobs,
                       const int8_t *d_state, const double2 *d_stats, const double2 *d_basis,
                       double2 *d_z, double2 *d_y, const FitOptions &opt, const int *d_fit_list,
                       FitResult *d_results, double *d_trace) {
    if (nfits <= 0) return;
    if (opt.flags & 2u)
        k_fit_direct<true><<<nfits, FIT_THREADS, 0, L.stream>>>(tv, d_jobs, d_state, d_stats,
                                                               d_basis, d_z, d_y, opt, d_fit_list,
                                                               d_results, d_trace);
    else
        k_fit_direct<false><<<nfits, FIT_THREADS, 0, L.stream>>>(tv, d_jobs, d_state, d_stats,
                                                                d_basis, d_z, d_y, opt, d_fit_list,
                                                                d_results, d_trace);
    *L.counter += 1;
}

}  // namespace gppd
