// newuoa2.cuh -- device-side Powell NEWUOA for n = 2, npt = 5.
//
// The reference minimises chi2(b, phi) with
//   newuoa(x -> lkl(scratch, x), xinit, 1, 1e-3; check=false)
// (reference src/Modulation.jl:332-336), i.e. OptimPackNextGen's NEWUOA with
// its defaults npt = 2n+1 = 5, maxeval = 30n = 60.  This is that algorithm
// (routines NEWUOB, TRSAPP, BIGLAG, BIGDEN, UPDATE of Powell's 2004 report),
// specialised to two variables and written as a resumable state machine:
//
//     Newuoa2 s;  s.start(x0, rhobeg, rhoend, maxfun);
//     while (s.step(f)) f = objective(s.x[1], s.x[2]);   // s.x = next trial point
//     // s.x = best point, s.f = its value, s.nf = objective calls
//
// so that a whole warp of independent fits can evaluate their objectives in
// lock step while the (divergent) solver algebra runs in between, and so that
// a thread block that shares one fit can run the algebra redundantly in every
// thread with the block-wide reduction inside `objective`.
//
// NEWUOA has rounding-level ties (e.g. the test SUM > DISTSQ directly after
// DELTA = HALF*DNORM), so every expression here keeps one fixed evaluation
// order and this translation unit is compiled with -fmad=false: given the
// same objective values the device solver takes exactly the steps of the CPU
// oracle (oracle/newuoa.c).  tests/test_newuoa_device.py asserts that
// bit for bit.
#pragma once

namespace gppd {

// Portable sin/cos for the solver's angle searches (|x| <~ 8): Cody-Waite
// reduction by pi/2 + fdlibm kernel polynomials in a fixed operation order.
__host__ __device__ inline void nu_sincos(double x, double *sn, double *cs) {
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;
    const double pio2_1t = 6.07710050650619224932e-11;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double fn = floor(x * invpio2 + 0.5);
    int k = (int)fn;
    double r = (x - fn * pio2_1) - fn * pio2_1t;
    double z = r * r;
    double v = z * r;
    double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double ks = r + v * (S1 + z * ps);
    double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    double kc = w + (((1.0 - w) - hz) + z * pc);
    switch (k & 3) {
    case 0: *sn = ks; *cs = kc; break;
    case 1: *sn = kc; *cs = -ks; break;
    case 2: *sn = -ks; *cs = -kc; break;
    default: *sn = -kc; *cs = ks; break;
    }
}

enum { NU_SUCCESS = 0, NU_ROUNDING_ERRORS = -2, NU_TOO_MANY_EVALUATIONS = -3 };

// Loops over the NPT = 5 points / NDIM = 7 rows are kept rolled on the device: fully
// unrolled, the solver is ~140 KB of code and the fit kernels stall on instruction
// fetch (the arithmetic and its order are the same either way).
#ifdef __CUDA_ARCH__
#define NU_ROLLED _Pragma("unroll 1")
#else
#define NU_ROLLED
#endif
// rarely executed, large routines: out of line on the device
#ifdef __CUDACC__
#define NU_COLD __noinline__
#else
#define NU_COLD
#endif
// NU_OUTLINE (experiment builds): the three hot routines out of line too, to read their sizes
#if defined(__CUDACC__) && defined(NU_OUTLINE)
#define NU_HOT __noinline__
#else
#define NU_HOT
#endif

// The three angle searches (TRSAPP, BIGLAG, BIGDEN) always probe the same 49
// angles 2 pi i / 50: their sines and cosines are tabulated once with nu_sincos
// (so the tabulated values are the very bits the oracle computes in its loops).
constexpr int NU_ANGLES = 49;
struct NuSinCos {
    double s, c;
};
__host__ __device__ inline void nu_angle_entry(int i, NuSinCos *e) {
    const double twopi = 6.283185307179586476925;
    const double temp = twopi / (double)(NU_ANGLES + 1);
    const double angle = (double)i * temp;
    nu_sincos(angle, &e->s, &e->c);
}

// WARP = false: one thread owns the solver (host build, thread-per-fit and
//               block-replicated device use); the angle searches are serial loops.
// WARP = true : the 32 lanes of a converged warp run the same solver (the object
//               may live in shared memory, every lane storing identical values);
//               the angle searches are split over the lanes, two angles each, and
//               reduced with the serial loop's own selection rule, so the result
//               is bit for bit the serial one.
template <bool WARP>
struct Newuoa2T {
    static constexpr int N = 2, NPT = 5, NP = 3, NH = 3, NPTM = 2, NDIM = 7;

    // ---- interface ----
    double x[N + 1];  // 1-based: trial point to evaluate / final best point
    double f;         // final best value
    int nf;           // objective calls so far
    int status;

    // ---- persistent solver state (all vectors 1-based, slot 0 unused) ----
    double xbase[N + 1], xopt[N + 1], xnew[N + 1];
    double fval[NPT + 1], gq[N + 1], hq[NH + 1], pq[NPT + 1];
    double d[N + 1], vlag[NDIM + 1], w[2 * NDIM + 2 * NPT + 1];
    double xpt_[NPT * N], bmat_[NDIM * N], zmat_[NPT * NPTM];
    double rhobeg, rhoend, rho, delta, rhosq, recip, reciq;
    double fbeg, fopt, xipt, xjpt, diffa, diffb, diffc, xoptsq, dsq, dnorm;
    double ratio, crvmin, beta, alpha, dstep, fcur;
    int nftest, nfm, nfmm, kopt, idz, itest, nfsav, knew, ipt, jpt;
    int phase;  // 0 = not started, 1 = waiting for an objective value, 2 = done
    const NuSinCos *ang;  // [0..49]: nu_angle_entry(i), set by the owner before start()
    // BIGDEN work space (members so that a shared-memory solver keeps them there)
    double den[10], denex[10], par[10], wvec_[NDIM * 5], prod_[NDIM * 5];

#define XPT(k, j) xpt_[((k)-1) + ((j)-1) * NPT]
#define BMAT(i, j) bmat_[((i)-1) + ((j)-1) * NDIM]
#define ZMAT(k, j) zmat_[((k)-1) + ((j)-1) * NPT]

    __host__ __device__ static double dmax(double a, double b) { return a > b ? a : b; }
    __host__ __device__ static double dmin(double a, double b) { return a < b ? a : b; }

    __host__ __device__ void start(double b0, double phi0, double rhobeg_, double rhoend_,
                                   int maxfun) {
        x[1] = b0;
        x[2] = phi0;
        rhobeg = rhobeg_;
        rhoend = rhoend_;
        nftest = maxfun > 1 ? maxfun : 1;
        phase = 0;
        status = NU_SUCCESS;
        nf = 0;
    }

    // HD = (second derivative matrix of the model) * D
    __host__ __device__ void hess_mul(const double *dd_, double *hd) const {
        for (int i = 1; i <= N; ++i) hd[i] = 0.0;
        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            double temp = 0.0;
            for (int j = 1; j <= N; ++j) temp += XPT(k, j) * dd_[j];
            temp *= pq[k];
            for (int i = 1; i <= N; ++i) hd[i] += temp * XPT(k, i);
        }
        int ih = 0;
        for (int j = 1; j <= N; ++j)
            for (int i = 1; i <= j; ++i) {
                ++ih;
                if (i < j) hd[j] += hq[ih] * dd_[i];
                hd[i] += hq[ih] * dd_[j];
            }
    }

    // ------------------------------------------------------------------
    // Angle search shared by TRSAPP / BIGLAG / BIGDEN: v_i = val(sin, cos) of the
    // 49 tabulated angles; Powell's loop keeps the first best value (MODE 0: the
    // smallest, MODE 1: the largest magnitude) if it beats v0, together with its
    // two neighbours (v_0 = v_50 = v0) for the parabolic refinement:
    //     for i: if better(v_i, best) {best = v_i; isave = i; tempa = v_{i-1};}
    //            else if (i == isave + 1) tempb = v_i;
    //     if (isave == 0) tempa = v_49;  if (isave == 49) tempb = v0;
    template <int MODE>
    __host__ __device__ static bool better(double a, double b) {
        return MODE == 0 ? (a < b) : (fabs(a) > fabs(b));
    }
    template <int MODE, class F>
    __host__ __device__ void angle_search(F val, double v0, int &isave, double &best, double &tempa,
                                          double &tempb) const {
        const int iu = NU_ANGLES;
#ifdef __CUDA_ARCH__
        if (WARP) {
            const unsigned full = 0xffffffffu;
            const int lane = threadIdx.x & 31;
            const int i1 = lane + 1, i2 = lane + 33;
            const bool has2 = i2 <= iu;
            const NuSinCos a1 = ang[i1], a2 = ang[has2 ? i2 : i1];
            const double v1 = val(a1.s, a1.c), v2 = val(a2.s, a2.c);
            double bv = v0;
            int bi = 0;
            if (better<MODE>(v1, bv)) { bv = v1; bi = i1; }
            if (has2 && better<MODE>(v2, bv)) { bv = v2; bi = i2; }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(full, bv, o);
                const int oi = __shfl_xor_sync(full, bi, o);
                if (better<MODE>(ov, bv) || (!better<MODE>(bv, ov) && oi < bi)) { bv = ov; bi = oi; }
            }
            isave = bi;
            best = bv;
            // neighbours v_{ja}, v_{jb} of the winner
            const int ja = bi == 0 ? iu : bi - 1, jb = bi + 1;          // 0..49, 1..50
            const int sa = ja == 0 ? 1 : ja, sb = jb == iu + 1 ? 1 : jb;  // a valid table slot
            const double a_lo = __shfl_sync(full, v1, (sa - 1) & 31), a_hi = __shfl_sync(full, v2, (sa - 1) & 31);
            const double b_lo = __shfl_sync(full, v1, (sb - 1) & 31), b_hi = __shfl_sync(full, v2, (sb - 1) & 31);
            tempa = ja == 0 ? v0 : (sa <= 32 ? a_lo : a_hi);
            tempb = jb == iu + 1 ? v0 : (sb <= 32 ? b_lo : b_hi);
            // the only lane-dependent branches of the solver are above: re-converge before
            // the lanes go back to updating the (shared) solver state with identical values
            __syncwarp(full);
            return;
        }
#elif defined(NU_HOST_EMULATE_WARP)
        // test scaffolding (tests/native/newuoa2_host.cpp): the lane-parallel
        // selection above, emulated lane by lane on the host
        if (WARP) {
            double v1[32], v2[32], bv[32];
            int bi[32];
            for (int lane = 0; lane < 32; ++lane) {
                const int i1 = lane + 1, i2 = lane + 33;
                const bool has2 = i2 <= iu;
                v1[lane] = val(ang[i1].s, ang[i1].c);
                v2[lane] = val(ang[has2 ? i2 : i1].s, ang[has2 ? i2 : i1].c);
                bv[lane] = v0;
                bi[lane] = 0;
                if (better<MODE>(v1[lane], bv[lane])) { bv[lane] = v1[lane]; bi[lane] = i1; }
                if (has2 && better<MODE>(v2[lane], bv[lane])) { bv[lane] = v2[lane]; bi[lane] = i2; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                double nv[32];
                int ni[32];
                for (int lane = 0; lane < 32; ++lane) {
                    const double ov = bv[lane ^ o];
                    const int oi = bi[lane ^ o];
                    nv[lane] = bv[lane];
                    ni[lane] = bi[lane];
                    if (better<MODE>(ov, bv[lane]) || (!better<MODE>(bv[lane], ov) && oi < bi[lane])) {
                        nv[lane] = ov;
                        ni[lane] = oi;
                    }
                }
                for (int lane = 0; lane < 32; ++lane) { bv[lane] = nv[lane]; bi[lane] = ni[lane]; }
            }
            isave = bi[7];
            best = bv[7];
            const int ja = isave == 0 ? iu : isave - 1, jb = isave + 1;
            const int sa = ja == 0 ? 1 : ja, sb = jb == iu + 1 ? 1 : jb;
            tempa = ja == 0 ? v0 : (sa <= 32 ? v1[(sa - 1) & 31] : v2[(sa - 1) & 31]);
            tempb = jb == iu + 1 ? v0 : (sb <= 32 ? v1[(sb - 1) & 31] : v2[(sb - 1) & 31]);
            return;
        }
#endif
        double prev = v0, v = v0;
        best = v0;
        isave = 0;
        for (int i = 1; i <= iu; ++i) {
            v = val(ang[i].s, ang[i].c);
            if (better<MODE>(v, best)) {
                best = v;
                isave = i;
                tempa = prev;
            } else if (i == isave + 1) {
                tempb = v;
            }
            prev = v;
        }
        if (isave == 0) tempa = v;
        if (isave == iu) tempb = v0;
    }

    // ------------------------------------------------------------------
    __host__ __device__ NU_HOT void trsapp(double *step) {
        const double half = 0.5, zero = 0.0;
        const double twopi = 6.283185307179586476925;
        double dd_[N + 1], g[N + 1], hd[N + 1], hs[N + 1];
        double delsq = delta * delta;
        int iterc = 0;
        const int itermax = N;
        double qred, dd, ds, ss, gg, ggbeg, bstep, dhd, alph, temp, qadd, ggsav;
        double sg, shs, sgk, angtest, tempa = 0, tempb = 0, dg, dhs, cf, qbeg;
        double qmin, angle, cth, sth, reduc, rat;
        int isave;

        for (int i = 1; i <= N; ++i) dd_[i] = xopt[i];
        hess_mul(dd_, hd);
        qred = zero;
        dd = zero;
        for (int i = 1; i <= N; ++i) {
            step[i] = zero;
            hs[i] = zero;
            g[i] = gq[i] + hd[i];
            dd_[i] = -g[i];
            dd += dd_[i] * dd_[i];
        }
        crvmin = zero;
        if (dd == zero) return;
        ds = zero;
        ss = zero;
        gg = dd;
        ggbeg = gg;

        for (;;) {
            ++iterc;
            temp = delsq - ss;
            bstep = temp / (ds + sqrt(ds * ds + dd * temp));
            hess_mul(dd_, hd);
            dhd = zero;
            for (int j = 1; j <= N; ++j) dhd += dd_[j] * hd[j];
            alph = bstep;
            if (dhd > zero) {
                temp = dhd / dd;
                if (iterc == 1) crvmin = temp;
                crvmin = dmin(crvmin, temp);
                alph = dmin(alph, gg / dhd);
            }
            qadd = alph * (gg - half * alph * dhd);
            qred += qadd;
            ggsav = gg;
            gg = zero;
            for (int i = 1; i <= N; ++i) {
                step[i] += alph * dd_[i];
                hs[i] += alph * hd[i];
                double t = g[i] + hs[i];
                gg += t * t;
            }
            if (alph < bstep) {
                if (qadd <= 0.01 * qred) return;
                if (gg <= 1.0e-4 * ggbeg) return;
                if (iterc == itermax) return;
                temp = gg / ggsav;
                dd = zero;
                ds = zero;
                ss = zero;
                for (int i = 1; i <= N; ++i) {
                    dd_[i] = temp * dd_[i] - g[i] - hs[i];
                    dd += dd_[i] * dd_[i];
                    ds += dd_[i] * step[i];
                    ss += step[i] * step[i];
                }
                if (ds <= zero) return;
                if (ss < delsq) continue;
            }
            break;
        }
        crvmin = zero;

        for (;;) {
            if (gg <= 1.0e-4 * ggbeg) return;
            sg = zero;
            shs = zero;
            for (int i = 1; i <= N; ++i) {
                sg += step[i] * g[i];
                shs += step[i] * hs[i];
            }
            sgk = sg + shs;
            angtest = sgk / sqrt(gg * delsq);
            if (angtest <= -0.99) return;
            ++iterc;
            temp = sqrt(delsq * gg - sgk * sgk);
            tempa = delsq / temp;
            tempb = sgk / temp;
            for (int i = 1; i <= N; ++i) dd_[i] = tempa * (g[i] + hs[i]) - tempb * step[i];
            hess_mul(dd_, hd);
            dg = zero;
            dhd = zero;
            dhs = zero;
            for (int i = 1; i <= N; ++i) {
                dg += dd_[i] * g[i];
                dhd += hd[i] * dd_[i];
                dhs += hd[i] * step[i];
            }
            cf = half * (shs - dhd);
            qbeg = sg + cf;
            const int iu = NU_ANGLES;
            temp = twopi / (double)(iu + 1);
            angle_search<0>(
                [=](double sn, double cs) { return (sg + cf * cs) * cs + (dg + dhs * cs) * sn; }, qbeg,
                isave, qmin, tempa, tempb);
            angle = zero;
            if (tempa != tempb) {
                tempa -= qmin;
                tempb -= qmin;
                angle = half * (tempa - tempb) / (tempa + tempb);
            }
            angle = temp * ((double)isave + angle);
            nu_sincos(angle, &sth, &cth);
            reduc = qbeg - (sg + cf * cth) * cth - (dg + dhs * cth) * sth;
            gg = zero;
            for (int i = 1; i <= N; ++i) {
                step[i] = cth * step[i] + sth * dd_[i];
                hs[i] = cth * hs[i] + sth * hd[i];
                double t = g[i] + hs[i];
                gg += t * t;
            }
            qred += reduc;
            rat = reduc / qred;
            if (iterc < itermax && rat > 0.01) continue;
            return;
        }
    }

    // ------------------------------------------------------------------
    // hcol = vlag[1..NPT], gc = vlag[NPT+1..], as in Powell's call
    __host__ __device__ NU_HOT void biglag(double dlt) {
        const double half = 0.5, one = 1.0, zero = 0.0;
        const double twopi = 6.283185307179586476925;
        double *hcol = vlag, *gc = vlag + NPT;
        double gd[N + 1], s[N + 1], ww[N + 1];
        double delsq = dlt * dlt;
        int iterc = 0, isave;
        double temp, sum, dd, gg, sp, dhd, scale, tau = 0, ss, denom;
        double cf1, cf2, cf3, cf4, cf5, taubeg, taumax, angle, cth, sth;
        double tempa = 0, tempb = 0, step;

        NU_ROLLED for (int k = 1; k <= NPT; ++k) hcol[k] = zero;
        for (int j = 1; j <= NPTM; ++j) {
            temp = ZMAT(knew, j);
            if (j < idz) temp = -temp;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) hcol[k] += temp * ZMAT(k, j);
        }
        alpha = hcol[knew];
        dd = zero;
        for (int i = 1; i <= N; ++i) {
            d[i] = XPT(knew, i) - xopt[i];
            gc[i] = BMAT(knew, i);
            gd[i] = zero;
            dd += d[i] * d[i];
        }
        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            temp = zero;
            sum = zero;
            for (int j = 1; j <= N; ++j) {
                temp += XPT(k, j) * xopt[j];
                sum += XPT(k, j) * d[j];
            }
            temp = hcol[k] * temp;
            sum = hcol[k] * sum;
            for (int i = 1; i <= N; ++i) {
                gc[i] += temp * XPT(k, i);
                gd[i] += sum * XPT(k, i);
            }
        }
        gg = zero;
        sp = zero;
        dhd = zero;
        for (int i = 1; i <= N; ++i) {
            gg += gc[i] * gc[i];
            sp += d[i] * gc[i];
            dhd += d[i] * gd[i];
        }
        scale = dlt / sqrt(dd);
        if (sp * dhd < zero) scale = -scale;
        temp = zero;
        if (sp * sp > 0.99 * dd * gg) temp = one;
        tau = scale * (fabs(sp) + half * scale * fabs(dhd));
        if (gg * delsq < 0.01 * tau * tau) temp = one;
        for (int i = 1; i <= N; ++i) {
            d[i] = scale * d[i];
            gd[i] = scale * gd[i];
            s[i] = gc[i] + temp * gd[i];
        }
        for (;;) {
            ++iterc;
            dd = zero;
            sp = zero;
            ss = zero;
            for (int i = 1; i <= N; ++i) {
                dd += d[i] * d[i];
                sp += d[i] * s[i];
                ss += s[i] * s[i];
            }
            temp = dd * ss - sp * sp;
            if (temp <= 1.0e-8 * dd * ss) return;
            denom = sqrt(temp);
            for (int i = 1; i <= N; ++i) {
                s[i] = (dd * s[i] - sp * d[i]) / denom;
                ww[i] = zero;
            }
            NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                sum = zero;
                for (int j = 1; j <= N; ++j) sum += XPT(k, j) * s[j];
                sum = hcol[k] * sum;
                for (int i = 1; i <= N; ++i) ww[i] += sum * XPT(k, i);
            }
            cf1 = cf2 = cf3 = cf4 = cf5 = zero;
            for (int i = 1; i <= N; ++i) {
                cf1 += s[i] * ww[i];
                cf2 += d[i] * gc[i];
                cf3 += s[i] * gc[i];
                cf4 += d[i] * gd[i];
                cf5 += s[i] * gd[i];
            }
            cf1 = half * cf1;
            cf4 = half * cf4 - cf1;
            taubeg = cf1 + cf2 + cf4;
            const int iu = NU_ANGLES;
            temp = twopi / (double)(iu + 1);
            angle_search<1>(
                [=](double sn, double cs) { return cf1 + (cf2 + cf4 * cs) * cs + (cf3 + cf5 * cs) * sn; },
                taubeg, isave, taumax, tempa, tempb);
            step = zero;
            if (tempa != tempb) {
                tempa -= taumax;
                tempb -= taumax;
                step = half * (tempa - tempb) / (tempa + tempb);
            }
            angle = temp * ((double)isave + step);
            nu_sincos(angle, &sth, &cth);
            tau = cf1 + (cf2 + cf4 * cth) * cth + (cf3 + cf5 * cth) * sth;
            for (int i = 1; i <= N; ++i) {
                d[i] = cth * d[i] + sth * s[i];
                gd[i] = cth * gd[i] + sth * ww[i];
                s[i] = gc[i] + gd[i];
            }
            if (fabs(tau) <= 1.1 * fabs(taubeg)) return;
            if (iterc >= N) return;
        }
    }

    // ------------------------------------------------------------------
    __host__ __device__ NU_COLD void bigden() {
        const double half = 0.5, one = 1.0, quart = 0.25, two = 2.0, zero = 0.0;
        const double twopi = 6.283185307179586476925;
        double s[N + 1];
#define WVEC(k, j) wvec_[((k)-1) + ((j)-1) * NDIM]
#define PROD(k, j) prod_[((k)-1) + ((j)-1) * NDIM]
        double temp, alph, dd, ds, ss, xosq, dtest, dstemp, sstemp, diff;
        double ssden, densav, xoptd, xopts, tempa = 0, tempb = 0, tempc, sum;
        double denold, denmax, angle, step, tau;
        int ksav, iterc, isave, nw;

        NU_ROLLED for (int k = 1; k <= NPT; ++k) w[N + k] = zero;
        for (int j = 1; j <= NPTM; ++j) {
            temp = ZMAT(knew, j);
            if (j < idz) temp = -temp;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) w[N + k] += temp * ZMAT(k, j);
        }
        alph = w[N + knew];
        dd = ds = ss = xosq = zero;
        for (int i = 1; i <= N; ++i) {
            dd += d[i] * d[i];
            s[i] = XPT(knew, i) - xopt[i];
            ds += d[i] * s[i];
            ss += s[i] * s[i];
            xosq += xopt[i] * xopt[i];
        }
        if (ds * ds > 0.99 * dd * ss) {
            ksav = knew;
            dtest = ds * ds / ss;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                if (k != kopt) {
                    dstemp = zero;
                    sstemp = zero;
                    for (int i = 1; i <= N; ++i) {
                        diff = XPT(k, i) - xopt[i];
                        dstemp += d[i] * diff;
                        sstemp += diff * diff;
                    }
                    if (dstemp * dstemp / sstemp < dtest) {
                        ksav = k;
                        dtest = dstemp * dstemp / sstemp;
                        ds = dstemp;
                        ss = sstemp;
                    }
                }
            }
            for (int i = 1; i <= N; ++i) s[i] = XPT(ksav, i) - xopt[i];
        }
        ssden = dd * ss - ds * ds;
        iterc = 0;
        densav = zero;
        for (;;) {
            ++iterc;
            temp = one / sqrt(ssden);
            xoptd = zero;
            xopts = zero;
            for (int i = 1; i <= N; ++i) {
                s[i] = temp * (dd * s[i] - ds * d[i]);
                xoptd += xopt[i] * d[i];
                xopts += xopt[i] * s[i];
            }
            tempa = half * xoptd * xoptd;
            tempb = half * xopts * xopts;
            den[1] = dd * (xosq + half * dd) + tempa + tempb;
            den[2] = two * xoptd * dd;
            den[3] = two * xopts * dd;
            den[4] = tempa - tempb;
            den[5] = xoptd * xopts;
            for (int i = 6; i <= 9; ++i) den[i] = zero;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                tempa = tempb = tempc = zero;
                for (int i = 1; i <= N; ++i) {
                    tempa += XPT(k, i) * d[i];
                    tempb += XPT(k, i) * s[i];
                    tempc += XPT(k, i) * xopt[i];
                }
                WVEC(k, 1) = quart * (tempa * tempa + tempb * tempb);
                WVEC(k, 2) = tempa * tempc;
                WVEC(k, 3) = tempb * tempc;
                WVEC(k, 4) = quart * (tempa * tempa - tempb * tempb);
                WVEC(k, 5) = half * tempa * tempb;
            }
            for (int i = 1; i <= N; ++i) {
                int ip = i + NPT;
                WVEC(ip, 1) = zero;
                WVEC(ip, 2) = d[i];
                WVEC(ip, 3) = s[i];
                WVEC(ip, 4) = zero;
                WVEC(ip, 5) = zero;
            }
            NU_ROLLED for (int jc = 1; jc <= 5; ++jc) {
                nw = NPT;
                if (jc == 2 || jc == 3) nw = NDIM;
                NU_ROLLED for (int k = 1; k <= NPT; ++k) PROD(k, jc) = zero;
                for (int j = 1; j <= NPTM; ++j) {
                    sum = zero;
                    NU_ROLLED for (int k = 1; k <= NPT; ++k) sum += ZMAT(k, j) * WVEC(k, jc);
                    if (j < idz) sum = -sum;
                    NU_ROLLED for (int k = 1; k <= NPT; ++k) PROD(k, jc) += sum * ZMAT(k, j);
                }
                if (nw == NDIM) {
                    NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                        sum = zero;
                        for (int j = 1; j <= N; ++j) sum += BMAT(k, j) * WVEC(NPT + j, jc);
                        PROD(k, jc) += sum;
                    }
                }
                for (int j = 1; j <= N; ++j) {
                    sum = zero;
                    NU_ROLLED for (int i = 1; i <= nw; ++i) sum += BMAT(i, j) * WVEC(i, jc);
                    PROD(NPT + j, jc) = sum;
                }
            }
            NU_ROLLED for (int k = 1; k <= NDIM; ++k) {
                sum = zero;
                for (int i = 1; i <= 5; ++i) {
                    par[i] = half * PROD(k, i) * WVEC(k, i);
                    sum += par[i];
                }
                den[1] = den[1] - par[1] - sum;
                tempa = PROD(k, 1) * WVEC(k, 2) + PROD(k, 2) * WVEC(k, 1);
                tempb = PROD(k, 2) * WVEC(k, 4) + PROD(k, 4) * WVEC(k, 2);
                tempc = PROD(k, 3) * WVEC(k, 5) + PROD(k, 5) * WVEC(k, 3);
                den[2] = den[2] - tempa - half * (tempb + tempc);
                den[6] = den[6] - half * (tempb - tempc);
                tempa = PROD(k, 1) * WVEC(k, 3) + PROD(k, 3) * WVEC(k, 1);
                tempb = PROD(k, 2) * WVEC(k, 5) + PROD(k, 5) * WVEC(k, 2);
                tempc = PROD(k, 3) * WVEC(k, 4) + PROD(k, 4) * WVEC(k, 3);
                den[3] = den[3] - tempa - half * (tempb - tempc);
                den[7] = den[7] - half * (tempb + tempc);
                tempa = PROD(k, 1) * WVEC(k, 4) + PROD(k, 4) * WVEC(k, 1);
                den[4] = den[4] - tempa - par[2] + par[3];
                tempa = PROD(k, 1) * WVEC(k, 5) + PROD(k, 5) * WVEC(k, 1);
                tempb = PROD(k, 2) * WVEC(k, 3) + PROD(k, 3) * WVEC(k, 2);
                den[5] = den[5] - tempa - half * tempb;
                den[8] = den[8] - par[4] + par[5];
                tempa = PROD(k, 4) * WVEC(k, 5) + PROD(k, 5) * WVEC(k, 4);
                den[9] = den[9] - half * tempa;
            }
            sum = zero;
            for (int i = 1; i <= 5; ++i) {
                par[i] = half * PROD(knew, i) * PROD(knew, i);
                sum += par[i];
            }
            denex[1] = alph * den[1] + par[1] + sum;
            tempa = two * PROD(knew, 1) * PROD(knew, 2);
            tempb = PROD(knew, 2) * PROD(knew, 4);
            tempc = PROD(knew, 3) * PROD(knew, 5);
            denex[2] = alph * den[2] + tempa + tempb + tempc;
            denex[6] = alph * den[6] + tempb - tempc;
            tempa = two * PROD(knew, 1) * PROD(knew, 3);
            tempb = PROD(knew, 2) * PROD(knew, 5);
            tempc = PROD(knew, 3) * PROD(knew, 4);
            denex[3] = alph * den[3] + tempa + tempb - tempc;
            denex[7] = alph * den[7] + tempb + tempc;
            tempa = two * PROD(knew, 1) * PROD(knew, 4);
            denex[4] = alph * den[4] + tempa + par[2] - par[3];
            tempa = two * PROD(knew, 1) * PROD(knew, 5);
            denex[5] = alph * den[5] + tempa + PROD(knew, 2) * PROD(knew, 3);
            denex[8] = alph * den[8] + par[4] - par[5];
            denex[9] = alph * den[9] + PROD(knew, 4) * PROD(knew, 5);

            sum = denex[1] + denex[2] + denex[4] + denex[6] + denex[8];
            denold = sum;
            const int iu = NU_ANGLES;
            temp = twopi / (double)(iu + 1);
            par[1] = one;
            {
                const double e1 = denex[1], e2 = denex[2], e3 = denex[3], e4 = denex[4], e5 = denex[5],
                             e6 = denex[6], e7 = denex[7], e8 = denex[8], e9 = denex[9];
                angle_search<1>(
                    [=](double sn, double cs) {
                        const double p4 = cs * cs - sn * sn, p5 = cs * sn + sn * cs;
                        const double p6 = cs * p4 - sn * p5, p7 = cs * p5 + sn * p4;
                        const double p8 = cs * p6 - sn * p7, p9 = cs * p7 + sn * p6;
                        double sm = 0.0;
                        sm += e1 * 1.0;
                        sm += e2 * cs;
                        sm += e3 * sn;
                        sm += e4 * p4;
                        sm += e5 * p5;
                        sm += e6 * p6;
                        sm += e7 * p7;
                        sm += e8 * p8;
                        sm += e9 * p9;
                        return sm;
                    },
                    denold, isave, denmax, tempa, tempb);
            }
            step = zero;
            if (tempa != tempb) {
                tempa -= denmax;
                tempb -= denmax;
                step = half * (tempa - tempb) / (tempa + tempb);
            }
            angle = temp * ((double)isave + step);
            nu_sincos(angle, &par[3], &par[2]);
            for (int j = 4; j <= 8; j += 2) {
                par[j] = par[2] * par[j - 2] - par[3] * par[j - 1];
                par[j + 1] = par[2] * par[j - 1] + par[3] * par[j - 2];
            }
            beta = zero;
            denmax = zero;
            for (int j = 1; j <= 9; ++j) {
                beta += den[j] * par[j];
                denmax += denex[j] * par[j];
            }
            NU_ROLLED for (int k = 1; k <= NDIM; ++k) {
                vlag[k] = zero;
                for (int j = 1; j <= 5; ++j) vlag[k] += PROD(k, j) * par[j];
            }
            tau = vlag[knew];
            dd = zero;
            tempa = zero;
            tempb = zero;
            for (int i = 1; i <= N; ++i) {
                d[i] = par[2] * d[i] + par[3] * s[i];
                w[i] = xopt[i] + d[i];
                dd += d[i] * d[i];
                tempa += d[i] * w[i];
                tempb += w[i] * w[i];
            }
            if (iterc >= N) break;
            if (iterc > 1) densav = dmax(densav, denold);
            if (fabs(denmax) <= 1.1 * fabs(densav)) break;
            densav = denmax;
            for (int i = 1; i <= N; ++i) {
                temp = tempa * xopt[i] + tempb * d[i] - vlag[NPT + i];
                s[i] = tau * BMAT(knew, i) + alph * temp;
            }
            NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                sum = zero;
                for (int j = 1; j <= N; ++j) sum += XPT(k, j) * w[j];
                temp = (tau * w[N + k] - alph * vlag[k]) * sum;
                for (int i = 1; i <= N; ++i) s[i] += temp * XPT(k, i);
            }
            ss = zero;
            ds = zero;
            for (int i = 1; i <= N; ++i) {
                ss += s[i] * s[i];
                ds += d[i] * s[i];
            }
            ssden = dd * ss - ds * ds;
            if (ssden >= 1.0e-8 * dd * ss) continue;
            break;
        }
        NU_ROLLED for (int k = 1; k <= NDIM; ++k) {
            w[k] = zero;
            for (int j = 1; j <= 5; ++j) w[k] += WVEC(k, j) * par[j];
        }
        vlag[kopt] += one;
#undef WVEC
#undef PROD
    }

    // ------------------------------------------------------------------
    __host__ __device__ NU_HOT void update() {
        const double one = 1.0, zero = 0.0;
        int jl = 1, iflag, ja, jb;
        double temp, tempa, tempb = 0, alph, tau, tausq, denom, scala, scalb;
        for (int j = 2; j <= NPTM; ++j) {
            if (j == idz) {
                jl = idz;
            } else if (ZMAT(knew, j) != zero) {
                temp = sqrt(ZMAT(knew, jl) * ZMAT(knew, jl) + ZMAT(knew, j) * ZMAT(knew, j));
                tempa = ZMAT(knew, jl) / temp;
                tempb = ZMAT(knew, j) / temp;
                NU_ROLLED for (int i = 1; i <= NPT; ++i) {
                    temp = tempa * ZMAT(i, jl) + tempb * ZMAT(i, j);
                    ZMAT(i, j) = tempa * ZMAT(i, j) - tempb * ZMAT(i, jl);
                    ZMAT(i, jl) = temp;
                }
                ZMAT(knew, j) = zero;
            }
        }
        tempa = ZMAT(knew, 1);
        if (idz >= 2) tempa = -tempa;
        if (jl > 1) tempb = ZMAT(knew, jl);
        NU_ROLLED for (int i = 1; i <= NPT; ++i) {
            w[i] = tempa * ZMAT(i, 1);
            if (jl > 1) w[i] += tempb * ZMAT(i, jl);
        }
        alph = w[knew];
        tau = vlag[knew];
        tausq = tau * tau;
        denom = alph * beta + tausq;
        vlag[knew] -= one;
        iflag = 0;
        if (jl == 1) {
            temp = sqrt(fabs(denom));
            tempb = tempa / temp;
            tempa = tau / temp;
            NU_ROLLED for (int i = 1; i <= NPT; ++i) ZMAT(i, 1) = tempa * ZMAT(i, 1) - tempb * vlag[i];
            // Powell's published tests use TEMP (>= 0) here, kept as published
            if (idz == 1 && temp < zero) idz = 2;
            if (idz >= 2 && temp >= zero) iflag = 1;
        } else {
            ja = 1;
            if (beta >= zero) ja = jl;
            jb = jl + 1 - ja;
            temp = ZMAT(knew, jb) / denom;
            tempa = temp * beta;
            tempb = temp * tau;
            temp = ZMAT(knew, ja);
            scala = one / sqrt(fabs(beta) * temp * temp + tausq);
            scalb = scala * sqrt(fabs(denom));
            NU_ROLLED for (int i = 1; i <= NPT; ++i) {
                ZMAT(i, ja) = scala * (tau * ZMAT(i, ja) - temp * vlag[i]);
                ZMAT(i, jb) = scalb * (ZMAT(i, jb) - tempa * w[i] - tempb * vlag[i]);
            }
            if (denom <= zero) {
                if (beta < zero) idz = idz + 1;
                if (beta >= zero) iflag = 1;
            }
        }
        if (iflag == 1) {
            idz = idz - 1;
            NU_ROLLED for (int i = 1; i <= NPT; ++i) {
                temp = ZMAT(i, 1);
                ZMAT(i, 1) = ZMAT(i, idz);
                ZMAT(i, idz) = temp;
            }
        }
        for (int j = 1; j <= N; ++j) {
            int jp = NPT + j;
            w[jp] = BMAT(knew, j);
            tempa = (alph * vlag[jp] - tau * w[jp]) / denom;
            tempb = (-beta * w[jp] - tau * vlag[jp]) / denom;
            NU_ROLLED for (int i = 1; i <= jp; ++i) {
                BMAT(i, j) = BMAT(i, j) + tempa * vlag[i] + tempb * w[i];
                if (i > NPT) BMAT(jp, i - NPT) = BMAT(i, j);
            }
        }
    }

    // ------------------------------------------------------------------
    // Shift XBASE to XOPT (taken when the step is small against |xopt|): rare and
    // large, kept out of line so that the hot path of step() stays compact.
    __host__ __device__ NU_COLD void shift_xbase() {
        const double half = 0.5, zero = 0.0;
        double temp, tempq, sum, sumz;
        int ih, ip;
        tempq = 0.25 * xoptsq;
        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            sum = zero;
            for (int i = 1; i <= N; ++i) sum += XPT(k, i) * xopt[i];
            temp = pq[k] * sum;
            sum -= half * xoptsq;
            w[NPT + k] = sum;
            for (int i = 1; i <= N; ++i) {
                gq[i] += temp * XPT(k, i);
                XPT(k, i) -= half * xopt[i];
                vlag[i] = BMAT(k, i);
                w[i] = sum * XPT(k, i) + tempq * xopt[i];
                ip = NPT + i;
                for (int j = 1; j <= i; ++j)
                    BMAT(ip, j) = BMAT(ip, j) + vlag[i] * w[j] + w[i] * vlag[j];
            }
        }
        NU_ROLLED for (int k = 1; k <= NPTM; ++k) {
            sumz = zero;
            NU_ROLLED for (int i = 1; i <= NPT; ++i) {
                sumz += ZMAT(i, k);
                w[i] = w[NPT + i] * ZMAT(i, k);
            }
            for (int j = 1; j <= N; ++j) {
                sum = tempq * sumz * xopt[j];
                NU_ROLLED for (int i = 1; i <= NPT; ++i) sum += w[i] * XPT(i, j);
                vlag[j] = sum;
                if (k < idz) sum = -sum;
                NU_ROLLED for (int i = 1; i <= NPT; ++i) BMAT(i, j) = BMAT(i, j) + sum * ZMAT(i, k);
            }
            for (int i = 1; i <= N; ++i) {
                ip = i + NPT;
                temp = vlag[i];
                if (k < idz) temp = -temp;
                for (int j = 1; j <= i; ++j) BMAT(ip, j) = BMAT(ip, j) + temp * vlag[j];
            }
        }
        ih = 0;
        for (int j = 1; j <= N; ++j) {
            w[j] = zero;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                w[j] += pq[k] * XPT(k, j);
                XPT(k, j) -= half * xopt[j];
            }
            for (int i = 1; i <= j; ++i) {
                ++ih;
                if (i < j) gq[j] += hq[ih] * xopt[i];
                gq[i] += hq[ih] * xopt[j];
                hq[ih] = hq[ih] + w[i] * xopt[j] + xopt[i] * w[j];
                BMAT(NPT + i, j) = BMAT(NPT + j, i);
            }
        }
        for (int j = 1; j <= N; ++j) {
            xbase[j] += xopt[j];
            xopt[j] = zero;
        }
        xoptsq = zero;
    }

    // ------------------------------------------------------------------
    // Advance the solver.  On the first call `fin` is ignored; afterwards it
    // is the objective value at the point x[] returned by the previous call.
    // Returns true when x[] must be evaluated, false when finished.
    // ------------------------------------------------------------------
    // NEWUOB as a resumable machine of SEGMENTS (Powell's labelled blocks: 50, 70, 90/100,
    // 120, 290, 310, after CALFUN, 410, 460, 490, 530).  advance() runs the segment `pc`
    // points at and sets pc to the next one; step() runs segments until the solver wants an
    // objective value (true) or has finished (false).  The arithmetic and its order are those
    // of the straight-line version; the segmentation only exists so that the threads of a warp
    // which each own a solver (thread-per-fit kernel) can be scheduled segment by segment
    // (step_coop): all lanes that are at the same segment run it together, instead of every
    // lane dragging the others through its own path from the first divergent branch on.
    // Segment numbers grow along the usual flow of one iteration (after CALFUN -> 410 -> 460 ->
    // 490 -> 90/100 -> 120 -> 290 -> 310): scheduling the smallest pc first lets the lanes that
    // take a detour catch up with the others before the expensive common segments.
    enum {
        PC_INIT = 0, PC_L50, PC_L70, PC_AFTER, PC_L410, PC_L460, PC_L490, PC_L90, PC_L100, PC_L120,
        PC_L290, PC_L310, PC_L530, PC_YIELD /* objective value wanted */, PC_DONE, PC_IDLE
    };
    int pc;
    // values that live from the segment after CALFUN into 410 / 460
    double s_diff, s_fsave, s_vquad;
    int s_ksave;

    __host__ __device__ void enter(double fin) {
        if (phase == 2) {
            pc = PC_DONE;
        } else if (phase == 1) {
            fcur = fin;
            pc = PC_AFTER;
        } else {
            pc = PC_INIT;
        }
    }

    __host__ __device__ bool step(double fin) {
        enter(fin);
        while (pc < PC_YIELD) advance();
        return pc == PC_YIELD;
    }

#ifdef __CUDACC__
    // All 32 lanes of a warp call this together, each with its own solver; a lane without
    // work passes idle = true.  Returns like step().
    __device__ bool step_coop(double fin, bool idle) {
        if (idle) pc = PC_IDLE;
        else enter(fin);
        for (;;) {
            const int cur = __reduce_min_sync(0xffffffffu, pc);
            if (cur >= PC_YIELD) break;
            if (pc == cur) advance();
        }
        return pc == PC_YIELD;
    }
#endif

    __host__ __device__ void advance() {
        const double half = 0.5, one = 1.0, tenth = 0.1, zero = 0.0;
        double temp, sum, suma, sumb, bsum, dx;
        double detrat, hdiag, distsq, gqsq, gisq;
        int ih, itemp, ktemp;

        switch (pc) {
        case PC_INIT:
        // ---- set-up (first call) ----
        for (int j = 1; j <= N; ++j) xbase[j] = x[j];
        for (int i = 0; i < NPT * N; ++i) xpt_[i] = zero;
        for (int i = 0; i < NDIM * N; ++i) bmat_[i] = zero;
        for (int i = 0; i < NPT * NPTM; ++i) zmat_[i] = zero;
        for (ih = 1; ih <= NH; ++ih) hq[ih] = zero;
        NU_ROLLED for (int k = 1; k <= NPT; ++k) pq[k] = zero;
        for (int k = 0; k <= NPT; ++k) fval[k] = zero;
        for (int k = 0; k <= N; ++k) gq[k] = zero;
        // the rest of the work space starts from zero as well (the oracle's is calloc'ed):
        // with maxfun < npt the solver returns xbase + xopt before xopt is ever assigned
        for (int k = 0; k <= N; ++k) xopt[k] = xnew[k] = d[k] = zero;
        for (int k = 0; k <= NDIM; ++k) vlag[k] = zero;
        for (int k = 0; k <= 2 * NDIM + 2 * NPT; ++k) w[k] = zero;
        rhosq = rhobeg * rhobeg;
        recip = one / rhosq;
        reciq = sqrt(half) / rhosq;
        nf = 0;
        kopt = 1;
        idz = 1;
        itest = 0;
        nfsav = 0;
        knew = 0;
        ipt = jpt = 0;
        xipt = xjpt = zero;
        fbeg = fopt = fcur = zero;
        rho = delta = diffa = diffb = diffc = xoptsq = dsq = dnorm = zero;
        ratio = crvmin = beta = alpha = dstep = zero;
        s_diff = s_fsave = s_vquad = zero;
        s_ksave = 0;
        pc = PC_L50;
        return;

        case PC_L50:
        nfm = nf;
        nfmm = nf - N;
        ++nf;
        if (nfm <= 2 * N) {
            if (nfm >= 1 && nfm <= N) {
                XPT(nf, nfm) = rhobeg;
            } else if (nfm > N) {
                XPT(nf, nfmm) = -rhobeg;
            }
        } else {
            itemp = (nfmm - 1) / N;
            jpt = nfm - itemp * N - N;
            ipt = jpt + itemp;
            if (ipt > N) {
                itemp = jpt;
                jpt = ipt - N;
                ipt = itemp;
            }
            xipt = rhobeg;
            if (fval[ipt + NP] < fval[ipt + 1]) xipt = -xipt;
            xjpt = rhobeg;
            if (fval[jpt + NP] < fval[jpt + 1]) xjpt = -xjpt;
            XPT(nf, ipt) = xipt;
            XPT(nf, jpt) = xjpt;
        }
        for (int j = 1; j <= N; ++j) x[j] = XPT(nf, j) + xbase[j];
        pc = PC_L310;
        return;

        case PC_L70:
        fval[nf] = fcur;
        if (nf == 1) {
            fbeg = fcur;
            fopt = fcur;
            kopt = 1;
        } else if (fcur < fopt) {
            fopt = fcur;
            kopt = nf;
        }
        if (nfm <= 2 * N) {
            if (nfm >= 1 && nfm <= N) {
                gq[nfm] = (fcur - fbeg) / rhobeg;
                if (NPT < nf + N) {
                    BMAT(1, nfm) = -one / rhobeg;
                    BMAT(nf, nfm) = one / rhobeg;
                    BMAT(NPT + nfm, nfm) = -half * rhosq;
                }
            } else if (nfm > N) {
                BMAT(nf - N, nfmm) = half / rhobeg;
                BMAT(nf, nfmm) = -half / rhobeg;
                ZMAT(1, nfmm) = -reciq - reciq;
                ZMAT(nf - N, nfmm) = reciq;
                ZMAT(nf, nfmm) = reciq;
                ih = (nfmm * (nfmm + 1)) / 2;
                temp = (fbeg - fcur) / rhobeg;
                hq[ih] = (gq[nfmm] - temp) / rhobeg;
                gq[nfmm] = half * (gq[nfmm] + temp);
            }
        } else {
            ih = (ipt * (ipt - 1)) / 2 + jpt;
            if (xipt < zero) ipt += N;
            if (xjpt < zero) jpt += N;
            ZMAT(1, nfmm) = recip;
            ZMAT(nf, nfmm) = recip;
            ZMAT(ipt + 1, nfmm) = -recip;
            ZMAT(jpt + 1, nfmm) = -recip;
            hq[ih] = (fbeg - fval[ipt + 1] - fval[jpt + 1] + fcur) / (xipt * xjpt);
        }
        if (nf < NPT) {
            pc = PC_L50;
            return;
        }

        rho = rhobeg;
        delta = rho;
        idz = 1;
        diffa = zero;
        diffb = zero;
        itest = 0;
        xoptsq = zero;
        for (int i = 1; i <= N; ++i) {
            xopt[i] = XPT(kopt, i);
            xoptsq += xopt[i] * xopt[i];
        }
        pc = PC_L90;
        return;

        case PC_L90:
        nfsav = nf;
        pc = PC_L100;
        return;

        case PC_L100:
        knew = 0;
        trsapp(d);
        dsq = zero;
        for (int i = 1; i <= N; ++i) dsq += d[i] * d[i];
        dnorm = dmin(delta, sqrt(dsq));
        if (dnorm < half * rho) {
            knew = -1;
            delta = tenth * delta;
            ratio = -1.0;
            if (delta <= 1.5 * rho) delta = rho;
            if (nf <= nfsav + 2) {
                pc = PC_L460;
                return;
            }
            temp = 0.125 * crvmin * rho * rho;
            if (temp <= dmax(diffa, dmax(diffb, diffc))) {
                pc = PC_L460;
                return;
            }
            pc = PC_L490;
            return;
        }
        pc = PC_L120;
        return;

        case PC_L120:
        if (dsq <= 1.0e-3 * xoptsq) shift_xbase();
        if (knew > 0) biglag(dstep);

        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            suma = zero;
            sumb = zero;
            sum = zero;
            for (int j = 1; j <= N; ++j) {
                suma += XPT(k, j) * d[j];
                sumb += XPT(k, j) * xopt[j];
                sum += BMAT(k, j) * d[j];
            }
            w[k] = suma * (half * suma + sumb);
            vlag[k] = sum;
        }
        beta = zero;
        NU_ROLLED for (int k = 1; k <= NPTM; ++k) {
            sum = zero;
            NU_ROLLED for (int i = 1; i <= NPT; ++i) sum += ZMAT(i, k) * w[i];
            if (k < idz) {
                beta += sum * sum;
                sum = -sum;
            } else {
                beta -= sum * sum;
            }
            NU_ROLLED for (int i = 1; i <= NPT; ++i) vlag[i] += sum * ZMAT(i, k);
        }
        bsum = zero;
        dx = zero;
        for (int j = 1; j <= N; ++j) {
            sum = zero;
            NU_ROLLED for (int i = 1; i <= NPT; ++i) sum += w[i] * BMAT(i, j);
            bsum += sum * d[j];
            int jp = NPT + j;
            for (int k = 1; k <= N; ++k) sum += BMAT(jp, k) * d[k];
            vlag[jp] = sum;
            bsum += sum * d[j];
            dx += d[j] * xopt[j];
        }
        beta = dx * dx + dsq * (xoptsq + dx + dx + half * dsq) + beta - bsum;
        vlag[kopt] += one;
        if (knew > 0) {
            temp = one + alpha * beta / (vlag[knew] * vlag[knew]);
            if (fabs(temp) <= 0.8) bigden();
        }
        pc = PC_L290;
        return;

        case PC_L290:
        for (int i = 1; i <= N; ++i) {
            xnew[i] = xopt[i] + d[i];
            x[i] = xbase[i] + xnew[i];
        }
        ++nf;
        pc = PC_L310;
        return;

        case PC_L310:
        if (nf > nftest) {
            --nf;
            status = NU_TOO_MANY_EVALUATIONS;
            pc = PC_L530;
            return;
        }
        phase = 1;
        pc = PC_YIELD;  // ---- objective call ----
        return;

        case PC_AFTER:
        if (nf <= NPT) {
            pc = PC_L70;
            return;
        }
        if (knew == -1) {
            pc = PC_L530;
            return;
        }

        s_vquad = zero;
        ih = 0;
        for (int j = 1; j <= N; ++j) {
            s_vquad += d[j] * gq[j];
            for (int i = 1; i <= j; ++i) {
                ++ih;
                temp = d[i] * xnew[j] + d[j] * xopt[i];
                if (i == j) temp = half * temp;
                s_vquad += temp * hq[ih];
            }
        }
        NU_ROLLED for (int k = 1; k <= NPT; ++k) s_vquad += pq[k] * w[k];
        s_diff = fcur - fopt - s_vquad;
        diffc = diffb;
        diffb = diffa;
        diffa = fabs(s_diff);
        if (dnorm > rho) nfsav = nf;
        s_fsave = fopt;
        if (fcur < fopt) {
            fopt = fcur;
            xoptsq = zero;
            for (int i = 1; i <= N; ++i) {
                xopt[i] = xnew[i];
                xoptsq += xopt[i] * xopt[i];
            }
        }
        s_ksave = knew;
        if (knew > 0) {
            pc = PC_L410;
            return;
        }
        if (s_vquad >= zero) {
            status = NU_ROUNDING_ERRORS;
            pc = PC_L530;
            return;
        }
        ratio = (fcur - s_fsave) / s_vquad;
        if (ratio <= tenth) {
            delta = half * dnorm;
        } else if (ratio <= 0.7) {
            delta = dmax(half * delta, dnorm);
        } else {
            delta = dmax(half * delta, dnorm + dnorm);
        }
        if (delta <= 1.5 * rho) delta = rho;
        rhosq = dmax(tenth * delta, rho);
        rhosq = rhosq * rhosq;
        ktemp = 0;
        detrat = zero;
        if (fcur >= s_fsave) {
            ktemp = kopt;
            detrat = one;
        }
        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            hdiag = zero;
            for (int j = 1; j <= NPTM; ++j) {
                temp = one;
                if (j < idz) temp = -one;
                hdiag += temp * ZMAT(k, j) * ZMAT(k, j);
            }
            temp = fabs(beta * hdiag + vlag[k] * vlag[k]);
            distsq = zero;
            for (int j = 1; j <= N; ++j) {
                double t = XPT(k, j) - xopt[j];
                distsq += t * t;
            }
            if (distsq > rhosq) {
                double r = distsq / rhosq;
                temp = temp * (r * r * r);
            }
            if (temp > detrat && k != ktemp) {
                detrat = temp;
                knew = k;
            }
        }
        pc = knew == 0 ? PC_L460 : PC_L410;
        return;

        case PC_L410:
        update();
        fval[knew] = fcur;
        ih = 0;
        for (int i = 1; i <= N; ++i) {
            temp = pq[knew] * XPT(knew, i);
            for (int j = 1; j <= i; ++j) {
                ++ih;
                hq[ih] += temp * XPT(knew, j);
            }
        }
        pq[knew] = zero;
        for (int j = 1; j <= NPTM; ++j) {
            temp = s_diff * ZMAT(knew, j);
            if (j < idz) temp = -temp;
            NU_ROLLED for (int k = 1; k <= NPT; ++k) pq[k] += temp * ZMAT(k, j);
        }
        gqsq = zero;
        for (int i = 1; i <= N; ++i) {
            gq[i] += s_diff * BMAT(knew, i);
            gqsq += gq[i] * gq[i];
            XPT(knew, i) = xnew[i];
        }
        if (s_ksave == 0 && delta == rho) {
            if (fabs(ratio) > 1.0e-2) {
                itest = 0;
            } else {
                NU_ROLLED for (int k = 1; k <= NPT; ++k) vlag[k] = fval[k] - fval[kopt];
                gisq = zero;
                for (int i = 1; i <= N; ++i) {
                    sum = zero;
                    NU_ROLLED for (int k = 1; k <= NPT; ++k) sum += BMAT(k, i) * vlag[k];
                    gisq += sum * sum;
                    w[i] = sum;
                }
                ++itest;
                if (gqsq < 1.0e2 * gisq) itest = 0;
                if (itest >= 3) {
                    for (int i = 1; i <= N; ++i) gq[i] = w[i];
                    for (ih = 1; ih <= NH; ++ih) hq[ih] = zero;
                    for (int j = 1; j <= NPTM; ++j) {
                        w[j] = zero;
                        NU_ROLLED for (int k = 1; k <= NPT; ++k) w[j] += vlag[k] * ZMAT(k, j);
                        if (j < idz) w[j] = -w[j];
                    }
                    NU_ROLLED for (int k = 1; k <= NPT; ++k) {
                        pq[k] = zero;
                        for (int j = 1; j <= NPTM; ++j) pq[k] += ZMAT(k, j) * w[j];
                    }
                    itest = 0;
                }
            }
        }
        if (fcur < s_fsave) kopt = knew;
        if (fcur <= s_fsave + tenth * s_vquad) {
            pc = PC_L100;
            return;
        }
        if (s_ksave > 0) {
            pc = PC_L100;
            return;
        }
        knew = 0;
        pc = PC_L460;
        return;

        case PC_L460:
        distsq = 4.0 * delta * delta;
        NU_ROLLED for (int k = 1; k <= NPT; ++k) {
            sum = zero;
            for (int j = 1; j <= N; ++j) {
                double t = XPT(k, j) - xopt[j];
                sum += t * t;
            }
            if (sum > distsq) {
                knew = k;
                distsq = sum;
            }
        }
        if (knew > 0) {
            dstep = dmax(dmin(tenth * sqrt(distsq), half * delta), rho);
            dsq = dstep * dstep;
            pc = PC_L120;
            return;
        }
        if (ratio > zero) {
            pc = PC_L100;
            return;
        }
        if (dmax(delta, dnorm) > rho) {
            pc = PC_L100;
            return;
        }
        pc = PC_L490;
        return;

        case PC_L490:
        if (rho > rhoend) {
            delta = half * rho;
            ratio = rho / rhoend;
            if (ratio <= 16.0) {
                rho = rhoend;
            } else if (ratio <= 250.0) {
                rho = sqrt(ratio) * rhoend;
            } else {
                rho = tenth * rho;
            }
            delta = dmax(delta, rho);
            pc = PC_L90;
            return;
        }
        pc = knew == -1 ? PC_L290 : PC_L530;
        return;

        case PC_L530:
        if (fopt <= fcur) {
            for (int i = 1; i <= N; ++i) x[i] = xbase[i] + xopt[i];
            fcur = fopt;
        }
        f = fcur;
        phase = 2;
        pc = PC_DONE;
        return;

        default:
            return;
        }
    }

#undef XPT
#undef BMAT
#undef ZMAT
};

typedef Newuoa2T<false> Newuoa2;

}  // namespace gppd
