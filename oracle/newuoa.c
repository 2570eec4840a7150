/*
 * oracle/newuoa.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  See newuoa.h.
 *
 * General-n restatement of Powell's NEWUOA.  The reference reaches it at
 * src/Modulation.jl:335 as  newuoa(f, xinit, 1, 1e-3; check=false)  with
 * n = 2; OptimPackNextGen's defaults then give npt = 2n+1 = 5 and
 * maxeval = 30n = 60 [third-party, not in the reference tree].
 *
 * Arrays are addressed 1-based through macros so that every formula can be
 * checked against Powell's report (DAMTP 2004/NA05) term by term:
 *   XPT(k,j)   k-th interpolation point, displacement from XBASE
 *   BMAT(i,j)  last n columns of the inverse KKT matrix H  (ndim x n)
 *   ZMAT(k,j)  factor of the leading npt x npt block of H  (npt x nptm)
 *   HQ(ih)     explicit second-derivative part of the model (packed upper)
 *   PQ(k)      implicit second-derivative coefficients
 * Compile with -ffp-contract=off: the reference stack does not fuse.
 */
#include "newuoa.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define XPT(k, j) xpt[((k)-1) + (size_t)((j)-1) * npt]
#define BMAT(i, j) bmat[((i)-1) + (size_t)((j)-1) * ndim]
#define ZMAT(k, j) zmat[((k)-1) + (size_t)((j)-1) * npt]
#define WVEC(k, j) wvec[((k)-1) + (size_t)((j)-1) * ndim]
#define PROD(k, j) prod[((k)-1) + (size_t)((j)-1) * ndim]

static newuoa_probe g_probe = 0;
static void *g_probe_data = 0;
void newuoa_oracle_set_probe(newuoa_probe p, void *data) {
    g_probe = p;
    g_probe_data = data;
}

/* call counters for tests: trsapp, biglag, bigden, update, xbase shifts */
static long g_counters[5];
static int g_force_bigden = 0; /* tests: take the BIGDEN branch on every model step */
void newuoa_oracle_force_bigden(int on) { g_force_bigden = on; }
void newuoa_oracle_counters(long *out5, int reset) {
    for (int i = 0; i < 5; ++i) {
        if (out5) out5[i] = g_counters[i];
        if (reset) g_counters[i] = 0;
    }
}

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }


/* ---------------------------------------------------------------------- */
/* Portable sin/cos for the solver's own angle searches (|x| <~ 8).
 * Powell's code calls DCOS/DSIN on i*2pi/50 and on one interpolated angle;
 * libm results differ in the last bit between platforms, and NEWUOA has
 * rounding-level ties (e.g. SUM > DISTSQ right after DELTA = HALF*DNORM), so
 * the oracle uses a fixed arithmetic here: two-term Cody-Waite reduction by
 * pi/2 and the classic fdlibm kernel polynomials, evaluated without fused
 * multiply-adds.  Error < 1 ulp; the product's device solver evaluates the
 * same expression order and is bit-identical given identical F values. */
static void nu_sincos(double x, double *sn, double *cs) {
    static const double invpio2 = 6.36619772367581382433e-01;
    static const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    static const double pio2_1t = 6.07710050650619224932e-11; /* pi/2 - pio2_1 */
    static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                        S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                        S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
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

/* ---------------------------------------------------------------------- */
/* HD = (second derivative matrix of Q) * D                                */
static void hess_mul(int n, int npt, const double *xpt, const double *hq,
                     const double *pq, const double *d, double *hd) {
    /* all vectors 1-based */
    for (int i = 1; i <= n; ++i) hd[i] = 0.0;
    for (int k = 1; k <= npt; ++k) {
        double temp = 0.0;
        for (int j = 1; j <= n; ++j) temp += XPT(k, j) * d[j];
        temp *= pq[k];
        for (int i = 1; i <= n; ++i) hd[i] += temp * XPT(k, i);
    }
    int ih = 0;
    for (int j = 1; j <= n; ++j) {
        for (int i = 1; i <= j; ++i) {
            ++ih;
            if (i < j) hd[j] += hq[ih] * d[i];
            hd[i] += hq[ih] * d[j];
        }
    }
}

/* ---------------------------------------------------------------------- */
/* TRSAPP: approximate solution of the trust region subproblem by truncated
 * conjugate gradients followed by two-dimensional searches on the boundary */
static void trsapp(int n, int npt, const double *xopt, const double *xpt,
                   const double *gq, const double *hq, const double *pq,
                   double delta, double *step, double *d, double *g,
                   double *hd, double *hs, double *crvmin) {
    const double half = 0.5, zero = 0.0;
    const double twopi = 6.283185307179586476925;
    double delsq = delta * delta;
    int iterc = 0, itermax = n, itersw = itermax;
    double qred = 0, dd = 0, ds = 0, ss = 0, gg = 0, ggbeg = 0, bstep = 0;
    double dhd, alpha, temp, qadd, ggsav, sg = 0, shs = 0, sgk, angtest;
    double tempa = 0, tempb = 0, dg, dhs, cf, qbeg, qsav, qmin, qnew, angle;
    double cth, sth, reduc, ratio;
    int isave, iu;

    for (int i = 1; i <= n; ++i) d[i] = xopt[i];
    hess_mul(n, npt, xpt, hq, pq, d, hd);

    /* Prepare for the first line search. */
    qred = zero;
    dd = zero;
    for (int i = 1; i <= n; ++i) {
        step[i] = zero;
        hs[i] = zero;
        g[i] = gq[i] + hd[i];
        d[i] = -g[i];
        dd += d[i] * d[i];
    }
    *crvmin = zero;
    if (dd == zero) return;
    ds = zero;
    ss = zero;
    gg = dd;
    ggbeg = gg;

    /* Conjugate gradient iterations inside the trust region. */
    for (;;) {
        ++iterc;
        temp = delsq - ss;
        bstep = temp / (ds + sqrt(ds * ds + dd * temp));
        hess_mul(n, npt, xpt, hq, pq, d, hd);
        dhd = zero;
        for (int j = 1; j <= n; ++j) dhd += d[j] * hd[j];

        /* Update CRVMIN and set the step-length ALPHA. */
        alpha = bstep;
        if (dhd > zero) {
            temp = dhd / dd;
            if (iterc == 1) *crvmin = temp;
            *crvmin = dmin(*crvmin, temp);
            alpha = dmin(alpha, gg / dhd);
        }
        qadd = alpha * (gg - half * alpha * dhd);
        qred += qadd;

        /* Update STEP and HS. */
        ggsav = gg;
        gg = zero;
        for (int i = 1; i <= n; ++i) {
            step[i] += alpha * d[i];
            hs[i] += alpha * hd[i];
            double t = g[i] + hs[i];
            gg += t * t;
        }

        /* Begin another conjugate direction iteration if required. */
        if (alpha < bstep) {
            if (qadd <= 0.01 * qred) return;
            if (gg <= 1.0e-4 * ggbeg) return;
            if (iterc == itermax) return;
            temp = gg / ggsav;
            dd = zero;
            ds = zero;
            ss = zero;
            for (int i = 1; i <= n; ++i) {
                d[i] = temp * d[i] - g[i] - hs[i];
                dd += d[i] * d[i];
                ds += d[i] * step[i];
                ss += step[i] * step[i];
            }
            if (ds <= zero) return;
            if (ss < delsq) continue;
        }
        break;
    }
    *crvmin = zero;
    itersw = iterc;
    (void)itersw;

    /* Alternative iterations: searches round the trust region boundary. */
    for (;;) {
        if (gg <= 1.0e-4 * ggbeg) return;
        sg = zero;
        shs = zero;
        for (int i = 1; i <= n; ++i) {
            sg += step[i] * g[i];
            shs += step[i] * hs[i];
        }
        sgk = sg + shs;
        angtest = sgk / sqrt(gg * delsq);
        if (angtest <= -0.99) return;

        /* New direction D in span{STEP, gradient}, orthogonal to STEP. */
        ++iterc;
        temp = sqrt(delsq * gg - sgk * sgk);
        tempa = delsq / temp;
        tempb = sgk / temp;
        for (int i = 1; i <= n; ++i)
            d[i] = tempa * (g[i] + hs[i]) - tempb * step[i];
        hess_mul(n, npt, xpt, hq, pq, d, hd);
        dg = zero;
        dhd = zero;
        dhs = zero;
        for (int i = 1; i <= n; ++i) {
            dg += d[i] * g[i];
            dhd += hd[i] * d[i];
            dhs += hd[i] * step[i];
        }

        /* Seek the value of the angle that minimizes Q. */
        cf = half * (shs - dhd);
        qbeg = sg + cf;
        qsav = qbeg;
        qmin = qbeg;
        isave = 0;
        iu = 49;
        temp = twopi / (double)(iu + 1);
        for (int i = 1; i <= iu; ++i) {
            angle = (double)i * temp;
            nu_sincos(angle, &sth, &cth);
            qnew = (sg + cf * cth) * cth + (dg + dhs * cth) * sth;
            if (qnew < qmin) {
                qmin = qnew;
                isave = i;
                tempa = qsav;
            } else if (i == isave + 1) {
                tempb = qnew;
            }
            qsav = qnew;
        }
        if (isave == 0) tempa = qnew;
        if (isave == iu) tempb = qbeg;
        angle = zero;
        if (tempa != tempb) {
            tempa -= qmin;
            tempb -= qmin;
            angle = half * (tempa - tempb) / (tempa + tempb);
        }
        angle = temp * ((double)isave + angle);

        /* Calculate the new STEP and HS. Then test for convergence. */
        nu_sincos(angle, &sth, &cth);
        reduc = qbeg - (sg + cf * cth) * cth - (dg + dhs * cth) * sth;
        gg = zero;
        for (int i = 1; i <= n; ++i) {
            step[i] = cth * step[i] + sth * d[i];
            hs[i] = cth * hs[i] + sth * hd[i];
            double t = g[i] + hs[i];
            gg += t * t;
        }
        qred += reduc;
        ratio = reduc / qred;
        if (iterc < itermax && ratio > 0.01) continue;
        return;
    }
}

/* ---------------------------------------------------------------------- */
/* BIGLAG: step D of length DELTA from XOPT that makes the modulus of the
 * KNEW-th Lagrange function large                                          */
static void biglag(int n, int npt, const double *xopt, const double *xpt,
                   const double *bmat, const double *zmat, int idz, int ndim,
                   int knew, double delta, double *d, double *alpha_out,
                   double *hcol, double *gc, double *gd, double *s, double *w) {
    const double half = 0.5, one = 1.0, zero = 0.0;
    const double twopi = 6.283185307179586476925;
    double delsq = delta * delta;
    int nptm = npt - n - 1;
    int iterc = 0;
    double temp, sum, dd, gg, sp, dhd, scale, tau, ss, denom;
    double cf1, cf2, cf3, cf4, cf5, taubeg, taumax, tauold, angle, cth, sth;
    double tempa = 0, tempb = 0, step;
    int isave, iu;

    /* Leading elements of the KNEW-th column of H. */
    for (int k = 1; k <= npt; ++k) hcol[k] = zero;
    for (int j = 1; j <= nptm; ++j) {
        temp = ZMAT(knew, j);
        if (j < idz) temp = -temp;
        for (int k = 1; k <= npt; ++k) hcol[k] += temp * ZMAT(k, j);
    }
    *alpha_out = hcol[knew];

    /* Unscaled initial direction D; gradient of the Lagrange function at
     * XOPT (GC) and D times its second derivative matrix (GD). */
    dd = zero;
    for (int i = 1; i <= n; ++i) {
        d[i] = XPT(knew, i) - xopt[i];
        gc[i] = BMAT(knew, i);
        gd[i] = zero;
        dd += d[i] * d[i];
    }
    for (int k = 1; k <= npt; ++k) {
        temp = zero;
        sum = zero;
        for (int j = 1; j <= n; ++j) {
            temp += XPT(k, j) * xopt[j];
            sum += XPT(k, j) * d[j];
        }
        temp = hcol[k] * temp;
        sum = hcol[k] * sum;
        for (int i = 1; i <= n; ++i) {
            gc[i] += temp * XPT(k, i);
            gd[i] += sum * XPT(k, i);
        }
    }

    /* Scale D and GD, with a sign change if required. Set S to another
     * vector in the initial two dimensional subspace. */
    gg = zero;
    sp = zero;
    dhd = zero;
    for (int i = 1; i <= n; ++i) {
        gg += gc[i] * gc[i];
        sp += d[i] * gc[i];
        dhd += d[i] * gd[i];
    }
    scale = delta / sqrt(dd);
    if (sp * dhd < zero) scale = -scale;
    temp = zero;
    if (sp * sp > 0.99 * dd * gg) temp = one;
    tau = scale * (fabs(sp) + half * scale * fabs(dhd));
    if (gg * delsq < 0.01 * tau * tau) temp = one;
    for (int i = 1; i <= n; ++i) {
        d[i] = scale * d[i];
        gd[i] = scale * gd[i];
        s[i] = gc[i] + temp * gd[i];
    }

    /* Iterations: rotate D in span{D,S} to maximise |Lagrange function|. */
    for (;;) {
        ++iterc;
        dd = zero;
        sp = zero;
        ss = zero;
        for (int i = 1; i <= n; ++i) {
            dd += d[i] * d[i];
            sp += d[i] * s[i];
            ss += s[i] * s[i];
        }
        temp = dd * ss - sp * sp;
        if (temp <= 1.0e-8 * dd * ss) return;
        denom = sqrt(temp);
        for (int i = 1; i <= n; ++i) {
            s[i] = (dd * s[i] - sp * d[i]) / denom;
            w[i] = zero;
        }

        /* Coefficients of the Lagrange function on the circle. */
        for (int k = 1; k <= npt; ++k) {
            sum = zero;
            for (int j = 1; j <= n; ++j) sum += XPT(k, j) * s[j];
            sum = hcol[k] * sum;
            for (int i = 1; i <= n; ++i) w[i] += sum * XPT(k, i);
        }
        cf1 = cf2 = cf3 = cf4 = cf5 = zero;
        for (int i = 1; i <= n; ++i) {
            cf1 += s[i] * w[i];
            cf2 += d[i] * gc[i];
            cf3 += s[i] * gc[i];
            cf4 += d[i] * gd[i];
            cf5 += s[i] * gd[i];
        }
        cf1 = half * cf1;
        cf4 = half * cf4 - cf1;

        /* Seek the value of the angle that maximizes the modulus of TAU. */
        taubeg = cf1 + cf2 + cf4;
        taumax = taubeg;
        tauold = taubeg;
        isave = 0;
        iu = 49;
        temp = twopi / (double)(iu + 1);
        for (int i = 1; i <= iu; ++i) {
            angle = (double)i * temp;
            nu_sincos(angle, &sth, &cth);
            tau = cf1 + (cf2 + cf4 * cth) * cth + (cf3 + cf5 * cth) * sth;
            if (fabs(tau) > fabs(taumax)) {
                taumax = tau;
                isave = i;
                tempa = tauold;
            } else if (i == isave + 1) {
                tempb = tau;
            }
            tauold = tau;
        }
        if (isave == 0) tempa = tau;
        if (isave == iu) tempb = taubeg;
        step = zero;
        if (tempa != tempb) {
            tempa -= taumax;
            tempb -= taumax;
            step = half * (tempa - tempb) / (tempa + tempb);
        }
        angle = temp * ((double)isave + step);

        /* Calculate the new D and GD. Then test for convergence. */
        nu_sincos(angle, &sth, &cth);
        tau = cf1 + (cf2 + cf4 * cth) * cth + (cf3 + cf5 * cth) * sth;
        for (int i = 1; i <= n; ++i) {
            d[i] = cth * d[i] + sth * s[i];
            gd[i] = cth * gd[i] + sth * w[i];
            s[i] = gc[i] + gd[i];
        }
        if (fabs(tau) <= 1.1 * fabs(taubeg)) return;
        if (iterc >= n) return;
    }
}

/* ---------------------------------------------------------------------- */
/* BIGDEN: alternative model step that makes |DENOM| = |alpha*beta+tau^2|
 * large when BIGLAG's step suffers cancellation.  On return D, VLAG, BETA
 * and W(1..ndim) (= w_check) correspond to the chosen step.                */
static void bigden(int n, int npt, const double *xopt, const double *xpt,
                   const double *bmat, const double *zmat, int idz, int ndim,
                   int kopt, int knew, double *d, double *w, double *vlag,
                   double *beta, double *s, double *wvec, double *prod) {
    const double half = 0.5, one = 1.0, quart = 0.25, two = 2.0, zero = 0.0;
    const double twopi = 6.283185307179586476925;
    int nptm = npt - n - 1;
    double den[10], denex[10], par[10];
    double temp, alpha, dd, ds, ss, xoptsq, dtest, dstemp, sstemp, diff;
    double ssden, densav, xoptd, xopts, tempa = 0, tempb = 0, tempc, sum;
    double denold, denmax, sumold, angle, step, tau;
    int ksav, iterc, isave, iu, nw;

    /* W(n+1..n+npt) <- leading elements of the KNEW-th column of H. */
    for (int k = 1; k <= npt; ++k) w[n + k] = zero;
    for (int j = 1; j <= nptm; ++j) {
        temp = ZMAT(knew, j);
        if (j < idz) temp = -temp;
        for (int k = 1; k <= npt; ++k) w[n + k] += temp * ZMAT(k, j);
    }
    alpha = w[n + knew];

    /* Initial S: direction from XOPT to X_KNEW unless nearly parallel to D. */
    dd = ds = ss = xoptsq = zero;
    for (int i = 1; i <= n; ++i) {
        dd += d[i] * d[i];
        s[i] = XPT(knew, i) - xopt[i];
        ds += d[i] * s[i];
        ss += s[i] * s[i];
        xoptsq += xopt[i] * xopt[i];
    }
    if (ds * ds > 0.99 * dd * ss) {
        ksav = knew;
        dtest = ds * ds / ss;
        for (int k = 1; k <= npt; ++k) {
            if (k != kopt) {
                dstemp = zero;
                sstemp = zero;
                for (int i = 1; i <= n; ++i) {
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
        for (int i = 1; i <= n; ++i) s[i] = XPT(ksav, i) - xopt[i];
    }
    ssden = dd * ss - ds * ds;
    iterc = 0;
    densav = zero;

    for (;;) {
        /* Overwrite S with a vector of the required length and direction. */
        ++iterc;
        temp = one / sqrt(ssden);
        xoptd = zero;
        xopts = zero;
        for (int i = 1; i <= n; ++i) {
            s[i] = temp * (dd * s[i] - ds * d[i]);
            xoptd += xopt[i] * d[i];
            xopts += xopt[i] * s[i];
        }

        /* Coefficients of the first two terms of BETA. */
        tempa = half * xoptd * xoptd;
        tempb = half * xopts * xopts;
        den[1] = dd * (xoptsq + half * dd) + tempa + tempb;
        den[2] = two * xoptd * dd;
        den[3] = two * xopts * dd;
        den[4] = tempa - tempb;
        den[5] = xoptd * xopts;
        for (int i = 6; i <= 9; ++i) den[i] = zero;

        /* Coefficients of w_check in WVEC. */
        for (int k = 1; k <= npt; ++k) {
            tempa = tempb = tempc = zero;
            for (int i = 1; i <= n; ++i) {
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
        for (int i = 1; i <= n; ++i) {
            int ip = i + npt;
            WVEC(ip, 1) = zero;
            WVEC(ip, 2) = d[i];
            WVEC(ip, 3) = s[i];
            WVEC(ip, 4) = zero;
            WVEC(ip, 5) = zero;
        }

        /* Coefficients of H * w_check in PROD. */
        for (int jc = 1; jc <= 5; ++jc) {
            nw = npt;
            if (jc == 2 || jc == 3) nw = ndim;
            for (int k = 1; k <= npt; ++k) PROD(k, jc) = zero;
            for (int j = 1; j <= nptm; ++j) {
                sum = zero;
                for (int k = 1; k <= npt; ++k) sum += ZMAT(k, j) * WVEC(k, jc);
                if (j < idz) sum = -sum;
                for (int k = 1; k <= npt; ++k) PROD(k, jc) += sum * ZMAT(k, j);
            }
            if (nw == ndim) {
                for (int k = 1; k <= npt; ++k) {
                    sum = zero;
                    for (int j = 1; j <= n; ++j)
                        sum += BMAT(k, j) * WVEC(npt + j, jc);
                    PROD(k, jc) += sum;
                }
            }
            for (int j = 1; j <= n; ++j) {
                sum = zero;
                for (int i = 1; i <= nw; ++i) sum += BMAT(i, j) * WVEC(i, jc);
                PROD(npt + j, jc) = sum;
            }
        }

        /* Include in DEN the part of BETA that depends on THETA. */
        for (int k = 1; k <= ndim; ++k) {
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

        /* Extend DEN so that it holds all the coefficients of DENOM. */
        sum = zero;
        for (int i = 1; i <= 5; ++i) {
            par[i] = half * PROD(knew, i) * PROD(knew, i);
            sum += par[i];
        }
        denex[1] = alpha * den[1] + par[1] + sum;
        tempa = two * PROD(knew, 1) * PROD(knew, 2);
        tempb = PROD(knew, 2) * PROD(knew, 4);
        tempc = PROD(knew, 3) * PROD(knew, 5);
        denex[2] = alpha * den[2] + tempa + tempb + tempc;
        denex[6] = alpha * den[6] + tempb - tempc;
        tempa = two * PROD(knew, 1) * PROD(knew, 3);
        tempb = PROD(knew, 2) * PROD(knew, 5);
        tempc = PROD(knew, 3) * PROD(knew, 4);
        denex[3] = alpha * den[3] + tempa + tempb - tempc;
        denex[7] = alpha * den[7] + tempb + tempc;
        tempa = two * PROD(knew, 1) * PROD(knew, 4);
        denex[4] = alpha * den[4] + tempa + par[2] - par[3];
        tempa = two * PROD(knew, 1) * PROD(knew, 5);
        denex[5] = alpha * den[5] + tempa + PROD(knew, 2) * PROD(knew, 3);
        denex[8] = alpha * den[8] + par[4] - par[5];
        denex[9] = alpha * den[9] + PROD(knew, 4) * PROD(knew, 5);

        /* Seek the value of the angle that maximizes the modulus of DENOM. */
        sum = denex[1] + denex[2] + denex[4] + denex[6] + denex[8];
        denold = sum;
        denmax = sum;
        isave = 0;
        iu = 49;
        temp = twopi / (double)(iu + 1);
        par[1] = one;
        for (int i = 1; i <= iu; ++i) {
            angle = (double)i * temp;
            nu_sincos(angle, &par[3], &par[2]);
            for (int j = 4; j <= 8; j += 2) {
                par[j] = par[2] * par[j - 2] - par[3] * par[j - 1];
                par[j + 1] = par[2] * par[j - 1] + par[3] * par[j - 2];
            }
            sumold = sum;
            sum = zero;
            for (int j = 1; j <= 9; ++j) sum += denex[j] * par[j];
            if (fabs(sum) > fabs(denmax)) {
                denmax = sum;
                isave = i;
                tempa = sumold;
            } else if (i == isave + 1) {
                tempb = sum;
            }
        }
        if (isave == 0) tempa = sum;
        if (isave == iu) tempb = denold;
        step = zero;
        if (tempa != tempb) {
            tempa -= denmax;
            tempb -= denmax;
            step = half * (tempa - tempb) / (tempa + tempb);
        }
        angle = temp * ((double)isave + step);

        /* New parameters of the denominator, new VLAG and new D. */
        nu_sincos(angle, &par[3], &par[2]);
        for (int j = 4; j <= 8; j += 2) {
            par[j] = par[2] * par[j - 2] - par[3] * par[j - 1];
            par[j + 1] = par[2] * par[j - 1] + par[3] * par[j - 2];
        }
        *beta = zero;
        denmax = zero;
        for (int j = 1; j <= 9; ++j) {
            *beta += den[j] * par[j];
            denmax += denex[j] * par[j];
        }
        for (int k = 1; k <= ndim; ++k) {
            vlag[k] = zero;
            for (int j = 1; j <= 5; ++j) vlag[k] += PROD(k, j) * par[j];
        }
        tau = vlag[knew];
        dd = zero;
        tempa = zero;
        tempb = zero;
        for (int i = 1; i <= n; ++i) {
            d[i] = par[2] * d[i] + par[3] * s[i];
            w[i] = xopt[i] + d[i];
            dd += d[i] * d[i];
            tempa += d[i] * w[i];
            tempb += w[i] * w[i];
        }
        if (iterc >= n) break;
        if (iterc > 1) densav = dmax(densav, denold);
        if (fabs(denmax) <= 1.1 * fabs(densav)) break;
        densav = denmax;

        /* S <- half the gradient of the denominator with respect to D. */
        for (int i = 1; i <= n; ++i) {
            temp = tempa * xopt[i] + tempb * d[i] - vlag[npt + i];
            s[i] = tau * BMAT(knew, i) + alpha * temp;
        }
        for (int k = 1; k <= npt; ++k) {
            sum = zero;
            for (int j = 1; j <= n; ++j) sum += XPT(k, j) * w[j];
            temp = (tau * w[n + k] - alpha * vlag[k]) * sum;
            for (int i = 1; i <= n; ++i) s[i] += temp * XPT(k, i);
        }
        ss = zero;
        ds = zero;
        for (int i = 1; i <= n; ++i) {
            ss += s[i] * s[i];
            ds += d[i] * s[i];
        }
        ssden = dd * ss - ds * ds;
        if (ssden >= 1.0e-8 * dd * ss) continue;
        break;
    }

    /* Set the vector W before the return. */
    for (int k = 1; k <= ndim; ++k) {
        w[k] = zero;
        for (int j = 1; j <= 5; ++j) w[k] += WVEC(k, j) * par[j];
    }
    vlag[kopt] += one;
}

/* ---------------------------------------------------------------------- */
/* UPDATE: revise BMAT, ZMAT, IDZ when the KNEW-th point moves             */
static void update(int n, int npt, double *bmat, double *zmat, int *idz,
                   int ndim, double *vlag, double beta, int knew, double *w) {
    const double one = 1.0, zero = 0.0;
    int nptm = npt - n - 1;
    int jl = 1, iflag, ja, jb;
    double temp, tempa, tempb = 0, alpha, tau, tausq, denom, scala, scalb;

    /* Rotations that put zeros in the KNEW-th row of ZMAT. */
    for (int j = 2; j <= nptm; ++j) {
        if (j == *idz) {
            jl = *idz;
        } else if (ZMAT(knew, j) != zero) {
            temp = sqrt(ZMAT(knew, jl) * ZMAT(knew, jl) +
                        ZMAT(knew, j) * ZMAT(knew, j));
            tempa = ZMAT(knew, jl) / temp;
            tempb = ZMAT(knew, j) / temp;
            for (int i = 1; i <= npt; ++i) {
                temp = tempa * ZMAT(i, jl) + tempb * ZMAT(i, j);
                ZMAT(i, j) = tempa * ZMAT(i, j) - tempb * ZMAT(i, jl);
                ZMAT(i, jl) = temp;
            }
            ZMAT(knew, j) = zero;
        }
    }

    /* First NPT components of the KNEW-th column of HLAG into W, and the
     * parameters of the updating formula. */
    tempa = ZMAT(knew, 1);
    if (*idz >= 2) tempa = -tempa;
    if (jl > 1) tempb = ZMAT(knew, jl);
    for (int i = 1; i <= npt; ++i) {
        w[i] = tempa * ZMAT(i, 1);
        if (jl > 1) w[i] += tempb * ZMAT(i, jl);
    }
    alpha = w[knew];
    tau = vlag[knew];
    tausq = tau * tau;
    denom = alpha * beta + tausq;
    vlag[knew] -= one;

    /* Complete the updating of ZMAT when there is only one nonzero element
     * in the KNEW-th row of the new matrix ZMAT. */
    iflag = 0;
    if (jl == 1) {
        temp = sqrt(fabs(denom));
        tempb = tempa / temp;
        tempa = tau / temp;
        for (int i = 1; i <= npt; ++i)
            ZMAT(i, 1) = tempa * ZMAT(i, 1) - tempb * vlag[i];
        /* Powell's code tests TEMP (= sqrt|denom| >= 0) here, not DENOM;
         * the restatement keeps his tests as published. */
        if (*idz == 1 && temp < zero) *idz = 2;
        if (*idz >= 2 && temp >= zero) iflag = 1;
    } else {
        /* The alternative case. */
        ja = 1;
        if (beta >= zero) ja = jl;
        jb = jl + 1 - ja;
        temp = ZMAT(knew, jb) / denom;
        tempa = temp * beta;
        tempb = temp * tau;
        temp = ZMAT(knew, ja);
        scala = one / sqrt(fabs(beta) * temp * temp + tausq);
        scalb = scala * sqrt(fabs(denom));
        for (int i = 1; i <= npt; ++i) {
            ZMAT(i, ja) = scala * (tau * ZMAT(i, ja) - temp * vlag[i]);
            ZMAT(i, jb) = scalb * (ZMAT(i, jb) - tempa * w[i] - tempb * vlag[i]);
        }
        if (denom <= zero) {
            if (beta < zero) *idz = *idz + 1;
            if (beta >= zero) iflag = 1;
        }
    }

    /* IDZ is reduced in the following case, and usually the first column
     * of ZMAT is exchanged with a later one. */
    if (iflag == 1) {
        *idz = *idz - 1;
        for (int i = 1; i <= npt; ++i) {
            temp = ZMAT(i, 1);
            ZMAT(i, 1) = ZMAT(i, *idz);
            ZMAT(i, *idz) = temp;
        }
    }

    /* Finally, update the matrix BMAT. */
    for (int j = 1; j <= n; ++j) {
        int jp = npt + j;
        w[jp] = BMAT(knew, j);
        tempa = (alpha * vlag[jp] - tau * w[jp]) / denom;
        tempb = (-beta * w[jp] - tau * vlag[jp]) / denom;
        for (int i = 1; i <= jp; ++i) {
            BMAT(i, j) = BMAT(i, j) + tempa * vlag[i] + tempb * w[i];
            if (i > npt) BMAT(jp, i - npt) = BMAT(i, j);
        }
    }
}

/* ---------------------------------------------------------------------- */
int newuoa_oracle(int n, int npt, newuoa_objfun calfun, void *data, double *xio,
                  double rhobeg, double rhoend, int maxfun, double *fout,
                  int *nfout, newuoa_observer obs, void *obsdata) {
    const double half = 0.5, one = 1.0, tenth = 0.1, zero = 0.0;
    int np = n + 1, nh = (n * np) / 2, nptm = npt - np, ndim = npt + n;
    int status = NEWUOA_SUCCESS;
    if (npt < n + 2 || npt > ((n + 2) * np) / 2) return NEWUOA_BAD_NPT;

    /* one zeroed block; every vector is 1-based (slot 0 unused) */
    size_t nvec = 0;
#define TAKE(len) (nvec += (size_t)(len) + 1, nvec - ((size_t)(len) + 1))
    size_t o_x = TAKE(n), o_xbase = TAKE(n), o_xopt = TAKE(n), o_xnew = TAKE(n);
    size_t o_fval = TAKE(npt), o_gq = TAKE(n), o_hq = TAKE(nh), o_pq = TAKE(npt);
    size_t o_d = TAKE(n), o_vlag = TAKE(ndim), o_w = TAKE(2 * ndim + 2 * npt);
    size_t o_t1 = TAKE(n), o_t2 = TAKE(n), o_t3 = TAKE(n), o_t4 = TAKE(n);
    size_t o_s = TAKE(n);
    size_t o_xpt = TAKE((size_t)npt * n), o_bmat = TAKE((size_t)ndim * n);
    size_t o_zmat = TAKE((size_t)npt * nptm), o_wvec = TAKE((size_t)ndim * 5);
    size_t o_prod = TAKE((size_t)ndim * 5);
#undef TAKE
    double *mem = (double *)calloc(nvec, sizeof(double));
    if (!mem) return NEWUOA_NO_MEMORY;
    double *x = mem + o_x, *xbase = mem + o_xbase, *xopt = mem + o_xopt;
    double *xnew = mem + o_xnew, *fval = mem + o_fval, *gq = mem + o_gq;
    double *hq = mem + o_hq, *pq = mem + o_pq, *d = mem + o_d;
    double *vlag = mem + o_vlag, *w = mem + o_w;
    double *t1 = mem + o_t1, *t2 = mem + o_t2, *t3 = mem + o_t3, *t4 = mem + o_t4;
    double *svec = mem + o_s;
    double *xpt = mem + o_xpt, *bmat = mem + o_bmat, *zmat = mem + o_zmat;
    double *wvec = mem + o_wvec, *prod = mem + o_prod;

    int nftest = maxfun > 1 ? maxfun : 1;
    double rhosq, recip, reciq, f = 0, fbeg = 0, fopt = 0, xipt = 0, xjpt = 0;
    double rho = 0, delta = 0, diffa = 0, diffb = 0, diffc = 0, xoptsq = 0, dsq = 0, dnorm = 0;
    double ratio = 0, crvmin = 0, temp, tempq, sum, sumz, suma, sumb, bsum, dx;
    double beta = 0, alpha = 0, dstep = 0, vquad = 0, diff = 0, fsave = 0;
    double detrat, hdiag, distsq, gqsq, gisq;
    int nf, nfm, nfmm, kopt = 1, idz = 1, itest = 0, nfsav = 0, knew = 0, ksave = 0, ktemp;
    int ipt = 0, jpt = 0, itemp, ih, ip;

    for (int j = 1; j <= n; ++j) {
        x[j] = xio[j - 1];
        xbase[j] = x[j];
    }

    /* ---- initial interpolation set ---- */
    rhosq = rhobeg * rhobeg;
    recip = one / rhosq;
    reciq = sqrt(half) / rhosq;
    nf = 0;
L50:
    nfm = nf;
    nfmm = nf - n;
    ++nf;
    if (nfm <= 2 * n) {
        if (nfm >= 1 && nfm <= n) {
            XPT(nf, nfm) = rhobeg;
        } else if (nfm > n) {
            XPT(nf, nfmm) = -rhobeg;
        }
    } else {
        itemp = (nfmm - 1) / n;
        jpt = nfm - itemp * n - n;
        ipt = jpt + itemp;
        if (ipt > n) {
            itemp = jpt;
            jpt = ipt - n;
            ipt = itemp;
        }
        xipt = rhobeg;
        if (fval[ipt + np] < fval[ipt + 1]) xipt = -xipt;
        xjpt = rhobeg;
        if (fval[jpt + np] < fval[jpt + 1]) xjpt = -xjpt;
        XPT(nf, ipt) = xipt;
        XPT(nf, jpt) = xjpt;
    }
    for (int j = 1; j <= n; ++j) x[j] = XPT(nf, j) + xbase[j];
    goto L310;
L70:
    fval[nf] = f;
    if (nf == 1) {
        fbeg = f;
        fopt = f;
        kopt = 1;
    } else if (f < fopt) {
        fopt = f;
        kopt = nf;
    }
    /* nonzero initial elements of BMAT and of the quadratic model */
    if (nfm <= 2 * n) {
        if (nfm >= 1 && nfm <= n) {
            gq[nfm] = (f - fbeg) / rhobeg;
            if (npt < nf + n) {
                BMAT(1, nfm) = -one / rhobeg;
                BMAT(nf, nfm) = one / rhobeg;
                BMAT(npt + nfm, nfm) = -half * rhosq;
            }
        } else if (nfm > n) {
            BMAT(nf - n, nfmm) = half / rhobeg;
            BMAT(nf, nfmm) = -half / rhobeg;
            ZMAT(1, nfmm) = -reciq - reciq;
            ZMAT(nf - n, nfmm) = reciq;
            ZMAT(nf, nfmm) = reciq;
            ih = (nfmm * (nfmm + 1)) / 2;
            temp = (fbeg - f) / rhobeg;
            hq[ih] = (gq[nfmm] - temp) / rhobeg;
            gq[nfmm] = half * (gq[nfmm] + temp);
        }
    } else {
        /* off-diagonal second derivatives */
        ih = (ipt * (ipt - 1)) / 2 + jpt;
        if (xipt < zero) ipt += n;
        if (xjpt < zero) jpt += n;
        ZMAT(1, nfmm) = recip;
        ZMAT(nf, nfmm) = recip;
        ZMAT(ipt + 1, nfmm) = -recip;
        ZMAT(jpt + 1, nfmm) = -recip;
        hq[ih] = (fbeg - fval[ipt + 1] - fval[jpt + 1] + f) / (xipt * xjpt);
    }
    if (nf < npt) goto L50;

    /* ---- iterative procedure ---- */
    rho = rhobeg;
    delta = rho;
    idz = 1;
    diffa = zero;
    diffb = zero;
    itest = 0;
    xoptsq = zero;
    for (int i = 1; i <= n; ++i) {
        xopt[i] = XPT(kopt, i);
        xoptsq += xopt[i] * xopt[i];
    }
L90:
    nfsav = nf;

    /* next trust region step */
L100:
    knew = 0;
    ++g_counters[0];
    trsapp(n, npt, xopt, xpt, gq, hq, pq, delta, d, t1, t2, t3, t4, &crvmin);
    dsq = zero;
    for (int i = 1; i <= n; ++i) dsq += d[i] * d[i];
    dnorm = dmin(delta, sqrt(dsq));
    if (dnorm < half * rho) {
        knew = -1;
        delta = tenth * delta;
        ratio = -1.0;
        if (delta <= 1.5 * rho) delta = rho;
        if (nf <= nfsav + 2) goto L460;
        temp = 0.125 * crvmin * rho * rho;
        if (temp <= dmax(diffa, dmax(diffb, diffc))) goto L460;
        goto L490;
    }

    /* shift XBASE if XOPT may be too far from XBASE */
L120:
    if (dsq <= 1.0e-3 * xoptsq) {
        tempq = 0.25 * xoptsq;
        ++g_counters[4];
        for (int k = 1; k <= npt; ++k) {
            sum = zero;
            for (int i = 1; i <= n; ++i) sum += XPT(k, i) * xopt[i];
            temp = pq[k] * sum;
            sum -= half * xoptsq;
            w[npt + k] = sum;
            for (int i = 1; i <= n; ++i) {
                gq[i] += temp * XPT(k, i);
                XPT(k, i) -= half * xopt[i];
                vlag[i] = BMAT(k, i);
                w[i] = sum * XPT(k, i) + tempq * xopt[i];
                ip = npt + i;
                for (int j = 1; j <= i; ++j)
                    BMAT(ip, j) = BMAT(ip, j) + vlag[i] * w[j] + w[i] * vlag[j];
            }
        }
        /* revisions of BMAT that depend on ZMAT */
        for (int k = 1; k <= nptm; ++k) {
            sumz = zero;
            for (int i = 1; i <= npt; ++i) {
                sumz += ZMAT(i, k);
                w[i] = w[npt + i] * ZMAT(i, k);
            }
            for (int j = 1; j <= n; ++j) {
                sum = tempq * sumz * xopt[j];
                for (int i = 1; i <= npt; ++i) sum += w[i] * XPT(i, j);
                vlag[j] = sum;
                if (k < idz) sum = -sum;
                for (int i = 1; i <= npt; ++i)
                    BMAT(i, j) = BMAT(i, j) + sum * ZMAT(i, k);
            }
            for (int i = 1; i <= n; ++i) {
                ip = i + npt;
                temp = vlag[i];
                if (k < idz) temp = -temp;
                for (int j = 1; j <= i; ++j)
                    BMAT(ip, j) = BMAT(ip, j) + temp * vlag[j];
            }
        }
        /* complete the shift, including the model parameters */
        ih = 0;
        for (int j = 1; j <= n; ++j) {
            w[j] = zero;
            for (int k = 1; k <= npt; ++k) {
                w[j] += pq[k] * XPT(k, j);
                XPT(k, j) -= half * xopt[j];
            }
            for (int i = 1; i <= j; ++i) {
                ++ih;
                if (i < j) gq[j] += hq[ih] * xopt[i];
                gq[i] += hq[ih] * xopt[j];
                hq[ih] = hq[ih] + w[i] * xopt[j] + xopt[i] * w[j];
                BMAT(npt + i, j) = BMAT(npt + j, i);
            }
        }
        for (int j = 1; j <= n; ++j) {
            xbase[j] += xopt[j];
            xopt[j] = zero;
        }
        xoptsq = zero;
    }

    /* model step when KNEW is positive */
    if (knew > 0) {
        /* HCOL = vlag(1..npt), GC = vlag(npt+1..), GD, S, W work vectors */
        ++g_counters[1];
        biglag(n, npt, xopt, xpt, bmat, zmat, idz, ndim, knew, dstep, d, &alpha,
               vlag, vlag + npt, t1, t2, t3);
    }

    /* VLAG and BETA for the current D; first NPT components of w_check in W */
    for (int k = 1; k <= npt; ++k) {
        suma = zero;
        sumb = zero;
        sum = zero;
        for (int j = 1; j <= n; ++j) {
            suma += XPT(k, j) * d[j];
            sumb += XPT(k, j) * xopt[j];
            sum += BMAT(k, j) * d[j];
        }
        w[k] = suma * (half * suma + sumb);
        vlag[k] = sum;
    }
    beta = zero;
    for (int k = 1; k <= nptm; ++k) {
        sum = zero;
        for (int i = 1; i <= npt; ++i) sum += ZMAT(i, k) * w[i];
        if (k < idz) {
            beta += sum * sum;
            sum = -sum;
        } else {
            beta -= sum * sum;
        }
        for (int i = 1; i <= npt; ++i) vlag[i] += sum * ZMAT(i, k);
    }
    bsum = zero;
    dx = zero;
    for (int j = 1; j <= n; ++j) {
        sum = zero;
        for (int i = 1; i <= npt; ++i) sum += w[i] * BMAT(i, j);
        bsum += sum * d[j];
        int jp = npt + j;
        for (int k = 1; k <= n; ++k) sum += BMAT(jp, k) * d[k];
        vlag[jp] = sum;
        bsum += sum * d[j];
        dx += d[j] * xopt[j];
    }
    beta = dx * dx + dsq * (xoptsq + dx + dx + half * dsq) + beta - bsum;
    vlag[kopt] += one;

    /* alternative model step if the cancellation in DENOM is unacceptable */
    if (knew > 0) {
        temp = one + alpha * beta / (vlag[knew] * vlag[knew]);
        if (fabs(temp) <= 0.8 || g_force_bigden) {
            ++g_counters[2];
            bigden(n, npt, xopt, xpt, bmat, zmat, idz, ndim, kopt, knew, d, w,
                   vlag, &beta, svec, wvec, prod);
        }
    }

    /* next value of the objective function */
L290:
    for (int i = 1; i <= n; ++i) {
        xnew[i] = xopt[i] + d[i];
        x[i] = xbase[i] + xnew[i];
    }
    ++nf;
L310:
    if (nf > nftest) {
        --nf;
        status = NEWUOA_TOO_MANY_EVALUATIONS;
        goto L530;
    }
    f = calfun(n, x + 1, data);
    if (obs) obs(nf, n, x + 1, f, obsdata);
    if (nf <= npt) goto L70;
    if (knew == -1) goto L530;

    /* predicted change VQUAD and the error DIFF of the prediction */
    vquad = zero;
    ih = 0;
    for (int j = 1; j <= n; ++j) {
        vquad += d[j] * gq[j];
        for (int i = 1; i <= j; ++i) {
            ++ih;
            temp = d[i] * xnew[j] + d[j] * xopt[i];
            if (i == j) temp = half * temp;
            vquad += temp * hq[ih];
        }
    }
    for (int k = 1; k <= npt; ++k) vquad += pq[k] * w[k];
    diff = f - fopt - vquad;
    diffc = diffb;
    diffb = diffa;
    diffa = fabs(diff);
    if (dnorm > rho) nfsav = nf;

    /* update FOPT and XOPT if the new F is the least value so far */
    fsave = fopt;
    if (f < fopt) {
        fopt = f;
        xoptsq = zero;
        for (int i = 1; i <= n; ++i) {
            xopt[i] = xnew[i];
            xoptsq += xopt[i] * xopt[i];
        }
    }
    ksave = knew;
    if (knew > 0) goto L410;

    /* next DELTA after a trust region step */
    if (vquad >= zero) {
        status = NEWUOA_ROUNDING_ERRORS;
        goto L530;
    }
    ratio = (f - fsave) / vquad;
    if (ratio <= tenth) {
        delta = half * dnorm;
    } else if (ratio <= 0.7) {
        delta = dmax(half * delta, dnorm);
    } else {
        delta = dmax(half * delta, dnorm + dnorm);
    }
    if (delta <= 1.5 * rho) delta = rho;

    /* index KNEW of the interpolation point to be deleted */
    rhosq = dmax(tenth * delta, rho);
    rhosq = rhosq * rhosq;
    ktemp = 0;
    detrat = zero;
    if (f >= fsave) {
        ktemp = kopt;
        detrat = one;
    }
    for (int k = 1; k <= npt; ++k) {
        hdiag = zero;
        for (int j = 1; j <= nptm; ++j) {
            temp = one;
            if (j < idz) temp = -one;
            hdiag += temp * ZMAT(k, j) * ZMAT(k, j);
        }
        temp = fabs(beta * hdiag + vlag[k] * vlag[k]);
        distsq = zero;
        for (int j = 1; j <= n; ++j) {
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
    if (knew == 0) goto L460;

    /* update BMAT, ZMAT, IDZ and the quadratic model */
L410:
    ++g_counters[3];
    update(n, npt, bmat, zmat, &idz, ndim, vlag, beta, knew, w);
    fval[knew] = f;
    ih = 0;
    for (int i = 1; i <= n; ++i) {
        temp = pq[knew] * XPT(knew, i);
        for (int j = 1; j <= i; ++j) {
            ++ih;
            hq[ih] += temp * XPT(knew, j);
        }
    }
    pq[knew] = zero;
    for (int j = 1; j <= nptm; ++j) {
        temp = diff * ZMAT(knew, j);
        if (j < idz) temp = -temp;
        for (int k = 1; k <= npt; ++k) pq[k] += temp * ZMAT(k, j);
    }
    gqsq = zero;
    for (int i = 1; i <= n; ++i) {
        gq[i] += diff * BMAT(knew, i);
        gqsq += gq[i] * gq[i];
        XPT(knew, i) = xnew[i];
    }

    /* least Frobenius norm interpolant test after small trust region steps */
    if (ksave == 0 && delta == rho) {
        if (fabs(ratio) > 1.0e-2) {
            itest = 0;
        } else {
            for (int k = 1; k <= npt; ++k) vlag[k] = fval[k] - fval[kopt];
            gisq = zero;
            for (int i = 1; i <= n; ++i) {
                sum = zero;
                for (int k = 1; k <= npt; ++k) sum += BMAT(k, i) * vlag[k];
                gisq += sum * sum;
                w[i] = sum;
            }
            ++itest;
            if (gqsq < 1.0e2 * gisq) itest = 0;
            if (itest >= 3) {
                for (int i = 1; i <= n; ++i) gq[i] = w[i];
                for (ih = 1; ih <= nh; ++ih) hq[ih] = zero;
                for (int j = 1; j <= nptm; ++j) {
                    w[j] = zero;
                    for (int k = 1; k <= npt; ++k) w[j] += vlag[k] * ZMAT(k, j);
                    if (j < idz) w[j] = -w[j];
                }
                for (int k = 1; k <= npt; ++k) {
                    pq[k] = zero;
                    for (int j = 1; j <= nptm; ++j) pq[k] += ZMAT(k, j) * w[j];
                }
                itest = 0;
            }
        }
    }
    if (f < fsave) kopt = knew;

    if (g_probe) {
        newuoa_state_view sv = {n,    npt,      idz,      kopt,     nf,
                                xbase + 1, xopt + 1, xpt, fval + 1, gq + 1,
                                hq + 1,    pq + 1,   bmat, zmat,    rho, delta};
        g_probe(&sv, g_probe_data);
    }

    /* sufficient decrease, or a model step: another trust region step */
    if (f <= fsave + tenth * vquad) goto L100;
    if (ksave > 0) goto L100;

    /* are the interpolation points close enough to the best point? */
    knew = 0;
L460:
    distsq = 4.0 * delta * delta;
    for (int k = 1; k <= npt; ++k) {
        sum = zero;
        for (int j = 1; j <= n; ++j) {
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
        goto L120;
    }
    if (ratio > zero) goto L100;
    if (dmax(delta, dnorm) > rho) goto L100;

    /* next values of RHO and DELTA */
L490:
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
        goto L90;
    }

    /* one last Newton-Raphson step if it was too short to be tried before */
    if (knew == -1) goto L290;
L530:
    if (fopt <= f) {
        for (int i = 1; i <= n; ++i) x[i] = xbase[i] + xopt[i];
        f = fopt;
    }
    for (int i = 1; i <= n; ++i) xio[i - 1] = x[i];
    if (fout) *fout = f;
    if (nfout) *nfout = nf;
    free(mem);
    return status;
}

/* test hook: the solver's portable sin/cos */
void newuoa_oracle_sincos(double x, double *s, double *c) { nu_sincos(x, s, c); }
