/*
 * oracle/gppd_oracle.c -- TEST INFRASTRUCTURE ONLY.  See gppd_oracle.h.
 *
 * Restates, function by function, the reference's demodulateall path.  Every
 * function cites the reference lines it follows.  The floating-point order of
 * the reference is kept where it decides results (separate multiply and add in
 * omega*t+phi, src/Modulation.jl:137; psi = (b*sin+alpha)-alpha, :66-69,:421).
 * Sums are accumulated in index order (the reference's @simd / pairwise / BLAS
 * orders cannot be matched bit for bit and differ at the 1e-16 level).
 */
#include "gppd_oracle.h"
#include "newuoa.h"

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

static const double ORA_OMEGA = 6.283185; /* M_2PI, src/Modulation.jl:11 */
#define ORA_PI 3.14159265358979323846       /* Float64(pi) */

/* src/Modulation.jl:17-22 */
int ora_idx(int side, int telescope, int diode) {
    if (diode == 5) return 32 + side / 4 + (telescope - 1) + 1;
    return side + (diode - 1) + (telescope - 1) * 4 + 1;
}

/* range(-pi, pi, 8), src/Modulation.jl:360: Julia evaluates the range in
 * twice precision, so each element is the correctly rounded value of
 * -pi_f + k*(2*pi_f/7); long double arithmetic reproduces that rounding. */
void ora_phirange(double *phi8) {
    long double lo = -(long double)ORA_PI, hi = (long double)ORA_PI;
    for (int k = 0; k < 8; ++k)
        phi8[k] = (double)(((long double)(7 - k) * lo + (long double)k * hi) / 7.0L);
    phi8[0] = -ORA_PI;
    phi8[7] = ORA_PI;
}

/* ---------------------------------------------------------------------- */
/* src/Faint.jl:21-73, statement by statement (queues as index cursors)    */
int ora_buildstates(long n, const double *t, const double *timer1, long n1,
                    const double *timer2, long n2, long lag, double pre,
                    double post, int8_t *state) {
    if (n < 2 || n1 < 1 || n2 < 1) return -1;
    double timestep = t[1] - t[0];                   /* :24 */
    double shift = (double)lag * timestep;           /* :25-26 */
    long premax = (long)ceil(pre / timestep);        /* :29 */
    long postmax = (long)ceil(post / timestep);      /* :30 */
    double tlast = t[n - 1];
    int current = ORA_NORMAL;                        /* :32 */
    long i1 = 0, i2 = 0;
    double first1 = timer1[i1++] + shift;            /* :33 */
    double first2 = timer2[i2++] + shift;            /* :34 */
    long forget = 0;
    for (long k = 0; k < n; ++k) {                   /* :37 */
        double time = t[k];
        if (time >= first1) {                        /* :40 HIGH */
            current = ORA_HIGH;
            forget = premax;
            if (i1 >= n1) {
                first1 = tlast;
                if (first2 == tlast) current = ORA_NORMAL;
            } else {
                first1 = timer1[i1++] + shift;
            }
        }
        if (time >= first2) {                        /* :54 LOW */
            current = ORA_LOW;
            forget = postmax;
            if (i2 >= n2) {
                first2 = tlast;
                if (first1 == tlast) current = ORA_NORMAL;
            } else {
                first2 = timer2[i2++] + shift;
            }
        }
        if (forget > 0) {                            /* :66 */
            state[k] = ORA_TRANSIENT;
            --forget;
        } else {
            state[k] = (int8_t)current;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------------- */
/* src/Faint.jl:89-100; abs(::Complex) is hypot; var uses the n-1 divisor  */
void ora_mean_var_power(long n, const int8_t *state, const double *data,
                        double *power, double *weight) {
    static const int states[5] = {ORA_OFF, ORA_LOW, ORA_NORMAL, ORA_HIGH, ORA_TRANSIENT};
    for (long i = 0; i < n; ++i) power[i] = weight[i] = 0.0;
    for (int s = 0; s < 5; ++s) {
        long cnt = 0;
        double sum = 0.0;
        for (long i = 0; i < n; ++i)
            if (state[i] == states[s]) {
                sum += hypot(data[2 * i], data[2 * i + 1]);
                ++cnt;
            }
        if (cnt == 0) continue; /* mean(empty)=NaN is assigned to no sample */
        double m = sum / (double)cnt;
        double ss = 0.0;
        for (long i = 0; i < n; ++i)
            if (state[i] == states[s]) {
                double r = hypot(data[2 * i], data[2 * i + 1]) - m;
                ss += r * r;
            }
        double var = ss / (double)(cnt - 1); /* cnt==1 -> 0/0 = NaN as in Julia */
        double wgt = 1.0 / var;
        for (long i = 0; i < n; ++i)
            if (state[i] == states[s]) {
                power[i] = m;
                weight[i] = wgt;
            }
    }
}

/* ---------------------------------------------------------------------- */
typedef struct {
    long n;
    const double *t;   /* valid timestamps */
    const cplx *d;     /* valid data */
    const double *w;   /* weights or NULL (scalar weight 1.0) */
    const cplx *p;     /* power * FCphasor */
    cplx *model;       /* scratch, n */
    int fitoffsets;
    double omega;
    /* the mutable Modulation struct, src/Modulation.jl:26-39 */
    cplx c, a;
    double b, phi;
    int nfev;
} ora_cost;

/* updatemodulation! with power, src/Modulation.jl:122-148, followed by the
 * weighted residual norm, :323-326 and :299-315 */
static double cost_eval(ora_cost *self, double b, double phi) {
    long n = self->n;
    const double *t = self->t;
    const cplx *d = self->d, *p = self->p;
    const double *w = self->w;
    cplx *model = self->model;
    double omega = self->omega;
    self->b = b;
    self->phi = phi;
    self->nfev++;
    for (long i = 0; i < n; ++i) { /* :137 */
        double wt = omega * t[i];
        double arg = wt + phi;
        double u = b * sin(arg);
        double su, cu;
        sincos(u, &su, &cu);
        double pr = creal(p[i]), pi = cimag(p[i]);
        model[i] = CMPLX(pr * cu - pi * su, pr * su + pi * cu);
    }
    if (self->fitoffsets) { /* linearregression :174-195 / :197-215 */
        double a11 = 0.0, a22 = 0.0;
        double a12r = 0, a12i = 0, b1r = 0, b1i = 0, b2r = 0, b2i = 0;
        for (long i = 0; i < n; ++i) {
            double wi = w ? w[i] : 1.0;
            double gr = creal(model[i]), gi = cimag(model[i]);
            double dr = creal(d[i]), di = cimag(d[i]);
            a11 += wi;
            a12r += wi * gr;
            a12i += wi * gi;
            a22 += wi * (gr * gr + gi * gi);
            b1r += wi * dr;
            b1i += wi * di;
            /* w*conj(g)*d */
            double cr = wi * gr, ci = -(wi * gi);
            b2r += cr * dr - ci * di;
            b2i += cr * di + ci * dr;
        }
        cplx A12 = CMPLX(a12r, a12i), A21 = conj(A12);
        cplx B1 = CMPLX(b1r, b1i), B2 = CMPLX(b2r, b2i);
        /* StaticArrays 2x2 solve: Cramer with det = a11*a22 - a12*a21 */
        cplx det = a11 * a22 - A12 * A21;
        self->c = (a22 * B1 - A12 * B2) / det;
        self->a = (a11 * B2 - A21 * B1) / det;
        for (long i = 0; i < n; ++i) model[i] = self->c + self->a * model[i]; /* :141 */
    } else { /* :143-145 */
        double nr = 0, ni = 0, qr = 0, qi = 0;
        for (long i = 0; i < n; ++i) {
            double wi = w ? w[i] : 1.0;
            double gr = creal(model[i]), gi = cimag(model[i]);
            double mr = gr * wi, mi = gi * wi; /* mw = model .* weight */
            double dr = creal(d[i]), di = cimag(d[i]);
            /* conj(mw)*d and conj(mw)*g */
            nr += mr * dr + mi * di;
            ni += mr * di - mi * dr;
            qr += mr * gr + mi * gi;
            qi += mr * gi - mi * gr;
        }
        self->a = CMPLX(nr, ni) / CMPLX(qr, qi);
        self->c = 0.0;
        for (long i = 0; i < n; ++i) model[i] = self->a * model[i];
    }
    double s = 0.0; /* weighted_norm2(model .- data, weight) ./ N, :325 */
    for (long i = 0; i < n; ++i) {
        double rr = creal(model[i]) - creal(d[i]);
        double ri = cimag(model[i]) - cimag(d[i]);
        double wi = w ? w[i] : 1.0;
        s += wi * (rr * rr + ri * ri);
    }
    return s / (double)n;
}

static double cost_newuoa(int n, const double *x, void *data) {
    (void)n;
    return cost_eval((ora_cost *)data, x[0], x[1]);
}

/* minimize!, src/Modulation.jl:332-336: newuoa(f, xinit, 1, 1e-3; check=false)
 * -> npt = 2n+1, maxeval = 30n [OptimPackNextGen defaults, not in tree] */
static void cost_minimize(ora_cost *self, double *x, int maxfun) {
    double f;
    int nf;
    newuoa_oracle(2, 5, cost_newuoa, self, x, 1.0, 1e-3, maxfun, &f, &nf, 0, 0);
}

double ora_chi2(long n, const double *t, const double *d, const double *w,
                const double *pw, const double *fc, int fitoffsets, double b,
                double phi, double omega, double *ca) {
    cplx *p = (cplx *)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
    cplx *model = (cplx *)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
    for (long i = 0; i < n; ++i) {
        double pwi = pw ? pw[i] : 1.0;
        p[i] = CMPLX(pwi * fc[2 * i], pwi * fc[2 * i + 1]);
    }
    ora_cost cf = {n, t, (const cplx *)d, w, p, model, fitoffsets, omega, 0, 0, 0, 0, 0};
    double f = cost_eval(&cf, b, phi);
    if (ca) {
        ca[0] = creal(cf.c);
        ca[1] = cimag(cf.c);
        ca[2] = creal(cf.a);
        ca[3] = cimag(cf.a);
    }
    free(p);
    free(model);
    return f;
}

/* ---------------------------------------------------------------------- */
typedef struct {
    long n;
    const double *t;
    const cplx *data;
    const int8_t *state;
    const unsigned char *valid; /* NULL => all rows */
    long nvalid;
    int fitoffsets, recenter, maxfun;
    const double *xinit;
    cplx *out;
    double *params, *chi2;
    int *nfev;
    /* work queue over the 8 (telescope, side) groups, :387 */
    int next_group;
    pthread_mutex_t lock;
} ora_job;

static void do_group(ora_job *job, int tel, int side) {
    long n = job->n, nv = job->nvalid;
    const double *t = job->t;
    double phirange[8];
    ora_phirange(phirange);

    double *tv = (double *)malloc(sizeof(double) * (size_t)nv);
    cplx *fcv = (cplx *)malloc(sizeof(cplx) * (size_t)nv);
    cplx *dv = (cplx *)malloc(sizeof(cplx) * (size_t)nv);
    cplx *p = (cplx *)malloc(sizeof(cplx) * (size_t)nv);
    cplx *model = (cplx *)malloc(sizeof(cplx) * (size_t)nv);
    double *power = (double *)malloc(sizeof(double) * (size_t)nv);
    double *weight = (double *)malloc(sizeof(double) * (size_t)nv);
    int8_t *sv = (int8_t *)malloc((size_t)nv);

    /* :388 FCphasor = exp.(1im .* angle.(data[:, idx(k,j,FC)])), then [valid] */
    const cplx *fc = job->data + (size_t)(ora_idx(side, tel, 5) - 1) * n;
    long m = 0;
    for (long i = 0; i < n; ++i) {
        if (job->valid && !job->valid[i]) continue;
        double ang = atan2(cimag(fc[i]), creal(fc[i]));
        fcv[m] = CMPLX(cos(ang), sin(ang));
        tv[m] = t[i];
        if (job->state) sv[m] = job->state[i];
        ++m;
    }

    for (int diode = 1; diode <= 4; ++diode) { /* :389 */
        int ch = ora_idx(side, tel, diode) - 1;
        const cplx *d = job->data + (size_t)ch * n;
        m = 0;
        for (long i = 0; i < n; ++i)
            if (!job->valid || job->valid[i]) dv[m++] = d[i];

        const double *w = 0;
        if (job->state) { /* :391-392 */
            ora_mean_var_power(nv, sv, (const double *)dv, power, weight);
            w = weight;
            for (long i = 0; i < nv; ++i) p[i] = power[i] * fcv[i]; /* :396 */
        } else {                                                    /* :394 */
            for (long i = 0; i < nv; ++i) p[i] = 1.0 * fcv[i];
        }

        ora_cost lkl = {nv, tv, dv, w, p, model, job->fitoffsets, ORA_OMEGA, 0, 0, 0, 0, 0};
        double x[2];
        if (!job->xinit) { /* :402-406 */
            double binit = 0.1, best = 0;
            int kbest = 0, have_nan = 0;
            for (int k = 0; k < 8; ++k) {
                double f = cost_eval(&lkl, binit, phirange[k]);
                if (have_nan) continue;
                if (isnan(f)) { /* Julia argmin returns the first NaN */
                    kbest = k;
                    have_nan = 1;
                } else if (k == 0 || f < best) {
                    best = f;
                    kbest = k;
                }
            }
            x[0] = binit;
            x[1] = phirange[kbest];
        } else {
            x[0] = job->xinit[0];
            x[1] = job->xinit[1];
        }
        cost_minimize(&lkl, x, job->maxfun);                /* :407 */
        double lklval = cost_eval(&lkl, x[0], x[1]);        /* :408 */
        double phipi = x[1] + (x[1] < 0 ? ORA_PI : -ORA_PI); /* :409 */
        if (lklval > cost_eval(&lkl, x[0], phipi)) {        /* :411 */
            x[1] = phipi;
            cost_minimize(&lkl, x, job->maxfun);            /* :413 */
        }
        job->chi2[ch] = cost_eval(&lkl, x[0], x[1]);        /* :416 */

        /* :417-425 over ALL rows */
        cplx *o = job->out + (size_t)ch * n;
        if (job->recenter) {
            double alpha = carg(lkl.a);
            for (long i = 0; i < n; ++i) {
                double wt = ORA_OMEGA * t[i];
                double arg = wt + lkl.phi;
                double gp = lkl.b * sin(arg) + alpha; /* getphase :66-69 */
                double psi = gp - alpha;
                cplx e = CMPLX(cos(psi), -sin(psi)); /* exp(-1im*psi) */
                cplx v = job->fitoffsets ? d[i] - lkl.c : d[i];
                o[i] = v * e;
            }
        } else { /* :424 */
            for (long i = 0; i < n; ++i) {
                double wt = ORA_OMEGA * t[i];
                double arg = wt + lkl.phi;
                double u = lkl.b * sin(arg);
                cplx mdl = lkl.a * CMPLX(cos(u), sin(u));
                if (job->fitoffsets) mdl = lkl.c + mdl;
                double ang = carg(mdl);
                o[i] = d[i] * CMPLX(cos(ang), -sin(ang));
            }
        }
        /* :426-431 */
        if (lkl.b < 0) {
            lkl.b = -lkl.b;
            lkl.phi += (lkl.phi < 0 ? ORA_PI : -ORA_PI);
        }
        double *pr = job->params + 6 * ch;
        pr[0] = creal(lkl.c);
        pr[1] = cimag(lkl.c);
        pr[2] = creal(lkl.a);
        pr[3] = cimag(lkl.a);
        pr[4] = lkl.b;
        pr[5] = lkl.phi;
        if (job->nfev) job->nfev[ch] = lkl.nfev;
    }
    free(tv);
    free(fcv);
    free(dv);
    free(p);
    free(model);
    free(power);
    free(weight);
    free(sv);
}

static void *worker(void *arg) {
    ora_job *job = (ora_job *)arg;
    for (;;) {
        pthread_mutex_lock(&job->lock);
        int g = job->next_group++;
        pthread_mutex_unlock(&job->lock);
        if (g >= 8) break;
        /* product(1:4, (FT,SC)): telescope fastest, :387 */
        do_group(job, g % 4 + 1, (g / 4) ? 16 : 0);
    }
    return 0;
}

int ora_demodulateall(long n, const double *t, const double *data,
                      const int8_t *state, int onlyhigh, int fitoffsets,
                      int recenter, const double *xinit, int maxfun, double *out,
                      double *params, double *chi2, int *nfev, int nthreads) {
    if (n < 1) return -1;
    memcpy(out, data, sizeof(double) * 2 * 40 * (size_t)n); /* :353 */
    unsigned char *valid = 0;
    long nvalid = n;
    if (state) { /* :374-383 */
        valid = (unsigned char *)malloc((size_t)n);
        int any_transient = 0;
        for (long i = 0; i < n; ++i) {
            valid[i] = onlyhigh ? (state[i] == ORA_HIGH || state[i] == ORA_NORMAL) : 1;
            if (state[i] == ORA_TRANSIENT) any_transient = 1;
        }
        if (any_transient)
            for (long i = 0; i < n; ++i)
                if (state[i] == ORA_TRANSIENT) valid[i] = 0;
        nvalid = 0;
        for (long i = 0; i < n; ++i) nvalid += valid[i];
    }
    ora_job job = {n, t, (const cplx *)data, state, valid, nvalid, fitoffsets,
                   recenter, maxfun, xinit, (cplx *)out, params, chi2, nfev, 0,
                   PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 8) nthreads = 8;
    if (nthreads == 1) {
        worker(&job);
    } else {
        pthread_t th[8];
        for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], 0, worker, &job);
        for (int i = 0; i < nthreads; ++i) pthread_join(th[i], 0);
    }
    free(valid);
    return 0;
}
