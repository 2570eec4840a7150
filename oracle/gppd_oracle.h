/*
 * oracle/gppd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, FP64, -ffp-contract=off) of the reference's
 * demodulateall hot path: src/Modulation.jl:344-435 and its callees, plus
 * src/Faint.jl:21-73 (buildstates) and :89-100 (compute_mean_var_power).
 * PARITY UNPINNED: the reference ships no tests or golden vectors and its
 * optimiser (OptimPackNextGen NEWUOA) is not in the tree (see newuoa.h); no
 * Julia toolchain exists in this image, so the reference cannot be run here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may use this library.  The product path never links it.
 */
#ifndef GPPD_ORACLE_H
#define GPPD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* MetState values, src/Faint.jl:1 */
enum { ORA_OFF = 0, ORA_LOW = 1, ORA_NORMAL = 2, ORA_HIGH = 3, ORA_TRANSIENT = -1 };

/* src/Modulation.jl:17-22; side 0=FT 16=SC, telescope 1..4, diode 1..4 or 5=FC;
 * returns the 1-based channel number */
int ora_idx(int side, int telescope, int diode);

/* src/Faint.jl:21-73.  timer1 is the HIGH series (the FaintStates constructor
 * swap, src/Faint.jl:12-19, is the caller's job).  Returns 0, or -1 on bad
 * arguments (n < 2 or an empty timer). */
int ora_buildstates(long n, const double *t, const double *timer1, long n1,
                    const double *timer2, long n2, long lag, double preswitchdelay,
                    double postswitchdelay, int8_t *state);

/* src/Faint.jl:89-100 on already-selected samples: per-sample mean |d| of the
 * sample's state and 1/var(|d|) (n-1 denominator). data is interleaved re,im */
void ora_mean_var_power(long n, const int8_t *state, const double *data,
                        double *power, double *weight);

/* One objective evaluation, src/Modulation.jl:323-326 -> :122-148 -> :174-215.
 * w and pw may be NULL (bright: w=1, power=1); fc is the unit FC phasor
 * (interleaved).  Writes (c,a) into ca[4] and returns chi2. */
double ora_chi2(long n, const double *t, const double *d, const double *w,
                const double *pw, const double *fc, int fitoffsets, double b,
                double phi, double omega, double *ca);

/*
 * demodulateall, src/Modulation.jl:344-435.
 *   t[n] seconds; data n x 40 complex128 column-major (channel-major), interleaved
 *   state: NULL (faintparam=nothing) or n MetState values
 *   xinit: NULL (init=:auto) or 2 doubles
 *   out   n x 40 complex128; params 32 x 6 = (c.re,c.im,a.re,a.im,b,phi);
 *   chi2 32; nfev 32 (objective calls per diode, may be NULL)
 *   nthreads: worker threads over the 8 (telescope,side) groups, :387
 */
int ora_demodulateall(long n, const double *t, const double *data,
                      const int8_t *state, int onlyhigh, int fitoffsets,
                      int recenter, const double *xinit, int maxfun, double *out,
                      double *params, double *chi2, int *nfev, int nthreads);

/* the 8-point phase scan grid range(-pi, pi, 8), src/Modulation.jl:360 */
void ora_phirange(double *phi8);

#ifdef __cplusplus
}
#endif
#endif
