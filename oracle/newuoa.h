/*
 * oracle/newuoa.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of M.J.D. Powell's NEWUOA (2004), "The NEWUOA software for
 * unconstrained optimization without derivatives", as used by the reference
 * through OptimPackNextGen.Powell.Newuoa.newuoa (reference call site
 * src/Modulation.jl:335; import src/Modulation.jl:2).  The solver's source is
 * NOT in the reference tree and its version is not pinned (no Manifest.toml,
 * no [compat] in Project.toml:1-15), so this file restates the published
 * algorithm (routines NEWUOB, TRSAPP, BIGLAG, BIGDEN, UPDATE).
 * PARITY UNPINNED: no golden vector of the reference pins these iterates.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may link or call this file.  The product (libgppd.so) never does.
 */
#ifndef GPPD_ORACLE_NEWUOA_H
#define GPPD_ORACLE_NEWUOA_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double (*newuoa_objfun)(int n, const double *x, void *data);

/* status codes (check=false in the reference means none of them is an error,
 * src/Modulation.jl:335) */
#define NEWUOA_SUCCESS 0
#define NEWUOA_BAD_NPT (-1)
#define NEWUOA_ROUNDING_ERRORS (-2)      /* trust region step failed to reduce Q */
#define NEWUOA_TOO_MANY_EVALUATIONS (-3) /* objective called MAXFUN times */
#define NEWUOA_NO_MEMORY (-4)

/* Observer called after every objective evaluation (tests use it to record
 * the sequence of trial points and to check the solver's invariants). */
typedef void (*newuoa_observer)(int nf, int n, const double *x, double f,
                                void *data);

/*
 * Minimise f over R^n from x (in/out).  npt interpolation conditions
 * (n+2 <= npt <= (n+1)(n+2)/2), trust radii rhobeg >= rhoend, at most maxfun
 * objective calls.  On return x holds the best point, *fout its value and
 * *nfout the number of objective calls.
 */
int newuoa_oracle(int n, int npt, newuoa_objfun f, void *data, double *x,
                  double rhobeg, double rhoend, int maxfun, double *fout,
                  int *nfout, newuoa_observer obs, void *obsdata);

/*
 * Invariant probe for tests: when non-NULL, called once per iteration right
 * after the model/inverse-matrix update with read-only views of the state.
 */
typedef struct {
    int n, npt, idz, kopt, nf;
    const double *xbase, *xopt, *xpt, *fval, *gq, *hq, *pq, *bmat, *zmat;
    double rho, delta;
} newuoa_state_view;
typedef void (*newuoa_probe)(const newuoa_state_view *s, void *data);
void newuoa_oracle_set_probe(newuoa_probe p, void *data);

/* test counters: calls of trsapp, biglag, bigden, update and xbase shifts
 * (process-global, not thread safe; tests only) */
void newuoa_oracle_counters(long *out5, int reset);
/* tests: force the (rarely taken) BIGDEN branch on every model step */
void newuoa_oracle_force_bigden(int on);

#ifdef __cplusplus
}
#endif
#endif
