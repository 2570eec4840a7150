// Host build of the product's device solver (csrc/newuoa2.cuh) so that the
// CPU test-suite can compare it bit for bit with the oracle (oracle/newuoa.c)
// without a GPU.  g++ -O2 -ffp-contract=off, mirroring nvcc -fmad=false.
#define __host__
#define __device__
#define NU_HOST_EMULATE_WARP
#include <cmath>
using std::fabs;
using std::floor;
using std::sqrt;
#include "newuoa2.cuh"

typedef double (*objfun)(int n, const double *x, void *data);
typedef void (*observer)(int nf, int n, const double *x, double f, void *data);

template <bool WARP>
static int run_solver(objfun f, void *data, double *x, double rhobeg, double rhoend, int maxfun,
                      double *fout, int *nfout, observer obs, void *obsdata) {
    static gppd::NuSinCos angles[gppd::NU_ANGLES + 1];
    for (int i = 0; i <= gppd::NU_ANGLES; ++i) gppd::nu_angle_entry(i, &angles[i]);
    gppd::Newuoa2T<WARP> s;
    s.ang = angles;
    s.start(x[0], x[1], rhobeg, rhoend, maxfun);
    double fv = 0.0;
    while (s.step(fv)) {
        fv = f(2, &s.x[1], data);
        if (obs) obs(s.nf, 2, &s.x[1], fv, obsdata);
    }
    x[0] = s.x[1];
    x[1] = s.x[2];
    *fout = s.f;
    *nfout = s.nf;
    return s.status;
}
extern "C" {
// serial angle searches (what a single device thread runs)
int newuoa2_host(objfun f, void *data, double *x, double rhobeg, double rhoend,
                 int maxfun, double *fout, int *nfout, observer obs, void *obsdata) {
    return run_solver<false>(f, data, x, rhobeg, rhoend, maxfun, fout, nfout, obs, obsdata);
}
// lane-parallel angle searches of the warp-per-fit kernel, emulated lane by lane
int newuoa2_host_warp(objfun f, void *data, double *x, double rhobeg, double rhoend,
                      int maxfun, double *fout, int *nfout, observer obs, void *obsdata) {
    return run_solver<true>(f, data, x, rhobeg, rhoend, maxfun, fout, nfout, obs, obsdata);
}
void nu_sincos_host(double x, double *s, double *c) { gppd::nu_sincos(x, s, c); }
}
