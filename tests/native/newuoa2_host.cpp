// Host build of the product's device solver (csrc/newuoa2.cuh) so that the
// CPU test-suite can compare it bit for bit with the oracle (oracle/newuoa.c)
// without a GPU.  g++ -O2 -ffp-contract=off, mirroring nvcc -fmad=false.
#define __host__
#define __device__
#include <cmath>
using std::fabs;
using std::floor;
using std::sqrt;
#include "newuoa2.cuh"

extern "C" {
typedef double (*objfun)(int n, const double *x, void *data);
typedef void (*observer)(int nf, int n, const double *x, double f, void *data);

int newuoa2_host(objfun f, void *data, double *x, double rhobeg, double rhoend,
                 int maxfun, double *fout, int *nfout, observer obs, void *obsdata) {
    gppd::Newuoa2 s;
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
void nu_sincos_host(double x, double *s, double *c) { gppd::nu_sincos(x, s, c); }
}
