// fit_driver.cuh -- the per-diode fit procedure of demodulateall as a resumable
// state machine around "evaluate chi2 at (b, phi)":
//
//   8-point phase scan at b = 0.1            reference src/Modulation.jl:402-406
//   x = minimize!(lkl, xinit)  (NEWUOA)      :407  -> :332-336
//   lklval = lkl(x)                          :408
//   phipi = x[2] + (x[2] < 0 ? +pi : -pi)    :409
//   if lklval > lkl(x[1], phipi): NEWUOA again from [x[1], phipi]   :411-414
//   likelihood = lkl(x)                      :416   (this call fixes a and c)
//
// Every objective call is requested through step(): the caller evaluates
// chi2(b, phi) however it likes (one thread, a block-wide reduction, ...).
#pragma once
#include "gppd_device.cuh"
#include "newuoa2.cuh"

namespace gppd {

// WARP: see Newuoa2T (true: the object is shared by the 32 converged lanes of a warp).
template <bool WARP>
struct FitDriverT {
    enum { SCAN, NEWUOA1, LKL_X, LKL_FLIP, NEWUOA2, FINAL, DONE };
    Newuoa2T<WARP> nu;
    double b, phi;         // point to evaluate next
    double x1, x2;         // current solution
    double best, lklval, phipi, chi2;
    double rhobeg, rhoend;
    int maxfun, phase, k, kbest, have_nan, nfev, second, status;

    __device__ void start(const FitOptions &o, const NuSinCos *angles) {
        nu.ang = angles;
        rhobeg = o.rhobeg;
        rhoend = o.rhoend;
        maxfun = o.maxfun;
        nfev = 0;
        second = 0;
        status = 0;
        k = 0;
        kbest = 0;
        have_nan = 0;
        best = 0.0;
        if (o.has_xinit) {  // init = [b, phi], reference :362-364
            x1 = o.xinit[0];
            x2 = o.xinit[1];
            begin_solver(NEWUOA1, x1, x2);
        } else {
            phase = SCAN;
            b = 0.1;  // binit, reference :403
            phi = o.phi8[0];
        }
    }

    // NEWUOA's first objective call is at its start point: hand that point out
    // directly; the solver itself is set up when the value comes back (so that the
    // solver code has a single call site, see step()).
    __device__ void begin_solver(int solver_phase, double b0, double phi0) {
        phase = solver_phase;
        nu.start(b0, phi0, rhobeg, rhoend, maxfun);
        b = b0;
        phi = phi0;
    }

    // f = chi2 at the (b, phi) handed out by the previous call.  The part of a step before
    // the solver: returns 0 = the fit is finished, 1 = evaluate at (this->b, this->phi),
    // 2 = the solver has to digest f first (after_solver() then finishes the step).
    __device__ int before_solver(const FitOptions &o, double f, int &next_phase) {
        ++nfev;
        next_phase = DONE;
        switch (phase) {
        case SCAN:
            // argmin over the scan; Julia's argmin returns the first NaN
            if (!have_nan) {
                if (f != f) {
                    kbest = k;
                    have_nan = 1;
                } else if (k == 0 || f < best) {
                    best = f;
                    kbest = k;
                }
            }
            ++k;
            if (k < 8) {
                phi = o.phi8[k];
                return 1;
            }
            x1 = 0.1;
            x2 = o.phi8[kbest];
            begin_solver(NEWUOA1, x1, x2);
            return 1;
        case NEWUOA1:
            next_phase = LKL_X;
            return 2;
        case NEWUOA2:
            next_phase = FINAL;
            return 2;
        case LKL_X:
            lklval = f;
            phipi = x2 + (x2 < 0 ? PI_F64 : -PI_F64);
            phase = LKL_FLIP;
            b = x1;
            phi = phipi;
            return 1;
        case LKL_FLIP:
            if (lklval > f) {  // "bad minima", strict >
                second = 1;
                begin_solver(NEWUOA2, x1, phipi);
                return 1;
            }
            phase = FINAL;
            b = x1;
            phi = x2;
            return 1;
        case FINAL:
            chi2 = f;
            phase = DONE;
            return 0;
        default:
            return 0;
        }
    }

    __device__ void after_solver(bool more, int next_phase) {
        if (more) {
            b = nu.x[1];
            phi = nu.x[2];
        } else {
            status = nu.status;
            x1 = nu.x[1];
            x2 = nu.x[2];
            phase = next_phase;
            b = x1;
            phi = x2;
        }
    }

    // Returns true while another evaluation (at this->b, this->phi) is needed.
    __device__ bool step(const FitOptions &o, double f) {
        int next_phase;
        const int r = before_solver(o, f, next_phase);
        if (r != 2) return r == 1;
        // ---- the one call site of the solver (phase NEWUOA1 / NEWUOA2) ----
        // A solver that has not run yet first sets itself up and asks for its start
        // point, which is the point f was just evaluated at: feed f straight back.
        bool more = true;
        const int calls = nu.phase == 0 ? 2 : 1;
#pragma unroll 1
        for (int c = 0; c < calls; ++c) more = nu.step(f);
        after_solver(more, next_phase);
        return true;
    }

#ifdef __CUDACC__
    // The same step for kernels in which every lane of a warp owns a fit: ALL 32 lanes call it
    // together (`active` = this lane has a value f for its fit), and the solvers of the warp
    // are advanced segment by segment (Newuoa2T::step_coop), lanes at the same segment
    // together.  A lane's arithmetic is exactly that of step().
    __device__ bool step_coop(const FitOptions &o, double f, bool active) {
        int next_phase = DONE;
        const int r = active ? before_solver(o, f, next_phase) : 0;
        const int calls = r == 2 ? (nu.phase == 0 ? 2 : 1) : 0;
        bool more = true;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            if (!__any_sync(0xffffffffu, c < calls)) break;          // warp-uniform
            const bool m = nu.step_coop(f, !(c < calls));
            if (c < calls) more = m;
        }
        if (r != 2) return r == 1;
        after_solver(more, next_phase);
        return true;
    }
#endif
};

typedef FitDriverT<false> FitDriver;

}  // namespace gppd
