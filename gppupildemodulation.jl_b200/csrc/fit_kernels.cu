// fit_kernels.cu -- the per-diode modulation fit (reference
// src/Modulation.jl:387-416: objective :323-326, model update :122-148,
// closed-form linear parameters :174-215, NEWUOA :332-336).
//
// Compiled with -fmad=false: the solver state machine must take exactly the
// oracle's floating-point steps (see newuoa2.cuh).  Hot loops therefore spell
// out their fused multiply-adds with fma().
//
// Two evaluators of chi2(b, phi) drive the same fit procedure (fit_driver.cuh):
//
//  k_fit_harmonic_warp   one WARP per fit (the default).  chi2 from the per-fit
//                  harmonic table (harm_kernels.cu): S_gd = sum_k J_k(b) e^{-jkq} Z_k,
//                  lane k holding harmonic k, butterfly-summed in a fixed order.
//                  The solver state lives once per warp in shared memory; all
//                  lanes run the same (uniform) solver algebra and split the
//                  NEWUOA angle searches between them.  Low latency: a whole-file
//                  job has only 32 fits.
//  k_fit_harmonic  one THREAD per fit (very large batches of small windows, where
//                  throughput matters more than latency).  All fits of a warp
//                  evaluate in lock step (the solver is a resumable state machine),
//                  the solver algebra in between diverges.
//                  Both give up (fallback flag) when |b| > 5 or the job has no
//                  uniform phase quantum for the requested phi.
//
//  k_fit_direct    one BLOCK per fit, the reference's own formulation: one pass
//                  over the rows per objective call,
//                      e_n = exp(j b sin(theta_n + q)),  S_gd = sum conj(e_n) z_n
//                  with z_n = w_n conj(p_n)(d_n - mu) either staged in HBM scratch
//                  (method = direct) or recomputed from the table (fallback of
//                  the harmonic path: no scratch needed).  Fixed-order block
//                  reductions (deterministic).
//
// Both compute chi2 = (S_dd - Re(conj(c') S_d' + conj(a) S_gd)) / N, the
// reference's sum w |c + a g - d|^2 / N at its own least-squares (c, a).
#include <cstdlib>

#include "fit_driver.cuh"
#include "fit_math.cuh"
#include "gppd_device.cuh"
#include "kernels.h"

namespace gppd {

constexpr int FIT_THREADS = 256;
// up to this many fits per batch the harmonic fit runs one warp per fit; above, one
// thread per fit (measured cross-over on a B200: 3200 fits 0.8 ms against 1.5 ms, 12800
// fits 3.2 ms against 1.65 ms; GPPD_FIT_WARP_MAX_FITS overrides, for the tests)
static int fit_warp_max_fits() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("GPPD_FIT_WARP_MAX_FITS");
        v = e ? atoi(e) : 6000;
    }
    return v;
}
#define FIT_WARP_MAX_FITS fit_warp_max_fits()

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

template <class Driver>
__device__ __forceinline__ void store_result(FitResult *results, int fit, const Driver &drv,
                                             const JobInfo &ji, double cre, double cim, double are,
                                             double aim, int method) {
    FitResult r;
    r.cre = cre; r.cim = cim; r.are = are; r.aim = aim;
    r.b = drv.b; r.phi = drv.phi;
    r.alpha = atan2(aim, are);
    r.chi2 = drv.chi2;
    PhaseQ pq = make_phaseq(drv.phi, ji.thmin, ji.thmax);
    r.q = pq.q; r.cq = pq.cq; r.sq = pq.sq; r.uniform = pq.uniform;
    r.nfev = drv.nfev; r.status = drv.status; r.method = method; r.second = drv.second;
    r.fallback = 0;
    results[fit] = r;
}

// ===========================================================================
// Harmonic evaluator: one thread per fit
// ===========================================================================
// Phase quantum usable by the harmonic form: the job's uniform q, or phi itself
// when every |theta| < 4096 (ulp <= 2^-41: the per-row rounding of theta + phi
// is below 2.3e-13 rad and is neglected).
__device__ __forceinline__ bool harm_quantum(double phi, const JobInfo &ji, double &q) {
    if (phase_quantum(phi, ji.thmin, ji.thmax, q)) return true;
    if (fabs(ji.thmin) < 4096.0 && fabs(ji.thmax) < 4096.0 && fabs(phi) < 4096.0) {
        q = phi;
        return true;
    }
    return false;
}

// Both harmonic evaluators (one warp per fit, one thread per fit) compute the SAME
// 32 "lane terms" with the same operations and add them in the same (butterfly) order,
// so a fit's chi2 values -- hence its trajectory -- do not depend on which kernel the
// size of the batch selected.
struct HarmLane {       // lane k = 1..HK: (A, B, C, D) of harmonic k; lane 0: (Z_0.re, Z_0.im)
    double zA, zB, zC, zD;
    double yA, yB, yC, yD;
};

// powers e^{j 2^i q}, i = 0..4
struct QPowers {
    double pr[5], pi[5];
};
__device__ __forceinline__ void q_powers(double cq, double sq, QPowers &P) {
    P.pr[0] = cq;
    P.pi[0] = sq;
#pragma unroll
    for (int i = 1; i < 5; ++i) {
        P.pr[i] = fma(P.pr[i - 1], P.pr[i - 1], -(P.pi[i - 1] * P.pi[i - 1]));
        P.pi[i] = 2.0 * (P.pr[i - 1] * P.pi[i - 1]);
    }
}

// term of harmonic `lane` (0 = DC, 1..HK) of S_gd (tr, ti) and S_g (ur, ui); J = J_lane(b)
template <bool OFFS>
__device__ __forceinline__ void lane_terms(int lane, double J, const QPowers &P, const HarmLane &h,
                                           double &tr, double &ti, double &ur, double &ui) {
    // (ck, sk) = (cos lane q, sin lane q): product of the powers selected by the bits of lane
    double ck = 1.0, sk = 0.0;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
        if ((lane >> bit) & 1) {
            const double nr = fma(ck, P.pr[bit], -(sk * P.pi[bit])), ni = fma(ck, P.pi[bit], sk * P.pr[bit]);
            ck = nr;
            sk = ni;
        }
    }
    tr = ti = ur = ui = 0.0;
    if (lane == 0) {
        tr = J * h.zA;
        ti = J * h.zB;
        if (OFFS) { ur = J * h.yA; ui = J * h.yB; }
    } else if (lane <= HK) {
        const double tj = 2.0 * J;
        if ((lane & 1) == 0) {
            tr = tj * fma(ck, h.zA, -(sk * h.zD));
            ti = tj * fma(ck, h.zC, -(sk * h.zB));
            if (OFFS) {
                ur = tj * fma(ck, h.yA, -(sk * h.yD));
                ui = tj * fma(ck, h.yC, -(sk * h.yB));
            }
        } else {
            tr = tj * fma(ck, h.zB, sk * h.zC);
            ti = tj * -fma(ck, h.zD, sk * h.zA);
            if (OFFS) {
                ur = tj * -fma(ck, h.yB, sk * h.yC);
                ui = tj * fma(ck, h.yD, sk * h.yA);
            }
        }
    }
}

// the harmonic table entries lane `lane` needs
template <bool OFFS>
__device__ __forceinline__ void load_lane(const double *H, int nfits, int lane, HarmLane &h) {
    h.zA = h.zB = h.zC = h.zD = h.yA = h.yB = h.yC = h.yD = 0.0;
    if (lane == 0) {
        h.zA = H[(long long)HV_Z0R * nfits];
        h.zB = H[(long long)HV_Z0I * nfits];
        if (OFFS) {
            h.yA = H[(long long)HV_Y0R * nfits];
            h.yB = H[(long long)HV_Y0I * nfits];
        }
    } else if (lane <= HK) {
        const double *z = H + (long long)(HV_ZK + 4 * (lane - 1)) * nfits;
        h.zA = z[0]; h.zB = z[nfits]; h.zC = z[2 * (long long)nfits]; h.zD = z[3 * (long long)nfits];
        if (OFFS) {
            const double *y = H + (long long)(HV_YK + 4 * (lane - 1)) * nfits;
            h.yA = y[0]; h.yB = y[nfits]; h.yC = y[2 * (long long)nfits]; h.yD = y[3 * (long long)nfits];
        }
    }
}

// ---- one thread per fit: the 32 lane terms one after the other, added in the order of
// the warp kernel's xor butterfly (o = 1, 2, 4, 8, 16)
template <bool OFFS>
__device__ __forceinline__ bool eval_harmonic(const double *H, int nfits, const FitConsts &kc,
                                              const JobInfo &ji, double b, double phi, double &f,
                                              double &cre, double &cim, double &are, double &aim) {
    double q;
    if (!(fabs(b) <= HARM_BMAX) || !(kc.sdd == kc.sdd) || !harm_quantum(phi, ji, q)) return false;
    double J[HK + 1];
    bessel_j(b, J);
    double sq, cq;
    sincos_moderate(q, &sq, &cq);
    QPowers P;
    q_powers(cq, sq, P);
    // pairwise summation in the order of the warp kernel's butterfly (xor 1, 2, 4, 8, 16):
    // lanes stream in, a 6-level carry stack merges equal-sized partial sums
    // (fully unrolled: `lane` is a compile-time constant in every copy, so the powers of
    // e^{jq} it needs, its slot of J[] and its place in the carry stack are all resolved by
    // the compiler and everything stays in registers)
    double st[6][4];
#pragma unroll
    for (int lane = 0; lane < 32; ++lane) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};     // the warp kernel's idle lanes contribute +0.0
        if (lane <= HK) {
            HarmLane h;
            load_lane<OFFS>(H, nfits, lane, h);
            lane_terms<OFFS>(lane, J[lane], P, h, v[0], v[1], v[2], v[3]);
        }
        int lvl = 0;
#pragma unroll
        for (int idx = lane; idx & 1; idx >>= 1, ++lvl) {
#pragma unroll
            for (int c = 0; c < (OFFS ? 4 : 2); ++c) v[c] = st[lvl][c] + v[c];
        }
#pragma unroll
        for (int c = 0; c < (OFFS ? 4 : 2); ++c) st[lvl][c] = v[c];
    }
    f = solve_linear(kc, OFFS, st[5][0], st[5][1], OFFS ? st[5][2] : 0.0, OFFS ? st[5][3] : 0.0, cre,
                     cim, are, aim);
    return true;
}

// the solver's 49-angle table (newuoa2.cuh), filled by the first 50 threads of a block
__device__ __forceinline__ void fill_angle_table(NuSinCos *tab) {
    for (int i = threadIdx.x; i <= NU_ANGLES; i += blockDim.x) nu_angle_entry(i, &tab[i]);
}

// ---- one warp per fit ---------------------------------------------------------
template <bool OFFS>
__device__ __forceinline__ bool eval_harmonic_warp(const HarmLane &h, const FitConsts &kc,
                                                   const JobInfo &ji, int lane, double b, double phi,
                                                   double &f, double &cre, double &cim, double &are,
                                                   double &aim) {
    double q;
    if (!(fabs(b) <= HARM_BMAX) || !(kc.sdd == kc.sdd) || !harm_quantum(phi, ji, q)) return false;   // warp-uniform
    const double J = bessel_j_lane(b, lane);
    double sq, cq;
    sincos_moderate(q, &sq, &cq);
    QPowers P;
    q_powers(cq, sq, P);
    double tr, ti, ur, ui;
    lane_terms<OFFS>(lane, J, P, h, tr, ti, ur, ui);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {   // fixed-order butterfly: every lane gets the same bits
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
        ti += __shfl_xor_sync(0xffffffffu, ti, o);
        if (OFFS) {
            ur += __shfl_xor_sync(0xffffffffu, ur, o);
            ui += __shfl_xor_sync(0xffffffffu, ui, o);
        }
    }
    f = solve_linear(kc, OFFS, tr, ti, ur, ui, cre, cim, are, aim);
    return true;
}

constexpr int FITW_WARPS = 1;   // fits per block: fits differ 4x in length (25..131 objective
                                // calls), a block would wait for its slowest one

template <bool OFFS>
__global__ void __launch_bounds__(FITW_WARPS * 32, 22)
k_fit_harmonic_warp(const TableDesc *tabs, const JobInfo *jobs, const double *htab, int nfits,
                    FitOptions opt, FitResult *results, double *trace, int *fbq) {
    __shared__ FitDriverT<true> s_drv[FITW_WARPS];
    __shared__ NuSinCos s_ang[NU_ANGLES + 1];
    fill_angle_table(s_ang);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fit = blockIdx.x * FITW_WARPS + warp;
    if (fit >= nfits) return;
    const int job = fit / NDIODE, ch = fit % NDIODE;
    if (!group_on(opt.flags, ch >> 2)) return;   // gppd_options.group_mask: not this call's fit
    const JobInfo ji = jobs[job];
    const TableDesc &tb = tabs[ji.table];
    const double *H = htab + fit;

    FitConsts kc;
    kc.sw = H[(long long)HV_SW * nfits];
    kc.sdd = H[(long long)HV_SDD * nfits];
    kc.sgg = H[(long long)HV_SGG * nfits];
    kc.sdr = H[(long long)HV_SDR * nfits];
    kc.sdi = H[(long long)HV_SDI * nfits];
    kc.nvalid = (double)ji.nvalid;
    kc.mur = kc.mui = 0.0;
    if (OFFS) {
        double2 mu = row_sample(tb.tv, ji.row0, ch);
        kc.mur = mu.x;
        kc.mui = mu.y;
    }
    HarmLane h;
    load_lane<OFFS>(H, nfits, lane, h);

    FitDriverT<true> &drv = s_drv[warp];
    drv.start(opt, s_ang);
    double cre = 0, cim = 0, are = 0, aim = 0, f = 0;
    double *tr = trace ? trace + (long long)fit * (3 * 160) : nullptr;
    bool failed = false;
    for (;;) {
        const double b = drv.b, phi = drv.phi;
        if (!eval_harmonic_warp<OFFS>(h, kc, ji, lane, b, phi, f, cre, cim, are, aim)) {
            failed = true;
            break;
        }
        if (tr && lane == 0 && drv.nfev < 160) {
            tr[3 * drv.nfev] = b;
            tr[3 * drv.nfev + 1] = phi;
            tr[3 * drv.nfev + 2] = f;
        }
        // The 32 lanes advance ONE solver object in shared memory, every lane computing and
        // storing the same values.  That is only sound while the warp is converged (a lane
        // that ran ahead would read state another lane has already advanced), so the warp is
        // re-converged around every step; inside the solver the lane-dependent code (the
        // angle searches) ends with a __syncwarp of its own.
        __syncwarp();
        const bool more = drv.step(opt, f);
        __syncwarp();
        if (!more) break;
    }
    if (lane != 0) return;
    if (failed) {
        FitResult r;
        r.cre = r.cim = r.are = r.aim = r.b = r.phi = r.alpha = r.chi2 = 0.0;
        r.q = r.cq = r.sq = 0.0;
        r.uniform = 0; r.nfev = 0; r.status = 0; r.method = 0; r.second = 0;
        r.fallback = 1;
        results[fit] = r;
        fbq[1 + atomicAdd(fbq, 1)] = fit;   // queue for the direct evaluator
        return;
    }
    store_result(results, fit, drv, ji, cre, cim, are, aim, 2);
}

// ---- one thread per fit ---------------------------------------------------------
#ifndef FITT_MINB
#define FITT_MINB 1
#endif
template <bool OFFS>
__global__ void __launch_bounds__(128, FITT_MINB)
k_fit_harmonic(const TableDesc *tabs, const JobInfo *jobs, const double *htab, int nfits,
               FitOptions opt, FitResult *results, double *trace, int *fbq) {
    __shared__ NuSinCos s_ang[NU_ANGLES + 1];
    fill_angle_table(s_ang);
    __syncthreads();
    const int fit = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = fit < nfits && group_on(opt.flags, (fit % NDIODE) >> 2);
    const int fidx = live ? fit : 0;
    const int job = fidx / NDIODE, ch = fidx % NDIODE;
    const JobInfo ji = jobs[job];
    const TableDesc &tb = tabs[ji.table];
    const double *H = htab + fidx;

    FitConsts kc;
    kc.sw = H[(long long)HV_SW * nfits];
    kc.sdd = H[(long long)HV_SDD * nfits];
    kc.sgg = H[(long long)HV_SGG * nfits];
    kc.sdr = H[(long long)HV_SDR * nfits];
    kc.sdi = H[(long long)HV_SDI * nfits];
    kc.nvalid = (double)ji.nvalid;
    kc.mur = kc.mui = 0.0;
    if (OFFS) {
        double2 mu = row_sample(tb.tv, ji.row0, ch);
        kc.mur = mu.x;
        kc.mui = mu.y;
    }

    FitDriver drv;
    drv.start(opt, s_ang);
    double cre = 0, cim = 0, are = 0, aim = 0, f = 0;
    double *tr = (trace && live) ? trace + (long long)fit * (3 * 160) : nullptr;
    bool running = live, failed = false;
    // lock-step loop: every fit of the warp evaluates, then the warp advances its solvers
    // together, segment by segment (FitDriverT::step_coop: lanes whose solvers are at the same
    // place of the algorithm run it in one pass instead of one after the other -- measured
    // on 1e6 fits of 512 rows: the solver part of this kernel was 73 % of its instructions
    // with a third of the lanes active)
    while (__any_sync(0xffffffffu, running)) {
        bool feed = false;
        if (running) {
            const double b = drv.b, phi = drv.phi;
            if (!eval_harmonic<OFFS>(H, nfits, kc, ji, b, phi, f, cre, cim, are, aim)) {
                failed = true;
                running = false;
            } else {
                if (tr && drv.nfev < 160) {
                    tr[3 * drv.nfev] = b;
                    tr[3 * drv.nfev + 1] = phi;
                    tr[3 * drv.nfev + 2] = f;
                }
                feed = true;
            }
        }
        const bool more = drv.step_coop(opt, f, feed);
        if (feed) running = more;
    }
    if (!live) return;
    if (failed) {
        FitResult r;
        r.cre = r.cim = r.are = r.aim = r.b = r.phi = r.alpha = r.chi2 = 0.0;
        r.q = r.cq = r.sq = 0.0;
        r.uniform = 0; r.nfev = 0; r.status = 0; r.method = 0; r.second = 0;
        r.fallback = 1;
        results[fit] = r;
        fbq[1 + atomicAdd(fbq, 1)] = fit;   // queue for the direct evaluator
        return;
    }
    store_result(results, fit, drv, ji, cre, cim, are, aim, 2);
}

void launch_fit_harmonic(const Launcher &L, const TableDesc *d_tabs, const JobInfo *d_jobs,
                         const double *d_htab, int nfits, const FitOptions &opt,
                         FitResult *d_results, double *d_trace, int *d_fbq) {
    if (nfits <= 0) return;
    const bool offs = (opt.flags & 2u) != 0;
    if (nfits <= FIT_WARP_MAX_FITS) {   // latency matters: one warp per fit
        const int blocks = (nfits + FITW_WARPS - 1) / FITW_WARPS;
        if (offs)
            k_fit_harmonic_warp<true><<<blocks, FITW_WARPS * 32, 0, L.stream>>>(
                d_tabs, d_jobs, d_htab, nfits, opt, d_results, d_trace, d_fbq);
        else
            k_fit_harmonic_warp<false><<<blocks, FITW_WARPS * 32, 0, L.stream>>>(
                d_tabs, d_jobs, d_htab, nfits, opt, d_results, d_trace, d_fbq);
    } else {                            // throughput matters: one thread per fit
        const int blocks = (nfits + 127) / 128;
        if (offs)
            k_fit_harmonic<true><<<blocks, 128, 0, L.stream>>>(d_tabs, d_jobs, d_htab, nfits, opt,
                                                              d_results, d_trace, d_fbq);
        else
            k_fit_harmonic<false><<<blocks, 128, 0, L.stream>>>(d_tabs, d_jobs, d_htab, nfits, opt,
                                                               d_results, d_trace, d_fbq);
    }
    *L.counter += 1;
}

// ===========================================================================
// Direct evaluator: one block per fit
// ===========================================================================
// SCRATCH = true : z / y staged in HBM by the prologue (method = direct)
// SCRATCH = false: z / y recomputed from the table at every call (fallback of the
//                  harmonic path; only fits flagged `fallback` run)
template <bool OFFS, bool SCRATCH>
__global__ void __launch_bounds__(FIT_THREADS)
k_fit_direct(const TableDesc *tabs, const JobInfo *jobs, int SP, const double *spart1,
             const double *spart2, FitOptions opt, FitResult *results, double *trace,
             const int *fbq) {
    __shared__ double red[7 * 8];
    __shared__ double2 st4[4];
    __shared__ NuSinCos s_ang[NU_ANGLES + 1];
    fill_angle_table(s_ang);
    // SCRATCH: block = fit.  Fallback: the blocks share the queue of the fits the
    // harmonic evaluator gave up on (fbq[0] = count, fbq[1..] = fit numbers).
  for (int q = blockIdx.x; SCRATCH ? q == (int)blockIdx.x : q < fbq[0]; q += gridDim.x) {
    const int fit = SCRATCH ? q : fbq[1 + q];
    __syncthreads();   // shared scratch of the previous fit is no longer in use
    const int job = fit / NDIODE, ch = fit % NDIODE;
    if (!group_on(opt.flags, ch >> 2)) continue;   // gppd_options.group_mask (block-uniform)
    const int group = ch >> 2, fcch = fc_channel(group);
    const JobInfo ji = jobs[job];
    const TableDesc &tb = tabs[ji.table];
    const int8_t *state = tb.state;
    const unsigned flags = opt.flags;
    double2 *z = SCRATCH ? tb.z + (long long)ch * tb.tv.n + ji.row0 : nullptr;
    double2 *y = (SCRATCH && OFFS) ? tb.y + (long long)ch * tb.tv.n + ji.row0 : nullptr;
    const double2 *bas = tb.basis + ji.row0;

    // per-state (mean |d|, 1/var |d|), reference compute_mean_var_power
    if (threadIdx.x < 4)
        st4[threadIdx.x] = state ? stats_mean_weight(spart2, job * NGROUP + group, ch & 3, threadIdx.x)
                                 : make_double2(1.0, 1.0);
    __syncthreads();

    FitConsts kc;
    kc.nvalid = (double)ji.nvalid;
    kc.mur = kc.mui = 0.0;
    if (OFFS) {
        double2 mu = row_sample(tb.tv, ji.row0, ch);
        kc.mur = mu.x;
        kc.mui = mu.y;
    }

    // z_n = w conj(p)(d - mu), y_n = w p, p = power .* FCphasor (:396)
    auto row_zy = [&](int i, double2 &zz, double2 &yy, double &w, double &dr, double &di,
                      double &pp) -> bool {
        const long long r = ji.row0 + i;
        double m = 1.0;
        w = 1.0;
        if (state) {
            const int st = state[r];
            if (!row_valid(st, flags)) return false;
            w = st4[st & 3].y;
            m = st4[st & 3].x;
        }
        const double2 d = row_sample(tb.tv, r, ch);
        const double2 fc = fc_phasor(row_sample(tb.tv, r, fcch));
        dr = d.x - kc.mur;
        di = d.y - kc.mui;
        const double pr = m * fc.x, pi = m * fc.y;
        const double wpr = w * pr, wpi = w * pi;
        zz.x = fma(wpr, dr, wpi * di);
        zz.y = fma(wpr, di, -(wpi * dr));
        yy.x = wpr;
        yy.y = wpi;
        pp = fma(pr, pr, pi * pi);
        return true;
    };

    {
        double acc[5] = {0, 0, 0, 0, 0};  // sw, sdd, sgg, sdr, sdi
        for (int i = threadIdx.x; i < ji.nrows; i += FIT_THREADS) {
            double2 zz = make_double2(0.0, 0.0), yy = zz;
            double w, dr, di, pp;
            if (row_zy(i, zz, yy, w, dr, di, pp)) {
                acc[0] += w;
                acc[1] = fma(w, fma(dr, dr, di * di), acc[1]);
                acc[2] = fma(w, pp, acc[2]);
                acc[3] = fma(w, dr, acc[3]);
                acc[4] = fma(w, di, acc[4]);
            }
            if (SCRATCH) {
                z[i] = zz;
                if (OFFS) y[i] = yy;
            }
        }
        block_sum_vec<5>(acc, red);
        kc.sw = acc[0];
        kc.sdd = acc[1];
        kc.sgg = acc[2];
        kc.sdr = acc[3];
        kc.sdi = acc[4];
    }
    __syncthreads();  // z / y visible to the whole block

    FitDriver drv;
    drv.start(opt, s_ang);
    double cre = 0, cim = 0, are = 0, aim = 0, f = 0;
    double *tr = trace ? trace + (long long)fit * (3 * 160) : nullptr;
    for (;;) {
        const double b = drv.b, phi = drv.phi;
        const PhaseQ pq = make_phaseq(phi, ji.thmin, ji.thmax);
        double acc[4] = {0, 0, 0, 0};
        for (int i = threadIdx.x; i < ji.nrows; i += FIT_THREADS) {
            double2 zz, yy = make_double2(0.0, 0.0);
            if (SCRATCH) {
                zz = z[i];
                if (OFFS) yy = y[i];
            } else {
                double w, dr, di, pp;
                zz = make_double2(0.0, 0.0);
                if (!row_zy(i, zz, yy, w, dr, di, pp)) continue;
            }
            const double2 sc = bas[i];
            double sn;
            if (pq.uniform) {
                sn = fma(sc.x, pq.cq, sc.y * pq.sq);
            } else {
                sn = sin_arg(pq, phi, row_theta(tb.tv, ji.row0 + i), sc);
            }
            double su, cu;
            sincos(b * sn, &su, &cu);
            // conj(e) z
            acc[0] = fma(cu, zz.x, fma(su, zz.y, acc[0]));
            acc[1] = fma(cu, zz.y, fma(-su, zz.x, acc[1]));
            if (OFFS) {  // e y
                acc[2] = fma(cu, yy.x, fma(-su, yy.y, acc[2]));
                acc[3] = fma(cu, yy.y, fma(su, yy.x, acc[3]));
            }
        }
        block_sum_vec<4>(acc, red);
        f = solve_linear(kc, OFFS, acc[0], acc[1], acc[2], acc[3], cre, cim, are, aim);
        if (tr && threadIdx.x == 0 && drv.nfev < 160) {
            tr[3 * drv.nfev] = b;
            tr[3 * drv.nfev + 1] = phi;
            tr[3 * drv.nfev + 2] = f;
        }
        if (!drv.step(opt, f)) break;
    }
    if (threadIdx.x == 0) store_result(results, fit, drv, ji, cre, cim, are, aim, 1);
  }
}

void launch_fit_direct(const Launcher &L, const TableDesc *d_tabs, const JobInfo *d_jobs,
                       int nfits, int SP, const double *d_spart1, const double *d_spart2,
                       const FitOptions &opt, bool scratch, FitResult *d_results, double *d_trace,
                       const int *d_fbq) {
    if (nfits <= 0) return;
    const bool offs = (opt.flags & 2u) != 0;
    // fallback: a fixed grid walks the (usually empty) queue
    const int grid = scratch ? nfits : (nfits < 1184 ? nfits : 1184);
#define GPPD_LAUNCH_DIRECT(O, S)                                                            \
    k_fit_direct<O, S><<<grid, FIT_THREADS, 0, L.stream>>>(d_tabs, d_jobs, SP, d_spart1,    \
                                                           d_spart2, opt, d_results, d_trace, d_fbq)
    if (offs && scratch) GPPD_LAUNCH_DIRECT(true, true);
    else if (offs) GPPD_LAUNCH_DIRECT(true, false);
    else if (scratch) GPPD_LAUNCH_DIRECT(false, true);
    else GPPD_LAUNCH_DIRECT(false, false);
#undef GPPD_LAUNCH_DIRECT
    *L.counter += 1;
}

}  // namespace gppd
