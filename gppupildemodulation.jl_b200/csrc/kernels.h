// kernels.h -- host-callable launchers of the device passes (one per pass).
#pragma once
#include "gppd_device.cuh"

namespace gppd {

// FAINT segmentation (reference buildstates, src/Faint.jl:21-73)
void launch_segmentation(const Launcher &L, const TableView &tv, const double *d_timer1,
                         int n1, const double *d_timer2, int n2, long long lag, double pre,
                         double post, long long *d_lb, void *d_events, int max_events,
                         int *d_flags, int8_t *d_state);
constexpr int SEG_EVENT_BYTES = 24;

// per-row basis + per-job theta range / valid count
void launch_basis(const Launcher &L, const TableView &tv, long long wrows, int njobs,
                  const int8_t *d_state, unsigned flags, double2 *d_basis,
                  unsigned long long *d_thkeys, int *d_nvalid, JobInfo *d_jobs);

// per-state mean / weight (reference compute_mean_var_power, src/Faint.jl:89-100)
void launch_stats(const Launcher &L, const TableView &tv, int njobs, const JobInfo *d_jobs,
                  const int8_t *d_state, unsigned flags, double2 *d_stats);

// the fit with the direct O(N)-per-call evaluator (one block per fit);
// d_fit_list == nullptr: fits 0..nfits-1, else the listed fit ids
void launch_fit_direct(const Launcher &L, const TableView &tv, int nfits, const JobInfo *d_jobs,
                       const int8_t *d_state, const double2 *d_stats, const double2 *d_basis,
                       double2 *d_z, double2 *d_y, const FitOptions &opt, const int *d_fit_list,
                       FitResult *d_results, double *d_trace);

// demodulation + repack (reference src/Modulation.jl:417-425)
void launch_demod(const Launcher &L, const TableView &tv, const OutView &ov, long long wrows,
                  const double2 *d_basis, const FitResult *d_results, unsigned flags);

// FitResult -> params / chi2 / info in the caller's layout
void launch_export(const Launcher &L, int nfits, const FitResult *d_results, double *d_params,
                   double *d_chi2, int *d_info);

// measured FP64 FMA throughput of the device (TFLOP/s), for the fit's roofline
double measure_dfma_tflops(cudaStream_t stream, double *d_scratch);

}  // namespace gppd
