// api.cu -- the C ABI of libgppd.so (include/gppd.h): handle, pipeline slots,
// host<->device staging and the launch sequence of one table.
//
// Launch sequence per table (all on the slot's stream, no host sync inside):
//   [segmentation] -> basis -> [stats] -> fit -> demod -> export
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gppd.h"
#include "kernels.h"

using namespace gppd;

namespace {

thread_local std::string g_last_error;

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);     \
            return GPPD_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)

// GPPD_DEBUG_SYNC=1: synchronise after every pass and name the one that faulted
bool debug_sync() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GPPD_DEBUG_SYNC");
        on = (e && *e && *e != '0') ? 1 : 0;
    }
    return on == 1;
}
#define DBG(stream, name)                                                          \
    do {                                                                           \
        if (debug_sync()) {                                                        \
            cudaError_t e_ = cudaStreamSynchronize(stream);                        \
            if (e_ == cudaSuccess) e_ = cudaGetLastError();                        \
            if (e_ != cudaSuccess) {                                               \
                g_last_error = std::string("pass ") + name + ": " + cudaGetErrorString(e_); \
                return GPPD_ERR_CUDA;                                              \
            }                                                                      \
        }                                                                          \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return GPPD_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            g_last_error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
            p = nullptr;
            return GPPD_ERR_NOMEM;
        }
        cap = want;
        return GPPD_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr int NSLOTS = 4;
constexpr int NPASS = GPPD_NPASS;

struct PassTimer {
    std::vector<cudaEvent_t> ev;      // begin/end pairs
    std::vector<int> pass;            // pass id per pair
    double ms[NPASS] = {0};
    long long count[NPASS] = {0};
};
constexpr int MAX_TIMER = 1 << 16;

struct Slot {
    cudaStream_t stream = nullptr;
    // staging of caller data
    DevBuf time, volt, volt_out, t, data, out, state_in;
    // scratch
    DevBuf state, basis, z, y, thkeys, nvalid, jobs, stats, results;
    DevBuf params, chi2, info, trace;
    DevBuf timers, lb, events, flags, offsets;
    PassTimer timer;
    bool busy = false;
};

}  // namespace

struct gppd_handle_s {
    int device = 0;
    bool timing = false;
    Slot slots[NSLOTS];
    long long launches = 0;
};

namespace {

void fill_options(const gppd_options *o, FitOptions &f) {
    gppd_options d;
    memset(&d, 0, sizeof d);
    if (o) d = *o;
    f.flags = d.flags;
    f.maxfun = d.maxfun > 0 ? d.maxfun : 60;
    f.has_xinit = d.has_xinit;
    f.xinit[0] = d.xinit[0];
    f.xinit[1] = d.xinit[1];
    f.rhobeg = d.rhobeg > 0 ? d.rhobeg : 1.0;
    f.rhoend = d.rhoend > 0 ? d.rhoend : 1e-3;
    gppd_phirange(f.phi8);
}

// cudaEvent pair around one pass (only when gppd_enable_timing is on)
struct PassScope {
    Slot *s;
    cudaStream_t st;
    bool on;
    PassScope(gppd_handle h, Slot &slot, cudaStream_t stream, int pass) : s(&slot), st(stream), on(h->timing) {
        if (!on) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        s->timer.ev.push_back(a);
        s->timer.ev.push_back(b);
        s->timer.pass.push_back(pass);
    }
    ~PassScope() {
        if (on) cudaEventRecord(s->timer.ev.back(), st);
    }
};

struct RunArgs {
    TableView tv;
    OutView ov;
    long long wrows;
    const int8_t *d_state_in;  // device states or nullptr
    const double *timer1, *timer2;  // host timers or nullptr
    long long n1, n2;
    double *d_params, *d_chi2;
    int *d_info;
    double *d_trace;
    int8_t *d_state_out;  // where the states end up (may be nullptr)
};

// Enqueue all passes of one table on `stream`.  Device pointers only.
int run_table(gppd_handle h, Slot &s, cudaStream_t stream, RunArgs &a, const gppd_options *opt) {
    FitOptions fo;
    fill_options(opt, fo);
    const long long n = a.tv.n;
    if (n < 2) {
        g_last_error = "need at least 2 rows";
        return GPPD_ERR_ARG;
    }
    long long wrows = (a.wrows <= 0 || a.wrows >= n) ? n : a.wrows;
    long long njobs_ll = (n + wrows - 1) / wrows;
    if (njobs_ll * NDIODE > 0x7fffffffll || wrows > 0x7fffffffll) {
        g_last_error = "too many windows / rows per window";
        return GPPD_ERR_ARG;
    }
    int njobs = (int)njobs_ll, nfits = njobs * NDIODE;
    Launcher L{stream, &h->launches};
    int rc;

    const int8_t *d_state = a.d_state_in;
    if (!d_state && a.n1 > 0 && a.n2 > 0) {
        if (a.n1 > MAX_TIMER || a.n2 > MAX_TIMER) {
            g_last_error = "timer series too long";
            return GPPD_ERR_ARG;
        }
        int n1 = (int)a.n1, n2 = (int)a.n2;
        int8_t *st = a.d_state_out;
        if (!st) {
            if ((rc = s.state.ensure((size_t)n))) return rc;
            st = s.state.as<int8_t>();
        }
        if ((rc = s.timers.ensure(sizeof(double) * (size_t)(n1 + n2)))) return rc;
        if ((rc = s.lb.ensure(sizeof(long long) * (size_t)(n1 + n2 + 2)))) return rc;
        int max_events = n1 + n2 + 1024;
        if ((rc = s.events.ensure((size_t)SEG_EVENT_BYTES * max_events))) return rc;
        if ((rc = s.flags.ensure(2 * sizeof(int)))) return rc;
        // small pageable copies: the runtime stages them before returning, so the
        // caller's arrays may be reused at once and tables of one slot cannot race
        CK(cudaMemcpyAsync(s.timers.p, a.timer1, sizeof(double) * (size_t)n1,
                           cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(s.timers.as<double>() + n1, a.timer2, sizeof(double) * (size_t)n2,
                           cudaMemcpyHostToDevice, stream));
        {
            PassScope ps(h, s, stream, GPPD_PASS_SEGMENT);
            launch_segmentation(L, a.tv, s.timers.as<double>(), n1, s.timers.as<double>() + n1, n2,
                                0, 0.0, 0.0, s.lb.as<long long>(), s.events.p, max_events,
                                s.flags.as<int>(), st);
        }
        DBG(stream, "segmentation");
        d_state = st;
    } else if (d_state && a.d_state_out && a.d_state_out != d_state) {
        CK(cudaMemcpyAsync(a.d_state_out, d_state, (size_t)n, cudaMemcpyDeviceToDevice, stream));
    }

    if ((rc = s.basis.ensure(sizeof(double2) * (size_t)n))) return rc;
    if ((rc = s.thkeys.ensure(sizeof(unsigned long long) * 2 * (size_t)njobs))) return rc;
    if ((rc = s.nvalid.ensure(sizeof(int) * (size_t)njobs))) return rc;
    if ((rc = s.jobs.ensure(sizeof(JobInfo) * (size_t)njobs))) return rc;
    if ((rc = s.results.ensure(sizeof(FitResult) * (size_t)nfits))) return rc;
    if ((rc = s.z.ensure(sizeof(double2) * (size_t)n * NDIODE))) return rc;
    if (fo.flags & GPPD_FITOFFSETS)
        if ((rc = s.y.ensure(sizeof(double2) * (size_t)n * NDIODE))) return rc;
    if (d_state)
        if ((rc = s.stats.ensure(sizeof(double2) * 4 * (size_t)nfits))) return rc;

    {
        PassScope ps(h, s, stream, GPPD_PASS_BASIS);
        launch_basis(L, a.tv, wrows, njobs, d_state, fo.flags, s.basis.as<double2>(),
                     s.thkeys.as<unsigned long long>(), s.nvalid.as<int>(), s.jobs.as<JobInfo>());
    }
    DBG(stream, "basis");
    if (d_state) {
        PassScope ps(h, s, stream, GPPD_PASS_STATS);
        launch_stats(L, a.tv, njobs, s.jobs.as<JobInfo>(), d_state, fo.flags, s.stats.as<double2>());
    }
    DBG(stream, "stats");
    {
        PassScope ps(h, s, stream, GPPD_PASS_FIT);
        launch_fit_direct(L, a.tv, nfits, s.jobs.as<JobInfo>(), d_state, s.stats.as<double2>(),
                          s.basis.as<double2>(), s.z.as<double2>(), s.y.as<double2>(), fo, nullptr,
                          s.results.as<FitResult>(), a.d_trace);
    }
    DBG(stream, "fit_direct");
    {
        PassScope ps(h, s, stream, GPPD_PASS_DEMOD);
        launch_demod(L, a.tv, a.ov, wrows, s.basis.as<double2>(), s.results.as<FitResult>(), fo.flags);
    }
    DBG(stream, "demod");
    {
        PassScope ps(h, s, stream, GPPD_PASS_EXPORT);
        launch_export(L, nfits, s.results.as<FitResult>(), a.d_params, a.d_chi2, a.d_info);
    }
    DBG(stream, "export");
    CK(cudaGetLastError());
    return GPPD_OK;
}

int check_handle(gppd_handle h) {
    if (!h) {
        g_last_error = "null handle";
        return GPPD_ERR_ARG;
    }
    CK(cudaSetDevice(h->device));
    return GPPD_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

int gppd_version(void) { return GPPD_VERSION; }

const char *gppd_strerror(int status) {
    switch (status) {
    case GPPD_OK: return "ok";
    case GPPD_ERR_ARG: return "bad argument";
    case GPPD_ERR_CUDA: return "CUDA error";
    case GPPD_ERR_NO_DEVICE: return "no sm_100 CUDA device (libgppd has no CPU fallback)";
    case GPPD_ERR_NOMEM: return "out of device memory";
    case GPPD_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

const char *gppd_last_error(void) { return g_last_error.c_str(); }

int gppd_idx(int side, int telescope, int diode) {
    if ((side != 0 && side != 16) || telescope < 1 || telescope > 4 || diode < 1 || diode > 5)
        return -1;
    if (diode == 5) return 32 + side / 4 + (telescope - 1) + 1;
    return side + (diode - 1) + (telescope - 1) * 4 + 1;
}

int gppd_phirange(double *phi8) {
    if (!phi8) return GPPD_ERR_ARG;
    // Julia's range(-pi, pi, 8) is evaluated in twice precision: each element is
    // the correctly rounded -pi_f + k (2 pi_f / 7); long double reproduces it.
    const long double lo = -(long double)PI_F64, hi = (long double)PI_F64;
    for (int k = 0; k < 8; ++k)
        phi8[k] = (double)(((long double)(7 - k) * lo + (long double)k * hi) / 7.0L);
    phi8[0] = -PI_F64;
    phi8[7] = PI_F64;
    return GPPD_OK;
}

int gppd_create(int device, gppd_handle *out) {
    if (!out) return GPPD_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        g_last_error = e != cudaSuccess ? cudaGetErrorString(e) : "no such CUDA device";
        return GPPD_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        g_last_error = std::string("device is sm_") + std::to_string(prop.major) +
                       std::to_string(prop.minor) + ", libgppd is built for sm_100a only";
        return GPPD_ERR_NO_DEVICE;
    }
    CK(cudaSetDevice(device));
    gppd_handle h = new gppd_handle_s;
    h->device = device;
    for (int i = 0; i < NSLOTS; ++i) {
        cudaError_t es = cudaStreamCreateWithFlags(&h->slots[i].stream, cudaStreamNonBlocking);
        if (es != cudaSuccess) {
            g_last_error = cudaGetErrorString(es);
            delete h;
            return GPPD_ERR_CUDA;
        }
    }
    *out = h;
    return GPPD_OK;
}

int gppd_destroy(gppd_handle h) {
    if (!h) return GPPD_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < NSLOTS; ++i) {
        Slot &s = h->slots[i];
        if (s.stream) {
            cudaStreamSynchronize(s.stream);
            cudaStreamDestroy(s.stream);
        }
        DevBuf *bufs[] = {&s.time, &s.volt, &s.volt_out, &s.t, &s.data, &s.out, &s.state_in,
                          &s.state, &s.basis, &s.z, &s.y, &s.thkeys, &s.nvalid, &s.jobs,
                          &s.stats, &s.results, &s.params, &s.chi2, &s.info, &s.trace,
                          &s.timers, &s.lb, &s.events, &s.flags, &s.offsets};
        for (DevBuf *b : bufs) b->release();
    }
    delete h;
    return GPPD_OK;
}

int gppd_alloc_pinned(gppd_handle h, uint64_t bytes, void **out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!out) return GPPD_ERR_ARG;
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return GPPD_OK;
}

int gppd_free_pinned(gppd_handle h, void *p) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (p) CK(cudaFreeHost(p));
    return GPPD_OK;
}

int gppd_num_slots(gppd_handle) { return NSLOTS; }

int64_t gppd_launch_count(gppd_handle h) { return h ? h->launches : 0; }

int gppd_measure_fp64_peak(gppd_handle h, double *tflops) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!tflops) return GPPD_ERR_ARG;
    Slot &s = h->slots[0];
    if ((rc = s.flags.ensure(64))) return rc;
    *tflops = measure_dfma_tflops(s.stream, s.flags.as<double>());
    h->launches += 4;
    CK(cudaGetLastError());
    return GPPD_OK;
}

int gppd_enable_timing(gppd_handle h, int on) {
    if (!h) return GPPD_ERR_ARG;
    h->timing = on != 0;
    return GPPD_OK;
}

int gppd_pass_times(gppd_handle h, double *ms, int64_t *counts, int reset) {
    int rc = check_handle(h);
    if (rc) return rc;
    for (int p = 0; p < NPASS; ++p) {
        if (ms) ms[p] = 0.0;
        if (counts) counts[p] = 0;
    }
    for (int i = 0; i < NSLOTS; ++i) {
        PassTimer &t = h->slots[i].timer;
        for (size_t k = 0; k < t.pass.size(); ++k) {
            cudaEvent_t a = t.ev[2 * k], b = t.ev[2 * k + 1];
            CK(cudaEventSynchronize(b));
            float dt = 0.f;
            CK(cudaEventElapsedTime(&dt, a, b));
            t.ms[t.pass[k]] += dt;
            t.count[t.pass[k]] += 1;
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
        t.ev.clear();
        t.pass.clear();
        for (int p = 0; p < NPASS; ++p) {
            if (ms) ms[p] += t.ms[p];
            if (counts) counts[p] += t.count[p];
            if (reset) {
                t.ms[p] = 0.0;
                t.count[p] = 0;
            }
        }
    }
    return GPPD_OK;
}

int64_t gppd_num_windows(int64_t n, int64_t nwindow) {
    if (n <= 0) return 0;
    if (nwindow <= 0 || nwindow >= n) return 1;
    return (n + nwindow - 1) / nwindow;
}

// ---------------------------------------------------------------------------
int gppd_buildstates(gppd_handle h, int64_t n, const double *t, const double *timer1,
                     int64_t n1, const double *timer2, int64_t n2, int64_t lag,
                     double pre, double post, int8_t *state_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!t || !timer1 || !timer2 || !state_out || n < 2 || n1 < 1 || n2 < 1 ||
        n1 > MAX_TIMER || n2 > MAX_TIMER) {
        g_last_error = "buildstates: need n >= 2 and non-empty timers";
        return GPPD_ERR_ARG;
    }
    Slot &s = h->slots[0];
    cudaStream_t st = s.stream;
    if ((rc = s.t.ensure(sizeof(double) * (size_t)n))) return rc;
    if ((rc = s.state.ensure((size_t)n))) return rc;
    if ((rc = s.timers.ensure(sizeof(double) * (size_t)(n1 + n2)))) return rc;
    if ((rc = s.lb.ensure(sizeof(long long) * (size_t)(n1 + n2 + 2)))) return rc;
    int max_events = (int)(n1 + n2) + 1024;
    if ((rc = s.events.ensure((size_t)SEG_EVENT_BYTES * max_events))) return rc;
    if ((rc = s.flags.ensure(2 * sizeof(int)))) return rc;
    CK(cudaMemcpyAsync(s.t.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.timers.p, timer1, sizeof(double) * (size_t)n1, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.timers.as<double>() + n1, timer2, sizeof(double) * (size_t)n2,
                       cudaMemcpyHostToDevice, st));
    TableView tv;
    memset(&tv, 0, sizeof tv);
    tv.kind = 1;
    tv.n = n;
    tv.t = s.t.as<double>();
    Launcher L{st, &h->launches};
    launch_segmentation(L, tv, s.timers.as<double>(), (int)n1, s.timers.as<double>() + n1, (int)n2,
                        lag, pre, post, s.lb.as<long long>(), s.events.p, max_events,
                        s.flags.as<int>(), s.state.as<int8_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(state_out, s.state.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GPPD_OK;
}

// ---------------------------------------------------------------------------
int gppd_demodulate_f64(gppd_handle h, int64_t n, int64_t nwindow, const double *t,
                        const double *data, const int8_t *state, const gppd_options *opt,
                        double *out, double *params, double *chi2, int32_t *info,
                        double *trace) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!t || !data || !out || !params || !chi2 || n < 2) {
        g_last_error = "demodulate_f64: null buffer or n < 2";
        return GPPD_ERR_ARG;
    }
    Slot &s = h->slots[0];
    cudaStream_t st = s.stream;
    int64_t nwin = gppd_num_windows(n, nwindow);
    size_t nfits = (size_t)nwin * NDIODE;
    size_t cbytes = sizeof(double) * 2 * NCHAN * (size_t)n;
    if ((rc = s.t.ensure(sizeof(double) * (size_t)n))) return rc;
    if ((rc = s.data.ensure(cbytes))) return rc;
    if ((rc = s.out.ensure(cbytes))) return rc;
    if ((rc = s.params.ensure(sizeof(double) * 6 * nfits))) return rc;
    if ((rc = s.chi2.ensure(sizeof(double) * nfits))) return rc;
    if ((rc = s.info.ensure(sizeof(int) * GPPD_INFO_STRIDE * nfits))) return rc;
    if (trace)
        if ((rc = s.trace.ensure(sizeof(double) * 3 * GPPD_TRACE_MAX * nfits))) return rc;
    if (state)
        if ((rc = s.state_in.ensure((size_t)n))) return rc;
    CK(cudaMemcpyAsync(s.t.p, t, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.data.p, data, cbytes, cudaMemcpyHostToDevice, st));
    if (state) CK(cudaMemcpyAsync(s.state_in.p, state, (size_t)n, cudaMemcpyHostToDevice, st));
    if (trace) CK(cudaMemsetAsync(s.trace.p, 0, sizeof(double) * 3 * GPPD_TRACE_MAX * nfits, st));

    RunArgs a;
    memset(&a, 0, sizeof a);
    a.tv.kind = 1;
    a.tv.n = n;
    a.tv.t = s.t.as<double>();
    a.tv.data = s.data.as<double2>();
    a.ov.kind = 1;
    a.ov.out = s.out.as<double2>();
    a.wrows = nwindow;
    a.d_state_in = state ? s.state_in.as<int8_t>() : nullptr;
    a.d_params = s.params.as<double>();
    a.d_chi2 = s.chi2.as<double>();
    a.d_info = s.info.as<int>();
    a.d_trace = trace ? s.trace.as<double>() : nullptr;
    if ((rc = run_table(h, s, st, a, opt))) return rc;

    CK(cudaMemcpyAsync(out, s.out.p, cbytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(params, s.params.p, sizeof(double) * 6 * nfits, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(chi2, s.chi2.p, sizeof(double) * nfits, cudaMemcpyDeviceToHost, st));
    if (info)
        CK(cudaMemcpyAsync(info, s.info.p, sizeof(int) * GPPD_INFO_STRIDE * nfits,
                           cudaMemcpyDeviceToHost, st));
    if (trace)
        CK(cudaMemcpyAsync(trace, s.trace.p, sizeof(double) * 3 * GPPD_TRACE_MAX * nfits,
                           cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GPPD_OK;
}

// ---------------------------------------------------------------------------
int gppd_table_windows(int64_t n, const int32_t *time_us, double mjd, double window_s,
                       int64_t *nwindow_rows, int64_t *nwin) {
    if (!nwindow_rows || !nwin || n < 2) return GPPD_ERR_ARG;
    if (!(window_s > 0.0)) {
        *nwindow_rows = n;
        *nwin = 1;
        return GPPD_OK;
    }
    if (!time_us) return GPPD_ERR_ARG;
    // nwindow = round(Int, window / (times[2] - times[1])), ties to even,
    // reference src/GPPupilDemodulation.jl:139,192 (host code is compiled with
    // -ffp-contract=off so these are the reference's roundings)
    volatile double tmjd = 86400.0 * mjd;
    volatile double t0 = (double)time_us[0] * 1e-6;
    volatile double t1 = (double)time_us[1] * 1e-6;
    t0 = t0 + tmjd;
    t1 = t1 + tmjd;
    double w = nearbyint(window_s / (t1 - t0));
    if (!(w >= 1.0) || w > 9.0e15) {
        g_last_error = "window shorter than one row";
        return GPPD_ERR_ARG;
    }
    *nwindow_rows = (int64_t)w;
    *nwin = gppd_num_windows(n, *nwindow_rows);
    return GPPD_OK;
}

static int table_views(int64_t n, double mjd, const int32_t *d_time, const float *d_volt,
                       const double *d_offsets, float *d_volt_out, uint32_t flags, RunArgs &a) {
    memset(&a, 0, sizeof a);
    a.tv.kind = 0;
    a.tv.big_endian = (flags & GPPD_BIG_ENDIAN) ? 1 : 0;
    a.tv.n = n;
    a.tv.time_us = d_time;
    a.tv.time_stride = 4;
    a.tv.volt = d_volt;
    a.tv.volt_stride = 320;
    a.tv.tmjd = 86400.0 * mjd;
    a.tv.offsets = reinterpret_cast<const double2 *>(d_offsets);
    a.ov.kind = 0;
    a.ov.big_endian = a.tv.big_endian;
    a.ov.keepraw = (flags & GPPD_KEEPRAW) ? 1 : 0;
    a.ov.volt = d_volt_out;
    a.ov.volt_stride = a.ov.keepraw ? 576 : 320;
    return GPPD_OK;
}

int gppd_submit_table_f32(gppd_handle h, int slot, int64_t n, const int32_t *time_us,
                          double mjd, const float *volt, const double *offsets,
                          const double *timer1, int64_t n1, const double *timer2, int64_t n2,
                          double window_s, const gppd_options *opt, float *volt_out,
                          double *params, double *chi2, int32_t *info, int8_t *state_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (slot < 0 || slot >= NSLOTS || !time_us || !volt || !volt_out || !params || !chi2 || n < 2) {
        g_last_error = "process_table_f32: bad slot, null buffer or n < 2";
        return GPPD_ERR_ARG;
    }
    gppd_options o;
    memset(&o, 0, sizeof o);
    if (opt) o = *opt;
    if (!offsets) o.flags |= GPPD_FITOFFSETS;   // offsets === false  => fitoffsets, :156
    else o.flags &= ~GPPD_FITOFFSETS;
    int64_t wrows = n, nwin = 1;
    if ((rc = gppd_table_windows(n, time_us, mjd, window_s, &wrows, &nwin))) return rc;
    size_t nfits = (size_t)nwin * NDIODE;
    Slot &s = h->slots[slot];
    cudaStream_t st = s.stream;
    CK(cudaStreamSynchronize(st));  // slot reuse: previous table of this slot must be done
    size_t vbytes = sizeof(float) * 80 * (size_t)n;
    size_t obytes = sizeof(float) * ((o.flags & GPPD_KEEPRAW) ? 144 : 80) * (size_t)n;
    if ((rc = s.time.ensure(sizeof(int32_t) * (size_t)n))) return rc;
    if ((rc = s.volt.ensure(vbytes))) return rc;
    if ((rc = s.volt_out.ensure(obytes))) return rc;
    if ((rc = s.params.ensure(sizeof(double) * 6 * nfits))) return rc;
    if ((rc = s.chi2.ensure(sizeof(double) * nfits))) return rc;
    if ((rc = s.info.ensure(sizeof(int) * GPPD_INFO_STRIDE * nfits))) return rc;
    if ((rc = s.state.ensure((size_t)n))) return rc;
    if ((rc = s.offsets.ensure(sizeof(double) * 80))) return rc;
    CK(cudaMemcpyAsync(s.time.p, time_us, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.volt.p, volt, vbytes, cudaMemcpyHostToDevice, st));
    if (offsets)
        CK(cudaMemcpyAsync(s.offsets.p, offsets, sizeof(double) * 80, cudaMemcpyHostToDevice, st));
    RunArgs a;
    table_views(n, mjd, s.time.as<int32_t>(), s.volt.as<float>(),
                offsets ? s.offsets.as<double>() : nullptr, s.volt_out.as<float>(), o.flags, a);
    a.wrows = wrows;
    a.timer1 = timer1;
    a.timer2 = timer2;
    a.n1 = (timer1 && timer2) ? n1 : 0;
    a.n2 = (timer1 && timer2) ? n2 : 0;
    a.d_params = s.params.as<double>();
    a.d_chi2 = s.chi2.as<double>();
    a.d_info = s.info.as<int>();
    a.d_state_out = s.state.as<int8_t>();
    if ((rc = run_table(h, s, st, a, &o))) return rc;
    CK(cudaMemcpyAsync(volt_out, s.volt_out.p, obytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(params, s.params.p, sizeof(double) * 6 * nfits, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(chi2, s.chi2.p, sizeof(double) * nfits, cudaMemcpyDeviceToHost, st));
    if (info)
        CK(cudaMemcpyAsync(info, s.info.p, sizeof(int) * GPPD_INFO_STRIDE * nfits,
                           cudaMemcpyDeviceToHost, st));
    if (state_out && a.n1 > 0)
        CK(cudaMemcpyAsync(state_out, s.state.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    s.busy = true;
    return GPPD_OK;
}

int gppd_wait(gppd_handle h, int slot) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (slot < 0 || slot >= NSLOTS) return GPPD_ERR_ARG;
    CK(cudaStreamSynchronize(h->slots[slot].stream));
    h->slots[slot].busy = false;
    return GPPD_OK;
}

int gppd_process_table_f32(gppd_handle h, int64_t n, const int32_t *time_us, double mjd,
                           const float *volt, const double *offsets, const double *timer1,
                           int64_t n1, const double *timer2, int64_t n2, double window_s,
                           const gppd_options *opt, float *volt_out, double *params,
                           double *chi2, int32_t *info, int8_t *state_out) {
    int rc = gppd_submit_table_f32(h, 0, n, time_us, mjd, volt, offsets, timer1, n1, timer2, n2,
                                   window_s, opt, volt_out, params, chi2, info, state_out);
    if (rc) return rc;
    return gppd_wait(h, 0);
}

int gppd_process_table_f32_dev(gppd_handle h, int slot, void *stream, int64_t n,
                               int64_t nwindow_rows, const int32_t *d_time_us, double mjd,
                               const float *d_volt, const double *d_offsets,
                               const double *timer1, int64_t n1, const double *timer2,
                               int64_t n2, const gppd_options *opt, float *d_volt_out,
                               double *d_params, double *d_chi2, int32_t *d_info,
                               int8_t *d_state_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (slot < 0 || slot >= NSLOTS || !d_time_us || !d_volt || !d_volt_out || !d_params ||
        !d_chi2 || n < 2) {
        g_last_error = "process_table_f32_dev: bad slot, null buffer or n < 2";
        return GPPD_ERR_ARG;
    }
    gppd_options o;
    memset(&o, 0, sizeof o);
    if (opt) o = *opt;
    if (!d_offsets) o.flags |= GPPD_FITOFFSETS;
    else o.flags &= ~GPPD_FITOFFSETS;
    Slot &s = h->slots[slot];
    cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
    RunArgs a;
    table_views(n, mjd, d_time_us, d_volt, d_offsets, d_volt_out, o.flags, a);
    a.wrows = nwindow_rows;
    a.timer1 = timer1;
    a.timer2 = timer2;
    a.n1 = (timer1 && timer2) ? n1 : 0;
    a.n2 = (timer1 && timer2) ? n2 : 0;
    a.d_params = d_params;
    a.d_chi2 = d_chi2;
    a.d_info = d_info;
    a.d_state_out = d_state_out;
    return run_table(h, s, st, a, &o);
}

}  // extern "C"
