"""CPU oracle for the demodulateall hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front end of ``oracle/build/liboracle.so`` (sources ``gppd_oracle.c``,
``newuoa.c``) plus NumPy restatements of the reference's array packing around
the hot path (``src/GPPupilDemodulation.jl:139-253``).  PARITY UNPINNED: the
reference has no tests/golden vectors, its optimiser is not in the tree and no
Julia exists in this image (see gppd_oracle.h).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
reference arm may import this package.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")

OFF, LOW, NORMAL, HIGH, TRANSIENT = 0, 1, 2, 3, -1  # src/Faint.jl:1
FT, SC = 0, 16                                      # src/Modulation.jl:9
D1, D2, D3, D4, FC = 1, 2, 3, 4, 5                  # src/Modulation.jl:10
M_2PI = 6.283185                                    # src/Modulation.jl:11
MJD_1970_1_1 = 40587.0                              # src/GPPupilDemodulation.jl:15
DAY_TO_SEC = 24 * 60 * 60                           # src/GPPupilDemodulation.jl:16


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no-op when up to date)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("newuoa.c", "gppd_oracle.c", "newuoa.h", "gppd_oracle.h", "Makefile")]
    if (force or not os.path.exists(_LIB_PATH) or
            any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_i8p = C.POINTER(C.c_int8)
_ip = C.POINTER(C.c_int)
OBJFUN = C.CFUNCTYPE(C.c_double, C.c_int, _dp, C.c_void_p)
OBSERVER = C.CFUNCTYPE(None, C.c_int, C.c_int, _dp, C.c_double, C.c_void_p)


class StateView(C.Structure):
    _fields_ = [("n", C.c_int), ("npt", C.c_int), ("idz", C.c_int),
                ("kopt", C.c_int), ("nf", C.c_int),
                ("xbase", _dp), ("xopt", _dp), ("xpt", _dp), ("fval", _dp),
                ("gq", _dp), ("hq", _dp), ("pq", _dp), ("bmat", _dp),
                ("zmat", _dp), ("rho", C.c_double), ("delta", C.c_double)]


PROBE = C.CFUNCTYPE(None, C.POINTER(StateView), C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ora_idx.restype = C.c_int
        L.ora_idx.argtypes = [C.c_int] * 3
        L.ora_buildstates.restype = C.c_int
        L.ora_buildstates.argtypes = [C.c_long, _dp, _dp, C.c_long, _dp, C.c_long,
                                      C.c_long, C.c_double, C.c_double, _i8p]
        L.ora_mean_var_power.restype = None
        L.ora_mean_var_power.argtypes = [C.c_long, _i8p, _dp, _dp, _dp]
        L.ora_chi2.restype = C.c_double
        L.ora_chi2.argtypes = [C.c_long, _dp, _dp, _dp, _dp, _dp, C.c_int,
                               C.c_double, C.c_double, C.c_double, _dp]
        L.ora_demodulateall.restype = C.c_int
        L.ora_demodulateall.argtypes = [C.c_long, _dp, _dp, _i8p, C.c_int, C.c_int,
                                        C.c_int, _dp, C.c_int, _dp, _dp, _dp, _ip,
                                        C.c_int]
        L.ora_phirange.restype = None
        L.ora_phirange.argtypes = [_dp]
        L.newuoa_oracle.restype = C.c_int
        L.newuoa_oracle.argtypes = [C.c_int, C.c_int, OBJFUN, C.c_void_p, _dp,
                                    C.c_double, C.c_double, C.c_int, _dp, _ip,
                                    OBSERVER, C.c_void_p]
        L.newuoa_oracle_set_probe.restype = None
        L.newuoa_oracle_set_probe.argtypes = [PROBE, C.c_void_p]
        L.newuoa_oracle_counters.restype = None
        L.newuoa_oracle_counters.argtypes = [C.POINTER(C.c_long), C.c_int]
        _lib = L
    return _lib


def _p(a, typ=_dp):
    return None if a is None else a.ctypes.data_as(typ)


# --------------------------------------------------------------------------
def idx(side: int, telescope: int, diode: int) -> int:
    """1-based channel number, src/Modulation.jl:17-22."""
    return lib().ora_idx(side, telescope, diode)


def phirange() -> np.ndarray:
    out = np.empty(8)
    lib().ora_phirange(_p(out))
    return out


def newuoa(fun, x0, rhobeg=1.0, rhoend=1e-3, npt=None, maxfun=None, record=None,
           probe=None):
    """Powell's NEWUOA as the reference calls it (src/Modulation.jl:335).
    Returns (status, x, f, nf).  ``record`` collects (x, f) per evaluation;
    ``probe`` receives a dict of the solver state after every model update."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    npt = 2 * n + 1 if npt is None else npt
    maxfun = 30 * n if maxfun is None else maxfun

    def _f(nn, xp, _):
        return float(fun(np.array([xp[i] for i in range(nn)])))

    def _obs(nf, nn, xp, f, _):
        record.append((np.array([xp[i] for i in range(nn)]), f))

    def _probe(svp, _):
        s = svp.contents
        nn, m = s.n, s.npt
        arr = lambda p, k: np.array([p[i] for i in range(k)])
        probe(dict(n=nn, npt=m, idz=s.idz, kopt=s.kopt, nf=s.nf, rho=s.rho,
                   delta=s.delta, xbase=arr(s.xbase, nn), xopt=arr(s.xopt, nn),
                   xpt=arr(s.xpt, m * nn).reshape(nn, m).T,
                   fval=arr(s.fval, m), gq=arr(s.gq, nn),
                   hq=arr(s.hq, nn * (nn + 1) // 2), pq=arr(s.pq, m),
                   bmat=arr(s.bmat, (m + nn) * nn).reshape(nn, m + nn).T,
                   zmat=arr(s.zmat, m * (m - nn - 1)).reshape(m - nn - 1, m).T))

    cf = OBJFUN(_f)
    co = OBSERVER(_obs) if record is not None else C.cast(None, OBSERVER)
    cp = PROBE(_probe) if probe is not None else C.cast(None, PROBE)
    fout, nf = C.c_double(0), C.c_int(0)
    L = lib()
    L.newuoa_oracle_set_probe(cp, None)
    try:
        st = L.newuoa_oracle(n, npt, cf, None, _p(x), rhobeg, rhoend, maxfun,
                             C.byref(fout), C.byref(nf), co, None)
    finally:
        L.newuoa_oracle_set_probe(C.cast(None, PROBE), None)
    return st, x, fout.value, nf.value


def newuoa_force_bigden(on: bool):
    lib().newuoa_oracle_force_bigden(int(on))


def newuoa_counters(reset=True):
    """(trsapp, biglag, bigden, update, xbase-shift) call counts since last reset."""
    out = (C.c_long * 5)()
    lib().newuoa_oracle_counters(out, int(reset))
    return tuple(out)


# --------------------------------------------------------------------------
class FaintStates:
    """src/Faint.jl:3-19: timer1 is always the lower-voltage (HIGH) series."""

    def __init__(self, timer1, timer2, voltage1, voltage2):
        timer1 = np.ascontiguousarray(timer1, dtype=np.float64)
        timer2 = np.ascontiguousarray(timer2, dtype=np.float64)
        if voltage1 > voltage2:  # LOW > HIGH, src/Faint.jl:14-16
            timer1, timer2, voltage1, voltage2 = timer2, timer1, voltage2, voltage1
        self.timer1, self.timer2 = timer1, timer2
        self.voltage1, self.voltage2 = float(voltage1), float(voltage2)
        self.state1, self.state2 = HIGH, LOW


def buildfaintparameters(hdr: dict) -> FaintStates:
    """src/GPPupilDemodulation.jl:64-81 on a plain dict of header keywords."""
    start1 = hdr["ESO INS ANLO3 TIMER1"] + MJD_1970_1_1 * DAY_TO_SEC
    start2 = hdr["ESO INS ANLO3 TIMER2"] + MJD_1970_1_1 * DAY_TO_SEC
    # Julia range `start .+ rate .* (0:(repeat-1))`
    # (a lazy twice-precision range: each element is rounded once, which long
    # double arithmetic reproduces)
    def _timer(start, rate, repeat):
        k = np.arange(int(repeat), dtype=np.longdouble)
        return (np.longdouble(start) + np.longdouble(rate) * k).astype(np.float64)
    timer1 = _timer(start1, hdr["ESO INS ANLO3 RATE1"], hdr["ESO INS ANLO3 REPEAT1"])
    timer2 = _timer(start2, hdr["ESO INS ANLO3 RATE2"], hdr["ESO INS ANLO3 REPEAT2"])
    return FaintStates(timer1, timer2, hdr["ESO INS ANLO3 VOLTAGE1"],
                       hdr["ESO INS ANLO3 VOLTAGE2"])


def buildstates(fs: FaintStates, timestamp, lag=0, preswitchdelay=0.0,
                postwitchdelay=0.0) -> np.ndarray:
    """src/Faint.jl:21-73 -> int8 MetState per sample."""
    t = np.ascontiguousarray(timestamp, dtype=np.float64)
    out = np.empty(t.size, dtype=np.int8)
    rc = lib().ora_buildstates(t.size, _p(t), _p(fs.timer1), fs.timer1.size,
                               _p(fs.timer2), fs.timer2.size, int(lag),
                               float(preswitchdelay), float(postwitchdelay),
                               _p(out, _i8p))
    if rc != 0:
        raise ValueError("buildstates: need >= 2 samples and non-empty timers")
    return out


def buildstates_py(fs: FaintStates, timestamp, lag=0, preswitchdelay=0.0,
                   postwitchdelay=0.0) -> np.ndarray:
    """Pure-Python twin of ora_buildstates (small cases; cross-check of the C)."""
    import math
    t = [float(v) for v in timestamp]
    n = len(t)
    timestep = t[1] - t[0]
    t1 = [float(v) + lag * timestep for v in fs.timer1]
    t2 = [float(v) + lag * timestep for v in fs.timer2]
    premax = math.ceil(preswitchdelay / timestep)
    postmax = math.ceil(postwitchdelay / timestep)
    cur = NORMAL
    first1, first2 = t1.pop(0), t2.pop(0)
    forget = 0
    out = np.empty(n, dtype=np.int8)
    for k in range(n):
        time = t[k]
        if time >= first1:
            cur, forget = HIGH, premax
            if not t1:
                first1 = t[-1]
                if first2 == t[-1]:
                    cur = NORMAL
            else:
                first1 = t1.pop(0)
        if time >= first2:
            cur, forget = LOW, postmax
            if not t2:
                first2 = t[-1]
                if first1 == t[-1]:
                    cur = NORMAL
            else:
                first2 = t2.pop(0)
        if forget > 0:
            out[k] = TRANSIENT
            forget -= 1
        else:
            out[k] = cur
    return out


def compute_mean_var_power(states, data):
    """src/Faint.jl:89-100 -> (power, weight) per sample."""
    s = np.ascontiguousarray(states, dtype=np.int8)
    d = np.ascontiguousarray(data, dtype=np.complex128)
    m = np.empty(d.size)
    w = np.empty(d.size)
    lib().ora_mean_var_power(d.size, _p(s, _i8p), _p(d.view(np.float64)), _p(m), _p(w))
    return m, w


def chi2(t, d, fcphasor, b, phi, weight=None, power=None, fitoffsets=False,
         omega=M_2PI):
    """One objective evaluation (src/Modulation.jl:323-326). Returns (chi2, c, a)."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.complex128)
    fc = np.ascontiguousarray(fcphasor, dtype=np.complex128)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
    pw = None if power is None else np.ascontiguousarray(power, dtype=np.float64)
    ca = np.empty(4)
    f = lib().ora_chi2(t.size, _p(t), _p(d.view(np.float64)), _p(w), _p(pw),
                       _p(fc.view(np.float64)), int(fitoffsets), float(b),
                       float(phi), float(omega), _p(ca))
    return f, complex(ca[0], ca[1]), complex(ca[2], ca[3])


def demodulateall(timestamp, data, init="auto", recenter=True, faintparam=None,
                  onlyhigh=False, fitoffsets=False, preswitchdelay=0.01,
                  postwitchdelay=0.3, maxfun=60, nthreads=1, return_nfev=False):
    """src/Modulation.jl:344-435.

    data: (N, 40) complex128.  faintparam: None | FaintStates | int8 state vector.
    Returns (output (N,40) complex128, param (32,6) = (c.re,c.im,a.re,a.im,b,phi),
    likelihood (32,))."""
    t = np.ascontiguousarray(timestamp, dtype=np.float64)
    dat = np.asfortranarray(data, dtype=np.complex128)  # channel-major
    n = t.size
    assert dat.shape == (n, 40)
    state = None
    if isinstance(faintparam, FaintStates):  # :366-367
        state = buildstates(faintparam, t, preswitchdelay=preswitchdelay,
                            postwitchdelay=postwitchdelay)
    elif faintparam is not None:             # :368-369
        state = np.ascontiguousarray(faintparam, dtype=np.int8)
    xinit = None if isinstance(init, str) else np.ascontiguousarray(init, dtype=np.float64)
    out = np.empty((n, 40), dtype=np.complex128, order="F")
    params = np.zeros((32, 6))
    like = np.zeros(32)
    nfev = np.zeros(32, dtype=np.int32)
    rc = lib().ora_demodulateall(
        n, _p(t), _p(dat.T.reshape(-1).view(np.float64)), _p(state, _i8p),
        int(onlyhigh), int(fitoffsets), int(recenter), _p(xinit), int(maxfun),
        _p(out.T.reshape(-1).view(np.float64)), _p(params), _p(like),
        _p(nfev, _ip), int(nthreads))
    if rc != 0:
        raise ValueError("ora_demodulateall failed")
    if return_nfev:
        return out, params, like, nfev
    return out, params, like


# --------------------------------------------------------------------------
def make_times(time_us, mjd):
    """src/GPPupilDemodulation.jl:139."""
    return np.asarray(time_us).astype(np.float64) * 1e-6 + (DAY_TO_SEC * float(mjd))


def circle_centre(x, y):
    """Centre of the least-squares circle through the points (x, y): the algebraic
    (Kasa) fit, minimise sum (x^2 + y^2 - 2 x0 x - 2 y0 y - c)^2, solved as the linear
    least-squares problem [2x 2y 1] (x0, y0, c)' = x^2 + y^2 by SVD.  No circle (fewer
    than 3 points, or all on one line) -> 0."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size < 3:
        return 0j
    # shift to the centroid first: lstsq on raw volts squares the offset / radius ratio
    xm, ym = x.mean(), y.mean()
    u, v = x - xm, y - ym
    A = np.stack([2 * u, 2 * v, np.ones_like(u)], axis=1)
    sol, _, rank, _ = np.linalg.lstsq(A, u * u + v * v, rcond=1e-9)
    if rank < 3:
        return 0j
    return complex(xm + sol[0], ym + sol[1])


def compute_offsets(cmplxV, state=None):
    """src/GPPupilDemodulation.jl:105-125: `fit(Circle, reim(channel)...)` for each of
    the 40 channels, over the HIGH samples when there are states (:108), over all
    samples otherwise (:120); returns the 40 complex centres.  `Circle` is not defined
    in the reference's environment (the call throws there); the fit restated here is the
    algebraic least-squares circle that call stands for (`circle_centre`)."""
    cmplxV = np.asarray(cmplxV)
    sel = slice(None) if state is None else (np.asarray(state) == HIGH)
    return np.array([circle_centre(cmplxV[sel, ch].real, cmplxV[sel, ch].imag)
                     for ch in range(40)], dtype=np.complex128)


def processmetrology(time_us, volt, mjd, window=None, faintparam=None,
                     keepraw=False, onlyhigh=False, offsets=True, nthreads=1,
                     maxfun=60):
    """src/GPPupilDemodulation.jl:128-255 on arrays instead of a FITS HDU.

    time_us: (N,) int32; volt: (N, 80) float32 (row n = the 80 VOLT values of
    table row n, i.e. Julia's 80 x N column-major matrix); offsets: (40,)
    complex128 centres, False (fit the centres) or True (``compute_offsets``: the
    reference throws there, see that function).
    Returns (table dict, hdr dict) with the reference's column/keyword names."""
    volt32 = np.asarray(volt, dtype=np.float32)
    n = volt32.shape[0]
    times = make_times(time_us, mjd)
    state = None if faintparam is None else buildstates(faintparam, times)  # :141-145
    v = volt32.astype(np.float64)                                           # :147
    cmplx = v[:, 0::2] + 1j * v[:, 1::2]                                    # :148
    fitoffsets = False
    if offsets is False:
        fitoffsets = True
    elif offsets is True:
        cmplx = cmplx - compute_offsets(cmplx, state).reshape(1, 40)             # :154
    else:
        cmplx = cmplx - np.asarray(offsets, dtype=np.complex128).reshape(1, 40)  # :152
    table, hdr = {}, {}
    names = [(s, sn, j, d) for s, sn in ((FT, "FT"), (SC, "SC")) for j in range(1, 5)
             for d in range(1, 5)]

    def norm_b_phi(b, phi):  # :177-180 (rem2pi(., RoundNearest))
        if b < 0:
            b = -b
            phi = float(np.remainder(phi + np.pi + np.pi, 2 * np.pi) - np.pi)
        return b, phi

    if window is None:
        out, param, _ = demodulateall(times, cmplx, faintparam=state, onlyhigh=onlyhigh,
                                      fitoffsets=fitoffsets, nthreads=nthreads,
                                      maxfun=maxfun)
        for s, sn, j, d in names:  # :174-189
            ch = idx(s, j, d) - 1
            b, phi = norm_b_phi(param[ch, 4], param[ch, 5])
            a = complex(param[ch, 2], param[ch, 3])
            suffix = f"{sn} T{j} D{d}"
            if fitoffsets:
                hdr[f"DEMODULATION CENTER X0 {suffix}"] = param[ch, 0]
                hdr[f"DEMODULATION CENTER Y0 {suffix}"] = param[ch, 1]
            hdr[f"DEMODULATION AMPLITUDE ABS {suffix}"] = abs(a)
            hdr[f"DEMODULATION AMPLITUDE ARG {suffix}"] = float(np.angle(a))
            hdr[f"DEMODULATION SIN AMPLITUDE {suffix}"] = b
            hdr[f"DEMODULATION SIN PHASE {suffix}"] = phi
    else:
        nwindow = int(np.round(window / (times[1] - times[0])))  # :192 (ties-to-even)
        out = np.empty((n, 40), dtype=np.complex128)
        cols = {k: np.empty((n, 32)) for k in ("ABSA", "ARGA", "B", "PHI", "X0", "Y0")}
        for lo in range(0, n, nwindow):  # Iterators.partition, :204
            I = slice(lo, min(lo + nwindow, n))
            o, param, _ = demodulateall(times[I], cmplx[I], faintparam=None if state is None else state[I],
                                        onlyhigh=onlyhigh, fitoffsets=fitoffsets,
                                        nthreads=nthreads, maxfun=maxfun)
            out[I] = o
            for ch in range(32):
                b, phi = norm_b_phi(param[ch, 4], param[ch, 5])
                a = complex(param[ch, 2], param[ch, 3])
                cols["B"][I, ch] = b
                cols["PHI"][I, ch] = phi
                cols["X0"][I, ch] = param[ch, 0]
                cols["Y0"][I, ch] = param[ch, 1]
                cols["ABSA"][I, ch] = abs(a)
                cols["ARGA"][I, ch] = np.angle(a)
        if fitoffsets:
            table["X0"] = cols["X0"].astype(np.float32)
            table["Y0"] = cols["Y0"].astype(np.float32)
        for k in ("ABSA", "ARGA", "B", "PHI"):
            table[k] = cols[k].astype(np.float32)
        if state is not None:
            table["STATE"] = state.astype(np.int8)
    if keepraw:  # :163-168
        s = np.empty((n, 144))
        s[:, :80] = v
        s[:, 80::2] = out[:, :32].real
        s[:, 81::2] = out[:, :32].imag
        v = s
    else:        # :170-171
        v = v.copy()
        v[:, 0::2] = out.real
        v[:, 1::2] = out.imag
    hdr["PROCSOFT"] = "GPPupilDemodulation.jl"  # :252
    table["VOLT"] = v.astype(np.float32)         # :253
    return table, hdr
