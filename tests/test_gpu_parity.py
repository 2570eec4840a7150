"""GPU parity tests: libgppd.so (through the C ABI / host mirror) against the
CPU oracle on the same seeded inputs.

How parity is established for the fit (PARITY UNPINNED against the real Julia
stack: no Julia here, no reference tests; see oracle/gppd_oracle.h):

 (1) objective parity  -- every chi2(b, phi) the GPU evaluated (its trace) equals
     the oracle's objective at the same point to <= 1e-10 relative (measured
     ~1e-12: the GPU uses the algebraically identical one-pass form), and the
     closed-form (c, a) to <= 1e-9;
 (2) solver parity     -- replaying the GPU's own objective values through the
     oracle's NEWUOA + fit procedure reproduces the GPU's trial points BIT FOR
     BIT, i.e. the device ran exactly the reference procedure on its values;
 (3) end to end        -- where the oracle's own trajectory coincides with the
     GPU's, parameters and demodulated output agree to 1e-9 (the north-star
     tolerance).  NEWUOA has rounding-level ties (e.g. SUM > DISTSQ right after
     DELTA = HALF*DNORM), so a 1e-15 difference in chi2 forks ~9 % of the fits
     (measured on the oracle against itself, tests/test_fork_envelope.py).  Every
     case must keep >= 22 of its 32 fits on the oracle's trajectory, every forked
     fit inside the measured envelope (tests/fitref.py), and the forks of the whole
     module are checked as a population (test_fork_population); the BASELINE-size
     configurations (1e5 rows, bright and FAINT, `-c stefan` and `-c fit`) are run
     against the oracle too (test_full_size_vs_oracle).
Integer/index work (segmentation, window partition, channel map) is bit-exact.
"""
import numpy as np
import pytest

import fitref
from conftest import make_case
from fitref import FORK_HARD, MIN_COINCIDE, REL_FIT, REL_OBJ

pytestmark = pytest.mark.gpu

# fork statistics of every end-to-end comparison of this module, checked as a population
# by test_fork_population at the end of the file (tolerances: tests/fitref.py, measured by
# tests/test_fork_envelope.py)
FORK_STATS = []


def _oracle_objective(ora, t, z, state, ch, onlyhigh, fitoffsets):
    return fitref.oracle_objective(ora, t, z, state, ch, onlyhigh, fitoffsets)


def _replay(ora, trace, nfev, maxfun=60, xinit=None):
    """The reference fit procedure (src/Modulation.jl:402-416) driven by the
    oracle's NEWUOA but fed the GPU's own objective values; asserts the GPU
    visited bit-identical points.  Returns (x, chi2, second_run)."""
    pos = [0]

    def f_at(x):
        k = pos[0]
        assert k < nfev, "GPU made fewer objective calls than the reference procedure"
        assert trace[k, 0] == x[0] and trace[k, 1] == x[1], (k, trace[k, :2], x)
        pos[0] += 1
        return trace[k, 2]

    if xinit is None:
        phi8 = ora.phirange()
        fs = [f_at((0.1, p)) for p in phi8]
        nan = [i for i, f in enumerate(fs) if f != f]
        k = nan[0] if nan else int(np.argmin(fs))
        x0 = [0.1, phi8[k]]
    else:
        x0 = list(xinit)
    _, x, _, _ = ora.newuoa(f_at, x0, maxfun=maxfun)
    lklval = f_at(x)
    phipi = x[1] + (np.pi if x[1] < 0 else -np.pi)
    second = False
    if lklval > f_at((x[0], phipi)):
        second = True
        _, x, _, _ = ora.newuoa(f_at, [x[0], phipi], maxfun=maxfun)
    chi2 = f_at(x)
    assert pos[0] == nfev, "GPU made more objective calls than the reference procedure"
    return x, chi2, second


CASES = {
    "bright": dict(faint=False, fitoffsets=False, onlyhigh=False),
    "bright_fitoffsets": dict(faint=False, fitoffsets=True, onlyhigh=False),
    "faint": dict(faint=True, fitoffsets=False, onlyhigh=False),
    "faint_onlyhigh_fitoffsets": dict(faint=True, fitoffsets=True, onlyhigh=True),
}


METHODS = ["direct", "auto"]   # auto = Jacobi-Anger harmonic evaluator (+ direct fallback)


@pytest.fixture(scope="module", params=[(c, m) for c in CASES for m in METHODS],
                ids=lambda p: f"{p[0]}-{p[1]}")
def fitcase(request, gp, ora):
    cfg = dict(CASES[request.param[0]], method=request.param[1])
    tab = make_case(gp.synthetic, 6000, k=11, faint=cfg["faint"], ora=ora)
    off = None if cfg["fitoffsets"] else gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    state = tab["state"]
    kw = dict(faintparam=state, onlyhigh=cfg["onlyhigh"], fitoffsets=cfg["fitoffsets"])
    out, par, like, info, trace = gp.demodulateall(t, z, raw=True, return_info=True,
                                                   return_trace=True, method=cfg["method"], **kw)
    assert np.all(info[:, 2] == (1 if cfg["method"] == "direct" else 2))   # evaluator used
    oo, op, ol, onf = ora.demodulateall(t, z, nthreads=8, return_nfev=True, **kw)
    return dict(cfg=cfg, t=t, z=z, state=state, out=out, par=par, like=like, info=info,
                trace=trace, oo=oo, op=op, ol=ol, onf=onf)


def test_objective_parity_along_trace(fitcase, ora):
    c = fitcase
    worst_f = worst_a = 0.0
    for ch in range(0, 32, 3):
        obj = _oracle_objective(ora, c["t"], c["z"], c["state"], ch, c["cfg"]["onlyhigh"],
                                c["cfg"]["fitoffsets"])
        nf = c["info"][ch, 0]
        for k in range(nf):
            b, phi, f = c["trace"][ch, k]
            fo, co, ao = obj(b, phi)
            worst_f = max(worst_f, abs(f - fo) / fo)
        # the last call fixes (c, a): compare with the oracle's closed form there
        b, phi, f = c["trace"][ch, nf - 1]
        fo, co, ao = obj(b, phi)
        a = complex(c["par"][ch, 2], c["par"][ch, 3])
        worst_a = max(worst_a, abs(a - ao) / abs(ao))
        if c["cfg"]["fitoffsets"]:
            cc = complex(c["par"][ch, 0], c["par"][ch, 1])
            worst_a = max(worst_a, abs(cc - co) / max(abs(co), abs(ao)))
        assert abs(c["like"][ch] - fo) <= REL_OBJ * fo
    assert worst_f <= REL_OBJ, worst_f
    assert worst_a <= REL_FIT, worst_a


def test_solver_replay_bitwise(fitcase, ora):
    c = fitcase
    for ch in range(32):
        nf = c["info"][ch, 0]
        x, chi2, second = _replay(ora, c["trace"][ch], nf)
        b, phi = x
        if b < 0:   # sign normalisation, src/Modulation.jl:427-430
            b, phi = -b, phi + (np.pi if phi < 0 else -np.pi)
        assert c["par"][ch, 4] == b and c["par"][ch, 5] == phi
        assert c["like"][ch] == chi2
        assert bool(c["info"][ch, 3]) == second


def test_demodulation_formula(fitcase, ora):
    # out = (d - c) exp(-j psi), psi = (b sin(w t + phi) + alpha) - alpha over ALL rows
    # (src/Modulation.jl:417-421, :66-69), from the GPU's own parameters
    c = fitcase
    t, z, par = c["t"], c["z"], c["par"]
    worst = 0.0
    for ch in range(32):
        nf = c["info"][ch, 0]
        b, phi = c["trace"][ch, nf - 1, :2]       # as fitted, before the sign flip
        a = complex(par[ch, 2], par[ch, 3])
        cc = complex(par[ch, 0], par[ch, 1]) if c["cfg"]["fitoffsets"] else 0.0
        alpha = np.angle(a)
        psi = (b * np.sin(ora.M_2PI * t + phi) + alpha) - alpha
        ref = (z[:, ch] - cc) * np.exp(-1j * psi)
        worst = max(worst, np.abs(c["out"][:, ch] - ref).max() / np.abs(z[:, ch]).max())
    assert worst <= 1e-12, worst
    assert np.array_equal(c["out"][:, 32:], z[:, 32:])     # output = copy(data), :353
    assert np.all(par[:, 4] >= 0)


def test_end_to_end_vs_oracle(fitcase, ora):
    c = fitcase
    par, op = c["par"], c["op"]
    # same trajectory <=> same number of objective calls and parameters equal to 1e-9;
    # compare_fits also holds every forked fit to the measured envelope (|delta b|,
    # |delta phi| <= FORK_HARD, relative chi2 difference <= FORK_CHI2_HARD)
    coincide, stats = fitref.compare_fits(par, c["like"], op, c["ol"], c["info"][:, 0], c["onf"])
    FORK_STATS.append(stats)
    print("forks %d / 32, max |db| %.1e |dphi| %.1e dchi2 %.1e"
          % (stats["forks"], stats["max_db"], stats["max_dphi"], stats["max_dchi2"]))
    assert coincide.sum() >= MIN_COINCIDE, "most fits must follow the oracle's trajectory exactly"
    a, ao = par[:, 2] + 1j * par[:, 3], op[:, 2] + 1j * op[:, 3]
    sel = coincide
    assert (np.abs(a - ao)[sel] <= REL_FIT * np.abs(ao)[sel]).all()
    assert (np.abs(c["like"] - c["ol"])[sel] <= REL_FIT * c["ol"][sel]).all()
    scale = np.abs(c["z"][:, :32]).max(axis=0)
    err = np.abs(c["out"][:, :32] - c["oo"][:, :32]).max(axis=0) / scale
    assert (err[sel] <= REL_FIT).all()
    # forked fits: the outputs differ by the phase the parameter difference makes,
    # |delta psi| <= |delta b| + b |delta phi|
    fork = ~sel
    if fork.any():
        assert err[fork].max() <= 4 * FORK_HARD


@pytest.mark.parametrize("method", METHODS)
def test_determinism(gp, ora, method):
    tab = make_case(gp.synthetic, 3000, k=3)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    r1 = gp.demodulateall(t, z, raw=True, method=method)
    r2 = gp.demodulateall(t, z, raw=True, method=method)
    for a, b in zip(r1, r2):
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("method", METHODS)
def test_init_vector_and_no_recenter(gp, ora, method):
    tab = make_case(gp.synthetic, 4000, k=5)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    out, par, like, info, trace = gp.demodulateall(t, z, init=[2.0, 0.0], recenter=False, raw=True,
                                                   return_info=True, return_trace=True,
                                                   method=method)
    # init=[b, phi]: no scan, the solver starts at the given point (src/Modulation.jl:362-364)
    assert np.all(trace[:, 0, 0] == 2.0) and np.all(trace[:, 0, 1] == 0.0)
    for ch in range(0, 32, 7):
        _replay(ora, trace[ch], info[ch, 0], xinit=[2.0, 0.0])
    # recenter=false: out = data * exp(-1im * angle(model(t))), :424
    for ch in (0, 13, 31):
        nf = info[ch, 0]
        b, phi = trace[ch, nf - 1, :2]
        a = complex(par[ch, 2], par[ch, 3])
        model = a * np.exp(1j * b * np.sin(ora.M_2PI * t + phi))
        ref = z[:, ch] * np.exp(-1j * np.angle(model))
        assert np.abs(out[:, ch] - ref).max() <= 1e-12 * np.abs(z[:, ch]).max()


@pytest.mark.parametrize("method", METHODS)
def test_relative_timestamps_nonuniform_quantum(gp, ora, method):
    """Timestamps starting at 0 put theta in many binades: the per-row phase
    quantum path.  The objective must still match the oracle (which adds phi
    to theta row by row like the reference)."""
    tab = make_case(gp.synthetic, 3000, k=6)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    trel = t - t[0]
    out, par, like, info, trace = gp.demodulateall(trel, z, raw=True, return_info=True,
                                                   return_trace=True, method=method)
    for ch in (0, 9, 22):
        obj = _oracle_objective(ora, trel, z, None, ch, False, False)
        for k in range(info[ch, 0]):
            b, phi, f = trace[ch, k]
            fo = obj(b, phi)[0]
            assert abs(f - fo) <= REL_OBJ * fo
        _replay(ora, trace[ch], info[ch, 0])


@pytest.mark.parametrize("seed", range(4))
def test_segmentation_bit_exact(gp, ora, seed):
    rng = np.random.default_rng(100 + seed)
    for trial in range(40):
        n = int(rng.integers(2, 3000))
        dt = rng.choice([1.0, 0.002, 0.5])
        steps = rng.choice([dt, dt, dt, 0.0], size=n) if trial % 3 else np.full(n, dt)
        t = np.cumsum(steps) + rng.choice([0.0, 5.2e9])
        if trial % 11 == 10:          # non-monotonic timestamps: the serial path
            j = int(rng.integers(1, n))
            t[j:] -= 3 * dt
        if not t[1] - t[0] > 0:
            t[1:] += dt
        n1, n2 = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        span = t.max() - t.min() + 4 * dt
        t1 = t.min() - 2 * dt + np.sort(rng.uniform(0, span, n1))
        t2 = t.min() - 2 * dt + np.sort(rng.uniform(0, span, n2))
        if trial % 5 == 0:
            t1[-1] = t[-1]            # a real event equal to the end-of-queue sentinel
        if trial % 7 == 0:
            t2[:] = t2[::-1]          # unsorted queue
        fs_o = ora.FaintStates(t1, t2, 1.0, 2.0)
        fs_g = gp.FaintStates(t1, t2, 1.0, 2.0)
        pre, post = rng.choice([0.0, 0.5 * dt, 3 * dt]), rng.choice([0.0, 2.2 * dt, 0.3])
        lag = int(rng.integers(-2, 3))
        a = ora.buildstates(fs_o, t, lag=lag, preswitchdelay=pre, postwitchdelay=post)
        b = gp.buildstates(fs_g, t, lag=lag, preswitchdelay=pre, postwitchdelay=post)
        assert np.array_equal(a, b), (seed, trial)


@pytest.mark.parametrize("method", METHODS)
def test_faintstates_struct_path(gp, ora, method):
    # faintparam::FaintStates -> buildstates with preswitchdelay=0.01, postwitchdelay=0.3
    # (src/Modulation.jl:366-367), which creates TRANSIENT rows that are excluded
    tab = make_case(gp.synthetic, 5000, k=8, faint=True, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    fo, fg = tab["faintstates"], gp.FaintStates(tab["faintstates"].timer1, tab["faintstates"].timer2, 1.0, 2.0)
    st = ora.buildstates(fo, t, preswitchdelay=0.01, postwitchdelay=0.3)
    assert (st == ora.TRANSIENT).sum() > 100
    out, par, like, info, trace = gp.demodulateall(t, z, faintparam=fg, raw=True, return_info=True,
                                                   return_trace=True, method=method)
    for ch in (2, 17):
        obj = _oracle_objective(ora, t, z, st, ch, False, False)
        for k in range(0, info[ch, 0], 3):
            b, phi, f = trace[ch, k]
            assert abs(f - obj(b, phi)[0]) <= REL_OBJ * f
        _replay(ora, trace[ch], info[ch, 0])


@pytest.mark.parametrize("method", METHODS)
def test_windows_equal_separate_calls(gp, ora, method):
    # the per-window loop of src/GPPupilDemodulation.jl:204-225 in one launch ==
    # one demodulateall per window; ragged last window included
    tab = make_case(gp.synthetic, 2300, k=9)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    out, par, like = gp.demodulateall(t, z, raw=True, nwindow=500, method=method)
    assert par.shape == (5 * 32, 6)
    for w, lo in enumerate(range(0, 2300, 500)):
        hi = min(lo + 500, 2300)
        o1, p1, l1 = gp.demodulateall(t[lo:hi], z[lo:hi], raw=True, method=method)
        assert o1.tobytes() == np.asfortranarray(out[lo:hi]).tobytes()
        assert p1.tobytes() == par[32 * w:32 * (w + 1)].tobytes()
        assert l1.tobytes() == like[32 * w:32 * (w + 1)].tobytes()


def _table_compare(gp, ora, tab, offsets, window, keepraw, faint, onlyhigh=False, method="auto"):
    fs_o = tab["faintstates"] if faint else None
    fs_g = gp.FaintStates(fs_o.timer1, fs_o.timer2, 1.0, 2.0) if faint else None
    tg, hg = gp.processmetrology({"TIME": tab["time_us"], "VOLT": tab["volt"]}, tab["mjd"],
                                 window=window, faintparam=fs_g, keepraw=keepraw,
                                 onlyhigh=onlyhigh, offsets=offsets, method=method)
    to, ho = ora.processmetrology(tab["time_us"], tab["volt"], tab["mjd"], window=window,
                                  faintparam=fs_o, keepraw=keepraw, onlyhigh=onlyhigh,
                                  offsets=offsets, nthreads=8)
    return tg, hg, to, ho


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("mode", ["stefan", "fit", "stefan_keepraw", "stefan_faint"])
def test_table_whole_file(gp, ora, mode, method):
    faint = mode.endswith("faint")
    tab = make_case(gp.synthetic, 5000, k=21, faint=faint, ora=ora)
    offsets = False if mode == "fit" else gp.synthetic.stefan_centres()
    tg, hg, to, ho = _table_compare(gp, ora, tab, offsets, None, "keepraw" in mode, faint, method=method)
    assert set(hg) == set(ho) and hg["PROCSOFT"] == "GPPupilDemodulation.jl"
    vg, vo = tg["VOLT"], to["VOLT"]
    assert vg.dtype == np.float32 and vg.shape == vo.shape
    base = 80 if "keepraw" in mode else 0
    if base:
        assert np.array_equal(vg[:, :80], tab["volt"])              # raw volts, :165
    else:
        assert np.array_equal(vg[:, 64:], vo[:, 64:])               # centred FC channels, bit-exact
    # diode channels: 1e-9 where the trajectories coincide (float32 storage: <= 1 ulp),
    # solver tolerance otherwise
    keys = [k for k in ho if "SIN AMPLITUDE" in k]
    ncoin = 0
    for k in keys:
        side, tel, dio = k.split()[-3:]
        ch = gp.idx(gp.Side[side], int(tel[1]), gp.Diode[dio]) - 1
        kphi = k.replace("SIN AMPLITUDE", "SIN PHASE")
        same = (abs(hg[k] - ho[k]) <= REL_FIT * abs(ho[k]) and
                abs(fitref.dphi(hg[kphi], ho[kphi])) <= REL_FIT * max(1.0, abs(ho[kphi])))
        cols = slice(base + 2 * ch, base + 2 * ch + 2)
        scale = np.abs(vo[:, cols]).max()
        err = np.abs(vg[:, cols].astype(np.float64) - vo[:, cols]).max() / scale
        if same:
            ncoin += 1
            assert err <= 2.0 ** -22, (k, err)
            for kk in ("AMPLITUDE ABS", "AMPLITUDE ARG"):
                name = k.replace("SIN AMPLITUDE", kk)
                assert abs(hg[name] - ho[name]) <= REL_FIT * max(1.0, abs(ho[name]))
        else:
            assert abs(hg[k] - ho[k]) <= FORK_HARD and abs(fitref.dphi(hg[kphi], ho[kphi])) <= FORK_HARD
            assert err <= 4 * FORK_HARD
    assert ncoin >= MIN_COINCIDE
    if faint:
        assert "STATE" not in tg   # whole-file mode writes no STATE column (:248 is window mode)


@pytest.mark.parametrize("method", METHODS)
def test_table_windowed_faint(gp, ora, method):
    tab = make_case(gp.synthetic, 4100, k=23, faint=True, ora=ora)
    tg, hg, to, ho = _table_compare(gp, ora, tab, False, 2.0, False, True, method=method)
    assert np.array_equal(tg["STATE"], to["STATE"]) and tg["STATE"].dtype == np.int8   # bit-exact
    for k in ("X0", "Y0", "ABSA", "ARGA", "B", "PHI"):
        assert tg[k].shape == to[k].shape == (4100, 32) and tg[k].dtype == np.float32
    # window partition is bit-exact: values change exactly at multiples of nwindow = 1000
    wrows, nwin = gp.table_windows(tab["time_us"], tab["mjd"], 2.0)
    assert (wrows, nwin) == (1000, 5)
    chg = np.nonzero(np.any(np.diff(tg["B"], axis=0) != 0, axis=1))[0] + 1
    assert set(chg) <= {1000, 2000, 3000, 4000}
    # the four full windows (1000 rows = two modulation periods): float32 columns, so
    # "coincide" is float32 equality of B and PHI; forks within the measured envelope
    full = slice(0, 4000, 1000)
    db = np.abs(tg["B"][full].astype(np.float64) - to["B"][full])
    dp = np.abs(fitref.dphi(tg["PHI"][full].astype(np.float64), to["PHI"][full]))
    close = (db <= 2.0 ** -22 * np.abs(to["B"][full])) & (dp <= 2.0 ** -22 * np.maximum(1.0, np.abs(to["PHI"][full])))
    assert close.mean() >= fitref.MIN_COINCIDE_FRACTION, close.mean()
    assert db.max() <= FORK_HARD and dp.max() <= FORK_HARD
    # the ragged last window has 100 rows = a fifth of a modulation period: (b, phi) are
    # not determined by the data there (the oracle against itself perturbed by 1e-15 moves
    # b by up to 0.17 and chi2 by 1 %), so only a loose bound applies
    assert np.abs(tg["B"][4000:] - to["B"][4000:]).max() <= 0.5
    assert np.array_equal(tg["VOLT"][:, 64:], to["VOLT"][:, 64:])


def test_big_endian_table_bytes(gp, ora, method="auto"):
    """Raw FITS byte order in, raw FITS byte order out (GPPD_BIG_ENDIAN)."""
    import ctypes as C
    tab = make_case(gp.synthetic, 1500, k=4)
    off = gp.synthetic.stefan_centres()
    v0, p0, c0, _, _ = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off,
                                        method=method)
    L, h = gp._lib.lib(), gp.default_handle()
    tu_be = tab["time_us"].astype(">i4")
    v_be = tab["volt"].astype(">f4")
    o = gp.api._options(method=method)
    o.flags |= gp._lib.BIG_ENDIAN
    vout = np.empty((1500, 80), dtype=">f4")
    par, chi2 = np.empty((32, 6)), np.empty(32)
    gp._lib.check(L.gppd_process_table_f32(
        h.raw, 1500, tu_be.ctypes.data_as(gp._lib._i32p), tab["mjd"],
        v_be.ctypes.data_as(gp._lib._fp), off.view(np.float64).ctypes.data_as(gp._lib._dp),
        None, 0, None, 0, 0.0, C.byref(o), vout.ctypes.data_as(gp._lib._fp),
        par.ctypes.data_as(gp._lib._dp), chi2.ctypes.data_as(gp._lib._dp), None, None))
    assert np.array_equal(vout.astype(np.float32), v0) and par.tobytes() == p0.tobytes()


def test_big_endian_faint_table_bytes(gp, ora):
    """The same for a FAINT table: segmentation, the statistics pass (its big-endian bulk
    kernel) and the weighted sums all read raw FITS byte order."""
    import ctypes as C
    n = 5003
    tab = make_case(gp.synthetic, n, k=6, faint=True, ora=ora)
    off = gp.synthetic.stefan_centres()
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0)
    v0, p0, c0, _, st0 = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, faintparam=fs_g)
    L, h = gp._lib.lib(), gp.default_handle()
    tu_be = tab["time_us"].astype(">i4")
    v_be = tab["volt"].astype(">f4")
    o = gp.api._options()
    o.flags |= gp._lib.BIG_ENDIAN
    vout = np.empty((n, 80), dtype=">f4")
    par, chi2, st = np.empty((32, 6)), np.empty(32), np.empty(n, np.int8)
    t1, t2 = np.ascontiguousarray(fs_g.timer1), np.ascontiguousarray(fs_g.timer2)
    gp._lib.check(L.gppd_process_table_f32(
        h.raw, n, tu_be.ctypes.data_as(gp._lib._i32p), tab["mjd"],
        v_be.ctypes.data_as(gp._lib._fp), off.view(np.float64).ctypes.data_as(gp._lib._dp),
        t1.ctypes.data_as(gp._lib._dp), t1.size, t2.ctypes.data_as(gp._lib._dp), t2.size, 0.0, C.byref(o),
        vout.ctypes.data_as(gp._lib._fp), par.ctypes.data_as(gp._lib._dp), chi2.ctypes.data_as(gp._lib._dp),
        None, st.ctypes.data_as(gp._lib._i8p)))
    assert np.array_equal(st, st0) and np.array_equal(st, tab["state"])
    assert np.array_equal(vout.astype(np.float32), v0) and par.tobytes() == p0.tobytes()


@pytest.mark.parametrize("n,wrows", [(30011, 0), (30011, 5003), (2100, 100), (9000, 1024)])
def test_statistics_kernels_agree(gp, ora, monkeypatch, n, wrows):
    """The statistics pass has two kernels (dense tables: warp-private bulk pipelines; other
    layouts: plain loads).  They add the same terms in a different order: the per-state
    (mean, weight) tables, seen through the weighted harmonic sums, agree to rounding."""
    from gppd_b200 import _lib
    tab = make_case(gp.synthetic, n, k=8, faint=True, ora=ora)
    off = gp.synthetic.stefan_centres()
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0)
    kw = {}
    nwin = 1
    if wrows:
        dt = float(np.diff(ora.make_times(tab["time_us"][:2], tab["mjd"]))[0])
        kw["window"] = wrows * dt
        nwin = -(-n // wrows)
    out = {}
    for plain in ("", "1"):
        if plain:
            monkeypatch.setenv("GPPD_STATS_PLAIN", "1")
        r = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, faintparam=fs_g, **kw)
        H = np.empty(103 * 32 * nwin)
        _lib.check(_lib.lib().gppd_debug_harmonics(_lib.default_handle().raw, 0, _lib.ptr(H), H.size))
        out[plain] = (H.reshape(103, -1), r)
    a, b = out[""][0], out["1"][0]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    scale = np.nanmax(np.abs(a), axis=0)
    scale[~(scale > 0)] = 1.0
    err = np.abs(np.where(ok, a - b, 0.0)) / scale
    assert err.max() <= 1e-11, err.max()
    assert np.array_equal(out[""][1][4], out["1"][1][4])              # STATE
    pa, pb = out[""][1][1], out["1"][1][1]
    fin = np.isfinite(pa[:, 4]) & np.isfinite(pb[:, 4])
    assert np.array_equal(np.isfinite(pa[:, 4]), np.isfinite(pb[:, 4]))
    if wrows == 0 or wrows >= 1000:      # (a 100-row window is a fifth of a modulation period: its
        # minimum is flat and the end point of NEWUOA follows the last bits of the sums)
        assert np.abs(pa[fin, 4:6] - pb[fin, 4:6]).max() <= 5 * FORK_HARD


@pytest.mark.parametrize("method", METHODS)
def test_edge_cases(gp, ora, method):
    # minimum size, NaN propagation like the reference (a state with one sample
    # has var = NaN -> weight NaN), no crash, no hang (maxfun bounds the solver)
    rng = np.random.default_rng(0)
    t = 5.2e9 + 0.002 * np.arange(4)
    z = rng.normal(size=(4, 40)) + 1j * rng.normal(size=(4, 40))
    out, par, like, info = gp.demodulateall(t, z, raw=True, return_info=True, method=method)
    assert np.all(info[:, 0] <= 8 + 60 + 3 + 60) and out.shape == (4, 40)
    st = np.array([ora.HIGH, ora.LOW, ora.LOW, ora.LOW], dtype=np.int8)
    out, par, like, info = gp.demodulateall(t, z, faintparam=st, raw=True, return_info=True,
                                            method=method)
    oo, op, ol = ora.demodulateall(t, z, faintparam=st)
    assert np.all(np.isnan(like)) and np.all(np.isnan(ol))
    with pytest.raises(ValueError):
        gp.demodulateall(t, z[:3])
    with pytest.raises(gp.GppdError):
        gp.demodulateall(t[:1], z[:1])


@pytest.mark.parametrize("method", METHODS)
def test_full_size_properties(gp, ora, method, monkeypatch):
    """BASELINE config sizes (1e5 rows): size-independent properties instead of
    an oracle run -- rotation preserves |d - c|, FC pass-through, b >= 0,
    recovery of the generating parameters, repeatability."""
    monkeypatch.setenv("GPPD_HARMONICS", "dmma")   # one kernel for both layouts: bit-identical fits
    tab = make_case(gp.synthetic, 100_000, k=1)
    off = gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    out, par, like, info = gp.demodulateall(t, z, raw=True, return_info=True, method=method)
    assert np.allclose(np.abs(out[:, :32]), np.abs(z[:, :32]), rtol=1e-13, atol=1e-16)
    assert np.array_equal(out[:, 32:], z[:, 32:]) and np.all(par[:, 4] >= 0)
    tr = tab["truth"]
    assert np.abs(par[:, 4] - tr["b"]).max() < 1e-3
    assert np.abs(np.angle(np.exp(1j * (par[:, 5] - tr["phi"])))).max() < 1e-3
    assert np.all(info[:, 0] <= 131)
    vout, p2, c2, i2, _ = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off,
                                           method=method)
    assert p2.tobytes() == par.tobytes()          # both boundaries run the same fit
    ref32 = np.empty((100_000, 80), np.float32)
    ref32[:, 0::2], ref32[:, 1::2] = out.real, out.imag
    # the table path evaluates exp(-j psi) to 1e-10 (enough for its float32 output), the array
    # path to full double precision: the float32 values agree except where the double result
    # sits within 1e-10 of a rounding boundary, and then by one unit in the last place
    assert np.array_equal(vout[:, 64:], ref32[:, 64:])
    diff = np.abs(vout[:, :64].astype(np.float64) - ref32[:, :64])
    assert diff.max() <= 2.0 ** -23 * np.abs(ref32[:, :64]).max()
    assert (vout[:, :64] != ref32[:, :64]).mean() < 0.02
    # default dispatch: the table's harmonic sums come from the int8 tensor-core kernel
    # (sums equal to ~1e-15, so a NEWUOA trajectory can fork at a rounding-level tie)
    monkeypatch.delenv("GPPD_HARMONICS")
    vout3, p3, c3, i3, _ = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off,
                                            method=method)
    same = np.abs(p3[:, 4] - par[:, 4]) <= REL_FIT * np.abs(par[:, 4])
    assert same.sum() >= MIN_COINCIDE and np.abs(p3[:, 4:6] - par[:, 4:6]).max() <= FORK_HARD
    assert np.abs(c3 - like).max() <= 1e-6 * np.abs(like).max()
    cols = np.repeat(same, 2)
    assert np.abs(vout3[:, :64][:, cols] - ref32[:, :64][:, cols]).max() <= 2.0 ** -22 * np.abs(ref32[:, :64]).max()


@pytest.mark.parametrize("n,window", [(5, None), (63, None), (257, None), (1237, 0.35), (2051, 1.0),
                                      (2500, None)])
@pytest.mark.parametrize("keepraw", [False, True])
def test_ragged_sizes_table(gp, ora, n, window, keepraw):
    """Row counts / windows that are not multiples of any tile size (4-row MMA
    k-steps, 32-row warps, 64-row demod tiles, 256-row harmonic tiles, 2048-row
    statistics segments): objective-level agreement with the oracle for every window."""
    tab = make_case(gp.synthetic, n, k=31, faint=False)
    off = gp.synthetic.stefan_centres()
    tg, hg, to, ho = _table_compare(gp, ora, tab, off, window, keepraw, False)
    vg, vo = tg["VOLT"], to["VOLT"]
    assert vg.shape == vo.shape == (n, 144 if keepraw else 80)
    if keepraw:
        assert np.array_equal(vg[:, :80], tab["volt"])
    else:
        assert np.array_equal(vg[:, 64:], vo[:, 64:])
    # |out| is invariant under the demodulation whatever trajectory the solver took
    base = 80 if keepraw else 0
    ag = np.hypot(vg[:, base:base + 64:2].astype(np.float64), vg[:, base + 1:base + 64:2])
    ao = np.hypot(vo[:, base:base + 64:2].astype(np.float64), vo[:, base + 1:base + 64:2])
    assert np.allclose(ag, ao, rtol=3e-7, atol=1e-9)
    if window is None:
        # (tables much shorter than a modulation period leave (b, phi) undetermined: the
        # objective is flat and trajectories fork, so parameters are not compared there)
        assert set(hg) == set(ho)
        if n >= 2000:
            nb = sum(abs(hg[k] - ho[k]) <= REL_FIT * abs(ho[k]) for k in ho if "SIN AMPLITUDE" in k)
            assert nb >= MIN_COINCIDE
    else:
        assert tg["B"].shape == to["B"].shape == (n, 32)
        wrows, nwin = gp.table_windows(tab["time_us"], tab["mjd"], window)
        chg = np.nonzero(np.any(np.diff(tg["B"], axis=0) != 0, axis=1))[0] + 1
        assert set(chg) <= set(range(wrows, n, wrows))


def test_fallback_queue_large_b(gp, ora):
    """A start vector with |b| > 5 is outside the harmonic evaluator's range: those
    fits go through the fallback queue to the direct evaluator (info[:, 2] == 1) and
    agree with a direct-only run bit for bit."""
    tab = make_case(gp.synthetic, 3000, k=5)
    off = gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    r_auto = gp.demodulateall(t, z, init=[7.0, 0.3], raw=True, return_info=True, method="auto")
    r_dir = gp.demodulateall(t, z, init=[7.0, 0.3], raw=True, return_info=True, method="direct")
    assert np.all(r_auto[3][:, 2] == 1)          # every fit fell back
    assert r_auto[1].tobytes() == r_dir[1].tobytes() and r_auto[2].tobytes() == r_dir[2].tobytes()
    assert np.array_equal(r_auto[0], r_dir[0])


def test_many_small_windows(gp, ora):
    """More (job, group) pairs than a CUDA grid's y dimension holds (65 535): 9 000
    windows of 12 rows; every window equals a separate call on its rows."""
    n, w = 108_000, 12
    tab = make_case(gp.synthetic, n, k=7)
    off = gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    out, par, like, info = gp.demodulateall(t, z, raw=True, nwindow=w, return_info=True)
    assert par.shape == (9000 * 32, 6) and np.all(info[:, 0] > 0)
    assert np.allclose(np.abs(out[:, :32]), np.abs(z[:, :32]), rtol=1e-12, atol=1e-15)
    for win in (0, 4321, 8999):
        lo = win * w
        o1, p1, l1 = gp.demodulateall(t[lo:lo + w], z[lo:lo + w], raw=True)
        assert p1.tobytes() == par[win * 32:(win + 1) * 32].tobytes()
        assert np.array_equal(o1, out[lo:lo + w])


# ---------------------------------------------------------------------------
# BASELINE.json configs 1 and 2 at their full size against the oracle
FULL = {"bright_stefan": (False, False, 41), "bright_fit": (False, True, 42),
        "faint_stefan": (True, False, 43), "faint_fit": (True, True, 44)}


@pytest.mark.parametrize("mode", list(FULL))
def test_full_size_vs_oracle(gp, ora, mode):
    """1e5-row tables (BASELINE configs 1 / 2; `-c stefan` = subtract the centres, `-c fit`
    = fit them): the oracle's demodulateall (reference src/Modulation.jl:344-435) against
    BOTH boundaries of the library -- gppd_demodulate_f64 (harmonic sums on the FP64 units)
    and the METROLOGY-table path (int8 tensor cores, int32 accumulation over 17 segments,
    sampled fixed-point scale, float32 re-interleave)."""
    faint, fit, k = FULL[mode]
    n = 100_000
    tab = make_case(gp.synthetic, n, k=k, faint=faint, jitter=True, ora=ora)
    off = None if fit else gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    state = tab["state"]                     # CLI path: buildstates with zero delays (:144)
    oo, op, ol, onf = ora.demodulateall(t, z, faintparam=state, fitoffsets=fit, nthreads=8,
                                        return_nfev=True)
    scale = np.abs(z[:, :32]).max(axis=0)

    # ---- the demodulateall boundary
    out, par, like, info, trace = gp.demodulateall(t, z, faintparam=state, fitoffsets=fit, raw=True,
                                                   return_info=True, return_trace=True)
    assert np.all(info[:, 2] == 2)           # harmonic evaluator, no fallback
    for ch in (3, 20):                       # objective parity along the GPU's own trace
        obj = _oracle_objective(ora, t, z, state, ch, False, fit)
        for kk in range(info[ch, 0]):
            b, phi, f = trace[ch, kk]
            fo = obj(b, phi)[0]
            assert abs(f - fo) <= REL_OBJ * fo, (ch, kk, f, fo)
    coincide, stats = fitref.compare_fits(par, like, op, ol, info[:, 0], onf)
    FORK_STATS.append(stats)
    print("%s array path: forks %d / 32, max |db| %.1e |dphi| %.1e dchi2 %.1e"
          % (mode, stats["forks"], stats["max_db"], stats["max_dphi"], stats["max_dchi2"]))
    assert coincide.sum() >= MIN_COINCIDE
    a, ao = par[:, 2] + 1j * par[:, 3], op[:, 2] + 1j * op[:, 3]
    assert (np.abs(a - ao)[coincide] <= REL_FIT * np.abs(ao)[coincide]).all()
    if fit:
        cc, co = par[:, 0] + 1j * par[:, 1], op[:, 0] + 1j * op[:, 1]
        assert (np.abs(cc - co)[coincide] <= REL_FIT * np.maximum(np.abs(co), np.abs(ao))[coincide]).all()
    assert (np.abs(like - ol)[coincide] <= REL_FIT * ol[coincide]).all()
    err = np.abs(out[:, :32] - oo[:, :32]).max(axis=0) / scale
    assert (err[coincide] <= REL_FIT).all() and err.max() <= 4 * FORK_HARD
    assert np.array_equal(out[:, 32:], z[:, 32:])

    # ---- the METROLOGY-table boundary
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    vout, p2, c2, i2, st2 = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off,
                                             faintparam=fs_g)
    if faint:
        assert np.array_equal(st2, state)    # segmentation: bit-exact
    assert np.all(i2[:, 2] == 2)
    coincide2, stats2 = fitref.compare_fits(p2, c2, op, ol, i2[:, 0], onf)
    FORK_STATS.append(stats2)
    print("%s table path: forks %d / 32, max |db| %.1e |dphi| %.1e dchi2 %.1e"
          % (mode, stats2["forks"], stats2["max_db"], stats2["max_dphi"], stats2["max_dchi2"]))
    assert coincide2.sum() >= MIN_COINCIDE
    a2 = p2[:, 2] + 1j * p2[:, 3]
    assert (np.abs(a2 - ao)[coincide2] <= REL_FIT * np.abs(ao)[coincide2]).all()
    assert (np.abs(c2 - ol)[coincide2] <= REL_FIT * ol[coincide2]).all()
    ref32 = np.empty((n, 80), np.float32)    # :170-171, :253
    ref32[:, 0::2], ref32[:, 1::2] = oo.real, oo.imag
    assert np.array_equal(vout[:, 64:], ref32[:, 64:])         # centred FC channels
    d = np.abs(vout[:, :64].astype(np.float64) - ref32[:, :64]).reshape(n, 32, 2).max(axis=(0, 2)) / scale
    assert (d[coincide2] <= 2.0 ** -22).all(), d[coincide2].max()
    assert d.max() <= 4 * FORK_HARD


def test_fork_population():
    """All end-to-end comparisons of this module together: the fraction of fits on the
    oracle's trajectory and the size of the forks are those of the reference procedure's
    own ambiguity (tests/test_fork_envelope.py), not of a defect."""
    if not FORK_STATS:
        pytest.skip("runs after the end-to-end comparisons of this module")
    forks, nfits = fitref.check_fork_population(FORK_STATS, "test_gpu_parity")
    print("module total: %d forks of %d fits (%.1f %%)" % (forks, nfits, 100.0 * forks / nfits))
