"""Known-answer tests of the CPU oracle against everything the reference tree
pins: channel indexing, VOLT interleave, Stefan centres,
cost-function symmetry, closed-form (c, a), rho schedule, segmentation hand
cases.  The reference ships no tests, so these are the only pins that exist
(PARITY UNPINNED for the fit itself)."""
import numpy as np
import pytest

from conftest import make_case


def test_idx_table(ora):
    # src/Modulation.jl:9-10,17-22
    k = 1
    for side in (ora.FT, ora.SC):
        for tel in range(1, 5):
            for dio in range(1, 5):
                assert ora.idx(side, tel, dio) == k
                k += 1
    assert [ora.idx(ora.FT, t, ora.FC) for t in range(1, 5)] == [33, 34, 35, 36]
    assert [ora.idx(ora.SC, t, ora.FC) for t in range(1, 5)] == [37, 38, 39, 40]


def test_stefan_centres(gp):
    # src/GPPupilDemodulation.jl:88-104: 1e-3*(VX + j VY) of the `avg` rows
    off = gp.synthetic.stefan_centres()
    assert off.shape == (40,)
    assert off[32] == 1e-3 * (0.5056591585058203 + 1j * 0.6414248583206548)   # FTT1FC -> 33
    assert off[0] == 1e-3 * (-3.9217663072871978 + 1j * -5.375368693370768)   # FTT1D1 -> 1
    assert np.all(off != 0)


def test_phirange(ora):
    # range(-pi, pi, 8), src/Modulation.jl:360
    p = ora.phirange()
    assert p[0] == -np.pi and p[7] == np.pi
    assert np.allclose(np.diff(p), 2 * np.pi / 7, rtol=0, atol=1e-15)
    assert np.all(np.abs(p + p[::-1]) <= 4.5e-16)


def test_times_worked_example(ora):
    # TIME=0,2000,.. us, MJD 59949.0
    t = ora.make_times(np.arange(0, 6000, 2000, dtype=np.int32), 59949.0)
    assert t[0] == 5179593600.0
    assert t[1] - t[0] == 0.0019998550415039062
    assert int(np.round(100.0 / (t[1] - t[0]))) == 50004


def _one_diode(ora, gp, n=3000, k=1, faint=False):
    tab = make_case(gp.synthetic, n, k=k, faint=faint, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    ch, fcch = ora.idx(ora.SC, 2, 3) - 1, ora.idx(ora.SC, 2, ora.FC) - 1
    fc = np.exp(1j * np.angle(z[:, fcch]))
    return t, z[:, ch], fc, tab


def test_cost_symmetry(ora, gp):
    # f(b, phi) = f(-b, phi + pi), tex/GPPupilDemodulation.tex:189; with the
    # reference's absolute times the identity holds to the phase quantum 2^-18
    t, d, fc, _ = _one_diode(ora, gp)
    for b, phi in ((0.7, 0.3), (1.9, -2.0), (0.1, 3.0)):
        f1 = ora.chi2(t, d, fc, b, phi)[0]
        f2 = ora.chi2(t, d, fc, -b, phi + np.pi)[0]
        assert abs(f1 - f2) <= 2e-5 * f1
    # exactly, with small time stamps (no quantisation of phi)
    trel = t - t[0]
    f1 = ora.chi2(trel, d, fc, 0.7, 0.3)[0]
    f2 = ora.chi2(trel, d, fc, -0.7, 0.3 + np.pi)[0]
    assert abs(f1 - f2) <= 1e-12 * f1


@pytest.mark.parametrize("fitoffsets", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_closed_form_linear(ora, gp, fitoffsets, weighted):
    # (c+, a+) of src/Modulation.jl:174-215 / :143-145 against a dense weighted
    # least-squares solve of the same model
    t, d, fc, _ = _one_diode(ora, gp, n=800)
    rng = np.random.default_rng(3)
    w = rng.uniform(0.5, 2.0, t.size) if weighted else None
    pw = rng.uniform(0.5, 3.0, t.size) if weighted else None
    b, phi = 1.3, -0.4
    f, c, a = ora.chi2(t, d, fc, b, phi, weight=w, power=pw, fitoffsets=fitoffsets)
    g = (1.0 if pw is None else pw) * fc * np.exp(1j * b * np.sin(ora.M_2PI * t + phi))
    sw = np.sqrt(np.ones(t.size) if w is None else w)
    A = np.stack([np.ones_like(g), g], axis=1) if fitoffsets else g[:, None]
    sol = np.linalg.lstsq(A * sw[:, None], d * sw, rcond=None)[0]
    if fitoffsets:
        assert abs(c - sol[0]) <= 1e-10 * abs(sol[0]) + 1e-13
        assert abs(a - sol[1]) <= 1e-10 * abs(sol[1])
    else:
        assert c == 0 and abs(a - sol[0]) <= 1e-11 * abs(sol[0])
    res = A @ sol - d
    assert abs(f - np.sum(sw ** 2 * np.abs(res) ** 2) / t.size) <= 1e-9 * f


def test_rho_schedule_and_defaults(ora):
    # rhobeg=1, rhoend=1e-3 -> stages 1, 0.1, 0.01, 0.001 (src/Modulation.jl:335 +
    # NEWUOA's reduction rule); npt=2n+1, maxeval=30n
    rhos = []
    f = lambda x: (x[0] - 0.3) ** 2 + 2 * (x[1] + 0.2) ** 2 + 0.1 * np.sin(3 * x[0])
    st, x, fx, nf = ora.newuoa(f, [1.0, 1.0], probe=lambda s: rhos.append(s["rho"]))
    assert st == 0 and nf <= 60
    uniq = sorted(set(rhos), reverse=True)
    assert np.allclose(uniq, [1.0, 0.1, 0.01, 0.001][:len(uniq)], rtol=1e-12)
    assert np.allclose(uniq[-1], 1e-3, rtol=1e-12)
    # first five trial points: x0, x0 + e1, x0 + e2, x0 - e1, x0 - e2
    rec = []
    ora.newuoa(f, [0.1, -0.5], record=rec)
    pts = np.array([r[0] for r in rec[:5]])
    assert np.array_equal(pts, np.array([[0.1, -0.5], [1.1, -0.5], [0.1, 0.5], [-0.9, -0.5], [0.1, -1.5]]))
    # the evaluation cap is part of the result (check=false)
    st, x, fx, nf = ora.newuoa(lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2,
                               [-1.2, 1.0], rhoend=1e-9, maxfun=60)
    assert st == -3 and nf == 60


def test_newuoa_invariants(ora):
    """After every update the factored matrix H (BMAT, ZMAT, IDZ) is the inverse
    of the KKT matrix of the current interpolation set and the model
    interpolates all FVAL: pins the algebra of UPDATE/BIGLAG/BIGDEN/shift."""
    rng = np.random.default_rng(1)

    def probe_factory(errs):
        def probe(s):
            n, m = s["n"], s["npt"]
            X = s["xpt"]
            W = np.zeros((m + n + 1, m + n + 1))
            W[:m, :m] = 0.5 * (X @ X.T) ** 2
            W[:m, m] = 1; W[m, :m] = 1; W[:m, m + 1:] = X; W[m + 1:, :m] = X.T
            Z = s["zmat"]; S = np.ones(Z.shape[1]); S[:s["idz"] - 1] = -1
            B = s["bmat"]
            H = np.zeros((m + n, m + n))
            H[:m, :m] = (Z * S) @ Z.T; H[:m, m:] = B[:m]; H[m:, :m] = B[:m].T; H[m:, m:] = B[m:]
            ii = list(range(m)) + list(range(m + 1, m + n + 1))
            Hs = np.linalg.inv(W)[np.ix_(ii, ii)]
            e1 = np.abs(H - Hs).max() / max(1.0, np.abs(Hs).max())
            G = np.zeros((n, n)); ih = 0
            for j in range(n):
                for i in range(j + 1):
                    G[i, j] = G[j, i] = s["hq"][ih]; ih += 1
            G = G + (X.T * s["pq"]) @ X
            q = X @ s["gq"] + 0.5 * np.einsum("ki,ij,kj->k", X, G, X)
            k0 = s["kopt"] - 1
            fv = s["fval"]
            e2 = np.abs((q - q[k0]) - (fv - fv[k0])).max() / max(1e-300, np.abs(fv - fv.min()).max())
            errs.append((e1, e2))
        return probe

    for force in (False, True):
        ora.newuoa_force_bigden(force)
        try:
            for n in (2, 3, 5):
                A = rng.normal(size=(n, n)); Q = A @ A.T + 0.1 * np.eye(n); c = rng.normal(size=n)
                f = lambda x: 0.5 * (x - c) @ Q @ (x - c) + 0.05 * np.sum((x - c) ** 4)
                for npt in sorted({2 * n + 1, n + 2, (n + 1) * (n + 2) // 2}):
                    errs = []
                    ora.newuoa_counters()
                    st, x, fx, nf = ora.newuoa(f, rng.normal(size=n), rhoend=1e-6, npt=npt,
                                               maxfun=3000, probe=probe_factory(errs))
                    cnt = ora.newuoa_counters()
                    assert st == 0 and np.abs(x - c).max() < 1e-4
                    assert np.array(errs).max() < 1e-6
                    if force:
                        assert cnt[2] == cnt[1] > 0   # BIGDEN taken on every model step
        finally:
            ora.newuoa_force_bigden(False)


def _fs(ora, t1, t2, v1=1.0, v2=2.0):
    return ora.FaintStates(np.array(t1, float), np.array(t2, float), v1, v2)


def test_buildstates_hand_cases(ora):
    # src/Faint.jl:21-73 by hand: dt = 1
    t = np.arange(10.0)
    H, L, N, T = ora.HIGH, ora.LOW, ora.NORMAL, ora.TRANSIENT
    # HIGH at 2 (queue 1 now empty -> sentinel first1 = last(t)); LOW at 5 finds
    # first1 == last(t) and falls back to NORMAL (:58-62); the last row fires
    # both sentinels and ends NORMAL
    s = ora.buildstates(_fs(ora, [2.0], [5.0]), t)
    assert s.tolist() == [N, N, H, H, H, N, N, N, N, N]
    # a later HIGH event keeps queue 1 alive, so LOW at 5 sticks; the HIGH event
    # at 7 empties queue 1 while queue 2 is at its sentinel -> NORMAL
    s = ora.buildstates(_fs(ora, [2.0, 7.0], [5.0]), t)
    assert s.tolist() == [N, N, H, H, H, L, L, N, N, N]
    # constructor swap: voltage1 > voltage2 exchanges the timer series
    s2 = ora.buildstates(_fs(ora, [5.0], [2.0, 7.0], v1=3.0, v2=1.0), t)
    assert s2.tolist() == s.tolist()
    # delays: premax = ceil(1.5/1) = 2 rows TRANSIENT after HIGH, postmax = 1 after LOW;
    # the sentinels re-arm `forget` on the last row
    s = ora.buildstates(_fs(ora, [2.0], [5.0]), t, preswitchdelay=1.5, postwitchdelay=0.5)
    assert s.tolist() == [N, N, T, T, H, T, N, N, N, T]
    # at most one event per queue per row: three HIGH events inside one sample step
    s = ora.buildstates(_fs(ora, [2.0, 2.1, 2.2], [100.0]), t, preswitchdelay=1.0)
    assert s.tolist() == [N, N, T, T, T, H, H, H, H, T]
    # lag shifts the timers by lag * timestep
    s = ora.buildstates(_fs(ora, [2.0], [5.0]), t, lag=2)
    assert s.tolist() == [N, N, N, N, H, H, H, N, N, N]


def test_buildstates_c_vs_python(ora):
    rng = np.random.default_rng(7)
    for trial in range(60):
        n = int(rng.integers(2, 200))
        dt = rng.choice([1.0, 0.002, 0.5])
        t = np.cumsum(rng.choice([dt, dt, dt, 0.0], size=n)) + rng.choice([0.0, 5.2e9])
        n1, n2 = int(rng.integers(1, 8)), int(rng.integers(1, 8))
        span = t[-1] - t[0] + 2 * dt
        t1 = t[0] - dt + np.sort(rng.uniform(0, span, n1))
        t2 = t[0] - dt + np.sort(rng.uniform(0, span, n2))
        if trial % 5 == 0:
            t1[-1] = t[-1]           # a real event equal to the end-of-queue sentinel
        if t[1] - t[0] <= 0:
            t[1] = t[0] + dt
            t = np.maximum.accumulate(t)
        fs = _fs(ora, t1, t2)
        pre, post = rng.choice([0.0, 0.5 * dt, 3 * dt]), rng.choice([0.0, 2.2 * dt])
        lag = int(rng.integers(-2, 3))
        a = ora.buildstates(fs, t, lag=lag, preswitchdelay=pre, postwitchdelay=post)
        b = ora.buildstates_py(fs, t, lag=lag, preswitchdelay=pre, postwitchdelay=post)
        assert np.array_equal(a, b)


def test_mean_var_power(ora):
    # src/Faint.jl:89-100: mean |d| and 1/var(|d|) with the n-1 divisor, per state
    rng = np.random.default_rng(2)
    st = rng.choice([ora.LOW, ora.NORMAL, ora.HIGH], size=500).astype(np.int8)
    d = rng.normal(size=500) + 1j * rng.normal(size=500)
    m, w = ora.compute_mean_var_power(st, d)
    for s in (ora.LOW, ora.NORMAL, ora.HIGH):
        sel = st == s
        assert np.allclose(m[sel], np.abs(d[sel]).mean(), rtol=1e-13)
        assert np.allclose(w[sel], 1 / np.abs(d[sel]).var(ddof=1), rtol=1e-12)
    # a state with one sample: var = NaN -> weight NaN, as in Julia
    st2 = np.array([ora.HIGH, ora.LOW, ora.LOW], dtype=np.int8)
    m, w = ora.compute_mean_var_power(st2, d[:3])
    assert np.isnan(w[0]) and m[0] == abs(d[0])


@pytest.mark.parametrize("fitoffsets", [False, True])
def test_synthetic_recovery_bright(ora, gp, fitoffsets):
    tab = make_case(gp.synthetic, 6000, k=2, noise=0.005)
    t, z = gp.synthetic.to_complex(tab, None if fitoffsets else gp.synthetic.stefan_centres())
    out, par, like, nfev = ora.demodulateall(t, z, fitoffsets=fitoffsets, nthreads=8, return_nfev=True)
    tr = tab["truth"]
    assert np.all(nfev <= 8 + 60 + 3 + 60)
    assert np.abs(par[:, 4] - tr["b"]).max() < 2e-3
    assert np.abs(np.angle(np.exp(1j * (par[:, 5] - tr["phi"])))).max() < 2e-3
    a = par[:, 2] + 1j * par[:, 3]
    assert (np.abs(a - tr["a"]) / np.abs(tr["a"])).max() < 5e-3
    if fitoffsets:
        assert np.abs((par[:, 0] + 1j * par[:, 1]) - tr["c"]).max() < 2e-3
    # b >= 0 after the sign normalisation (src/Modulation.jl:427-430)
    assert np.all(par[:, 4] >= 0)
    # FC columns are returned unchanged (output = copy(data), :353)
    assert np.array_equal(out[:, 32:], z[:, 32:])
    # demodulation is a pure rotation of (d - c)
    c = (par[:, 0] + 1j * par[:, 1]) if fitoffsets else 0
    assert np.allclose(np.abs(out[:, :32]), np.abs(z[:, :32] - c), rtol=1e-12, atol=1e-15)
    # the demodulated signal has lost the modulation: residual spread ~ noise
    dem = out[:, :32] * np.exp(-1j * np.angle(z[:, 32 + np.arange(32) // 4]))
    spread = np.std(dem, axis=0) / np.abs(tr["a"])
    assert spread.max() < 0.02


def test_synthetic_recovery_faint(ora, gp):
    tab = make_case(gp.synthetic, 8000, k=4, faint=True, noise=0.005, ora=ora)
    t, z = gp.synthetic.to_complex(tab, gp.synthetic.stefan_centres())
    st = tab["state"]
    assert set(np.unique(st)) >= {ora.HIGH, ora.LOW, ora.NORMAL}
    out, par, like = ora.demodulateall(t, z, faintparam=st, nthreads=8)
    tr = tab["truth"]
    assert np.abs(par[:, 4] - tr["b"]).max() < 5e-3
    # the fitted amplitude is relative to the per-state mean power: |a| * mean|d| ~ |a_true| * P
    m_high = np.abs(z[st == ora.HIGH][:, :32]).mean(axis=0)
    assert np.allclose(np.abs(par[:, 2] + 1j * par[:, 3]) * m_high, 3.0 * np.abs(tr["a"]), rtol=0.05)
    # onlyhigh restricts the fit to HIGH / NORMAL rows
    out2, par2, like2 = ora.demodulateall(t, z, faintparam=st, onlyhigh=True, nthreads=8)
    assert np.abs(par2[:, 4] - tr["b"]).max() < 5e-3


def test_processmetrology_packing(ora, gp):
    # src/GPPupilDemodulation.jl:139-253: interleave, centre subtraction,
    # float32 repack, keepraw layout, window tables
    tab = make_case(gp.synthetic, 1500, k=5)
    off = gp.synthetic.stefan_centres()
    table, hdr = ora.processmetrology(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, nthreads=8)
    assert table["VOLT"].dtype == np.float32 and table["VOLT"].shape == (1500, 80)
    assert hdr["PROCSOFT"] == "GPPupilDemodulation.jl"
    assert "DEMODULATION SIN AMPLITUDE SC T4 D4" in hdr and "DEMODULATION CENTER X0 FT T1 D1" not in hdr
    # FC columns: float32(double(v) - centre)
    v = tab["volt"].astype(np.float64)
    assert np.array_equal(table["VOLT"][:, 64], (v[:, 64] - off[32].real).astype(np.float32))
    tk, _ = ora.processmetrology(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, keepraw=True, nthreads=8)
    assert tk["VOLT"].shape == (1500, 144)
    assert np.array_equal(tk["VOLT"][:, :80], tab["volt"])
    assert np.array_equal(tk["VOLT"][:, 80:], table["VOLT"][:, :64])
    tw, _ = ora.processmetrology(tab["time_us"], tab["volt"], tab["mjd"], offsets=False, window=1.0, nthreads=8)
    assert tw["B"].shape == (1500, 32) and tw["B"].dtype == np.float32 and "X0" in tw
    assert np.all(tw["B"][0] == tw["B"][499]) and np.any(tw["B"][0] != tw["B"][500])   # nwindow = 500
