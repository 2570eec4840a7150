"""The two implementations of the harmonic sums (csrc/harm_kernels.cu: FP64 DMMA;
csrc/harm_tc_kernels.cu: int8 tensor cores, 48-bit fixed point) held against each other
through the `gppd_debug_harmonics` hook, and the tensor kernel's overflow fallback."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu

SUM_TOL = 2e-14      # |tensor - dmma| / largest sum of the fit (48-bit operands, exact products)


def _htab(gp, monkeypatch, mode, tab, faint, offsets, nvals, **kw):
    from gppd_b200 import _lib
    monkeypatch.setenv("GPPD_HARMONICS", mode)
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    res = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=offsets, faintparam=fs_g, **kw)
    out = np.empty(nvals)
    _lib.check(_lib.lib().gppd_debug_harmonics(_lib.default_handle().raw, 0, _lib.ptr(out), nvals))
    return out.reshape(-1, 32), res


@pytest.mark.parametrize("n,faint,fit,onlyhigh", [(700, False, False, False), (5000, False, False, False),
                                                  (20011, True, False, False), (20011, True, False, True),
                                                  (30000, False, True, False), (30000, True, True, False)])
def test_tensor_sums_equal_dmma_sums(gp, ora, monkeypatch, n, faint, fit, onlyhigh):
    tab = make_case(gp.synthetic, n, k=7, faint=faint, ora=ora)
    off = None if fit else gp.synthetic.stefan_centres()
    nv = (201 if fit else 103) * 32
    a, ra = _htab(gp, monkeypatch, "dmma", tab, faint, off, nv, onlyhigh=onlyhigh)
    b, rb = _htab(gp, monkeypatch, "tensor", tab, faint, off, nv, onlyhigh=onlyhigh)
    assert not np.isnan(b).any()
    err = np.abs(a - b) / np.abs(a).max(axis=0)
    assert err.max() <= SUM_TOL, err.max()
    assert (rb[3][:, 2] == 2).all()                       # harmonic evaluator, no fallback
    assert np.abs(ra[1][:, 4:6] - rb[1][:, 4:6]).max() <= 2e-3


def test_tensor_is_run_to_run_and_batch_independent(gp, ora, monkeypatch):
    """Integer accumulation: the sums do not depend on the order of anything."""
    tab = make_case(gp.synthetic, 30011, k=9, faint=True, ora=ora)
    off = gp.synthetic.stefan_centres()
    a, ra = _htab(gp, monkeypatch, "tensor", tab, True, off, 103 * 32)
    b, rb = _htab(gp, monkeypatch, "tensor", tab, True, off, 103 * 32)
    assert a.tobytes() == b.tobytes() and ra[0].tobytes() == rb[0].tobytes()


def test_tensor_overflow_falls_back_to_direct(gp, ora, monkeypatch):
    """A sample far outside the range the block sampled would wrap the fixed point: the
    group's fits must go to the direct evaluator and still be right."""
    tab = make_case(gp.synthetic, 20000, k=11)
    off = gp.synthetic.stefan_centres()
    volt = tab["volt"].copy()
    row = 12345                                  # not one of the 128 sampled rows (multiples of n / 128)
    assert (row * 128) % 20000 != 0
    volt[row, 2 * 5] += 3000.0                   # diode 6 (group 2): |V| jumps by > 2^4 x the sampled maximum
    tab2 = dict(tab, volt=volt)
    a, ra = _htab(gp, monkeypatch, "dmma", tab2, False, off, 103 * 32)
    b, rb = _htab(gp, monkeypatch, "tensor", tab2, False, off, 103 * 32)
    grp = slice(4, 8)
    assert np.isnan(b[1, grp]).all()                                   # poisoned sums of group 2
    assert (rb[3][grp, 2] == 1).all() and (rb[3][:4, 2] == 2).all()    # direct evaluator for them only
    # the direct evaluator agrees with the harmonic one on the other kernel to solver tolerance
    assert np.abs(ra[1][grp, 4:6] - rb[1][grp, 4:6]).max() <= 2e-3
    assert np.abs(ra[2][grp] - rb[2][grp]).max() <= 1e-6 * np.abs(ra[2][grp]).max()


@pytest.mark.parametrize("wrows", [4100, 5003, 7001])
def test_tensor_faint_windows_with_rare_states(gp, ora, monkeypatch, wrows):
    """Window boundaries cut state runs: a window can hold a handful of rows of a state,
    whose weight 1 / var is then huge.  The fixed-point scale must cover them (from the
    per-state table), not send the fits to the fallback."""
    tab = make_case(gp.synthetic, 30011, k=13, faint=True, ora=ora)
    off = gp.synthetic.stefan_centres()
    dt = float(np.diff(ora.make_times(tab["time_us"][:2], tab["mjd"]))[0])
    window = wrows * dt
    nwin = -(-30011 // wrows)
    a, ra = _htab(gp, monkeypatch, "dmma", tab, True, off, 103 * 32 * nwin, window=window)
    b, rb = _htab(gp, monkeypatch, "tensor", tab, True, off, 103 * 32 * nwin, window=window)
    assert rb[3].shape[0] == 32 * nwin and (rb[3][:, 2] == ra[3][:, 2]).all()   # same evaluator per fit
    ok = ~np.isnan(a)
    assert (np.isnan(b) == np.isnan(a)).all()
    scale = np.nanmax(np.abs(a).reshape(103, -1), axis=0)
    err = np.abs(np.where(ok, a - b, 0.0)).reshape(103, -1) / scale
    assert np.nanmax(err) <= 10 * SUM_TOL, np.nanmax(err)


@pytest.mark.parametrize("n", [5000, 40000])
def test_objective_from_tensor_sums_matches_oracle(gp, ora, monkeypatch, n):
    """The objective rebuilt (in numpy) from the tensor kernel's harmonic table,
        sum_n conj(e_n) z_n = sum_k J_k(b) e^{-jkq} Z_k,   chi2 N = S_dd - |S_gd|^2 / S_gg,
    against the oracle's O(N) evaluation of the reference objective (src/Modulation.jl:323-326)
    at 1e-10, for (b, phi) over the range the solver visits."""
    from scipy.special import jv
    tab = make_case(gp.synthetic, n, k=17)
    off = gp.synthetic.stefan_centres()
    H, _ = _htab(gp, monkeypatch, "tensor", tab, False, off, 103 * 32)
    t, z = gp.synthetic.to_complex(tab, off)
    theta = ora.M_2PI * t                                 # fl(omega * t), like the reference
    rng = np.random.default_rng(3)
    worst = 0.0
    for ch in (0, 7, 18, 31):
        fc = np.exp(1j * np.angle(z[:, 32 + ch // 4]))
        h = H[:, ch]
        sdd, sgg = h[1], h[2]
        k = np.arange(1, 25)
        A, B, C, D = (h[7 + 4 * (k - 1) + s] for s in range(4))
        Zp, Zm, Z0 = (A + B) + 1j * (C - D), (A - B) + 1j * (C + D), h[5] + 1j * h[6]
        for _ in range(6):
            b, phi = rng.uniform(0.05, 3.0), rng.uniform(-np.pi, np.pi)
            q = (theta + phi) - theta                     # the phase quantum: uniform over the table
            assert np.ptp(q) == 0.0
            q = q[0]
            S = jv(0, b) * Z0 + np.sum(jv(k, b) * (np.exp(-1j * k * q) * Zp + (-1.0) ** k * np.exp(1j * k * q) * Zm))
            f = (sdd - abs(S) ** 2 / sgg) / n
            fo = ora.chi2(t, z[:, ch], fc, b, phi)[0]
            worst = max(worst, abs(f - fo) / fo)
    assert worst <= 1e-10, worst


def _fit_both(gp, monkeypatch, tab, volt, off):
    out = {}
    for mode in ("dmma", "tensor"):
        monkeypatch.setenv("GPPD_HARMONICS", mode)
        out[mode] = gp.process_table(tab["time_us"], volt, tab["mjd"], offsets=off)
    return out["dmma"], out["tensor"]


def test_tensor_unusual_inputs(gp, ora, monkeypatch):
    """Inputs outside the comfortable range: the tensor kernel either agrees with the FP64
    kernel or hands the fit to the direct evaluator -- it never returns something else."""
    tab = make_case(gp.synthetic, 9000, k=19)
    off = gp.synthetic.stefan_centres()
    base = tab["volt"]
    # (1) diodes of very different amplitude in one group (per-diode scales)
    v = base.copy()
    v[:, 0:2] = (v[:, 0:2] - off[0].real) * 1e-4 + off[0].real
    v[:, 2:4] *= 300.0
    a, b = _fit_both(gp, monkeypatch, tab, v, off)
    assert np.isfinite(b[1]).all()
    assert np.abs(a[1][:, 4:6] - b[1][:, 4:6]).max() <= 2e-3
    # (2) a NaN and an Inf sample: the group goes to the direct evaluator, like the FP64 path
    #     the results are whatever the reference arithmetic gives (NaN), in both kernels
    v = base.copy()
    v[1234, 8] = np.nan
    v[4321, 17] = np.inf
    a, b = _fit_both(gp, monkeypatch, tab, v, off)
    # the NaN sits in channel 4 (floats 8, 9: group 1), the Inf in channel 8 (floats 16, 17: group 2)
    for ch in (4, 8):
        assert np.isnan(a[1][ch, 4]) == np.isnan(b[1][ch, 4])
    ok = [ch for ch in range(32) if ch // 4 not in (1, 2)]
    assert np.abs(a[1][ok, 4:6] - b[1][ok, 4:6]).max() <= 2e-3
    # (3) a table of zeros and a constant table: no crash, same evaluator decisions
    for v in (np.zeros_like(base), np.full_like(base, 0.25)):
        a, b = _fit_both(gp, monkeypatch, tab, v, off)
        assert np.array_equal(np.isnan(a[1]), np.isnan(b[1]))
        assert np.array_equal(np.isfinite(a[0]), np.isfinite(b[0]))


# ---- complex128 arrays (the demodulateall boundary): the same kernel with a loader and a V-producer
# lane mapping of its own (csrc/harm_tc_kernels.cu, template parameter ARR)
def _htab_arrays(gp, monkeypatch, mode, t, z, state, nvals, min_groups=None, **kw):
    from gppd_b200 import _lib
    monkeypatch.setenv("GPPD_HARMONICS", mode)
    if min_groups is not None:
        monkeypatch.setenv("GPPD_TENSOR_MIN_GROUPS", str(min_groups))
    res = gp.demodulateall(t, z, faintparam=state, raw=True, return_info=True, **kw)
    out = np.empty(nvals)
    _lib.check(_lib.lib().gppd_debug_harmonics(_lib.default_handle().raw, 0, _lib.ptr(out), nvals))
    return out.reshape(-1, 32), res


@pytest.mark.parametrize("n,faint,fit,onlyhigh,nwindow", [
    (700, False, False, False, 0), (6144, False, False, False, 0), (20011, True, False, False, 0),
    (20011, True, False, True, 0), (30000, False, True, False, 0), (30000, True, True, False, 0),
    (30011, True, False, False, 5003), (100000, False, False, False, 0)])
def test_tensor_sums_on_arrays_equal_dmma_sums(gp, ora, monkeypatch, n, faint, fit, onlyhigh, nwindow):
    tab = make_case(gp.synthetic, n, k=7, faint=faint, ora=ora)
    off = None if fit else gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    nwin = -(-n // nwindow) if nwindow else 1
    nv = (201 if fit else 103) * 32 * nwin
    kw = dict(fitoffsets=fit, onlyhigh=onlyhigh, nwindow=nwindow)
    a, ra = _htab_arrays(gp, monkeypatch, "dmma", t, z, tab["state"], nv, **kw)
    b, rb = _htab_arrays(gp, monkeypatch, "tensor", t, z, tab["state"], nv, **kw)
    assert a.tobytes() != b.tobytes()                        # the other kernel did run
    assert np.array_equal(np.isnan(a), np.isnan(b))
    per = 201 if fit else 103
    a2, b2 = a.reshape(per, -1), b.reshape(per, -1)
    err = np.nanmax(np.abs(a2 - b2) / np.nanmax(np.abs(a2), axis=0))
    assert err <= 10 * SUM_TOL, err
    assert (rb[3][:, 2] == ra[3][:, 2]).all()                # same evaluator per fit
    assert np.abs(ra[1][:, 4:6] - rb[1][:, 4:6]).max() <= 2e-3
    assert np.array_equal(ra[0][:, 32:], rb[0][:, 32:])      # FC columns: copies


@pytest.mark.parametrize("mask", [0x0f, 0xf0, 0x35, 0x80])
def test_tensor_on_arrays_with_a_group_mask(gp, ora, monkeypatch, mask):
    """gppd_options.group_mask on the tensor kernel: the groups of the mask get the bits of the
    unmasked call, the other columns are neither read (they hold NaN here) nor written."""
    n = 20011
    tab = make_case(gp.synthetic, n, k=9, faint=True, ora=ora)
    off = gp.synthetic.stefan_centres()
    t, z = gp.synthetic.to_complex(tab, off)
    full, rf = _htab_arrays(gp, monkeypatch, "tensor", t, z, tab["state"], 103 * 32, min_groups=1)
    zz = z.copy()
    on = [g for g in range(8) if (mask >> g) & 1]
    offg = [g for g in range(8) if not (mask >> g) & 1]
    for g in offg:
        zz[:, 4 * g:4 * g + 4] = np.nan
        zz[:, 32 + g] = np.nan
    part, rp = _htab_arrays(gp, monkeypatch, "tensor", t, zz, tab["state"], 103 * 32, min_groups=1, groups=mask)
    cols = np.concatenate([np.arange(4 * g, 4 * g + 4) for g in on])
    assert np.array_equal(full[:, cols], part[:, cols])
    assert rf[1][cols].tobytes() == rp[1][cols].tobytes() and rf[2][cols].tobytes() == rp[2][cols].tobytes()
    assert np.array_equal(rf[0][:, cols], rp[0][:, cols])
