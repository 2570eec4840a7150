"""`--center empirical`: circle centres (reference compute_offsets,
src/GPPupilDemodulation.jl:105-125, used by processmetrology at :153-154).

The reference throws on this path (`Circle` is undefined there), so there is no reference
output to pin against: the oracle restates the least-squares circle that call stands
for, the CPU tests pin the oracle on circles with known centres, and the GPU tests hold
the device reduction to the oracle and to the explicit-centres path."""
import numpy as np
import pytest

from conftest import make_case

CENTRE_TOL = 1e-10   # |centre_gpu - centre_oracle| / radius scale (float64 sums vs SVD)


# ---------------------------------------------------------------------------- CPU
def test_oracle_circle_known_centres(ora):
    rng = np.random.default_rng(5)
    for c, r, arc in [(3 + 4j, 2.0, 2 * np.pi), (-0.25 + 0.01j, 0.1, 1.0), (1e3 - 2e3j, 0.5, 3.0)]:
        th = rng.uniform(0.3, 0.3 + arc, size=400)
        p = c + r * np.exp(1j * th)
        got = ora.circle_centre(p.real, p.imag)
        assert abs(got - c) <= 1e-9 * max(1.0, abs(c)), (c, got)
    # noisy full circle: centre within the noise
    th = rng.uniform(0, 2 * np.pi, size=20000)
    p = (1 - 2j) + 0.7 * np.exp(1j * th) + 0.01 * (rng.normal(size=th.size) + 1j * rng.normal(size=th.size))
    assert abs(ora.circle_centre(p.real, p.imag) - (1 - 2j)) < 1e-3


def test_oracle_circle_degenerate(ora):
    assert ora.circle_centre([1.0, 2.0], [0.0, 1.0]) == 0                       # < 3 points
    x = np.linspace(0, 1, 50)
    assert ora.circle_centre(x, 2 * x + 1) == 0                                 # one line
    assert ora.circle_centre(np.ones(10), np.ones(10)) == 0                     # one point


def test_oracle_offsets_use_high_samples(ora):
    rng = np.random.default_rng(11)
    n = 3000
    state = rng.choice(np.array([ora.OFF, ora.LOW, ora.NORMAL, ora.HIGH, ora.TRANSIENT]), size=n)
    th = rng.uniform(0, 2 * np.pi, size=(n, 40))
    c_high = (np.arange(40) - 20) * (0.01 + 0.02j)
    v = np.where((state == ora.HIGH)[:, None], c_high[None, :] + 0.3 * np.exp(1j * th),
                 5.0 + 0.1 * np.exp(1j * th))          # other states: another circle
    assert np.abs(ora.compute_offsets(v, state) - c_high).max() < 1e-9
    allc = ora.compute_offsets(v, None)
    assert np.abs(allc - c_high).min() > 0.1           # all samples: a different answer


# ---------------------------------------------------------------------------- GPU
def _centres_gpu(gp, tab, faint, **kw):
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    c = np.zeros(40, dtype=np.complex128)
    res = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=True, faintparam=fs_g,
                           centres_out=c, **kw)
    return c, res, fs_g


@pytest.mark.gpu
@pytest.mark.parametrize("faint", [False, True])
@pytest.mark.parametrize("n", [700, 5000, 20011])
def test_centres_match_oracle(gp, ora, n, faint):
    tab = make_case(gp.synthetic, n, k=31, faint=faint, ora=ora)
    c, _, _ = _centres_gpu(gp, tab, faint)
    v = tab["volt"].astype(np.float64)
    z = v[:, 0::2] + 1j * v[:, 1::2]
    want = ora.compute_offsets(z, tab["state"])
    scale = np.abs(z - want[None, :]).max(axis=0)       # ~ radius of each channel
    assert np.abs(c - want).max() < 1.0 and (np.abs(c - want) <= CENTRE_TOL * scale).all()
    # and they are the generating centres (short arcs of a short table pin them less well)
    assert np.median(np.abs(c[:32] - gp.synthetic.stefan_centres()[:32])) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("faint,window", [(False, None), (True, None), (True, 1.0)])
def test_empirical_equals_explicit_centres(gp, ora, faint, window):
    """offsets = true is `cmplxV .-= compute_offsets(...)` followed by the ordinary path
    (:153-154): bit-identical to passing the same centres as a vector."""
    tab = make_case(gp.synthetic, 6000, k=32, faint=faint, ora=ora)
    c, res, fs_g = _centres_gpu(gp, tab, faint, window=window)
    ref = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=c, faintparam=fs_g,
                           window=window)
    for a, b in zip(res[:3], ref[:3]):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_empirical_processmetrology_vs_oracle(gp, ora):
    tab = make_case(gp.synthetic, 5000, k=33, faint=True, ora=ora)
    fs_o = tab["faintstates"]
    fs_g = gp.FaintStates(fs_o.timer1, fs_o.timer2, 1.0, 2.0)
    tg, hg = gp.processmetrology({"TIME": tab["time_us"], "VOLT": tab["volt"]}, tab["mjd"],
                                 faintparam=fs_g, offsets=True)
    to, ho = ora.processmetrology(tab["time_us"], tab["volt"], tab["mjd"], faintparam=fs_o,
                                  offsets=True, nthreads=8)
    assert set(hg) == set(ho) and not any("CENTER" in k for k in hg)
    # FC channels are only centred: float32 of (volt - centre), centres equal to 1e-10
    assert np.abs(tg["VOLT"][:, 64:].astype(np.float64) - to["VOLT"][:, 64:]).max() <= 1e-6
    keys = [k for k in ho if "SIN AMPLITUDE" in k]
    close = sum(abs(hg[k] - ho[k]) <= 1e-6 * abs(ho[k]) for k in keys)
    # (the centres themselves agree to 1e-10 of the radius only, hence 1e-6 rather than 1e-9;
    # forked fits inside the measured envelope of the reference procedure, tests/fitref.py)
    import fitref
    assert close >= fitref.MIN_COINCIDE and all(abs(hg[k] - ho[k]) <= fitref.FORK_HARD for k in keys)


@pytest.mark.gpu
def test_centres_degenerate_and_batch_independent(gp, ora):
    """A channel with no circle keeps centre 0; a table's centres do not depend on n
    being a multiple of the segment length."""
    tab = make_case(gp.synthetic, 4096 + 17, k=34, faint=False, ora=ora)
    volt = tab["volt"].copy()
    volt[:, 2 * 39] = 0.25            # FC channel 40: one point
    volt[:, 2 * 39 + 1] = -0.5
    x = np.linspace(-1, 1, volt.shape[0], dtype=np.float32)
    volt[:, 2 * 38], volt[:, 2 * 38 + 1] = x, 0.5 * x   # FC channel 39: a line
    c = np.zeros(40, dtype=np.complex128)
    gp.process_table(tab["time_us"], volt, tab["mjd"], offsets=True, centres_out=c)
    assert c[39] == 0 and c[38] == 0
    z = volt.astype(np.float64)
    want = ora.compute_offsets(z[:, 0::2] + 1j * z[:, 1::2], None)
    assert want[39] == 0 and want[38] == 0
    assert np.abs(c[:38] - want[:38]).max() <= 1e-9
