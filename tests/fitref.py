"""Shared pieces of the parity tests: the reference's per-diode fit procedure driven
by the oracle's NEWUOA on an arbitrary objective, the oracle objective of one diode,
and the tolerances every parity test uses (with where each number comes from).

TEST INFRASTRUCTURE: imports ``oracle``; nothing in the product does.
"""
import numpy as np

# ---- tolerances -----------------------------------------------------------------
REL_OBJ = 1e-10     # chi2(b, phi) on the GPU vs the oracle's objective at the same point
REL_FIT = 1e-9      # north-star FP64 tolerance: fits that follow the oracle's trajectory

# NEWUOA has rounding-level ties, so a chi2 that differs in its last
# bits can fork a trajectory; both branches are outputs of the reference procedure.  The
# numbers below are the measured envelope of that ambiguity -- the oracle against ITSELF
# with chi2 multiplied by (1 +- 1e-15), tests/test_fork_envelope.py (record:
# tests/golden/fork_envelope.json): 1 024 perturbed fits, bright / FAINT / fitted centres,
# 1 000 and 6 000 rows:
#   95 forks = 9.3 % of the fits (7-14 % per configuration);
#   max(|delta b|, |delta phi|): median 1.8e-5, 90th percentile 1.2e-4, largest 9.2e-4;
#   relative chi2 difference: median 2.0e-7, 90th percentile 6.4e-6, largest 5.6e-4.
# (The largest ones are fits that NEWUOA leaves at its rho_end = 1e-3 resolution.)
FORK_RATE_MAX = 0.25      # fraction of forked fits a case may show (measured 0.07-0.14)
FORK_HARD = 2e-3          # every forked fit: |delta b|, |delta phi| (2 rho_end; measured max 9.2e-4)
FORK_TYPICAL = 5e-4       # 90 % of the forked fits of a run (measured 1.2e-4)
FORK_CHI2_HARD = 1e-3     # every forked fit: relative chi2 difference (measured max 5.6e-4)
FORK_CHI2_TYPICAL = 1e-4  # 90 % of the forked fits (measured 6.4e-6)
# 32 fits per call at a fork rate p <= 0.14: mean 4.5 forks, sigma 2.0 -> 10 forks is the
# 3-sigma point; a case with fewer than 22 coinciding fits is a regression, not a tie
MIN_COINCIDE = 22
# ... and over a whole module (hundreds of fits) the coinciding fraction must be the measured one
MIN_COINCIDE_FRACTION = 0.80


def reference_fit(ora, f, maxfun=60, xinit=None):
    """reference src/Modulation.jl:402-416 and :426-431 on the objective ``f(b, phi)``:
    8-point phase scan at b = 0.1, NEWUOA (rhobeg 1, rhoend 1e-3), the pi-flip check
    with its optional second run, the final call, the sign normalisation.
    Returns dict(b, phi, chi2, nfev, second)."""
    nfev = [0]

    def F(x):
        nfev[0] += 1
        return f(float(x[0]), float(x[1]))

    if xinit is None:
        phi8 = ora.phirange()
        fs = [F((0.1, p)) for p in phi8]
        nan = [i for i, v in enumerate(fs) if v != v]
        k = nan[0] if nan else int(np.argmin(fs))
        x0 = [0.1, phi8[k]]
    else:
        x0 = list(xinit)
    _, x, _, _ = ora.newuoa(F, x0, maxfun=maxfun)
    lklval = F(x)
    phipi = x[1] + (np.pi if x[1] < 0 else -np.pi)
    second = False
    if lklval > F((x[0], phipi)):
        second = True
        _, x, _, _ = ora.newuoa(F, [x[0], phipi], maxfun=maxfun)
    chi2 = F(x)
    b, phi = float(x[0]), float(x[1])
    if b < 0:
        b, phi = -b, phi + (np.pi if phi < 0 else -np.pi)
    return dict(b=b, phi=phi, chi2=chi2, nfev=nfev[0], second=second)


def valid_rows(ora, state, onlyhigh):
    if state is None:
        return slice(None)
    v = (state != ora.TRANSIENT)
    if onlyhigh:
        v &= (state == ora.HIGH) | (state == ora.NORMAL)
    return v


def oracle_objective(ora, t, z, state, ch, onlyhigh=False, fitoffsets=False):
    """chi2(b, phi) -> (chi2, c, a) of diode channel ``ch`` exactly as demodulateall sets
    it up (src/Modulation.jl:388-399)."""
    g = ch // 4
    v = valid_rows(ora, state, onlyhigh)
    fc = np.exp(1j * np.angle(z[:, 32 + g]))[v]
    d = np.ascontiguousarray(z[:, ch][v])
    tt = np.ascontiguousarray(t[v])
    w = pw = None
    if state is not None:
        pw, w = ora.compute_mean_var_power(state[v], d)
    return lambda b, phi: ora.chi2(tt, d, fc, b, phi, weight=w, power=pw, fitoffsets=fitoffsets)


def dphi(a, b):
    """phase difference wrapped to (-pi, pi]"""
    return np.angle(np.exp(1j * (np.asarray(a) - np.asarray(b))))


def compare_fits(par, like, opar, olike, nfev=None, onfev=None):
    """GPU (par, like) against oracle (opar, olike), both (nfits, 6) / (nfits,) after the sign
    normalisation.  Returns (coincide mask, dict of fork statistics) and asserts the bounds
    every forked fit must keep."""
    par, opar = np.asarray(par), np.asarray(opar)
    db = np.abs(par[:, 4] - opar[:, 4])
    dp = np.abs(dphi(par[:, 5], opar[:, 5]))
    coincide = (db <= REL_FIT * np.abs(opar[:, 4])) & (dp <= REL_FIT * np.maximum(1.0, np.abs(opar[:, 5])))
    if nfev is not None and onfev is not None:
        coincide &= np.asarray(nfev) == np.asarray(onfev)
    fork = ~coincide
    dchi = np.abs(np.asarray(like) - np.asarray(olike)) / np.abs(olike)
    stats = dict(nfits=int(par.shape[0]), forks=int(fork.sum()),
                 max_db=float(db[fork].max()) if fork.any() else 0.0,
                 max_dphi=float(dp[fork].max()) if fork.any() else 0.0,
                 max_dchi2=float(dchi[fork].max()) if fork.any() else 0.0)
    if fork.any():
        assert db[fork].max() <= FORK_HARD and dp[fork].max() <= FORK_HARD, stats
        assert dchi[fork].max() <= FORK_CHI2_HARD, stats
    stats["db"], stats["dphi"], stats["dchi2"] = db[fork], dp[fork], dchi[fork]
    return coincide, stats


def check_fork_population(stats_list, what=""):
    """Over the forked fits of a whole run: the coinciding fraction is the measured one and
    90 % of the forks are far inside the hard bound."""
    nfits = sum(s["nfits"] for s in stats_list)
    forks = sum(s["forks"] for s in stats_list)
    assert nfits - forks >= MIN_COINCIDE_FRACTION * nfits, (what, forks, nfits)
    if forks >= 10:
        d = np.concatenate([np.maximum(s["db"], s["dphi"]) for s in stats_list])
        c = np.concatenate([s["dchi2"] for s in stats_list])
        assert np.quantile(d, 0.9) <= FORK_TYPICAL, (what, np.sort(d))
        assert np.quantile(c, 0.9) <= FORK_CHI2_TYPICAL, (what, np.sort(c))
    return forks, nfits
