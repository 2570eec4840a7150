"""The product's device solver (csrc/newuoa2.cuh, a resumable n=2 / npt=5
state machine) against the oracle's general-n NEWUOA: built for the host with
g++ -ffp-contract=off (what nvcc -fmad=false gives on the device) the two must
agree BIT FOR BIT -- every trial point, the returned x, f, nf and status --
because NEWUOA has rounding-level ties and any arithmetic difference would
split trajectories.  The same comparison runs on the GPU in test_gpu_parity."""
import ctypes as C

import numpy as np
import pytest


def _run2(L, ora, fun, x0, rhobeg=1.0, rhoend=1e-3, maxfun=60, entry="newuoa2_host"):
    dp = C.POINTER(C.c_double)
    solver = getattr(L, entry)
    solver.argtypes = [ora.OBJFUN, C.c_void_p, dp, C.c_double, C.c_double, C.c_int, dp,
                       C.POINTER(C.c_int), ora.OBSERVER, C.c_void_p]
    x = np.array(x0, float)
    rec = []
    cf = ora.OBJFUN(lambda n, xp, _: float(fun(np.array([xp[0], xp[1]]))))
    co = ora.OBSERVER(lambda nf, n, xp, f, _: rec.append((xp[0], xp[1], f)))
    fo, nf = C.c_double(), C.c_int()
    st = solver(cf, None, x.ctypes.data_as(dp), rhobeg, rhoend, maxfun, C.byref(fo),
                        C.byref(nf), co, None)
    return st, x, fo.value, nf.value, rec


def _same_bits(a, b):
    """bitwise equality, except that any NaN equals any NaN (sign/payload of a
    NaN is a compiler artefact)"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(nan | (a.view(np.int64) == b.view(np.int64))))


def _objectives(rng):
    p = rng.normal(size=8)
    return [
        lambda x: (x[0] - p[0]) ** 2 * (1 + p[1] ** 2) + (x[1] - p[2]) ** 2 * (0.5 + p[3] ** 2)
        + 0.3 * p[4] * (x[0] - p[0]) * (x[1] - p[2]),
        lambda x: 1 - np.cos(x[1] - p[0]) * np.exp(-(x[0] - 1 - 0.3 * p[1]) ** 2) + 0.01 * x[0] ** 2,
        lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + p[0] * 0.01 * x[0],
        lambda x: abs(x[0] - p[0]) ** 1.5 + np.sin(3 * x[1] + p[1]) * np.cos(x[0]) + 0.1 * x[1] ** 2,
        lambda x: 1.0,                                   # flat: trust-region step cannot reduce Q
        lambda x: float("nan") if x[0] > 1.5 else x[0] ** 2 + x[1] ** 2,
    ]


@pytest.mark.parametrize("entry", ["newuoa2_host", "newuoa2_host_warp"])
def test_device_solver_bitwise_equals_oracle(ora, newuoa2_host, entry):
    """entry = newuoa2_host: serial angle searches (thread-per-fit / block-replicated
    kernels); newuoa2_host_warp: the lane-parallel searches of the warp-per-fit
    kernel, emulated lane by lane (same selection code path, shuffles -> arrays)."""
    rng = np.random.default_rng(5)
    n_runs = 0
    for trial in range(120):
        for f in _objectives(rng):
            x0 = rng.normal(size=2)
            maxfun = int(rng.choice([60, 60, 25, 200, 7, 3]))
            rhoend = float(rng.choice([1e-3, 1e-6]))
            rec0 = []
            st0, xa, fa, nfa = ora.newuoa(f, x0, rhoend=rhoend, maxfun=maxfun, record=rec0)
            st1, xb, fb, nfb, rec1 = _run2(newuoa2_host, ora, f, x0, rhoend=rhoend, maxfun=maxfun,
                                           entry=entry)
            assert (st0, nfa) == (st1, nfb)
            assert _same_bits(xa, xb) and _same_bits([fa], [fb])
            assert len(rec0) == len(rec1)
            for (p0, f0), (b1, p1, f1) in zip(rec0, rec1):
                assert _same_bits(p0, [b1, p1])
            n_runs += 1
    assert n_runs == 720


def test_solver_sincos_portable(ora, newuoa2_host):
    # the solver's own sin/cos: identical bits in both implementations, < 1 ulp
    L = ora.lib()
    dp = C.POINTER(C.c_double)
    for fn in (L.newuoa_oracle_sincos, newuoa2_host.nu_sincos_host):
        fn.argtypes = [C.c_double, dp, dp]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-1, 8, 5000), 2 * np.pi * np.arange(51) / 50])
    s0, c0, s1, c1 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    for x in xs:
        L.newuoa_oracle_sincos(x, C.byref(s0), C.byref(c0))
        newuoa2_host.nu_sincos_host(x, C.byref(s1), C.byref(c1))
        assert s0.value == s1.value and c0.value == c1.value
        assert abs(s0.value - np.sin(x)) <= 2.3e-16 and abs(c0.value - np.cos(x)) <= 2.3e-16
