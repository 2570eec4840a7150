import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ora():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gp():
    import gppd_b200
    return gppd_b200


@pytest.fixture(scope="session")
def newuoa2_host():
    """Host build (g++ -ffp-contract=off) of the product's device solver."""
    import ctypes as C
    here = os.path.join(ROOT, "tests", "native")
    out = os.path.join(here, "build", "libnewuoa2_host.so")
    src = os.path.join(here, "newuoa2_host.cpp")
    hdr = os.path.join(ROOT, "gppupildemodulation.jl_b200", "csrc", "newuoa2.cuh")
    if (not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                        "-I", os.path.dirname(hdr), "-o", out, src], check=True)
    return C.CDLL(out)


def make_case(gp_synth, n, k=0, faint=False, centred=False, jitter=False, noise=0.02, ora=None):
    """Synthetic table (+ states for FAINT, built by the oracle's buildstates)."""
    tab = gp_synth.make_table(n, k=k, faint=False, jitter=jitter, noise=noise, centred=centred)
    state = None
    fs = None
    if faint:
        hdr = gp_synth.faint_header(tab["mjd"], t_first=1.0, rate=1.5, gap=0.4, repeat=max(2, int(n / 500 / 1.5)))
        fs = ora.buildfaintparameters(hdr)
        t = ora.make_times(tab["time_us"], tab["mjd"])
        state = ora.buildstates(fs, t)
        tab = gp_synth.make_table(n, k=k, faint=True, jitter=jitter, noise=noise, centred=centred, state=state)
        tab["header"] = hdr
    tab["state"] = state
    tab["faintstates"] = fs
    return tab
