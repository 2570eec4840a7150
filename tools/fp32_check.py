"""Measured accuracy of the optional GPPD_FP32 harmonic sums against the FP64 sums on the same
tables (GPU box; writes gpurun_out/r2_fp32.json).  The speed side is `bench.py --fp32`.
Usage: python tools/fp32_check.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gppd_b200 as gp                      # noqa: E402
from gppd_b200 import _lib                  # noqa: E402
import oracle as ora                        # noqa: E402  (checker only: builds the FAINT states)
from conftest import make_case              # noqa: E402


def run(tab, faint, off, per, fp32):
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    res = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, faintparam=fs_g, fp32=fp32)
    out = np.empty(per * 32)
    _lib.check(_lib.lib().gppd_debug_harmonics(_lib.default_handle().raw, 0, _lib.ptr(out), per * 32))
    return out.reshape(per, 32), res


def main():
    os.environ["GPPD_HARMONICS"] = "tensor"
    rows = []
    for n, faint, fit in [(1000, False, False), (10000, False, False), (100000, False, False),
                          (100000, True, False), (100000, False, True), (100000, True, True)]:
        worst = dict(sums=0.0, amp=0.0, b=[], chi2=0.0)
        for k in range(4):
            tab = make_case(gp.synthetic, n, k=31 + k, faint=faint, ora=ora)
            off = None if fit else gp.synthetic.stefan_centres()
            per = 201 if fit else 103
            a, ra = run(tab, faint, off, per, False)
            b, rb = run(tab, faint, off, per, True)
            worst["sums"] = max(worst["sums"], float(np.max(np.abs(a - b) / np.abs(a).max(axis=0))))
            pa, pb = ra[1], rb[1]
            amp = np.abs((pa[:, 2] + 1j * pa[:, 3]) - (pb[:, 2] + 1j * pb[:, 3])) / np.abs(pa[:, 2] + 1j * pa[:, 3])
            dx = np.maximum(np.abs(pa[:, 4] - pb[:, 4]), np.abs(np.angle(np.exp(1j * (pa[:, 5] - pb[:, 5])))))
            worst["amp"] = max(worst["amp"], float(amp.max()))
            worst["b"] += list(dx)
            st = rb[4]
            nvalid = n if st is None else int((st != ora.TRANSIENT).sum())
            worst["chi2"] = max(worst["chi2"], float(np.max(np.abs(ra[2] - rb[2]) / (a[1] / nvalid))))
        dx = np.array(worst["b"])
        rows.append(dict(rows=n, faint=faint, fitoffsets=fit, fits=int(dx.size),
                         max_rel_error_of_sums=worst["sums"], max_rel_error_of_amplitude=worst["amp"],
                         b_phi_difference=dict(median=float(np.median(dx)), p90=float(np.quantile(dx, 0.9)),
                                               max=float(dx.max())),
                         max_chi2_error_relative_to_Sdd_over_N=worst["chi2"]))
        print(rows[-1], flush=True)
    out = dict(what="GPPD_FP32 (harm_tc32_kernels.cu) against the FP64 tensor kernel on the same synthetic tables, "
                    "4 tables per line; b_phi_difference is the end point of NEWUOA (rho_end 1e-3), see "
                    "tests/test_fork_envelope.py for the same statistic under a 1e-15 perturbation", cases=rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_fp32.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
