"""Fork rate of the GPU against the oracle over a whole synthetic night (BASELINE.json
configs[2]: tables of 1e5 rows, 70 % bright / 30 % FAINT, whole-file fits, -c stefan):
how many of the 32 x N fits follow the oracle's NEWUOA trajectory (parameters equal to 1e-9),
and how far apart the others end.  TEST TOOL: uses the oracle.

    python tools/fork_rate_night.py [tables] [rows] > profiles/r2_fork_rate_night.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gppd_b200 as gp
import oracle
import fitref
from conftest import make_case

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
off = gp.synthetic.stefan_centres()
db, dp, dc, forks, per_table = [], [], [], 0, []
t0 = time.time()
for k in range(T):
    faint = (k % 10) in (3, 6, 9)
    tab = make_case(gp.synthetic, N, k=k, faint=faint, jitter=True, ora=oracle)
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    vout, par, chi2, info, st = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, faintparam=fs_g)
    t, z = gp.synthetic.to_complex(tab, off)
    oo, op, ol, onf = oracle.demodulateall(t, z, faintparam=tab["state"], nthreads=8, return_nfev=True)
    same, stats = fitref.compare_fits(par, chi2, op, ol, info[:, 0], onf)
    forks += stats["forks"]
    per_table.append(stats["forks"])
    db += list(stats["db"]); dp += list(stats["dphi"]); dc += list(stats["dchi2"])
    # coinciding fits: demodulated output within one float32 ulp of the oracle's
    ref = np.empty((N, 64), np.float32)
    ref[:, 0::2], ref[:, 1::2] = oo[:, :32].real, oo[:, :32].imag
    err = np.abs(vout[:, :64].astype(np.float64) - ref).reshape(N, 32, 2).max(axis=(0, 2)) / np.abs(z[:, :32]).max(axis=0)
    assert (err[same] <= 2.0 ** -22).all()
d = np.maximum(db, dp) if forks else np.zeros(1)
print(json.dumps({
    "workload": "night of %d tables x %d rows (30 %% FAINT), whole-file fits, -c stefan; GPU = METROLOGY-table path "
                "(int8 tensor-core harmonic sums), oracle = oracle.demodulateall" % (T, N),
    "fits": 32 * T, "forks": forks, "fork_rate": forks / (32.0 * T),
    "forks_per_table_max": int(max(per_table)), "forks_per_table_mean": float(np.mean(per_table)),
    "dpar_median": float(np.median(d)), "dpar_p90": float(np.quantile(d, 0.9)), "dpar_max": float(d.max()),
    "dchi2_median": float(np.median(dc)) if forks else 0.0, "dchi2_p90": float(np.quantile(dc, 0.9)) if forks else 0.0,
    "dchi2_max": float(np.max(dc)) if forks else 0.0,
    "coinciding_fits": "parameters to 1e-9, demodulated float32 output within 2^-22 of the column scale (asserted)",
    "seconds": time.time() - t0}))
