"""BASELINE.json configs[2] through the command line: a synthetic night directory of FITS
files (70 % bright / 30 % FAINT, 1e5 rows each) demodulated recursively by
bin/GPPupilDemodulation (-r), timed end to end (file read, H2D, kernels, D2H, file write)."""
import json, os, sys, time, shutil, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gppd_b200 as gp
import oracle
from conftest import make_case

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
d = "/tmp/night"; out = "/tmp/night_out"
shutil.rmtree(d, ignore_errors=True); shutil.rmtree(out, ignore_errors=True)
t0 = time.time()
for k in range(nfiles):
    faint = (k % 10) in (3, 6, 9)
    tab = make_case(gp.synthetic, rows, k=k, faint=faint, ora=oracle)
    sub = os.path.join(d, "sub%d" % (k % 4))
    os.makedirs(sub, exist_ok=True)
    gp.synthetic.make_fits(os.path.join(sub, "GRAVI_%03d.fits" % k), tab, tab["header"])
tgen = time.time() - t0
runs = []
for _ in range(3):                      # the box's disk / page cache makes single runs noisy
    shutil.rmtree(out, ignore_errors=True)
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "bin", "GPPupilDemodulation"), "-r", "-d", out, d],
                       capture_output=True, text=True)
    runs.append(time.time() - t0)
dt = min(runs)
nout = len(os.listdir(out)) if os.path.isdir(out) else 0
res = {"workload": "night directory of %d FITS files x %d rows (30 %% FAINT), bin/GPPupilDemodulation -r" % (nfiles, rows),
       "files_written": nout, "returncode": r.returncode, "seconds": dt, "files_per_s": nout / dt,
       "diode_samples_per_s": nout * rows * 32 / dt, "input_gb": nfiles * rows * 332 / 1e9,
       "generation_seconds": tgen, "seconds_of_each_run": runs, "note": "wall clock of the whole command including Python start-up, library load, file reads and writes on the box's local disk"}
print(json.dumps(res))
if r.returncode:
    print(r.stderr[-2000:], file=sys.stderr)
