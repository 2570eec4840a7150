"""Stress of the tensor harmonic kernel's hand-over protocol: many shapes, repeated, each
call under a watchdog; checks run-to-run bit stability of the fitted parameters."""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gppd_b200 as gp
import oracle
from conftest import make_case
off = gp.synthetic.stefan_centres()
t00 = time.time()
for n in (33, 64, 1000, 6143, 6144, 6145, 12289, 24607, 50000):
    for faint in (False, True):
        if faint and n < 1000:
            continue
        tab = make_case(gp.synthetic, n, k=n % 17, faint=faint, ora=oracle)
        fs = tab["faintstates"]
        fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
        for window in (None, 0.9, 7.3):
            ref = None
            for rep in range(6):
                r = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, faintparam=fs_g, window=window)
                key = (r[0].tobytes(), r[1].tobytes())
                assert ref is None or key == ref, (n, faint, window, rep)
                ref = key
    print("rows", n, "ok", "%.1f s" % (time.time() - t00), flush=True)
print("stress ok")
