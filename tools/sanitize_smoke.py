"""Small end-to-end calls for compute-sanitizer (memcheck): every kernel of the table path and
the array path at ragged sizes.  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gppd_b200 as gp
off = gp.synthetic.stefan_centres()
for n, faint, kw in ((2501, False, {}), (1237, True, dict(window=0.35)), (777, False, dict(keepraw=True)),
                     (6300, True, {}), (300, False, dict(offsets=None))):
    tab = gp.synthetic.make_table(n, k=3)
    fs = None
    if faint:
        fs = gp.buildfaintparameters(gp.synthetic.faint_header(tab["mjd"], t_first=0.3, rate=0.5, gap=0.2, repeat=4))
    o = kw.pop("offsets", off)
    r = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=o, faintparam=fs, **kw)
    assert np.isfinite(r[1]).all()
    t, z = gp.synthetic.to_complex(tab, off)
    out, par, like = gp.demodulateall(t, z, raw=True, groups=0x0f)
    out, par, like = gp.demodulateall(t, z, raw=True, nwindow=500, method="direct")
print("sanitize smoke ok")
