"""Debug: harmonic table of the tensor-core kernel against the DMMA kernel."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gppd_b200 as gp
from gppd_b200 import _lib
import oracle
from conftest import make_case

def htab(tab, faint, offsets, mode, nvals, **kw):
    os.environ["GPPD_HARMONICS"] = mode
    fs = tab["faintstates"]
    fs_g = gp.FaintStates(fs.timer1, fs.timer2, 1.0, 2.0) if faint else None
    res = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=offsets, faintparam=fs_g, **kw)
    h = _lib.default_handle()
    out = np.empty(nvals)
    _lib.check(_lib.lib().gppd_debug_harmonics(h.raw, 0, _lib.ptr(out), nvals))
    return out, res

for n, faint, fit in [(5000, False, False), (20011, False, False), (20011, True, False), (30000, False, True), (30000, True, True)]:
    tab = make_case(gp.synthetic, n, k=7, faint=faint, ora=oracle)
    off = None if fit else gp.synthetic.stefan_centres()
    nv = (201 if fit else 103) * 32
    a, ra = htab(tab, faint, off, "dmma", nv)
    b, rb = htab(tab, faint, off, "tensor", nv)
    a = a.reshape(-1, 32); b = b.reshape(-1, 32)
    scale = np.abs(a).max(axis=0)
    err = np.abs(a - b) / scale
    print(n, faint, fit, "max rel err (to column max)", err.max(), "at value", np.unravel_index(err.argmax(), err.shape),
          "nan:", np.isnan(b).sum(), "params diff", np.abs(ra[1] - rb[1]).max(), "fallbacks", int((rb[3][:, 2] == 1).sum()))
    if err.max() > 1e-9:
        v = np.unravel_index(err.argmax(), err.shape)[0]
        print("  worst rows:", np.argsort(-err.max(axis=1))[:10])
        print("  a", a[v, :4], "\n  b", b[v, :4])
