import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gppd_b200 as gp
from gppd_b200 import _lib
import oracle
from conftest import make_case
os.environ["GPPD_HARMONICS"] = "tensor"
tab = make_case(gp.synthetic, 100000, k=3)
off = gp.synthetic.stefan_centres()
for window in (100.008, 80.0, 50.0):
    res = gp.process_table(tab["time_us"], tab["volt"], tab["mjd"], offsets=off, window=window)
    info = res[3]
    nwin = info.shape[0] // 32
    h = np.empty(103 * 32 * nwin)
    _lib.check(_lib.lib().gppd_debug_harmonics(_lib.default_handle().raw, 0, _lib.ptr(h), h.size))
    h = h.reshape(103, nwin, 32)
    print("window", window, "nwin", nwin, "fallback fits", int((info[:, 2] == 1).sum()), "nan entries per window", np.isnan(h).sum(axis=(0, 2)),
          "nan values idx", np.unique(np.where(np.isnan(h))[0])[:10], "groups", np.unique(np.where(np.isnan(h))[2] // 4))
