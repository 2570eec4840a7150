#!/usr/bin/env python
"""Cycle breakdown of k_harm_tc from a -DTC_PROFILE build:
   GPPD_LIBRARY=.../libgppd_prof.so python tools/tc_profile.py [tables]
Runs a resident night through bench.py's generator and prints the per-block averages."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import gppd_b200 as gp
from gppd_b200 import _lib
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = 100_000
dev = torch.device("cuda", 0)
h, L = gp.Handle(0), _lib.lib()
time_us, volt, mjds, fss = bench.generate_night(torch, gp, dev, F, N, 0, plan=[False] * F)
out = torch.empty_like(volt)
params = torch.empty((F, 32, 6), dtype=torch.float64, device=dev)
chi2 = torch.empty((F, 32), dtype=torch.float64, device=dev)
offsets = torch.tensor(gp.synthetic.stefan_centres().view(np.float64), device=dev)
opt = gp.api._options()
vp = lambda ts: (C.c_void_p * F)(*[t.data_ptr() for t in ts])
b_n = (C.c_int64 * F)(*([N] * F))
b_mjd = (C.c_double * F)(*mjds)
args = (h.raw, 0, None, F, b_n, None, vp([time_us[k] for k in range(F)]), b_mjd, vp([volt[k] for k in range(F)]),
        C.c_void_p(offsets.data_ptr()), None, None, None, None, C.byref(opt), vp([out[k] for k in range(F)]),
        vp([params[k] for k in range(F)]), vp([chi2[k] for k in range(F)]), None, None)
cnt = (C.c_uint64 * 16)()
for _ in range(2):
    _lib.check(L.gppd_process_tables_f32_dev(*args))
_lib.check(L.gppd_debug_counters(h.raw, cnt, 1))
R = 5
for _ in range(R):
    _lib.check(L.gppd_process_tables_f32_dev(*args))
_lib.check(L.gppd_debug_counters(h.raw, cnt, 1))
c = [int(x) for x in cnt]
nb = max(c[0], 1)
names = ["blocks", "setup", "loop (control warp)", "epilogue", "ctl: wait operands", "ctl: issue MMAs",
         "ctl: load raw", "V0: wait raw", "V0: wait op stage", "V0: loop", "E0: wait raw", "E0: wait op stage",
         "E0: loop"]
print("blocks per launch", c[0] // R, " K-blocks per block ~", 192)
for i in range(1, 13):
    print("  %-22s %10.0f cycles per block  (%6.1f per K-block)" % (names[i], c[i] / nb, c[i] / nb / 192.0))
