"""Quick GPU-vs-oracle check (run on a GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gppd_b200 as gp
import oracle
from gppd_b200 import synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
method = sys.argv[2] if len(sys.argv) > 2 else "auto"
tab = syn.make_table(n, k=0)
t, z = syn.to_complex(tab, syn.stefan_centres())
t0 = time.time(); oo, op, ol, onf = oracle.demodulateall(t, z, nthreads=8, return_nfev=True); t_or = time.time() - t0
t0 = time.time(); go, gpar, gl, info, trace = gp.demodulateall(t, z, raw=True, return_info=True, return_trace=True, method=method); t_g = time.time() - t0
t0 = time.time(); go, gpar, gl, info = gp.demodulateall(t, z, raw=True, return_info=True, method=method); t_g2 = time.time() - t0
print("oracle s", t_or, "gpu s (first)", t_g, "second", t_g2)
print("method used", info[:, 2].tolist())
print("nfev oracle", onf.tolist()); print("nfev gpu   ", info[:, 0].tolist())
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
print("param maxrel b", rel(gpar[:, 4], op[:, 4]).max(), "phi", np.abs(gpar[:, 5] - op[:, 5]).max(),
      "a", (np.abs((gpar[:, 2] + 1j * gpar[:, 3]) - (op[:, 2] + 1j * op[:, 3])) / np.abs(op[:, 2] + 1j * op[:, 3])).max())
print("chi2 maxrel", rel(gl, ol).max())
print("out maxabs", np.abs(go - oo).max(), "scale", np.abs(oo).max())
same = info[:, 0] == onf
print("fits with same nfev:", same.sum(), "/32; max param diff on those:",
      np.abs(gpar[same] - op[same]).max() if same.any() else None)
# objective parity along the GPU trace
worst = 0
for ch in range(0, 32, 5):
    g = ch // 4
    fcph = np.exp(1j * np.angle(z[:, 32 + g]))
    for k in range(info[ch, 0]):
        b, phi, f = trace[ch, k]
        fo = oracle.chi2(t, z[:, ch], fcph, b, phi)[0]
        worst = max(worst, abs(f - fo) / fo)
print("objective parity along trace (max rel):", worst)
