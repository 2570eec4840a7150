#!/bin/bash
# A/B of experiment builds on the batched-fit sweep (1e6 fits of 512 rows): tools/ab_sweep.sh libgppd.so libgppd_x.so ...
cd "$(dirname "$0")/.."
for lib in "$@"; do
  GPPD_LIBRARY=$PWD/gppupildemodulation.jl_b200/$lib python bench.py --config sweep --sweep-max 1e6 --steps 3 > gpurun_out/abs_$lib.json 2> gpurun_out/abs_$lib.err || tail -3 gpurun_out/abs_$lib.err
  python - "$lib" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.load(open("gpurun_out/abs_%s.json" % lib))
    for v, c in d["sweep"].items():
        print("%-18s %-18s" % (lib, v), " ".join("F=%d: %.2f ms (fit %.2f)" % (r["fits"], r["ms"], r["pass_ms"]["fit"]) for r in c[2:]))
except Exception as e:
    print(lib, "failed", e)
PY
done
