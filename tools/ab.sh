#!/bin/bash
# A/B of kernel experiment builds on the night workload: tools/ab.sh libgppd.so libgppd_x.so ...
# (libraries next to gppupildemodulation.jl_b200/libgppd.so, built with `make BUILD=... OUT=... EXTRA=...`)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  GPPD_LIBRARY=$PWD/gppupildemodulation.jl_b200/$lib python bench.py $AB_FLAGS --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err || tail -3 gpurun_out/ab_$lib.err
  python - "$lib" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % lib))
    p = d["roofline"]["pass_ms_per_step"]
    print("%-22s step %.3f ms | %s" % (lib, d["ms_per_step"], " ".join("%s %.3f" % (k, v) for k, v in p.items())))
except Exception as e:
    print(lib, "failed", e)
PY
done
