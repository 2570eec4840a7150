#!/bin/bash
# A/B of experiment builds on the long-exposure stress workload: tools/ab_stress.sh libgppd.so libgppd_x.so ...
cd "$(dirname "$0")/.."
for lib in "$@"; do
  GPPD_LIBRARY=$PWD/gppupildemodulation.jl_b200/$lib python bench.py --config stress --stress-rows ${STRESS_ROWS:-20000000} --steps 3 --warmup 1 > gpurun_out/abs_$lib.json 2> gpurun_out/abs_$lib.err || tail -3 gpurun_out/abs_$lib.err
  python - "$lib" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.load(open("gpurun_out/abs_%s.json" % lib))
    p = d["variants"]["global_fit"]["pass_ms_per_step_rank0"]
    print("%-22s global %.3f ms windows %.3f ms | %s" % (lib, d["ms_per_step"], d["variants"]["windows_100s"]["ms_per_step"], " ".join("%s %.3f" % (k, v) for k, v in p.items())))
except Exception as e:
    print(lib, "failed", e)
PY
done
