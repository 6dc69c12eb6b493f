#!/bin/bash
# A/B on ONE box: remainder pairs off / on / off / on
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
for flag in "--no-remainder-pairs" ""; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample $flag > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - "$flag" <<PY
import json, sys
d=json.load(open("gpurun_out/r2g_bench.json")); print("%-22s ms/step %.1f e2e %.1f conv TF/s %.0f frac %.3f clk %s sha %s" % (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["out_sha256"][:16]))
PY
done
done
