#!/bin/bash
# bench after the remainder-pair kernel + launch list of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench.json")); print("value %.4g ms/step %.1f e2e %.1f conv-share %.3f conv TF/s %.0f frac %.3f sha %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["out_sha256"][:16]))
for k,v in d["roofline"]["layers"].items(): print("  %-32s %6.0f TF/s %7.3f ms" % (k, v["tflops"], v["ms_per_launch"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --size 256 256 256 --batch 9 --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar --no-train-sample > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2f_launches.csv | head -40
