#!/bin/bash
# quick GPU visit: kernel tests + short bench + per-kernel times of one 128^3 (8-cube) pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json")); print("value %.4g ms/step %.1f conv-share %.3f conv TF/s %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["share_of_step"], d["roofline"]["achieved"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_q.csv python bench.py --size 128 128 128 --batch 8 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_q.csv
