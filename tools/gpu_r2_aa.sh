#!/bin/bash
# im2col49 / col2im49 on shared-memory tiles: parity tests of the DeepLinear path, iteration time, launch list rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_deeplinear.py tests/test_gpu_apollo_step.py -q -x --timeout 1500 2>&1 | tail -3
python - <<PY
import sys, json, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
for crop in (108, 108, 148):
    ms, launches, _, model = bench.time_apollo_iterations(dev, crop, 10, 6, False)
    print("bench.time_apollo_iterations crop %d: %.2f ms" % (crop, ms))
    del model
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"im2col49|col2im49" -c 12 --csv --log-file gpurun_out/r2aa_launches.csv python tools/bench_apollo_step.py 108 4 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2aa_launches.csv 2>/dev/null | head -5
