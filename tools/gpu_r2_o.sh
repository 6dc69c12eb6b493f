#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests/test_gpu_deeplinear.py tests/test_gpu_unet_train.py tests/test_gpu_apollo_step.py -q --timeout 1500 2>&1 | tail -4
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1
python tools/bench_apollo_step.py 148 10 2>/dev/null | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r2o_train_launches.csv python tools/bench_apollo_step.py 108 4 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2o_train_launches.csv 2>/dev/null | head -32
python - <<PY
import sys, json, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
for crop in (148, 108, 108):
    ms, launches, _, model = bench.time_apollo_iterations(dev, crop, 5, 3, False)
    print("bench.time_apollo_iterations crop %d: %.2f ms" % (crop, ms))
    del model
PY
