#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_unet_train.py -q --timeout 900 -x 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --size 256 256 256 --batch 9 --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar --no-train-sample > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2p_launches.csv 2>/dev/null | grep -E "in_relu|total"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2p_bench.json")); print("ms/step %.1f e2e %.1f conv TF/s %.0f frac %.3f share %.3f clk %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["clocks"]["sm_mhz"]))
PY
