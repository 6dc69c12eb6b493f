#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
LOG=gpurun_out/r2k_probe.log
: > $LOG
P=build/probe_conv
for cap in 0 3; do for c in 0 1 2 3 4 5 6 7 8 9 10 11; do timeout 300 $P check $c $cap >> $LOG 2>&1 || echo "case $c cap $cap exit=$?" >> $LOG; done; done
for s in 3 4 7 8; do timeout 120 $P time $s 9 5 >> $LOG 2>&1; done
grep -c PASS $LOG; grep -c FAIL $LOG; grep TIME $LOG
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -q --timeout 900 -x 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -2 gpurun_out/r2k_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench.json")); print("ms/step %.1f e2e %.1f conv TF/s %.0f frac %.3f share %.3f clk %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["clocks"]["sm_mhz"]))
for k,v in d["roofline"]["layers"].items(): print("  %-32s %6.0f TF/s %7.3f ms" % (k, v["tflops"], v["ms_per_launch"]))
PY
done
