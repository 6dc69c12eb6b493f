#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "== unet_train fixture test, remainder pairs ON"; python -m pytest tests/test_gpu_unet_train.py -q -k test_gradients_match_reference_fixture -s 2>&1 | grep -E "t_conv2.bias|passed|failed"
echo "== remainder pairs OFF"; NEUROCLEAR_REMAINDER_PAIRS=0 python -m pytest tests/test_gpu_unet_train.py -q -k test_gradients_match_reference_fixture -s 2>&1 | grep -E "t_conv2.bias|passed|failed"
for i in 1 2; do
for st in 1 2 3; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample --streams $st > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - "$st" <<PY
import json, sys
d=json.load(open("gpurun_out/r2l_bench.json")); print("streams %s ms/step %.1f e2e %.1f conv(ev) %.0f share %.3f clk %s sha %s" % (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"], d["clocks"]["sm_mhz"], d["out_sha256"][:16]))
PY
done
done
