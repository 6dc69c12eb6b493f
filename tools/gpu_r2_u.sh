#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
bash tests/cuda/run_probe_grad.sh notime > /dev/null 2>&1; grep -c PASS gpurun_out/probe_grad.log; grep -E "FAIL|exit=" gpurun_out/probe_grad.log | head
NC_RP=1 build/probe_grad time 6 5 | grep dgrad; NC_RP=0 build/probe_grad time 6 5 | grep dgrad
python -m pytest tests/test_gpu_unet_train.py tests/test_gpu_apollo_step.py -q --timeout 1500 2>&1 | tail -3
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1
python tools/bench_apollo_step.py 148 10 2>/dev/null | tail -1
