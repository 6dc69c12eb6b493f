#!/bin/bash
# round 2, late: kernel-level tests of the PatchGAN convolutions + launch list of one training iteration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_discriminator.py tests/test_gpu_siblings.py -q --timeout 1500 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 900 --csv --log-file gpurun_out/r2y_train_launches.csv python tools/bench_apollo_step.py 108 8 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2y_train_launches.csv 2>/dev/null | head -30
