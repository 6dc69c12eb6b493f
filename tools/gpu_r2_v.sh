#!/bin/bash
# round 2, late: discriminator-path latency work (deep prefetch in conv2d_k4_*, Cout = 1 / Cin = 1 kernels, wider
# colsum_finalize / loss_fwd): parity tests, iteration time, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_discriminator.py tests/test_gpu_apollo_d_path.py tests/test_gpu_siblings.py tests/test_gpu_unet_train.py tests/test_gpu_apollo_step.py -q -x --timeout 1500 2>&1 | tail -6
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1
python tools/bench_apollo_step.py 148 10 2>/dev/null | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv --log-file gpurun_out/r2v_train_launches.csv python tools/bench_apollo_step.py 108 6 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2v_train_launches.csv 2>/dev/null | head -24
