#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_apollo_d_path.py tests/test_gpu_apollo_step.py tests/test_gpu_discriminator.py tests/test_gpu_siblings.py -q --timeout 1500 2>&1 | tail -8
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1
python tools/bench_apollo_step.py 148 10 2>/dev/null | tail -1
python tools/profile_apollo_step.py 2>/dev/null | tail -14
