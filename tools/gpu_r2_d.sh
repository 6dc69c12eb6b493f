#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin.py -q --timeout 1500 -s > gpurun_out/r2d_pytest.log 2>&1; tail -60 gpurun_out/r2d_pytest.log | cut -c1-600
