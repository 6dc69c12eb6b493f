#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_siblings.py tests/test_gpu_unet_vanilla.py tests/test_gpu_postprocess.py -q --timeout 1200 -s > gpurun_out/r2c_pytest.log 2>&1; tail -40 gpurun_out/r2c_pytest.log
