#!/bin/bash
# round 2, late: PatchGAN convolutions on mma.sync 3xTF32 — A/B probe, parity tests, iteration time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tests/cuda/run_probe_disc.sh 0 > /dev/null
python tools/debug_disc_grad.py 2>&1 | tail -5
python -m pytest tests/test_gpu_discriminator.py tests/test_gpu_apollo_d_path.py tests/test_gpu_siblings.py tests/test_gpu_apollo_step.py -q -x --timeout 1500 2>&1 | tail -6
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1
python - <<PY
import sys, json, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
for crop in (108, 108, 148):
    ms, launches, _, model = bench.time_apollo_iterations(dev, crop, 10, 6, False)
    print("bench.time_apollo_iterations crop %d: %.2f ms" % (crop, ms))
    del model
PY
