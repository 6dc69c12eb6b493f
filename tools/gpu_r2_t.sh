#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_zz_volume_io.py tests/test_gpu_multi.py -q --timeout 900 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -2 gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2t_bench.json")); print("N=1 ms/step %.1f e2e %.1f  sha %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["out_sha256"][:16]))
PY
