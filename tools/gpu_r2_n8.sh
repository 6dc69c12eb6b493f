#!/bin/bash
# 8 GPUs: the full bench line (900^3 + config 5 + data-parallel training at 148^3) and the multi-GPU test worker
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -3 gpurun_out/r2_bench_n8.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n8.json"))
print("N=8 value %.4g ms/step %.1f e2e %.1f sha %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["out_sha256"][:16]))
print("config5", json.dumps(d["config5"])[:600])
print("train_step_dp", json.dumps(d["train_step_dp"]))
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 900 2>&1 | tail -3
