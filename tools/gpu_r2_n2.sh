#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 900 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n2.json"))
print("N=2 value %.4g ms/step %.1f e2e %.1f sha %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["out_sha256"][:16]))
print("train_step_dp", json.dumps(d["train_step_dp"]))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-400
