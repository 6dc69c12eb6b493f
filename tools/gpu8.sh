#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -q --timeout 900 2>&1 | tail -2
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -c 400 gpurun_out/bench_n$n.err | tail -2
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$n.json")); print("N=$n value %.4g ms/step %.1f e2e %.4g ms %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done
# BASELINE config 5: 2048 x 2048 x 1024 (Z,Y,X) = (1024,2048,2048), 4000 cubes, 8 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 1 --warmup 1 --size 1024 2048 2048 --batch 10 > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
tail -2 gpurun_out/bench_c5_n8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c5_n8.json")); print("config5 N=8 value %.4g ms/step %.1f e2e ms %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
