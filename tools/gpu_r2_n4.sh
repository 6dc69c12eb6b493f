#!/bin/bash
cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; tail -2 gpurun_out/r2_bench_n4.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n4.json"))
print("N=4 value %.4g ms/step %.1f e2e %.1f sha %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["out_sha256"][:16]))
print("train_step_dp", json.dumps(d["train_step_dp"])[:500])
PY
