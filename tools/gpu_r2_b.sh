#!/bin/bash
# round 2, visit B (2 GPUs): new f3 tests, multi-GPU test, bench at N=1 (short) and N=2 -> out_sha256 equality, train_step_dp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ls -la oracle/_ref
python -m pytest tests/test_gpu_postprocess.py tests/test_gpu_multi.py tests/test_gpu_kernels.py tests/test_gpu_unet.py -q --timeout 1200 > gpurun_out/r2b_pytest.log 2>&1; tail -25 gpurun_out/r2b_pytest.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -2 gpurun_out/r2b_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -3 gpurun_out/r2b_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref_n2.json 2> gpurun_out/r2b_bench_ref_n2.err; tail -3 gpurun_out/r2b_bench_ref_n2.err
python tools/library_bar.py --json > gpurun_out/r2b_library_bar.json 2> gpurun_out/r2b_library_bar.err
python - <<PY
import json
for f in ("r2b_bench_n1","r2b_bench_n2","r2b_bench_ref_n2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d.get("out_sha256"), d.get("e2e",{}).get("ms_per_step"), json.dumps(d.get("train_step_dp")), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/r2b_library_bar.json").read()[-900:])
PY
