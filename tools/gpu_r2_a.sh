#!/bin/bash
# round 2, visit A: full GPU suite, smoke, bench line with all the new keys, reference arm, library bar
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi -L
python -m pytest tests -q -m gpu --timeout 1200 -x > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 2500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err; tail -c 1200 gpurun_out/r2a_bench_reference.json
