#!/bin/bash
# one GPU-box visit: tests, probe, bench, ncu launch list, ncu full capture of the conv kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
bash tests/cuda/run_probe.sh > /dev/null 2>&1; grep -c PASS gpurun_out/probe.log; grep TIME gpurun_out/probe.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 9 -c 9 -o gpurun_out/prof_conv python bench.py --size 128 128 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
