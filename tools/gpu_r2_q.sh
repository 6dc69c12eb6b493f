#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_cin1 -s 1 -c 1 -f -o gpurun_out/r2q_conv1 python bench.py --size 256 256 256 --batch 9 --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2q_ncu1.log 2>&1
ls -la gpurun_out/*.ncu-rep
