#!/bin/bash
# ncu --set full: one forward batch (8 cubes of a 128^3 volume = every kernel of the network incl. the remainder pairs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv|in_relu|head|in_stats' -s 28 -c 30 -f -o gpurun_out/r2j_full python bench.py --size 128 128 128 --batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar --no-train-sample > gpurun_out/r2j_ncu.log 2>&1
ls -la gpurun_out/r2j_full.ncu-rep
python tools/ncu_summary.py gpurun_out/r2j_full.ncu-rep gpurun_out/r2j_ncu_kernels.json > gpurun_out/r2j_ncu_kernels.txt 2>&1
cat gpurun_out/r2j_ncu_kernels.txt | cut -c1-230
