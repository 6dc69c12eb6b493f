#!/bin/bash
# Evidence run for the training path (profiles/r01_train_*): iteration benches, ncu launch list of the iteration,
# ncu --set full of the tensor-core / discriminator kernels inside one iteration, gradient-kernel exactness probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_apollo_step.py 108 10 2>/dev/null | tail -1 > gpurun_out/train_step_108.json; cat gpurun_out/train_step_108.json
python tools/bench_apollo_step.py 148 6 2>/dev/null | tail -1 > gpurun_out/train_step_148.json; cat gpurun_out/train_step_148.json
python tools/bench_train_step.py 108 5 2>/dev/null | tail -1 > gpurun_out/train_unet_108.json; cat gpurun_out/train_unet_108.json
python tools/profile_apollo_step.py 108 2>/dev/null | tail -11 > gpurun_out/train_phases_108.txt; cat gpurun_out/train_phases_108.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python tools/bench_apollo_step.py 108 1 > gpurun_out/train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/train_launches.csv > gpurun_out/train_launch_summary.txt 2>/dev/null; head -12 gpurun_out/train_launch_summary.txt
rm -f gpurun_out/train_launches.csv gpurun_out/*.ncu-rep
timeout 700 ncu --set full --clock-control none -k regex:'wgrad3d|conv3d_tc|conv2d_k4' -s 150 -c 70 -o gpurun_out/prof_train \
    python tools/bench_apollo_step.py 108 1 > gpurun_out/train_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_train.ncu-rep gpurun_out/train_ncu_kernels.json > gpurun_out/train_ncu_kernels.txt 2>&1
rm -f gpurun_out/prof_train.ncu-rep
head -5 gpurun_out/train_ncu_kernels.txt | cut -c1-300
bash tests/cuda/run_probe_grad.sh > /dev/null 2>&1; grep -c PASS gpurun_out/probe_grad.log; grep -c "FAIL\|exit=" gpurun_out/probe_grad.log
