#!/bin/bash
# remainder-pair kernel: bit-exact probe cases, kernel tests, per-layer timing with the kernel on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2e_probe.log
: > $LOG
P=build/probe_conv
for cap in 0 3; do
  for c in 1 2 4 9 10 11 12 13; do
    timeout 300 $P check $c $cap >> $LOG 2>&1 || echo "case $c cap $cap exit=$?" >> $LOG
  done
done
for s in 1 2 3 4 5; do
  NC_RP=1 timeout 120 $P time $s 9 5 >> $LOG 2>&1 || echo "time $s exit=$?" >> $LOG
  NC_RP=0 timeout 120 $P time $s 9 5 >> $LOG 2>&1 || echo "time $s exit=$?" >> $LOG
done
cat $LOG
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_pipeline.py -q --timeout 900 -x 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_dropin.py -q --timeout 1500 -s 2>&1 | tail -25 | cut -c1-400
