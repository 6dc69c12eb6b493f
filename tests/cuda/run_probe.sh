#!/bin/bash
# Runs the tcgen05 convolution probe on a GPU box; each case in its own process under a timeout so a hang or
# a sticky CUDA error cannot take the others down.  Output: gpurun_out/probe.log
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
P=build/probe_conv
for cap in 0 2; do
  for c in 0 1 2 3 4 5 6 7 8 9 10 11; do
    timeout 120 $P check $c $cap >> $LOG 2>&1 || echo "case $c cap $cap exit=$?" >> $LOG
  done
done
if [ "$1" != "notime" ]; then
  for s in 0 1 2 3 4 5 6 7 8; do
    timeout 120 $P time $s 1 5 >> $LOG 2>&1 || echo "time $s exit=$?" >> $LOG
  done
  for s in 2 4 5; do
    timeout 120 $P time $s 8 3 >> $LOG 2>&1 || echo "time $s nb8 exit=$?" >> $LOG
  done
fi
cat $LOG
