#!/bin/bash
# Runs the gradient-kernel probe on a GPU box; each case in its own process under a timeout so a hang or a sticky
# CUDA error cannot take the others down.  Output: gpurun_out/probe_grad.log
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
LOG=gpurun_out/probe_grad.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
P=build/probe_grad
for c in 0 1 2 3 4 5; do
  for f in "0 0" "1 1"; do
    timeout 120 $P wgrad $c $f >> $LOG 2>&1 || echo "wgrad case $c fmt $f exit=$?" >> $LOG
  done
done
for c in 0 1 2 3 5 10 11; do
  timeout 120 $P dgrad $c >> $LOG 2>&1 || echo "dgrad case $c exit=$?" >> $LOG
done
if [ "$1" != "notime" ]; then
  for c in 6 7 8 9; do
    timeout 120 $P time $c 5 >> $LOG 2>&1 || echo "time $c exit=$?" >> $LOG
  done
fi
cat $LOG
