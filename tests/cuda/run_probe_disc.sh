#!/bin/bash
# A/B timing of the PatchGAN convolution kernels (see probe_disc.cu).  Output: gpurun_out/probe_disc.log
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
LOG=gpurun_out/probe_disc.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
if [ -x build/probe_disc_old ]; then echo "== previous revision" >> $LOG; timeout 300 build/probe_disc_old >> $LOG 2>&1; fi
if [ -x build/probe_disc_g2 ]; then echo "== tree, two pipelines per CTA, auto cluster" >> $LOG; timeout 300 build/probe_disc_g2 0 >> $LOG 2>&1; fi
for c in "$@"; do
  echo "== tree, cluster override $c" >> $LOG
  timeout 300 build/probe_disc $c >> $LOG 2>&1 || echo "exit=$?" >> $LOG
done
cat $LOG
