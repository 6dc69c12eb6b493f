#!/bin/bash
# ncu --set full of the PatchGAN conv kernels on the 256 -> 512 layer (probe_disc), cluster 1 and 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1 8; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2d_k4_fwd_kernel -s 180 -c 1 -f -o gpurun_out/r2x_disc_fwd_c$c build/probe_disc $c > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2d_k4_dgrad_kernel -s 120 -c 1 -f -o gpurun_out/r2x_disc_dgrad_c8 build/probe_disc 8 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
