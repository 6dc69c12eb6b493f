#!/bin/bash
# ncu --set full of the PatchGAN kernels: probe_disc in quick mode launches every layer / direction once (N = 1, 2);
# the capture is summarised on the box (tools/ncu_summary.py), the .ncu-rep is not brought back
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"conv2d_k4" -f -o /tmp/r2ab_disc build/probe_disc 0 quick > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r2ab_disc.ncu-rep gpurun_out/r2ab_ncu_disc_kernels.json > gpurun_out/r2ab_ncu_disc_kernels.txt 2>&1
cat gpurun_out/r2ab_ncu_disc_kernels.txt | cut -c1-250
