#!/bin/bash
# compute-sanitizer over the late round-2 kernels: memcheck on the whole-pipeline workload (two apollo iterations
# incl. every PatchGAN kernel), memcheck + racecheck on the discriminator probe (every layer / direction once, cluster
# sizes auto and 8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitizer_workload.py > gpurun_out/r2_memcheck_late.log 2>&1
grep -E "ERROR SUMMARY|Invalid|inference|assembly|projections|psnr|training|Error" gpurun_out/r2_memcheck_late.log | head -20
for c in 0 8; do
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 build/probe_disc $c quick > gpurun_out/r2_memcheck_probe_disc_c$c.log 2>&1
  grep -E "ERROR SUMMARY" gpurun_out/r2_memcheck_probe_disc_c$c.log
  timeout 900 compute-sanitizer --tool racecheck --print-limit 20 build/probe_disc $c quick > gpurun_out/r2_racecheck_probe_disc_c$c.log 2>&1
  grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r2_racecheck_probe_disc_c$c.log | head -5
done
