#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitizer_workload.py > gpurun_out/r2_memcheck.log 2>&1
grep -E "ERROR SUMMARY|Invalid|inference|assembly|projections|psnr|training|Error" gpurun_out/r2_memcheck.log | head -20
