#!/bin/bash
# end-of-round evidence on one GPU: full GPU suite, smoke, the bench line as the driver runs it, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -m pytest tests -q -m gpu --timeout 1500 > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | grep "smoke ok"
python bench.py --steps 10 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("N=1 value %.4g ms/step %.1f e2e %.1f conv %.0f TF/s frac %.3f share %.3f whole %.3f clk %s sha %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["tensor_pipe_frac_whole_step"], d["clocks"]["sm_mhz"], d["out_sha256"][:16]))
print("train_step", json.dumps(d["train_step"])[:400])
print("train_step_dp", json.dumps(d["train_step_dp"])[:300])
print("library_bar", json.dumps(d["library_bar"])[:1200])
print("cpu_baseline", d["cpu_baseline"])
r=json.load(open("gpurun_out/r2_bench_reference.json")); print("reference", r["value"], r["ms_per_step"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
PY
