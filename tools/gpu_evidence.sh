#!/bin/bash
# Evidence run for profiles/: GPU tests, probe, smoke, default bench, reference arm, ncu launch list of the bench
# command, ncu --set full of every kernel of one 8-cube pass, compute-sanitizer memcheck of a small pipeline run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
bash tests/cuda/run_probe.sh > /dev/null 2>&1; grep -c PASS gpurun_out/probe.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; echo
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none -s 40 -c 40 -o gpurun_out/prof_all python bench.py --size 128 128 128 --batch 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep gpurun_out/ncu_all_kernels.json > gpurun_out/ncu_all_kernels.txt 2>&1
rm -f gpurun_out/prof_all.ncu-rep   # > 64 MiB: only the summary travels back
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/sanitizer.log 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
from neuroclear_b200.pipeline import DicedInference
from oracle import unet as ounet
vol = (np.random.default_rng(0).random((40, 41, 58)) ** 3 * 65535).astype(np.uint16)
got, _ = DicedInference(ounet.random_state_dict(0, 0.1), "cuda:0", 24, 6, 4, batch=4).run(vol)
print("sanitized run ok", got.shape, int(got.max()))
PY
echo "sanitizer exit $?"; tail -4 gpurun_out/sanitizer.log
ls -la gpurun_out | head -30
