#!/bin/bash
# Builds the discriminator probe twice (run here, nvcc cross-compiles): build/probe_disc from the tree and
# build/probe_disc_old from a previous revision of csrc/disc2d.cu (default: 7ccbe2c, the FFMA kernels of round 2), so
# that tests/cuda/run_probe_disc.sh measures both in one GPU call.   usage: build_probe_disc.sh [git revision]
cd "$(dirname "$0")/../.."
REV=${1:-7ccbe2c}
mkdir -p build/old
git show "$REV":neuroclear_b200/csrc/disc2d.cu > build/old/disc2d_old.cu || exit 1
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -Ineuroclear_b200/csrc"
nvcc $F -o build/probe_disc tests/cuda/probe_disc.cu || exit 1
nvcc $F -DOLD -o build/probe_disc_old tests/cuda/probe_disc.cu || exit 1
ls -la build/probe_disc build/probe_disc_old
