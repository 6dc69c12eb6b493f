"""Builds libneuroclear_b200.so in-tree with nvcc for sm_100a:  python -m neuroclear_b200.build [--force]"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SOURCES = ["runtime.cu", "conv3d_tc.cu", "wgrad3d_tc.cu", "conv1_tc.cu", "elementwise.cu", "unet_bwd.cu", "deeplinear.cu", "disc2d.cu", "patchgan.cu", "augment.cu", "api.cu"]
HEADERS = ["internal.h", "ptx.cuh", "augment_math.h"]
OUT = os.path.join(_HERE, "libneuroclear_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(_HERE, "csrc", f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(_ROOT, "include", "neuroclear_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(_HERE, "csrc", s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return OUT


def build_probe() -> str:
    """The standalone hardware probes (tests/cuda/probe_conv.cu, probe_grad.cu) used by tests/cuda/run_probe*.sh."""
    out_dir = os.path.join(_ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    for name in ("probe_conv", "probe_grad"):
        out = os.path.join(out_dir, name)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler",
                        "-fopenmp", "-o", out, os.path.join(_ROOT, "tests", "cuda", name + ".cu"), "-L", _HERE,
                        "-lneuroclear_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../neuroclear_b200"],
                       check=True)
    return os.path.join(out_dir, "probe_conv")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
