"""Recipe for oracle/_ref/: the REFERENCE ITSELF, compiled, so that it can travel to the GPU box.
Test infrastructure — see oracle/__init__.py.      python -m oracle.build_ref

The reference is pure Python; its "build" is byte-compilation.  Every module of the path is compiled with
py_compile FROM THE SOURCES WHERE THEY LIE under /root/reference, and only the resulting bytecode (.pyc, legacy
sourceless layout: <package>/<module>.pyc) is written into oracle/_ref/neuroclear/.  No reference source enters the
repository: oracle/_ref/ is git-ignored (it still ships to the GPU box with the gpurun snapshot, like our own .so).
The interpreter on the GPU box is the same image's python, so the bytecode loads there.

Consumers (the only ones): oracle/reference_harness.py, which bench.py's `--impl reference` leg and the drop-in
tests use when oracle/_ref/ is present (`cpu_baseline.kind = "reference"`); when it is absent they fall back to the
oracle port (`kind = "port"`).
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

REFERENCE_ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "neuroclear")
PACKAGES = ["models", "data", "util", "options"]
SCRIPTS = ["test_dice.py", "train_onecube.py"]


def build(verbose: bool = False) -> str | None:
    """Returns the output directory, or None when /root/reference is absent (the GPU box: prebuilt files are used)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return None
    shutil.rmtree(OUT, ignore_errors=True)
    n = 0
    for pkg in PACKAGES:
        src_dir = os.path.join(REFERENCE_ROOT, pkg)
        for name in sorted(os.listdir(src_dir)):
            if name.endswith(".py"):
                os.makedirs(os.path.join(OUT, pkg), exist_ok=True)
                py_compile.compile(os.path.join(src_dir, name), cfile=os.path.join(OUT, pkg, name + "c"),
                                   dfile="reference/%s/%s" % (pkg, name), doraise=True, quiet=1)
                n += 1
    for name in SCRIPTS:
        py_compile.compile(os.path.join(REFERENCE_ROOT, name), cfile=os.path.join(OUT, name + "c"),
                           dfile="reference/" + name, doraise=True, quiet=1)
        n += 1
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write("byte-compiled from %s by oracle/build_ref.py with python %s; %d modules\n"
                % (REFERENCE_ROOT, sys.version.split()[0], n))
    if verbose:
        print("oracle/_ref: %d reference modules byte-compiled into %s" % (n, OUT))
    return OUT


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "models", "networks.pyc"))


if __name__ == "__main__":
    build(verbose=True)
