"""Recipe for oracle/_ref/: the REFERENCE ITSELF, compiled, so that it can travel to the GPU box.
Test infrastructure — see oracle/__init__.py.      python -m oracle.build_ref

The reference is pure Python; its "build" is byte-compilation.  Every module of the path is compiled with
py_compile FROM THE SOURCES WHERE THEY LIE under /root/reference, and only the resulting bytecode (legacy sourceless
layout: <package>/<module>.pyc) is written — into ONE archive, oracle/_ref/neuroclear.zip, which python imports
directly (zipimport; loose .pyc files do not survive the gpurun snapshot).  No reference source enters the repository:
oracle/_ref/ is git-ignored (it still ships to the GPU box with the gpurun snapshot, like our own .so).  The
interpreter on the GPU box is the same image's python, so the bytecode loads there.

Consumers (the only ones): oracle/reference_harness.py, which bench.py's `--impl reference` leg and the drop-in
tests use when oracle/_ref/ is present (`cpu_baseline.kind = "reference"`); when it is absent they fall back to the
oracle port (`kind = "port"`).
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

REFERENCE_ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "neuroclear.zip")
PACKAGES = ["models", "data", "util", "options"]
SCRIPTS = ["test_dice.py", "train_onecube.py"]


def build(verbose: bool = False) -> str | None:
    """Returns the archive path, or None when /root/reference is absent (the GPU box: the prebuilt archive is used)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    shutil.rmtree(os.path.join(os.path.dirname(OUT), "neuroclear"), ignore_errors=True)     # the round-2 loose layout
    n = 0
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(OUT + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        def add(src, arc):
            nonlocal n
            cfile = os.path.join(tmp, "m%d.pyc" % n)
            py_compile.compile(src, cfile=cfile, dfile="reference/" + arc[:-1], doraise=True, quiet=1)
            z.write(cfile, arc)
            n += 1
        for pkg in PACKAGES:
            src_dir = os.path.join(REFERENCE_ROOT, pkg)
            for name in sorted(os.listdir(src_dir)):
                if name.endswith(".py"):
                    add(os.path.join(src_dir, name), "%s/%sc" % (pkg, name))
        for name in SCRIPTS:
            add(os.path.join(REFERENCE_ROOT, name), name + "c")
        z.writestr("BUILD_INFO", "byte-compiled from %s by oracle/build_ref.py with python %s; %d modules\n"
                   % (REFERENCE_ROOT, sys.version.split()[0], n))
    os.replace(OUT + ".tmp", OUT)
    if verbose:
        print("oracle/_ref: %d reference modules byte-compiled into %s" % (n, OUT))
    return OUT


def available() -> bool:
    return os.path.exists(OUT)


if __name__ == "__main__":
    build(verbose=True)
