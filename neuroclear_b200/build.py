"""Builds libneuroclear_b200.so in-tree with nvcc for sm_100a:  python -m neuroclear_b200.build [--force]"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SOURCES = ["runtime.cu", "conv3d_tc.cu", "wgrad3d_tc.cu", "conv1_tc.cu", "elementwise.cu", "unet_bwd.cu", "deeplinear.cu", "disc2d.cu", "patchgan.cu", "augment.cu", "postprocess.cu", "unet_infer.cu", "api.cu"]
HEADERS = ["internal.h", "ptx.cuh", "augment_math.h"]
OUT = os.path.join(_HERE, "libneuroclear_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def source_hash() -> str:
    """sha256 over every source, header and the compiler flags: what the binary is a function of."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(_HERE, "csrc", f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(_ROOT, "include", "neuroclear_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def binary_hash(path: str = OUT):
    """The digest embedded in the shared object (nc_build_source_hash), read from the file's bytes (no dlopen)."""
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    i = blob.find(b"NC_SOURCE_HASH=")
    if i < 0:
        return None
    return blob[i + 15:i + 15 + 64].decode("ascii", "replace")


def _stale() -> bool:
    """True when the .so is missing or was compiled from different sources / flags than the tree holds."""
    return binary_hash() != source_hash()


def _unit_hash(src: str) -> str:
    """digest of one translation unit: its source, every header, the flags"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in [os.path.join(_HERE, "csrc", f) for f in [src] + HEADERS] + [os.path.join(_ROOT, "include", "neuroclear_b200.h")]:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel, objects cached by content digest under build/obj) and link the
    shared object.  Rebuilds whenever the digest embedded in the .so differs from the digest of the tree."""
    if not force and not _stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = os.path.join(_ROOT, "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    digest = source_hash()
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        extra = ['-DNC_SOURCE_HASH="%s"' % digest] if src == "api.cu" else []
        tag = _unit_hash(src) + ("-" + digest[:16] if extra else "")
        obj = os.path.join(obj_dir, "%s.%s.o" % (src[:-3], tag))
        if force or not os.path.exists(obj):
            for old in os.listdir(obj_dir):                       # drop objects of earlier versions of this unit
                if old.startswith(src[:-3] + ".") and old.endswith(".o"):
                    os.remove(os.path.join(obj_dir, old))
            cmd = [nvcc] + compile_flags + extra + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-c", "-o", obj + ".tmp", os.path.join(_HERE, "csrc", src)]
            subprocess.run(cmd, check=True)
            os.replace(obj + ".tmp", obj)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                    "-o", OUT + ".tmp"] + objs, check=True)
    os.replace(OUT + ".tmp", OUT)
    if binary_hash() != digest:
        raise RuntimeError("libneuroclear_b200.so does not carry the digest of the sources it was built from")
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
