#!/usr/bin/env python
"""Counts the Blackwell-specific SASS instructions in libneuroclear_b200.so, per kernel — the static proof that the
hot kernels are tcgen05 / TMEM / TMA code (mnemonics from /opt/skills/guides/B200_PROFILING.md):

    UTCHMMA / UTCQMMA   tcgen05.mma (5th-gen tensor core, accumulators in TMEM)
    LDTM / STTM         tcgen05.ld / tcgen05.st (TMEM <-> registers)
    UTMALDG / UTMASTG   cp.async.bulk.tensor load / store (TMA)
    UBLKCP              cp.async.bulk (1-D bulk copy)
    UTCBAR              tcgen05.commit -> mbarrier
    SYNCS               mbarrier arrive / try_wait
    HMMA / IMMA         legacy mma.sync (should be absent from the conv kernels)

    python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neuroclear_b200", "libneuroclear_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    per[cur][mn] += 1
    try:
        names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    except OSError:
        names = list(per)
    total = collections.Counter()
    rows = []
    for (mangled, c), name in zip(per.items(), names):
        total.update(c)
        if any(c[m] for m in MNEMONICS):
            rows.append((re.sub(r"(?<!::)\(.*", "", name)[:78], c))
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print("# %s: %d kernels, %d SASS instructions, arch %s" % (os.path.relpath(lib, ROOT), len(per), total["_total"],
                                                            ",".join(sorted(set(arch)))))
    print("# build digest: %s" % _digest(lib))
    print("%-78s %s" % ("kernel", " ".join("%8s" % m for m in MNEMONICS)))
    for name, c in rows:
        print("%-78s %s" % (name, " ".join("%8d" % c[m] for m in MNEMONICS)))
    print("%-78s %s" % ("TOTAL (all %d kernels)" % len(per), " ".join("%8d" % total[m] for m in MNEMONICS)))


def _digest(lib):
    with open(lib, "rb") as f:
        blob = f.read()
    i = blob.find(b"NC_SOURCE_HASH=")
    return blob[i + 15:i + 79].decode("ascii", "replace") if i >= 0 else "unstamped"


if __name__ == "__main__":
    main()
