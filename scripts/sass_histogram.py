#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): the mnemonics that prove which hardware
paths a kernel uses -- UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP
(cp.async.bulk), SYNCS (mbarrier), MUFU (SFU), RED / ATOM, BAR.

    python scripts/sass_histogram.py [library.so] > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "diffusion_for_multi_scale_molecular_dynamics_b200", "csrc", "libb200md.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "ARRIVES", "SYNCS", "MUFU", "RED", "REDG", "ATOM", "ATOMG", "ATOMS", "BAR",
         "FFMA", "FFMA2", "FMUL2", "FADD2", "HFMA2", "LDG", "STG", "LDS", "STS", "SHFL", "USETMAXREG", "ELECT"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        kernels[name][m.group(1)] += 1

print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (sm_100a); instructions per kernel, static counts")
for name, counts in kernels.items():
    total = sum(counts.values())
    shown = "  ".join(f"{op}={counts[op]}" for op in WATCH if counts[op])
    print(f"{name}: {total} instructions\n    {shown}")
