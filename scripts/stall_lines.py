#!/usr/bin/env python
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines.
    python scripts/stall_lines.py gpurun_out/prof_X.ncu-rep <object file> <kernel substring> <source file>"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kernel, source = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sections = {}
cur_sec, cur_line = None, None
for line in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        cur_sec = m.group(1); sections[cur_sec] = {}; continue
    if cur_sec is None:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        sections[cur_sec][int(m.group(1), 16)] = cur_line
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
src_lines = open(source).read().splitlines()
for bi, b in enumerate(starts[:1]):
    name = rows[b][1]
    sec = [s for s in sections if kernel in s and all(tok in s for tok in re.findall(r"ILi\d+E", ""))]
    header = rows[b + 1]
    end = starts[bi + 1] if bi + 1 < len(starts) else len(rows)
    body = [r for r in rows[b + 2:end] if len(r) == len(header)]
    si = header.index("Warp Stall Sampling (All Samples)")
    base = int(body[0][0], 16)
    # pick the section whose instruction count matches
    cands = [s for s in sections if kernel in s and len(sections[s]) == len(body)]
    table = sections[cands[0]] if cands else {}
    counts = collections.Counter()
    total = 0
    for r in body:
        n = int(r[si]); total += n
        counts[table.get(int(r[0], 16) - base)] += n
    print(f"{name}: {total} samples, section {cands[0] if cands else '??'}")
    for key, n in counts.most_common(top):
        text = src_lines[key[1] - 1].strip()[:105] if key and key[0] == os.path.basename(source) else str(key)
        print(f"{100 * n / total:5.1f}%  {key[0] if key else '-'}:{key[1] if key else 0:<4d} {text}")
