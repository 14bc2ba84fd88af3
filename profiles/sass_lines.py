#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction counters to CUDA source lines.

    python profiles/sass_lines.py <report.ncu-rep> <kernel-substring> [<lib.so>] [--top N]

ncu's `--page source --csv` lists SASS instructions with executed counts but no line numbers;
`nvdisasm -g` lists the same instructions (same order) with `//## File ..., line N` markers.
This joins the two by instruction order and prints the hottest source lines."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2d-weather-sandbox_b200", "csrc", "libwsb200.so")
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = next(b for b in blocks if kern in b["name"])
hdr = blk["rows"][0]
iS, iI, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[iS].strip(), int(r[iI]), int(r[iSamp] or 0)) for r in blk["rows"][1:] if len(r) > iI and r[iI].isdigit()]

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, infunc, loc = [], False, ("?", 0)
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infunc = kern in m.group(1)
        continue
    if not infunc:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(loc)
if len(lines) != len(ins):
    print(f"warning: {len(lines)} disassembled vs {len(ins)} profiled instructions; joining the common prefix", file=sys.stderr)
agg, samp = collections.Counter(), collections.Counter()
tot = sum(n for _, n, _ in ins)
for (src, n, sm), l in zip(ins, lines):
    agg[l] += n
    samp[l] += sm
print(f"{blk['name'][:80]}: {tot} warp instructions, {len(ins)} static")
srcs = {}
for (f, l), n in agg.most_common(top):
    if f not in srcs:
        p = os.path.join(os.path.dirname(so), f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{n / tot * 100:5.1f}%  samples {samp[(f, l)]:6d}  {f}:{l:<4d} {text}")
