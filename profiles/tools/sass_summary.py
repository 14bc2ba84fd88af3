#!/usr/bin/env python
"""Static SASS summary of libwsb200.so (no GPU needed): per kernel the instruction count, registers / shared memory
from `cuobjdump -res-usage`, and the mnemonics that show how it moves data — UTMALDG (TMA box loads), SYNCS (mbarrier),
LDS / STS, LDG / STG, RED / ATOM, BAR — plus the fp32 instruction mix (FADD / FMUL / FFMA: -fmad=false leaves FFMA only
where the source calls fmaf or the division / sqrt sequences use it).

    python profiles/tools/sass_summary.py [lib.so] > profiles/r3_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "2d-weather-sandbox_b200", "csrc", "libwsb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2).strip()
KEYS = ["UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR", "FADD", "FMUL", "FFMA", "MUFU", "SHFL", "BRA"]
counts, total, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1).split(".")[0]
        for k in KEYS:
            if op.startswith(k):
                counts[fn][k] += 1
print(f"{os.path.relpath(so, ROOT)}: embedded cubins {arch}")
demangle = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
for fn, nice in sorted(zip(total, demangle), key=lambda t: -total[t[0]]):
    name = re.sub(r"\(.*", "", nice).replace("(anonymous namespace)::", "")
    print(f"\n{name}: {total[fn]} SASS instructions; {usage.get(fn, '')}")
    print("   " + "  ".join(f"{k} {counts[fn][k]}" for k in KEYS if counts[fn][k]))
