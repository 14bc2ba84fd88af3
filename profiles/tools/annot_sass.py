#!/usr/bin/env python
"""Static view of one kernel's SASS with source-line annotations (no GPU needed):

    python profiles/tools/annot_sass.py <kernel-substring> [<lib.so>] > out.txt

Each line: address, source file:line, instruction; labels are kept, so loop bodies can be counted."""
import os, re, subprocess, sys, tempfile
kern = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                       "2d-weather-sandbox_b200", "csrc", "libwsb200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
on, loc = False, ("?", 0)
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        on = kern in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (os.path.basename(m.group(1)).replace("wsb_", "").replace(".cuh", ""), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        print(f"{m.group(1)} {loc[0]}:{loc[1]:<4} {m.group(2)}")
    elif re.match(r"^\.L_x_\d+:", ln) or re.match(r"^\$", ln):
        print(ln.strip())
