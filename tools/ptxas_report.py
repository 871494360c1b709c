#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from `nvcc -Xptxas -v` (what `make ptxas-info` prints):
    make ptxas-info 2>&1 | python tools/ptxas_report.py [filter]"""
import re
import subprocess
import sys

flt = sys.argv[1] if len(sys.argv) > 1 else ""
text = sys.stdin.read()
rows = []
for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'(.*?)Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", text, re.S):
    name, body, regs, _bar, stack, smem = m.groups()
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", body)
    rows.append((name, int(regs), int(spill.group(1)) if spill else 0, int(spill.group(2)) if spill else 0, int(smem or 0)))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (raw, regs, ss, sl, smem), name in zip(rows, names):
    name = re.sub(r"\(gfb::EvalParams\)|gfb::|void ", "", name)
    if flt in name:
        print(f"{regs:4d} regs  spill {ss:3d}/{sl:3d} B  smem {smem:6d}  {name}")
