#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers, spills, shared memory per kernel (demangled)."""
import re, subprocess, sys
txt = sys.stdin.read()
names = re.findall(r"Compiling entry function '([^']+)'", txt)
blocks = re.split(r"ptxas info\s+: Compiling entry function", txt)[1:]
for name, b in zip(names, blocks):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("mswb::", "").replace("void ", "")
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    regs = re.search(r"Used (\d+) registers", b)
    smem = re.search(r"(\d+) bytes smem", b)
    print(f"{dem:70s} regs={regs.group(1) if regs else '?':>4s} spill={spill.group(1)+'/'+spill.group(2) if spill else '?':>9s} smem={smem.group(1) if smem else '0'}")
