#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp-instructions) and top stall lines from an ncu report.
usage: ncu_sass_hist.py report.ncu-rep [top_n]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
S, E, W = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
print(rows[0][1][:150])
ops = collections.Counter(); tot = 0; body = []
for r in rows[hi + 1:]:
    if len(r) <= max(S, E, W) or not r[E].isdigit():
        continue
    n = int(r[E]); tot += n
    op = r[S].split()[0] if not r[S].strip().startswith("@") else r[S].split()[1]
    ops[op.split(".")[0]] += n
    body.append((int(r[W]) if r[W].isdigit() else 0, n, r[S].strip()))
print("total warp-instructions:", tot)
for op, n in ops.most_common(22):
    print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%")
print("top stall lines (samples, executed, sass):")
for w, n, s in sorted(body, reverse=True)[:topn]:
    print(f"  {w:6d} {n:10d}  {s[:110]}")
