#!/usr/bin/env python
"""Registers / spills / static smem of every k_fused_step instantiation, from the make logs (csrc/*.ptxas.log).
   python tools/ptxas_summary.py [substring filter]"""
import glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "swalbe.jl_b200", "csrc", "*.ptxas.log"))):
    t = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers, used (\d+) barriers(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", t):
        rows.append((m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(4)), int(m.group(7) or 0)))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
for (mangled, regs, ss, sl, smem), name in zip(rows, names):
    name = name.replace("swalbe::", "").replace("(swalbe::FusedArgs)", "").replace("void ", "")
    if flt in name:
        print(f"{regs:4d} regs  spill {ss:4d}/{sl:4d} B  smem {smem:5d}  {name}")
