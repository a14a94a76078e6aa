#!/usr/bin/env python
"""Developer tool: compile accumulate_fast.cu with -Xptxas -v and print registers / spills per
kernel instantiation.  Usage: python scripts/ptxas_report.py [extra nvcc flags...]"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = ROOT / "isce3_b200" / "csrc" / "accumulate_fast.cu"
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
       "-std=c++17", "-Xcompiler", "-fPIC,-O2", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v", "-c", str(src),
       "-o", "/tmp/af_report.o"] + sys.argv[1:]
r = subprocess.run(cmd, capture_output=True, text=True)
txt = r.stderr + r.stdout
if r.returncode:
    print(txt)
    sys.exit(1)
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = m.group(1)
        k = re.search(r"kernelILi(\d+)ELi(\d+)ENS_(\d+)(CoefBank|CoefImmILi\d+E)", name)
        cur = f"K={k.group(1)} D={k.group(2)} {k.group(4)}" if k else name[:60]
        spill = None
        continue
    if cur and "spill" in line and spill is None:
        spill = line.strip()
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        print(f"{cur:28s} regs={m.group(1)}  {spill}")
        cur = None
