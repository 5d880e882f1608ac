#!/usr/bin/env python
"""Summarise registers / spills per kernel from the ptxas -v logs written by flatnav_b200/csrc/Makefile."""
import glob
import os
import re
import subprocess
import sys

build = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "flatnav_b200", "csrc", "build")
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(\d+) bytes stack frame, "
                 r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
for f in sorted(glob.glob(os.path.join(build, "*.ptxas.log"))):
    for name, stack, ss, sl, regs in pat.findall(open(f).read()):
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(fnb::.*|\(unsigned.*|\(float.*", "", dem)
        print(f"{os.path.basename(f)[:20]:20s} regs={regs:>3s} stack={stack:>4s} spill={ss:>4s}/{sl:<4s} {dem[-64:]}")
