#!/usr/bin/env python
"""Registers / spills per kernel from an `nvcc -Xptxas -v` log (argument: the log file)."""
import re
import subprocess
import sys

t = open(sys.argv[1]).read()
for b in re.split(r"ptxas info\s+: Compiling entry function ", t)[1:]:
    name = b.split("'")[1]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    m = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
    if len(sys.argv) < 3 or sys.argv[2] in dem:
        print(dem[:100], m.group(1) if m else "?", sp.group(0) if sp else "")
