#!/usr/bin/env python
"""Trim a compute-sanitizer log for profiles/: the case lines, the summaries, and the first instance of every distinct report
(same tool message at the same source line), without the host backtraces."""
import re
import sys

seen, out, block = set(), [], []


def flush():
    if not block:
        return
    head = block[0]
    where = next((l for l in block[1:] if " at " in l), "")
    key = (re.sub(r"0x[0-9a-f]+|\[\d+ hazards\]|thread \(\d+,\d+,\d+\)|block \(\d+,\d+,\d+\)", "", head),
           re.sub(r"\+0x[0-9a-f]+", "", where))
    if key not in seen:
        seen.add(key)
        out.extend(l for l in block if "Host Frame" not in l and "Saved host backtrace" not in l)
    block.clear()


for line in open(sys.argv[1]):
    line = line.rstrip("\n")
    if not line.startswith("========="):
        flush()
        out.append(line)
    elif re.match(r"=========\s*$", line):
        flush()
    elif re.match(r"========= (ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER|Target application)", line):
        flush()
        out.append(line)
    elif re.match(r"========= \S", line) and not line.startswith("=========     "):
        flush()
        block.append(line)
    else:
        block.append(line)
flush()
print("\n".join(out))
