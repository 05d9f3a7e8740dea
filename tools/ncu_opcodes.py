#!/usr/bin/env python
"""Executed warp instructions of one kernel of an .ncu-rep by SASS opcode (straight from the report's SASS page: no
object file, no line mapping).  Usage: ncu_opcodes.py rep.ncu-rep kernel-filter [top]   (filter: ncu --kernel-name syntax)"""
import collections
import csv
import re
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ii = h.index("Instructions Executed")
ops, tot, n_sass = collections.Counter(), 0, 0
for r in rows[hdr + 1:]:
    if r and r[0] == "Kernel Name":          # several launches: the first one only
        break
    if len(r) <= ii or not r[0]:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
    n = int(r[ii] or 0)
    ops[m.group(2) if m else r[1].strip()] += n
    tot += n
    n_sass += 1
print("kernel: %s" % rows[hdr - 1][1])
print("SASS instructions %d, executed warp instructions %d" % (n_sass, tot))
for k, v in ops.most_common(top):
    print("%-10s %12d %6.2f %%" % (k, v, 100.0 * v / tot))
