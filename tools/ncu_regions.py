#!/usr/bin/env python
"""Executed warp-instructions of inter_warp_kernel grouped by code region (marker strings in the source).
Usage: ncu_regions.py rep.ncu-rep obj.o source.cu n_records [kernel-filter [section-substring [nth]]]
section-substring: part of the mangled name of the instantiation (an object with several kernels has several .text sections
that all start at address 0); nth: which table of the report's source page to take (each launch appears twice)."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, srcpath, nrec = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
kernel = sys.argv[5] if len(sys.argv) > 5 and sys.argv[5] != "-" else None      # optional ncu --kernel-name filter for reports with several kernels
fsub = sys.argv[6] if len(sys.argv) > 6 else None
nth = int(sys.argv[7]) if len(sys.argv) > 7 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + (["--kernel-name", kernel] if kernel else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][nth]
print("kernel:", rows[hdr - 1][1][:110])
h = rows[hdr]
ii = h.index("Instructions Executed")
recs = []
for r in rows[hdr + 1:]:
    if r and r[0] == "Kernel Name":          # several launches: the first one only
        break
    if r and r[0] and len(r) > ii:
        recs.append((int(r[0], 16), int(r[ii] or 0)))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    dis = subprocess.run(["nvdisasm", "--print-line-info", glob.glob(td + "/*.cubin")[0]], capture_output=True, text=True).stdout
base_name = os.path.basename(srcpath)
line_at, cur, last_own, in_text = {}, None, None, False
for ln in dis.splitlines():
    if ln.lstrip().startswith(".section"):
        in_text = ".text." in ln and (fsub is None or fsub in ln)
        continue
    if not in_text:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        if cur[0] == base_name:
            last_own = cur[1]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", ln)
    if m and ".byte" not in ln:
        line_at[int(m.group(1), 16)] = last_own       # helpers/intrinsics are charged to the last own-source line
src = open(srcpath).read().split('\n')
marks = [(1, 'head')]
for i, l in enumerate(src):
    m = re.search(r'//@region (.+)$', l)
    if m:
        marks.append((i + 1, m.group(1).strip()))


def region(l):
    r = '?'
    for a, n in marks:
        if l is not None and l >= a:
            r = n
    return r


agg = collections.Counter()
base = recs[0][0]
for a, n in recs:
    agg[region(line_at.get(a - base))] += n
tot = sum(agg.values())
for k, v in agg.most_common():
    print("%-28s %12d %6.2f%%  %7.1f /record" % (k, v, 100 * v / tot, v / nrec))
print("total %d = %.1f /record" % (tot, tot / nrec))
