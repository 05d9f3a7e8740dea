#!/usr/bin/env python
"""Like ncu_lines.py, for reports that hold several kernels whose names do not survive ncu's --kernel-name filter:
per-source-line executed instructions / stall samples / shared-memory conflicts of the N-th kernel of the report.
Usage: ncu_lines_nth.py rep.ncu-rep build/x.o N function-substring [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, nth, fsub = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
tables = []
for si, st in enumerate(starts):
    h = rows[st]
    end = starts[si + 1] if si + 1 < len(starts) else len(rows)
    body = [r for r in rows[st + 1:end] if r and re.match(r"^(0x)?[0-9a-fA-F]+$", r[0]) and len(r) == len(h)]
    if body:
        tables.append((h, body, rows[st - 1][1] if st else ""))
h, body, name = tables[nth]
print("kernel:", name[:100])
col = {k: h.index(k) for k in ("Address", "Instructions Executed", "# Samples", "stall_no_inst", "stall_long_sb", "stall_short_sb")}
bank = h.index("L1 Conflicts Shared N-Way") if "L1 Conflicts Shared N-Way" in h else None
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    dis = subprocess.run(["nvdisasm", "--print-line-info", glob.glob(td + "/*.cubin")[0]], capture_output=True, text=True).stdout
cur, line_at, want = None, {}, False
for ln in dis.splitlines():
    if ln.lstrip().startswith(".section"):
        want = ".text." in ln and fsub in ln
        continue
    if not want:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", ln)
    if m and ".byte" not in ln:
        line_at[int(m.group(1), 16)] = cur
base = int(body[0][col["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0.0])
for r in body:
    v = agg[line_at.get(int(r[col["Address"]], 16) - base)]
    v[0] += int(r[col["Instructions Executed"]] or 0); v[1] += int(r[col["# Samples"]] or 0); v[2] += int(r[col["stall_no_inst"]] or 0)
    v[3] += int(r[col["stall_long_sb"]] or 0); v[4] += int(r[col["stall_short_sb"]] or 0); v[5] += 1
    v[6] += float(r[bank] or 0) if bank is not None else 0
tot = [sum(v[j] for v in agg.values()) for j in range(7)]
print("sass %d, executed %d, samples %d (no_inst %d, long_sb %d, short_sb %d), shared conflicts %.0f" % (len(body), tot[0], tot[1], tot[2], tot[3], tot[4], tot[6]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s sass %4d exec %10d %5.2f%% samples %6d no_inst %6d long %6d short %6d conflicts %9.0f" % (
        "%s:%d" % k if k else "?", v[5], v[0], 100.0 * v[0] / max(tot[0], 1), v[1], v[2], v[3], v[4], v[6]))
