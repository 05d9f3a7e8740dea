#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample totals of one kernel in an .ncu-rep.
Joins `ncu --page source --print-source sass` (per SASS address) with `nvdisasm --print-line-info`
of the object the kernel was built from.  Usage: ncu_lines.py rep.ncu-rep build/x.o [top [kernel-filter [section-substring]]]
An object that holds several kernels (template instantiations) needs the section substring - part of the mangled name, e.g.
itx_warp_kernelILi2E - or the sections' addresses, which all start at 0, are merged and the attribution is wrong."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile


def main(rep, obj, top=40, kernel=None, fsub=None):
    # kernel: optional name filter (ncu --kernel-name syntax, e.g. regex:itx_warp) for reports holding several kernels
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + (["--kernel-name", kernel] if kernel else []),
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ia, ii, isamp, ino = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("stall_no_inst")
    ilong, ishort, iwait = h.index("stall_long_sb"), h.index("stall_short_sb"), h.index("stall_wait")
    per_addr = []
    for r in rows[hdr + 1:]:
        if r and r[0] == "Kernel Name":         # a report with several launches of the kernel: the first one only
            break
        if len(r) <= ino or not r[ia]:
            continue
        per_addr.append((int(r[ii] or 0), int(r[isamp] or 0), int(r[ino] or 0), int(r[ilong] or 0), int(r[ishort] or 0), int(r[iwait] or 0), r[1], int(r[ia], 16)))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        cubin = glob.glob(os.path.join(td, "*.cubin"))[0]
        dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True, check=True).stdout
    line_at, cur, in_text = {}, None, False
    for ln in dis.splitlines():
        if ln.lstrip().startswith(".section"):
            in_text = ".text." in ln and (fsub is None or fsub in ln)
            continue
        if not in_text:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", ln)
        if m and ".byte" not in ln and ".dword" not in ln and ".short" not in ln:
            line_at[int(m.group(1), 16)] = cur
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
    base = per_addr[0][7]
    for rec in per_addr:
        a = agg[line_at.get(rec[7] - base)]
        for j in range(6):
            a[j] += rec[j]
        a[6] += 1
    line_of = line_at
    tot = [sum(v[j] for v in agg.values()) for j in range(7)]
    print("sass instrs %d (disasm %d), executed %d, samples %d (no_inst %d, long_sb %d, short_sb %d, wait %d)" % (
        len(per_addr), len(line_of), tot[0], tot[1], tot[2], tot[3], tot[4], tot[5]))
    print("%-28s %6s %12s %6s %8s %8s %8s %8s" % ("line", "sass", "executed", "%", "samples", "no_inst", "long_sb", "short_sb"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        print("%-28s %6d %12d %6.2f %8d %8d %8d %8d" % ("%s:%d" % key if key else "?", v[6], v[0], 100.0 * v[0] / max(tot[0], 1), v[1], v[2], v[3], v[4]))


if __name__ == "__main__":
    main(*sys.argv[1:])
