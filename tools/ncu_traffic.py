#!/usr/bin/env python
"""DRAM bytes per stage and picture from an ncu summary written by tools/ncu_summary.py (one whole reconstruction pass of
tools/profile_recon.py N 1): the source of bench.py's `roofline.traffic`.
Usage: ncu_traffic.py profiles/r02_recon16_ncu_full.csv 16 > profiles/r02_traffic.json"""
import csv
import json
import sys

path, pictures = sys.argv[1], int(sys.argv[2])
rows = list(csv.DictReader(open(path)))
rk = next(k for k in rows[0] if "dram__bytes_read" in k)
wk = next(k for k in rows[0] if "dram__bytes_write" in k)
tk = next(k for k in rows[0] if "gpu__time_duration" in k)


def stage_of(name):
    for key, st in (("inter_", "inter"), ("itx_", "residual"), ("lmcs", "lmcs"), ("deblock_kernel<1>", "deblock_v"),
                    ("deblock_kernel<0>", "deblock_h"), ("sao", "sao"), ("alf", "alf")):
        if key in name:
            return st
    return None


per = {}
for r in rows:
    st = stage_of(r["Kernel Name"])
    if st is None:
        continue
    p = per.setdefault(st, {"dram_read_mb_per_picture": 0.0, "dram_write_mb_per_picture": 0.0, "us_per_launch_under_ncu": 0.0})
    p["dram_read_mb_per_picture"] += float(r[rk]) / pictures
    p["dram_write_mb_per_picture"] += float(r[wk]) / pictures
    p["us_per_launch_under_ncu"] += float(r[tk])
for p in per.values():
    for k in p:
        p[k] = round(p[k], 2)
json.dump({"source": "%s: ncu --set full --clock-control none --import-source on of one whole reconstruction pass of "
                     "tools/profile_recon.py %d 1 (%d x 4K pictures per launch as in the bench, working set far above the 126 MB L2; "
                     "quantised levels in the 16-bit window layout), dram__bytes_read.sum / dram__bytes_write.sum per kernel, summed per "
                     "stage, divided by the pictures of the launch" % (path, pictures, pictures),
           "per_stage": per}, sys.stdout, indent=1)
