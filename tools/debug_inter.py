#!/usr/bin/env python
"""GPU-side triage for the inter kernels: run seeded cases, compare with the oracle, and report
mismatches grouped by record kind (which flags / sizes / planes fail).  Test infrastructure."""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib  # noqa: E402
from tests import util  # noqa: E402
from tests.test_oracle_vs_ref_inter import STRESS_MIX, make_case, run_inter  # noqa: E402


def kind(r):
    f = int(r["flags"])
    names = []
    for bit, n in ((1, "dmvr"), (2, "bdof"), (4, "prof0"), (8, "prof1"), (16, "gpm"), (32, "wp")):
        if f & bit:
            names.append(n)
    if r["bcw_idx"]:
        names.append("bcw")
    return "%dx%d pl%d pf%d %s" % (r["w"], r["h"], r["planes"], r["pred_flag"], "+".join(names) or "plain")


def main():
    ctx = lib.Context(0)
    torch.cuda.set_stream(ctx.torch_stream())
    cases = [(416, 240, 1, False), (416, 240, 2, True), (136, 72, 5, True), (832, 480, 4, False)]
    for (w, h, seed, uniform) in cases:
        gd, gr, refs, pbs, wp, prof = make_case(w, h, seed, mix=STRESS_MIX, uniform=uniform)
        dst = device.DeviceFrames(gd, planes=abi.alloc_planes(gd, fill=77))
        ref = device.DeviceFrames(gr, planes=refs)
        t1, p1 = device.to_device(pbs)
        t2, p2 = device.to_device(wp)
        t3, p3 = device.to_device(prof)
        t4, p4 = device.to_device(np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE))
        ctx.inter_frame(dst.desc, ref.desc, p1, len(pbs), p2, p3, p4)
        ctx.sync()
        got, go = dst.to_numpy(), t4.cpu().numpy().view(abi.DMVR_OUT_DTYPE)
        od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
        bad, good = collections.Counter(), collections.Counter()
        shown = 0
        for i, r in enumerate(pbs):
            k = kind(r)
            ok = True
            for c in range(3):
                if not (r["planes"] & (2 if c else 1)):
                    continue
                sh = 1 if c else 0
                x0, y0, bw, bh = r["x0"] >> sh, r["y0"] >> sh, r["w"] >> sh, r["h"] >> sh
                a = got[c][r["pic"], y0:y0 + bh, x0:x0 + bw].astype(int)
                b = od[c][r["pic"], y0:y0 + bh, x0:x0 + bw].astype(int)
                if not np.array_equal(a, b):
                    ok = False
                    if shown < 12:
                        shown += 1
                        yy, xx = np.argwhere(a != b)[0]
                        print("  rec %d %s plane %d at (%d,%d) mv=%s: %d of %d differ, first (%d,%d) got %d want %d" % (
                            i, k, c, r["x0"], r["y0"], r["mv"].tolist(), int((a != b).sum()), a.size, xx, yy, a[yy, xx], b[yy, xx]))
            if (r["flags"] & 1) and (r["planes"] & 1) and go[i] != oo[i]:
                ok = False
                if shown < 12:
                    shown += 1
                    print("  rec %d %s dmvr_out got %s want %s" % (i, k, go[i], oo[i]))
            (good if ok else bad)[k] += 1
        print("case %dx%d seed %d: %d records, %d bad" % (w, h, seed, len(pbs), sum(bad.values())))
        for k in sorted(set(bad) | set(good)):
            if bad[k]:
                print("   BAD %-40s %5d of %5d" % (k, bad[k], bad[k] + good[k]))


if __name__ == "__main__":
    main()
