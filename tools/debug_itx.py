#!/usr/bin/env python
"""GPU-side triage for the residual kernels: seeded TB lists vs the oracle, mismatches grouped by TB kind."""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffvvc_b200 import abi, device, lib, synth  # noqa: E402
from tests import util  # noqa: E402


def main():
    ctx = lib.Context(0)
    torch.cuda.set_stream(ctx.torch_stream())
    for (w, h, seed) in [(416, 240, 1), (832, 480, 3)]:
        geom = abi.FrameGeom(w, h)
        tbs, coeffs = synth.tb_list(geom, seed=seed, lfnst_set_of=util.oracle().vvco_lfnst_tr_set)
        pred = synth.uniform_planes(geom, seed=seed + 10)
        fr = device.DeviceFrames(geom, planes=pred)
        t1, p1 = device.to_device(coeffs)
        t2, p2 = device.to_device(tbs)
        ctx.itx_frame(fr.desc, p1, p2, len(tbs), 15)
        ctx.sync()
        got, gc = fr.to_numpy(), t1.cpu().numpy().view(np.int32)
        planes = [p.copy() for p in pred]
        co = coeffs.copy()
        util.oracle().vvco_itx_frame(abi.frame_from_numpy(geom, planes), co.ctypes.data, tbs.ctypes.data, len(tbs), 15)
        bad, good, shown = collections.Counter(), collections.Counter(), 0
        for i, t in enumerate(tbs):
            tw, th = 1 << int(t["log2_w"]), 1 << int(t["log2_h"])
            k = "%dx%d c%d tr%d%d fl%d lf%d" % (tw, th, t["c_idx"], t["trh"], t["trv"], t["flags"], 1 if t["lfnst"] else 0)
            if t["flags"] & abi.TB_STORE_RESIDUAL:
                o = int(t["coeff_offset"])
                a, b = gc[o:o + tw * th], co[o:o + tw * th]
            else:
                c = int(t["c_idx"])
                a = got[c][t["pic"], t["y0"]:t["y0"] + th, t["x0"]:t["x0"] + tw].astype(int).reshape(-1)
                b = planes[c][t["pic"], t["y0"]:t["y0"] + th, t["x0"]:t["x0"] + tw].astype(int).reshape(-1)
            ok = np.array_equal(a, b)
            if not ok and shown < 10:
                shown += 1
                j = int(np.argwhere(a != b)[0][0])
                print("  tb %d %s nz %dx%d: %d of %d differ, first idx %d got %d want %d" % (i, k, t["nzw"], t["nzh"], int((a != b).sum()), a.size, j, a[j], b[j]))
            (good if ok else bad)[k] += 1
        print("case %dx%d seed %d: %d TBs, %d bad" % (w, h, seed, len(tbs), sum(bad.values())))
        for k in sorted(bad):
            print("   BAD %-36s %5d of %5d" % (k, bad[k], bad[k] + good[k]))


if __name__ == "__main__":
    main()
