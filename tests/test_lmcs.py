"""LMCS: oracle vs compiled reference (CPU) and CUDA vs oracle (GPU)."""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def rects_for(geom, seed):
    rng = synth.LCG(seed)
    n = 64
    r = np.zeros(n, dtype=abi.RECT_DTYPE)
    r["w"] = 4 << rng.below(n, 5)
    r["h"] = 4 << rng.below(n, 5)
    r["x"] = (rng.below(n, 1 << 16) % np.maximum((geom.width - r["w"]) // 4, 1)) * 4
    r["y"] = (rng.below(n, 1 << 16) % np.maximum((geom.height - r["h"]) // 4, 1)) * 4
    r["pic"] = rng.below(n, geom.batch)
    return r


@pytest.mark.parametrize("w,h,bd", [(416, 240, 10), (200, 136, 10), (256, 128, 12)])
def test_lmcs_oracle_matches_reference(w, h, bd):
    geom = abi.FrameGeom(w, h, bit_depth=bd, batch=2)
    fwd, inv = synth.lmcs_luts(bd)
    planes = synth.uniform_planes(geom, seed=w)
    enable = (synth.LCG(1).below(geom.ctb_count * 2, 4) > 0).astype(np.uint8)
    rects = rects_for(geom, 3)
    a, b = [p.copy() for p in planes], [p.copy() for p in planes]
    util.oracle().vvco_lmcs_frame(abi.frame_from_numpy(geom, a), inv.ctypes.data, enable.ctypes.data)
    util.ref().vvcref_lmcs_frame(abi.frame_from_numpy(geom, b), inv.ctypes.data, enable.ctypes.data)
    util.assert_planes_equal(geom, a, b, "inverse LUT per CTU")
    assert not np.array_equal(a[0], planes[0])
    util.oracle().vvco_lmcs_rects(abi.frame_from_numpy(geom, a), fwd.ctypes.data, rects.ctypes.data, len(rects))
    util.ref().vvcref_lmcs_rects(abi.frame_from_numpy(geom, b), fwd.ctypes.data, rects.ctypes.data, len(rects))
    util.assert_planes_equal(geom, a, b, "forward LUT on rectangles")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bd", [(416, 240, 10), (200, 136, 10), (256, 128, 12), (3840, 2160, 10)])
def test_lmcs_cuda_bit_exact(w, h, bd):
    import torch
    from ffvvc_b200 import device, lib
    geom = abi.FrameGeom(w, h, bit_depth=bd, batch=2 if w < 1000 else 1)
    fwd, inv = synth.lmcs_luts(bd)
    planes = synth.uniform_planes(geom, seed=w)
    enable = (synth.LCG(1).below(geom.ctb_count * geom.batch, 4) > 0).astype(np.uint8)
    rects = rects_for(geom, 3)
    want = [p.copy() for p in planes]
    util.oracle().vvco_lmcs_frame(abi.frame_from_numpy(geom, want), inv.ctypes.data, enable.ctypes.data)
    ctx = lib.Context(0)
    with torch.cuda.stream(ctx.torch_stream()):
        fr = device.DeviceFrames(geom, planes=planes)
        t1, p1 = device.to_device(inv)
        t2, p2 = device.to_device(enable)
        t3, p3 = device.to_device(fwd)
        t4, p4 = device.to_device(rects)
        ctx.lmcs_frame(fr.desc, p1, p2)
        ctx.sync()
        util.assert_planes_equal(geom, fr.to_numpy(), want, "cuda vs oracle (frame)")
        # overlapping rectangles would race: keep the disjoint ones of picture order
        util.oracle().vvco_lmcs_rects(abi.frame_from_numpy(geom, want), fwd.ctypes.data, rects[:1].ctypes.data, 1)
        ctx.lmcs_rects(fr.desc, p3, p4, 1)
        ctx.sync()
        util.assert_planes_equal(geom, fr.to_numpy(), want, "cuda vs oracle (rect)")
        host = [p.copy() for p in planes]
        ctx.lmcs_frame_host(abi.frame_from_numpy(geom, host), inv.ctypes.data, enable.ctypes.data)
        want2 = [p.copy() for p in planes]
        util.oracle().vvco_lmcs_frame(abi.frame_from_numpy(geom, want2), inv.ctypes.data, enable.ctypes.data)
        util.assert_planes_equal(geom, host, want2, "host entry")
    ctx.close()
