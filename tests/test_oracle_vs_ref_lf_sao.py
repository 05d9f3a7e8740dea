"""Pin the oracle's deblocking and SAO restatements against the compiled reference (CPU only)."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util

GEOMS = [(256, 128, 7), (200, 136, 7), (416, 240, 7), (176, 144, 6), (128, 64, 5)]


def deblock_both(lib_fn_o, lib_fn_r, geom, planes, maps):
    src = abi.frame_from_numpy(geom, planes)
    desc = abi.deblock_maps_desc(geom, maps)
    res = []
    for fn in (lib_fn_o, lib_fn_r):
        tmp = abi.alloc_planes(geom)
        out = abi.alloc_planes(geom)
        fn(abi.frame_from_numpy(geom, tmp), src, C.byref(desc), 1)
        fn(abi.frame_from_numpy(geom, out), abi.frame_from_numpy(geom, tmp), C.byref(desc), 0)
        res.append((tmp, out))
    return res


@pytest.mark.parametrize("w,h,ctb_log2", GEOMS)
def test_deblock_frame_matches_reference(w, h, ctb_log2):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = synth.struct_planes(geom, seed=w + h)
    maps = synth.deblock_maps(geom, seed=w * 3 + h)
    (v_o, h_o), (v_r, h_r) = deblock_both(util.oracle().vvco_deblock_frame, util.ref().vvcref_deblock_frame, geom, planes, maps)
    util.assert_planes_equal(geom, v_o, v_r, "vertical pass oracle vs reference")
    util.assert_planes_equal(geom, h_o, h_r, "horizontal pass oracle vs reference")
    changed = sum(int((a != b).sum()) for a, b in zip(util.visible(geom, h_o), util.visible(geom, planes)))
    assert changed > 0.01 * w * h, "deblocking decisions never fired: %d samples changed" % changed


def test_deblock_1080p_and_filter_mix():
    geom = abi.FrameGeom(1920, 1080)
    planes = synth.struct_planes(geom, seed=5)
    maps = synth.deblock_maps(geom, seed=6, qp_base=27, qp_span=16)
    (v_o, h_o), (v_r, h_r) = deblock_both(util.oracle().vvco_deblock_frame, util.ref().vvcref_deblock_frame, geom, planes, maps)
    util.assert_planes_equal(geom, v_o, v_r, "vertical")
    util.assert_planes_equal(geom, h_o, h_r, "horizontal")


def test_deblock_uniform_noise_and_12bit():
    for bd in (10, 12):
        geom = abi.FrameGeom(320, 192, bit_depth=bd)
        planes = synth.uniform_planes(geom, seed=bd)
        maps = synth.deblock_maps(geom, seed=bd, qp_base=40, qp_span=24)
        (v_o, h_o), (v_r, h_r) = deblock_both(util.oracle().vvco_deblock_frame, util.ref().vvcref_deblock_frame, geom, planes, maps)
        util.assert_planes_equal(geom, v_o, v_r, "vertical bd%d" % bd)
        util.assert_planes_equal(geom, h_o, h_r, "horizontal bd%d" % bd)


def sao_both(geom, planes, params):
    src = abi.frame_from_numpy(geom, planes)
    o = abi.alloc_planes(geom, fill=0xAAAA)
    r = abi.alloc_planes(geom, fill=0x5555)
    util.oracle().vvco_sao_frame(abi.frame_from_numpy(geom, o), src, params.ctypes.data)
    util.ref().vvcref_sao_frame(abi.frame_from_numpy(geom, r), src, params.ctypes.data)
    return o, r


@pytest.mark.parametrize("w,h,ctb_log2", GEOMS + [(1920, 1080, 7)])
@pytest.mark.parametrize("restore", [False, True])
def test_sao_frame_matches_reference(w, h, ctb_log2, restore):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = synth.uniform_planes(geom, seed=w ^ h) if (w & 16) else synth.struct_planes(geom, seed=w ^ h)
    params = synth.sao_params(geom, seed=h, with_restore=restore)
    o, r = sao_both(geom, planes, params)
    util.assert_planes_equal(geom, o, r, "SAO oracle vs reference")
