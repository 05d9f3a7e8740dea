"""Pin the oracle's ALF restatement against the compiled reference (CPU only).

Reference side = loop of the reference's own table entries in the reference driver's order
(oracle/refbuild/ref_glue_alf.c); oracle side = oracle/src/alf.c.
"""
import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def run_both(geom, planes, ctbs, sets):
    src = abi.frame_from_numpy(geom, planes)
    out_o = abi.alloc_planes(geom, fill=0xFFFF)
    out_r = abi.alloc_planes(geom, fill=0xFFFF)
    util.oracle().vvco_alf_frame(abi.frame_from_numpy(geom, out_o), src, ctbs.ctypes.data, sets.ctypes.data, 0)
    util.ref().vvcref_alf_frame(abi.frame_from_numpy(geom, out_r), src, ctbs.ctypes.data, sets.ctypes.data, 0)
    return out_o, out_r


@pytest.mark.parametrize("w,h,ctb_log2", [(256, 128, 7), (200, 136, 7), (416, 240, 7), (176, 144, 6), (128, 64, 5)])
@pytest.mark.parametrize("dist", ["uniform", "struct"])
def test_alf_frame_matches_reference(w, h, ctb_log2, dist):
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2)
    planes = (synth.uniform_planes if dist == "uniform" else synth.struct_planes)(geom, seed=w * 7 + h)
    ctbs, sets = synth.alf_params(geom, seed=h + ctb_log2, all_on=(dist == "uniform"))
    out_o, out_r = run_both(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, out_o, out_r, "oracle vs reference")
    if dist == "uniform":  # every CTB flag on: the filter must have done something
        assert not np.array_equal(util.visible(geom, out_o)[0], util.visible(geom, planes)[0])


def test_alf_frame_interior_edges_and_batch():
    """Slice/tile style CTB edges in the interior + a 2-picture ring."""
    geom = abi.FrameGeom(384, 256, batch=2)
    planes = synth.struct_planes(geom, seed=99)
    ctbs, sets = synth.alf_params(geom, seed=5)
    ctbs["edges"] = synth.LCG(3).below(len(ctbs), 16)
    out_o, out_r = run_both(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, out_o, out_r, "oracle vs reference")


def test_alf_1080p_frame():
    geom = abi.FrameGeom(1920, 1080)
    planes = synth.struct_planes(geom, seed=12345)
    ctbs, sets = synth.alf_params(geom, seed=12345)
    out_o, out_r = run_both(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, out_o, out_r, "oracle vs reference")


@pytest.mark.parametrize("coeffs", ["full", "max"])
@pytest.mark.parametrize("w,h,ctb_log2,bd", [(256, 128, 7, 10), (176, 144, 6, 10), (128, 64, 5, 10), (256, 128, 7, 12)])
def test_alf_full_coefficient_range(w, h, ctb_log2, bd, coeffs):
    """Coefficients over the whole legal range -128..+128 (cbs_h266_syntax_template.c:2285,2314), every APS slot."""
    geom = abi.FrameGeom(w, h, ctb_log2=ctb_log2, bit_depth=bd)
    planes = synth.uniform_planes(geom, seed=w + h)
    ctbs, sets = synth.alf_params(geom, seed=h + ctb_log2, coeffs=coeffs)
    assert sets["luma_coeff"].max() == 128 and sets["chroma_coeff"].max() == 128 and ctbs["filt_set_idx_y"].max() > 16
    out_o, out_r = run_both(geom, planes, ctbs, sets)
    util.assert_planes_equal(geom, out_o, out_r, "oracle vs reference")
