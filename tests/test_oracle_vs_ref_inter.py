"""Pins the oracle's inter prediction stage against the compiled reference: the reference's own
table entries (put*, avg, w_avg, put_gpm, dmvr, sad, apply_bdof, apply_prof*, *fetch_samples),
emulated_edge_mc and ff_vvc_clip_mv, driven in the order of libavcodec/vvc/vvc_inter.c
(oracle/refbuild/ref_glue_inter.c)."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util

STRESS_MIX = dict(bi=70, dmvr=35, bdof_only=15, bcw=15, wp=15, affine=20, gpm=15)


def run_inter(fn, geom_dst, geom_ref, refs, pbs, wp, prof):
    dst = abi.alloc_planes(geom_dst, fill=77)
    out = np.zeros(len(pbs), dtype=abi.DMVR_OUT_DTYPE)
    fn(abi.frame_from_numpy(geom_dst, dst), abi.frame_from_numpy(geom_ref, refs), pbs.ctypes.data, len(pbs),
       wp.ctypes.data, prof.ctypes.data, out.ctypes.data)
    return dst, out


def make_case(w, h, seed, mix=None, uniform=False, bit_depth=10, n_refs=3, batch=1):
    gd = abi.FrameGeom(w, h, batch=batch, bit_depth=bit_depth)
    gr = abi.FrameGeom(w, h, batch=n_refs, bit_depth=bit_depth)
    refs = (synth.uniform_planes if uniform else synth.struct_planes)(gr, seed=seed)
    pbs, wp, prof = synth.pb_list(gd, n_refs=n_refs, seed=seed + 1, mix=mix)
    return gd, gr, refs, pbs, wp, prof


@pytest.mark.ref
@pytest.mark.parametrize("w,h,seed,uniform", [(416, 240, 1, False), (416, 240, 2, True), (256, 192, 3, False),
                                               (832, 480, 4, False), (136, 72, 5, True)])
def test_inter_stage_matches_reference(w, h, seed, uniform):
    gd, gr, refs, pbs, wp, prof = make_case(w, h, seed, mix=STRESS_MIX if seed != 4 else None, uniform=uniform)
    assert len(pbs) > 50
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    rd, ro = run_inter(util.ref().vvcref_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, od, rd, "oracle vs reference")
    dm = (pbs["flags"] & abi.PB_DMVR) != 0
    assert dm.any()
    assert np.array_equal(oo[dm], ro[dm]), "DMVR outputs differ"
    # the search really moved vectors and really switched BDOF off somewhere
    if not uniform:
        assert (oo["mv"][dm] != pbs["mv"][dm]).any()


@pytest.mark.ref
def test_inter_stage_12bit():
    gd, gr, refs, pbs, wp, prof = make_case(256, 128, 9, mix=STRESS_MIX, bit_depth=12)
    od, oo = run_inter(util.oracle().vvco_inter_frame, gd, gr, refs, pbs, wp, prof)
    rd, ro = run_inter(util.ref().vvcref_inter_frame, gd, gr, refs, pbs, wp, prof)
    util.assert_planes_equal(gd, od, rd, "oracle vs reference, 12 bit")
    assert np.array_equal(oo, ro)
