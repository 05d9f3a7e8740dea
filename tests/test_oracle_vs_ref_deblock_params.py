"""Pin the oracle's derivation of the deblocking parameters (boundary strengths, maximum filter lengths, QP -> beta / tc,
LADF) against the compiled reference's own drivers ff_vvc_deblock_vertical / _horizontal, through the pictures they
produce (CPU only)."""
import ctypes as C

import numpy as np
import pytest

from ffvvc_b200 import abi, synth
from tests import util


def empty_maps(geom):
    return [[np.zeros((geom.batch,) + abi.deblock_map_shape(geom, d, c), dtype=abi.DBK_EDGE_DTYPE) for c in range(3)] for d in range(2)]


def derive_and_filter(lib_params, geom, planes, tus, mvfs, ctbs, prm):
    """V parameters from the picture, V pass, H parameters from its output (LADF reads the pass's input), H pass."""
    o = util.oracle()
    maps = empty_maps(geom)
    md = abi.deblock_maps_desc(geom, maps)
    a, b = abi.alloc_planes(geom), abi.alloc_planes(geom)
    lib_params(abi.frame_from_numpy(geom, planes), tus, mvfs, ctbs, prm, md, 1)
    o.vvco_deblock_frame(abi.frame_from_numpy(geom, a), abi.frame_from_numpy(geom, planes), C.byref(md), 1)
    lib_params(abi.frame_from_numpy(geom, a), tus, mvfs, ctbs, prm, md, 0)
    o.vvco_deblock_frame(abi.frame_from_numpy(geom, b), abi.frame_from_numpy(geom, a), C.byref(md), 0)
    return b, maps


def oracle_params(frame, tus, mvfs, ctbs, prm, md, d):
    util.oracle().vvco_deblock_params_frame(frame, tus.ctypes.data, len(tus), mvfs.ctypes.data, len(mvfs), ctbs.ctypes.data, C.byref(prm), C.byref(md), d)


def smooth_planes(geom, seed):
    """gentle ramps with small steps at the 4x4 grid and +-1 noise: the long and strong filters are chosen on most edges,
    so the maximum filter lengths and the side decisions show in the samples"""
    rng = synth.LCG(seed)
    planes = abi.alloc_planes(geom)
    for c, p in enumerate(planes):
        b, hh, ww = p.shape
        y, x = np.mgrid[0:hh, 0:ww]
        base = 300 + (x + 2 * y) // 3 + 3 * ((x // 4 + y // 4) % 3)
        noise = rng.below(b * hh * ww, 3).reshape(b, hh, ww).astype(np.int64) - 1
        p[:] = np.clip(base[None] + noise, 0, (1 << geom.bit_depth) - 1)
    return planes


@pytest.mark.parametrize("w,h,batch,seed,bd,ctb_log2,ladf,smooth", [(416, 240, 2, 1, 10, 7, True, False), (256, 192, 2, 2, 10, 6, False, True),
                                                                    (832, 480, 1, 3, 10, 7, True, True), (200, 136, 2, 4, 12, 5, True, False),
                                                                    (1920, 1080, 1, 5, 10, 7, True, False), (416, 240, 3, 6, 10, 7, False, True)])
def test_derived_parameters_filter_like_the_reference(w, h, batch, seed, bd, ctb_log2, ladf, smooth):
    geom = abi.FrameGeom(w, h, batch=batch, bit_depth=bd, ctb_log2=ctb_log2)
    planes = smooth_planes(geom, seed + 30) if smooth else synth.struct_planes(geom, seed=seed + 30)
    tus, mvfs, ctbs, prm = synth.deblock_side_info(geom, seed=seed, ladf=ladf)
    got, maps = derive_and_filter(oracle_params, geom, planes, tus, mvfs, ctbs, prm)
    want = [p.copy() for p in planes]
    util.ref().vvcref_deblock_params_filter(abi.frame_from_numpy(geom, want), tus.ctypes.data, len(tus), mvfs.ctypes.data, len(mvfs),
                                            ctbs.ctypes.data, C.byref(prm))
    util.assert_planes_equal(geom, got, want, "oracle parameters + oracle filter vs the reference's drivers")
    assert not np.array_equal(want[0], planes[0]) and not np.array_equal(want[1], planes[1])
    # every kind of decision occurred
    ml = maps[0][0]["max_len"][maps[0][0]["tc"] > 0]
    assert set(np.unique(ml & 15)) >= {1, 3, 7} and (maps[0][0]["tc"] == 0).any()
    assert (tus["cu_flags"] & abi.DBK_CU_SUBBLOCK).any() and (mvfs["pred_flag"] == 0).any() and (mvfs["pred_flag"] == 3).any()
